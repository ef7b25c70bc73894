"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (ctypes), against the FP64 oracle on
the same inputs, against the committed README goldens, and -- at BASELINE.json's full size -- through
size-independent properties.  Tolerances (BASELINE.json north_star): sample indices bit-exact;
eigenvalues within 1e-5 relative; eigenvectors up to sign / subspace rotation; output image within
1 LSB on >= 99.9 % of pixels (the goldens themselves are reproduced by the oracle to <= 2 LSB)."""
import ctypes as C
import hashlib
import threading

import numpy as np
import pytest

from nle_testlib import load_case, manifest, oracle_sq_spread, oracle_stages, synth_lum, train_args
from oracle import nle_oracle as O

pytestmark = pytest.mark.gpu

SQ_RTOL = 1e-5          # north_star: eigenvalues within 1e-5 relative
SQ_SPREAD_FACTOR = 10.0  # ill-conditioned inputs only: accept a backward error of 10 eps ||Ka|| (LAPACK documents p(n) eps, p(n) = O(n))
PIX_FRAC = 0.999        # north_star: within 1 LSB on >= 99.9 % of pixels


def sq_close(S, S_ref, spread=None):
    """north_star: every eigenvalue within 1e-5 relative.  (Round 1 relaxed this below 1e-4 * max; profiles/sq_conditioning.md
    shows that two FP64 evaluation orders of the reference's own algebra agree to < 1e-6 on the README images, so no blanket
    relaxation.)  `spread` (nle_testlib.oracle_sq_spread: the per-eigenvalue movement of the reference algebra at THIS input
    under other LAPACK eigensolvers, the factor-form evaluation order and eps * ||Ka|| perturbations of Ka) widens the bound to
    SQ_SPREAD_FACTOR x spread only for eigenvalues whose spread itself is that large -- the hx = 5000, hy = 100 corner of the
    sweep, where LAPACK's QR-iteration and MRRR solvers differ by 3.8e-4 on the same FP64 matrices."""
    S, S_ref = np.asarray(S), np.asarray(S_ref)
    tol = np.full(S_ref.shape, SQ_RTOL) if spread is None else np.maximum(SQ_RTOL, SQ_SPREAD_FACTOR * np.asarray(spread))
    return S.shape == S_ref.shape and np.all(np.abs(S - S_ref) <= tol * np.abs(S_ref))


def subspace_gap(Va, Vb):
    """sin of the largest principal angle between the column spaces (Va, Vb need not be orthonormal)."""
    Qa, _ = np.linalg.qr(Va)
    Qb, _ = np.linalg.qr(Vb)
    s = np.linalg.svd(Qa.T @ Qb, compute_uv=False)
    return float(np.sqrt(max(0.0, 1.0 - s.min() ** 2)))


# ---- (1) sampling: bit-exact -----------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(736, 491, 20, 10), (100, 100, 40, 40), (7, 5, 7, 5), (33, 17, 4, 3),
                                   (267, 400, 10, 20), (50, 64, 1, 1), (9, 9, 9, 1), (1024, 1024, 40, 40),
                                   (584, 876, 50, 50)])
def test_sample_indices_bit_exact(nb, shape):
    sel, rest = nb.sampleIndices(*shape)
    so, ro = O.sample_pixels(*shape)
    assert np.array_equal(sel, so.astype(np.int32))
    assert np.array_equal(rest, ro.astype(np.int32))


def test_sample_indices_random_shapes(nb):
    rng = np.random.default_rng(5)
    for _ in range(25):
        rows, cols = int(rng.integers(1, 120)), int(rng.integers(1, 120))
        a, b = int(rng.integers(1, rows + 1)), int(rng.integers(1, cols + 1))
        sel, rest = nb.sampleIndices(rows, cols, a, b)
        so, ro = O.sample_pixels_loop(rows, cols, a, b)
        assert np.array_equal(sel, so) and np.array_equal(rest, ro), (rows, cols, a, b)


def test_too_many_samples_is_the_reference_error(nb):
    with pytest.raises(nb.NleError, match="Number of samples per row and col must be <= that of image."):
        nb.sampleIndices(10, 10, 11, 1)


# ---- eigenDecomposition (test_filter.cpp:42-68) ------------------------------------------------------
def test_eigen_decomposition_known_answer(nb):
    R = np.array([[2., -1, 0], [-1, 2, -1], [0, -1, 2]])
    U, D = nb.eigenDecomposition(R, 1e-10)
    assert np.linalg.norm(D - [3.41421356, 2., 0.58578644]) <= 1e-5 * np.linalg.norm(D)
    assert np.linalg.norm((U * D) @ U.T - R) <= 1e-10 * np.linalg.norm(R)
    assert np.linalg.norm(U.T @ U - np.eye(3)) <= 1e-10 * np.sqrt(3)


def test_eigen_decomposition_lower_triangle_and_truncation(nb):
    U, D = nb.eigenDecomposition(np.array([[2., 99.], [1., 2.]]))
    assert np.allclose(D, [3., 1.], atol=1e-12)
    U, D = nb.eigenDecomposition(np.diag([1., 1e-11, -3.]))
    assert D.size == 1 and U.shape == (3, 1) and abs(D[0] - 1) < 1e-12


@pytest.mark.parametrize("n", [1, 2, 5, 16, 17, 64, 130, 300])
def test_eigen_decomposition_random_symmetric(nb, n):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n)); A = (A + A.T) / 2           # indefinite: exercises the shift fallback
    U, D = nb.eigenDecomposition(A, eps=-1e300)
    w = np.linalg.eigvalsh(A)[::-1]
    scale = max(1.0, np.abs(w).max())
    assert D.size == n and np.abs(D - w).max() <= 1e-11 * scale * max(1, n / 16)
    assert np.abs(U.T @ U - np.eye(n)).max() <= 1e-11
    assert np.abs(A @ U - U * D).max() <= 1e-10 * scale


def test_eigen_rank_cut_matches_lapack_on_bench_Ka(nb):
    """The 1e-10 prefix cut (filter.cpp:213-216) on the p=1600 Ka of the bench workload."""
    from bench import synth_luminance
    lum = synth_luminance(1024, 1024).astype(np.float64)
    sel, _ = O.sample_pixels(1024, 1024, 40, 40)
    Ka = O.affinity_block(lum.ravel(), 1024, sel, sel, 500.0, 30.0)
    Uo, Do = O.eigen_decomposition(Ka)
    U, D = nb.eigenDecomposition(Ka)
    assert D.size == Do.size
    assert np.abs(D - Do).max() <= 1e-11
    big = Do > 1e-6
    assert np.all(np.abs(D[big] - Do[big]) <= 1e-9 * Do[big])
    # spectral projector of the kept subspace
    assert np.abs(U @ (U.T @ Ka[:, :8]) - Uo @ (Uo.T @ Ka[:, :8])).max() < 1e-7


# ---- computeKernel / nystromApproximation --------------------------------------------------------------
def test_compute_kernel_matches_oracle(nb):
    L = synth_lum(40, 56).astype(np.float64)
    perm, Ka, Kab = nb.computeKernel(L, 5, 7, 30.0, 12.0)
    po, Kao, Kabo = O.compute_kernel(L, 5, 7, 30.0, 12.0)
    assert np.array_equal(perm, po)
    assert np.abs(Ka - Kao).max() <= 4e-16
    # product of three table exponentials vs one exponential of the summed argument: |arg| * eps relative
    assert np.abs(Kab - Kabo).max() <= 1e-15 and np.all(np.abs(Kab - Kabo) <= 2e-13 * Kabo + 1e-300)


def test_compute_kernel_rejects_non_integer_luminance(nb):
    L = synth_lum(20, 20).astype(np.float64) + 0.5
    with pytest.raises(nb.NleError) as e:
        nb.computeKernel(L, 4, 4, 10.0, 10.0)
    assert e.value.code == -3


def test_nystrom_approximation_matches_oracle(nb):
    L = synth_lum(24, 32).astype(np.float64)
    _, Ka, Kab = O.compute_kernel(L, 4, 6, 15.0, 20.0)
    lam, phi = nb.nystromApproximation(Ka, Kab)
    lo, po = O.nystrom_approximation(Ka, Kab)
    assert lam.size == lo.size and np.allclose(lam, lo, rtol=1e-9, atol=1e-14)
    assert np.abs((phi * lam) @ phi.T - (po * lo) @ po.T).max() < 1e-8          # sign-invariant


# ---- sinkhorn / orthogonalize (test_filter.cpp:70-153) ------------------------------------------------
def test_sinkhorn_identity(nb):
    Wa, Wab = nb.sinkhorn(np.eye(2), np.ones(2), 10)
    assert Wab.shape == (2, 0)
    assert np.allclose(Wa, Wa.T) and np.allclose(Wa.sum(0), 1, atol=1e-10) and np.allclose(Wa.sum(1), 1, atol=1e-10)


def test_sinkhorn_matches_oracle_on_random_factors(nb):
    rng = np.random.default_rng(3)
    R = rng.uniform(0, 1, (5, 5))
    U, D = O.eigen_decomposition(R, 1e-10)
    Wa, Wab = nb.sinkhorn(U, D, 20)
    Wao, Wabo, _, _ = O.sinkhorn(U, D, 20)
    assert np.allclose(Wa, Wao, atol=1e-12) and np.allclose(Wab, Wabo, atol=1e-12)
    rng = np.random.default_rng(4)
    phi = np.abs(rng.standard_normal((60, 4))) + 0.1
    lam = np.array([3.0, 1.0, 0.5, 0.1])
    Wa, Wab = nb.sinkhorn(phi, lam, 15)
    Wao, Wabo, _, _ = O.sinkhorn(phi, lam, 15)
    assert np.allclose(Wa, Wao, rtol=1e-10, atol=1e-14) and np.allclose(Wab, Wabo, rtol=1e-10, atol=1e-14)
    W = np.hstack([Wa, Wab])
    assert np.allclose(W.sum(1), 1, atol=1e-8)


def test_orthogonalize_properties_and_oracle(nb):
    rng = np.random.default_rng(11)
    p, n, k = 10, 100, 5
    Wa = rng.uniform(0, 1, (p, p)); Wa = (Wa + Wa.T) / 2
    Wab = rng.uniform(0, 1, (p, n - p))
    V, S = nb.orthogonalize(Wa, Wab, k)
    Vo, So, _ = O.orthogonalize(Wa, Wab, k)
    assert S.size == V.shape[1] == So.size > 0
    assert np.linalg.norm(V.T @ V - np.eye(S.size)) <= 1e-8
    assert np.allclose(S, So, rtol=1e-9)
    assert np.abs((V * S) @ V.T - (Vo * So) @ Vo.T).max() < 1e-8


# ---- trainFilter: stage-wise parity on small images ---------------------------------------------------
SMALL = [
    dict(shape=(48, 64), args=(6, 8, 20.0, 25.0, 5, 6)),        # full rank
    dict(shape=(40, 56), args=(5, 7, 300.0, 12.0, 8, 40)),      # wide spatial kernel, k > rank available
    dict(shape=(64, 40), args=(9, 6, 50.0, 60.0, 3, 4)),        # p=54, ragged grid
    dict(shape=(33, 47), args=(33, 2, 8.0, 10.0, 2, 3)),        # every row sampled
]


@pytest.mark.parametrize("case", SMALL)
def test_train_stage_parity_small(nb, case):
    L = synth_lum(*case["shape"])
    a = case["args"]
    f = nb.NLEFilter().trainFilter(L, *a)
    fo = O.train_dense(L.astype(np.float64), *a)
    st, inf = fo.stages, f.info()
    assert (inf.p, inf.r, inf.r2, inf.k) == (st["p"], st["r"], st["r2"], fo.eigvals.size)
    Ka = f.stage(0).reshape(inf.p, inf.p, order="F")
    assert np.abs(Ka - st["Ka"]).max() <= 4e-16
    assert np.allclose(f.stage(1), st["lam"], rtol=1e-6, atol=1e-13)
    c = f.stage(3); co = np.empty_like(c); co[st["perm"]] = st["c"]
    assert np.allclose(c, co, rtol=1e-6)
    assert np.allclose(f.stage(2), st["rvec_head"], rtol=1e-6)
    Wa = f.stage(4).reshape(inf.r, inf.r, order="F")
    assert np.abs(Wa - st["Wa"]).max() <= 1e-8 * np.abs(st["Wa"]).max()
    Q = f.stage(5).reshape(inf.r, inf.r, order="F")
    assert np.abs(Q - st["Q"]).max() <= 1e-6 * np.abs(st["Q"]).max()
    assert sq_close(f.eigvals, fo.eigvals)
    # eigenvectors: compare the filter operator V f(S) V^T on the image (sign / rotation invariant)
    w = [2.0, 3.0, 4.0, 1.0]
    out = f.apply(L.astype(np.float64), nb.transformEigenValues(f.eigvals, w))
    ref = O.apply(fo, L.astype(np.float64), O.transform_eigenvalues(fo.eigvals, w))
    assert np.abs(out - ref).max() <= 1e-4
    d = np.abs(f.enhanceLuminance(L, w).astype(int) - O.enhance_luminance(fo, L, w).astype(int))
    assert d.max() <= 1 and (d <= 1).mean() >= PIX_FRAC and (d == 0).mean() >= 0.99


# hx / hy sweep (BASELINE.json configs[4] asks for one): the spatial and photometric widths move the ranks r, r2 and the
# conditioning of every eigensolve by orders of magnitude; the rank cuts and the output must follow the FP64 oracle.
@pytest.mark.parametrize("hx", [5.0, 100.0, 5000.0])
@pytest.mark.parametrize("hy", [3.0, 10.0, 30.0, 100.0])
def test_hx_hy_sweep_matches_oracle(nb, hx, hy):
    L = synth_lum(72, 88, seed=21)
    a = (9, 11, hx, hy, 6, 12)
    f = nb.NLEFilter().trainFilter(L, *a)
    fo = O.train_dense(L.astype(np.float64), *a)
    st, inf = fo.stages, f.info()
    assert (inf.p, inf.r, inf.r2, inf.k) == (st["p"], st["r"], st["r2"], fo.eigvals.size)
    _, spread = oracle_sq_spread(L, a)
    assert spread is not None                          # the reference algebra agrees with itself on the rank cuts
    assert sq_close(f.eigvals, fo.eigvals, spread)     # 1e-5, wider only where the reference's own evaluations spread more
    if spread.max() <= 0.1 * SQ_RTOL:
        assert sq_close(f.eigvals, fo.eigvals)
    w = [2.0, 3.0, 4.0, 1.0]
    d = np.abs(f.enhanceLuminance(L, w).astype(int) - O.enhance_luminance(fo, L, w).astype(int))
    assert d.max() <= 1 and (d <= 1).mean() >= PIX_FRAC


# degenerate inputs: constant image (Ka = all ones up to the spatial factor), two-level image, single-column image,
# every pixel a sample (no rest pixels at all), one weight / five weights
@pytest.mark.parametrize("kind", ["constant", "two_level", "single_column", "all_samples", "saturated"])
def test_degenerate_images_follow_the_oracle(nb, kind):
    if kind == "constant":
        L = np.full((40, 36), 137, np.uint8); a = (4, 4, 30.0, 20.0, 5, 6)
    elif kind == "two_level":
        y, x = np.mgrid[0:48, 0:40]
        L = np.where(((x // 8) + (y // 6)) % 2 == 0, 40, 210).astype(np.uint8); a = (6, 5, 25.0, 15.0, 6, 8)
    elif kind == "single_column":
        L = synth_lum(64, 1, seed=5); a = (8, 1, 20.0, 25.0, 4, 4)
    elif kind == "all_samples":
        L = synth_lum(6, 7, seed=6); a = (6, 7, 10.0, 30.0, 3, 5)
    else:
        L = synth_lum(40, 44, seed=8); L[:20] = 0; L[30:] = 255; a = (5, 4, 15.0, 10.0, 4, 6)
    try:
        fo = O.train_dense(L.astype(np.float64), *a)
    except Exception as e:                      # the reference itself has no defined result here: the GPU path must fail loudly too
        with pytest.raises(nb.NleError):
            nb.NLEFilter().trainFilter(L, *a)
        return
    f = nb.NLEFilter().trainFilter(L, *a)
    st, inf = fo.stages, f.info()
    assert (inf.p, inf.r, inf.r2, inf.k) == (st["p"], st["r"], st["r2"], fo.eigvals.size)
    assert sq_close(f.eigvals, fo.eigvals)
    for w in ([2.0], [2.0, 3.0, 4.0, 1.0], [1.5, 2.0, 2.5, 3.0, 1.0]):
        d = np.abs(f.enhanceLuminance(L, w).astype(int) - O.enhance_luminance(fo, L, w).astype(int))
        assert d.max() <= 1 and (d <= 1).mean() >= PIX_FRAC


def test_train_accepts_float64_channel_like_the_reference(nb):
    L = synth_lum(32, 40)
    a = (4, 5, 25.0, 20.0, 4, 5)
    f8 = nb.NLEFilter().trainFilter(L, *a)
    f64 = nb.NLEFilter().trainFilter(L.astype(np.float64), *a)       # CV_64F channel, filter.cpp:466
    assert np.array_equal(f8.eigvals, f64.eigvals)
    with pytest.raises(nb.NleError) as e:
        nb.NLEFilter().trainFilter(L.astype(np.float64) + 0.25, *a)
    assert e.value.code == -3


def test_apply_is_linear_and_matches_eigvecs(nb):
    L = synth_lum(36, 44)
    f = nb.NLEFilter().trainFilter(L, 4, 4, 30.0, 20.0, 4, 6)
    V, S = f.eigvecs, f.eigvals
    rng = np.random.default_rng(0)
    z1, z2 = rng.uniform(0, 255, L.shape), rng.uniform(0, 255, L.shape)
    g = rng.uniform(0.5, 2.0, S.size)
    a1, a2, a12 = f.apply(z1, g), f.apply(z2, g), f.apply(2.0 * z1 - 3.0 * z2, g)
    assert np.allclose(a12, 2.0 * a1 - 3.0 * a2, atol=1e-9)
    ref = (V @ (g * (V.T @ z1.ravel()))).reshape(L.shape)                 # filter.cpp:456
    assert np.allclose(a1, ref, atol=1e-10)
    with pytest.raises(nb.NleError, match="Number of values in channel must match that of training image."):
        f.apply(np.zeros((5, 5)), g)


def test_enhance_rounding_and_clamp(nb):
    """max(.,0), min(.,255), convertTo(CV_8U) = round-half-to-even (filter.cpp:434-436)."""
    L = synth_lum(36, 44)
    f = nb.NLEFilter().trainFilter(L, 4, 4, 30.0, 20.0, 4, 6)
    for w in ([2.0, 3.0, 4.0, 1.0], [40.0, 1.0], [-5.0, 2.0, 1.0]):
        ref = O.clamp_round_u8(f.apply(L.astype(np.float64), nb.transformEigenValues(f.eigvals, w)))
        got = f.enhanceLuminance(L, w)
        assert np.array_equal(got, ref)


def test_image_level_errors_mirror_the_reference(nb):
    import cv2
    L = synth_lum(32, 32)
    img = cv2.cvtColor(L, cv2.COLOR_GRAY2BGR)
    f = nb.NLEFilter()
    f.trainForEnhancement(img, 4, 4, 20.0, 20.0, 3, 3)
    with pytest.raises(nb.NleError, match="Can only enhance RGB image."):
        f.enhance(L, [1, 2])
    with pytest.raises(nb.NleError, match="Cannot apply filter on image with different size"):
        f.enhance(img[:16], [1, 2])
    with pytest.raises(nb.NleError, match="Number of samples per row and col must be <= that of image."):
        nb.NLEFilter().trainForEnhancement(img, 40, 4, 20.0, 20.0, 3, 3)


def test_denoise_path_matches_oracle(nb):
    """NLEFilter::trainForDenoise + denoise (filter.cpp:349-410, 521-538): apply on a/b with min(S,1)^k."""
    import cv2
    rng = np.random.default_rng(2)
    L = synth_lum(40, 48)
    img = np.stack([L, np.roll(L, 3, 0), np.roll(L, 5, 1)], axis=2)
    img = np.clip(img.astype(int) + rng.integers(-6, 7, img.shape), 0, 255).astype(np.uint8)
    f = nb.NLEFilter().trainForDenoise(img, 5, 6, 25.0, 30.0, 4, 6, 10, 10)
    out = f.denoise(img, 2.0, 10, 10)
    lab = cv2.cvtColor(img, cv2.COLOR_BGR2Lab)
    den = cv2.bilateralFilter(np.ascontiguousarray(lab[:, :, 0]), -1, 10, 10, borderType=cv2.BORDER_DEFAULT)
    fo = O.train_dense(den.astype(np.float64), 5, 6, 25.0, 30.0, 4, 6)
    te = np.power(np.minimum(fo.eigvals, 1.0), 2.0)
    ref = lab.copy()
    ref[:, :, 0] = den
    for c in (1, 2):
        ref[:, :, c] = O.clamp_round_u8(O.apply(fo, lab[:, :, c].astype(np.float64), te))
    ref = cv2.cvtColor(ref, cv2.COLOR_Lab2BGR)
    d = np.abs(out.astype(int) - ref.astype(int))
    assert (d <= 1).mean() >= PIX_FRAC


# ---- README goldens (BASELINE.json configs[0] and configs[1]) ------------------------------------------
@pytest.mark.parametrize("name", [m["name"] for m in manifest()])
def test_readme_examples(nb, name):
    import cv2
    from nle_testlib import GOLDEN
    m, img, gold = load_case(name)
    ref = oracle_stages()[name]
    f = nb.NLEFilter()
    f.trainForEnhancement(img, *train_args(m))
    inf = f.info()
    # sample indices: bit-exact (checksum of the oracle's indices recorded in the fixture)
    sel, _ = nb.sampleIndices(img.shape[0], img.shape[1], m["n_row_samples"], m["n_col_samples"], with_rest=False)
    assert hashlib.sha1(sel.astype(np.int32).tobytes()).hexdigest() == ref["sel_sha1"]
    assert (inf.p, inf.r, inf.r2, inf.k) == (ref["p"], ref["r"], ref["r2"], ref["k"])
    assert sq_close(f.eigvals, ref["Sq"])
    out = f.enhance(img, m["weights"])
    # vs the reference's own committed output
    dg = np.abs(out.astype(int) - gold.astype(int))
    assert dg.max() <= 2 and (dg <= 1).mean() >= PIX_FRAC
    # vs the FP64 oracle's L channel (fixture written by tests/golden/make_oracle_stages.py)
    Lo = cv2.imread(f"{GOLDEN}/{name}_oracle_L.png", cv2.IMREAD_GRAYSCALE)
    lab = cv2.cvtColor(img, cv2.COLOR_BGR2Lab)
    Lg = f.enhanceLuminance(np.ascontiguousarray(lab[:, :, 0]), m["weights"])
    dl = np.abs(Lg.astype(int) - Lo.astype(int))
    assert dl.max() <= 1 and (dl <= 1).mean() >= PIX_FRAC and (dl == 0).mean() >= 0.995


# ---- full-size properties at BASELINE.json configs[2] ---------------------------------------------------
def test_bench_workload_properties(nb):
    from bench import GRID, HX, HY, K_EIG, T_SINK, WEIGHTS, synth_luminance
    L = synth_luminance(1024, 1024)
    f = nb.NLEFilter().trainFilter(L, GRID[0], GRID[1], HX, HY, T_SINK, K_EIG)
    inf = f.info()
    assert inf.p == 1600 and inf.k == K_EIG
    # rank of Ka equals LAPACK's count (the cut that breaks when anything upstream is sloppier than FP64)
    sel, _ = O.sample_pixels(1024, 1024, *GRID)
    Ka = O.affinity_block(L.astype(np.float64).ravel(), 1024, sel, sel, HX, HY)
    assert inf.r == O.eigen_decomposition(Ka)[1].size
    S = f.eigvals
    assert np.all(np.diff(S) <= 1e-12) and abs(S[0] - 1.0) < 5e-3        # doubly-stochastic filter: top eigenvalue ~ 1
    # V f(S) V^T with f = 1 is a projector: applying it twice changes nothing
    z = L.astype(np.float64)
    ones = np.ones(S.size)
    p1 = f.apply(z, ones)
    p2 = f.apply(p1, ones)
    assert np.abs(p2 - p1).max() <= 2e-2 * np.abs(p1).max()              # the reference's V is only ~orthonormal (App. B)
    # linearity in the weights: fS is affine in w, so enhance pre-rounding is too
    a = f.apply(z, nb.transformEigenValues(S, WEIGHTS))
    b = f.apply(z, nb.transformEigenValues(S, [1.0, 1.0, 1.0, 1.0]))
    c = f.apply(z, nb.transformEigenValues(S, [3.0, 5.0, 7.0, 1.0]))
    assert np.allclose(c, 2 * a - b, atol=1e-6)
    # a constant image is (nearly) reproduced by the row-stochastic filter
    out = f.enhanceLuminance(L, WEIGHTS)
    assert out.shape == L.shape and out.dtype == np.uint8


# ---- row sharding on one GPU: two slabs, reduction through the callback ----------------------------------
def test_two_slab_sharding_equals_single(nb):
    """Two host threads own one row slab each; the all-reduce callback sums through host memory.
    Exercises nle_b200_train_u8_sharded / the allreduce hook without a second GPU."""
    import torch
    L = synth_lum(64, 48)
    a = (6, 6, 25.0, 25.0, 4, 6)
    full = nb.NLEFilter().trainFilter(L, *a)
    bar = threading.Barrier(2)
    stash = [None, None]
    res, err = [None, None], []

    class Arr:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 3}

    def run(rank):
        try:
            def allreduce(ptr, count, stream):
                t = torch.as_tensor(Arr(ptr, count), device="cuda")
                stash[rank] = t.clone()
                bar.wait()
                tot = stash[0] + stash[1]
                bar.wait()
                t.copy_(tot)
                torch.cuda.synchronize()
                bar.wait()
            f = nb.NLEFilter().trainFilter(L, *a, shard=((0, 32, allreduce) if rank == 0 else (32, 64, allreduce)))
            res[rank] = (f.eigvals, f.eigvecs, f)
        except Exception as e:   # pragma: no cover
            err.append(e)
            bar.abort()
    th = [threading.Thread(target=run, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not err, err
    assert np.allclose(res[0][0], full.eigvals, rtol=1e-9) and np.allclose(res[1][0], full.eigvals, rtol=1e-9)
    V = np.vstack([res[0][1], res[1][1]])
    Vf = full.eigvecs
    for j in range(Vf.shape[1]):
        s = np.sign(np.dot(V[:, j], Vf[:, j]))
        assert np.allclose(s * V[:, j], Vf[:, j], atol=1e-8)


# every width class of the level-major Sinkhorn GEMMs (sk_dot_gemm_nb_kernel / sk_reduce_gemm_nb_kernel are instantiated per
# number of 8-column tiles of the sample grid, 1 .. 8) and the m-tile blocking of the grid rows (7, 5, 4, 3 tiles per CTA)
@pytest.mark.parametrize("grid", [(6, 7), (5, 12), (9, 20), (60, 28), (4, 36), (7, 44), (34, 52), (3, 60), (58, 64)])
def test_sinkhorn_gemm_width_classes_match_oracle(nb, grid):
    nR, nC = grid
    L = synth_lum(max(64, nR + 3), max(96, nC + 5), seed=nR * 100 + nC)
    a = (nR, nC, 35.0, 28.0, 5, 8)
    f = nb.NLEFilter().trainFilter(L, *a)
    fo = O.train_dense(L.astype(np.float64), *a)
    st, inf = fo.stages, f.info()
    assert (inf.p, inf.r, inf.r2, inf.k) == (st["p"], st["r"], st["r2"], fo.eigvals.size)
    cg = f.stage(3)
    co = np.empty_like(cg)
    co[st["perm"]] = st["c"]
    mask = np.ones(cg.size, bool)
    mask[st["perm"][:inf.p]] = False
    # Sinkhorn scaling vector c on the rest pixels.  Densely sampled grids (60 x 28, 34 x 52 on 64 x 96 pixels) have an
    # ill-conditioned Ka: the dense oracle itself moves c by 2.0e-7 / 2.0e-6 relative when its LAPACK eigensolver is switched from
    # MRRR to divide & conquer (the algorithm family of the CUDA solver) -- the deviations seen on the GPU; well-conditioned
    # grids agree to 1e-11.  The contract below is on Sq and the output.
    assert np.abs(cg[mask] - co[mask]).max() <= 1e-5 * np.abs(co).max()
    assert sq_close(f.eigvals, fo.eigvals)
    w = [2.0, 3.0, 4.0, 1.0]
    d = np.abs(f.enhanceLuminance(L, w).astype(int) - O.enhance_luminance(fo, L, w).astype(int))
    assert d.max() <= 1 and (d <= 1).mean() >= PIX_FRAC
