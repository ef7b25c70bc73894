"""CPU FP64 oracle for the Nystrom spectral-filter hot path.  TEST INFRASTRUCTURE ONLY.

This file is a NumPy/SciPy restatement of the reference algorithm in
``/root/reference/src/filter.cpp`` (cited below as ``filter.cpp:LINE``) and
``/root/reference/include/utils.hpp``.  It is the checker the CUDA path is compared against.
It is NOT the product and NOT a fallback: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.

Why a restatement and not the reference binary: the reference needs Eigen >= 3.3 and the
OpenCV C++ SDK (``CMakeLists.txt:34,37``); neither exists in this image (no ``Eigen/Core``, no
``opencv2/core.hpp`` anywhere on the filesystem), so ``src/filter.cpp`` cannot be compiled here.

Third-party arithmetic the reference delegates to, and what stands in for it here:
  * Eigen ``SelfAdjointEigenSolver`` (``filter.cpp:207-210``) -> ``scipy.linalg.eigh(lower=True)``
    (LAPACK dsyevd).  Both read the LOWER triangle only and are backward stable, so eigenvalues
    agree to ~1e-15*||M|| and the ``>= 1e-10`` prefix cut picks the same rank.
  * OpenCV ``cvtColor(BGR2Lab / Lab2BGR)`` on 8-bit images and ``convertTo(CV_8U)``
    (``filter.cpp:423,436,440``) -> Python ``cv2`` 4.13 and ``np.rint`` (cvRound is
    round-half-to-even).

Parity pin: ``tests/test_oracle.py`` checks this oracle against (G1) the ten README
parameter rows and their committed outputs ``data/*-filtered.png`` (fixtures under
``tests/golden/``, made by ``tests/golden/make_golden.py``), (G2) the 3x3 known-answer
eigen-decomposition of ``test/test_filter.cpp:42-68``, (G3) the identity Sinkhorn case
``test_filter.cpp:70-94`` and (G4) the property tests ``test_filter.cpp:96-153`` with fixed seeds.

Two evaluation strategies are provided and tested against each other:
  * ``train_dense``     -- materialises Ka, Kab, phi, Wab exactly like the reference.
  * ``train_streaming`` -- same mathematics in factor form, tiled over
    pixels; never stores an O(p*N) array.  Used for configurations whose dense intermediates do
    not fit in host RAM, and it is the formulation the CUDA kernels implement.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.linalg

EPS = 1e-10  # include/filter.hpp:14


# --------------------------------------------------------------------------------------------
# filter.cpp:42-54
def inplace_reciprocal(v: np.ndarray, eps: float = EPS):
    """``v_i <- 1/v_i`` if ``|v_i| >= eps`` else 0; returns (result, #kept)."""
    v = np.asarray(v, dtype=np.float64)
    keep = np.abs(v) >= eps
    out = np.zeros_like(v)
    np.divide(1.0, v, out=out, where=keep)
    return out, int(keep.sum())


# --------------------------------------------------------------------------------------------
# filter.cpp:56-80
def sample_axis(n: int, k: int) -> np.ndarray:
    """Selected coordinates along one axis (the reference's test is separable per axis)."""
    step = n // k                                   # filter.cpp:58-59 (int division)
    offset = (step - 1 + (n - step * k)) // 2       # filter.cpp:60-61 (non-negative ints)
    r = np.arange(n)
    ok = (r >= offset) & (r <= n - offset) & ((r - offset) % step == 0)   # filter.cpp:68-70
    return r[ok].astype(np.int64)


def sample_pixels(nrows: int, ncols: int, n_row_samples: int, n_col_samples: int):
    """Returns (selected, rest) as 1-D raster indices (utils.hpp:11-14), both in raster order.

    Note: the count per axis can EXCEED the requested number (e.g. n=100,k=40 -> 41).
    """
    rows = sample_axis(nrows, n_row_samples)
    cols = sample_axis(ncols, n_col_samples)
    mask = np.zeros((nrows, ncols), dtype=bool)
    mask[np.ix_(rows, cols)] = True
    flat = mask.ravel()
    idx = np.arange(nrows * ncols, dtype=np.int64)
    return idx[flat], idx[~flat]


def sample_pixels_loop(nrows, ncols, n_row_samples, n_col_samples):
    """Literal double loop of filter.cpp:63-77 (slow; used by tests to pin ``sample_pixels``)."""
    row_step = nrows // n_row_samples
    col_step = ncols // n_col_samples
    row_off = (row_step - 1 + (nrows - row_step * n_row_samples)) // 2
    col_off = (col_step - 1 + (ncols - col_step * n_col_samples)) // 2
    sel, rest = [], []
    for r in range(nrows):
        for c in range(ncols):
            if (r >= row_off and c >= col_off and r <= nrows - row_off and c <= ncols - col_off
                    and (r - row_off) % row_step == 0 and (c - col_off) % col_step == 0):
                sel.append(r * ncols + c)
            else:
                rest.append(r * ncols + c)
    return np.array(sel, dtype=np.int64), np.array(rest, dtype=np.int64)


# --------------------------------------------------------------------------------------------
# filter.cpp:104-167
def affinity_block(lum_flat, ncols, idx_a, idx_b, hx, hy):
    """exp(negativeWeightedDistance) between pixel sets (filter.cpp:104-112,128-129,144-145).

    Spatial squared distance is integer arithmetic converted to double (filter.cpp:109);
    weights are 1/hx^2 and 1/hy^2 -- no factor 2.
    """
    sw = 1.0 / (hx * hx)
    pw = 1.0 / (hy * hy)
    ra, ca = np.divmod(idx_a, ncols)
    rb, cb = np.divmod(idx_b, ncols)
    d2 = ((ra[:, None] - rb[None, :]) ** 2 + (ca[:, None] - cb[None, :]) ** 2).astype(np.float64)
    dy = lum_flat[idx_a][:, None] - lum_flat[idx_b][None, :]
    return np.exp(-sw * d2 - pw * (dy * dy))


# filter.cpp:104-145 again, in C (oracle/nle_oracle_c.c): one scalar libm exp per entry, no temporaries, sample rows
# split over a thread pool.  Same expression and evaluation order; differs from the NumPy form above only by the
# last-ulp behaviour of the two exp implementations.  Used by train_streaming for the configurations the dense
# reference cannot hold in RAM (tests/golden/make_oracle_big.py); tests/test_oracle.py pins it to affinity_block.
_C = None


def _c_lib():
    global _C
    if _C is None:
        import ctypes
        import os
        from . import build_c
        lib = ctypes.CDLL(build_c.build())
        vp, i64, f64 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_double
        lib.nle_oracle_affinity_block.argtypes = [vp, vp, vp, i64, i64, vp, vp, vp, i64, f64, f64, vp]
        lib.nle_oracle_affinity_block.restype = None
        _C = lib
    return _C


def affinity_block_c(lum_flat, ncols, idx_a, idx_b, hx, hy, threads=None):
    import os
    from concurrent.futures import ThreadPoolExecutor
    lib = _c_lib()
    sw = 1.0 / (hx * hx)
    pw = 1.0 / (hy * hy)
    idx_a = np.ascontiguousarray(idx_a, dtype=np.int64)
    idx_b = np.ascontiguousarray(idx_b, dtype=np.int64)
    ra, ca = (np.ascontiguousarray(a) for a in np.divmod(idx_a, ncols))
    rb, cb = (np.ascontiguousarray(a) for a in np.divmod(idx_b, ncols))
    ya = np.ascontiguousarray(lum_flat[idx_a], dtype=np.float64)
    yb = np.ascontiguousarray(lum_flat[idx_b], dtype=np.float64)
    na, nb = idx_a.size, idx_b.size
    out = np.empty((na, nb), dtype=np.float64)
    nt = max(1, min(threads or (os.cpu_count() or 1), na))
    cuts = [na * t // nt for t in range(nt + 1)]

    def part(t):
        lib.nle_oracle_affinity_block(ra.ctypes.data, ca.ctypes.data, ya.ctypes.data, cuts[t], cuts[t + 1],
                                      rb.ctypes.data, cb.ctypes.data, yb.ctypes.data, nb, sw, pw, out.ctypes.data)

    if nt == 1:
        part(0)
    else:
        with ThreadPoolExecutor(nt) as ex:
            list(ex.map(part, range(nt)))
    return out


def compute_kernel(lum: np.ndarray, n_row_samples: int, n_col_samples: int, hx: float, hy: float):
    """filter.cpp:114-167 -> (perm, Ka, Kab); perm = [selected; rest] raster indices (:156-164)."""
    nrows, ncols = lum.shape
    if n_row_samples > nrows or n_col_samples > ncols:
        raise RuntimeError("Number of samples per row and col must be <= that of image.")  # :118
    sel, rest = sample_pixels(nrows, ncols, n_row_samples, n_col_samples)
    z = np.ascontiguousarray(lum, dtype=np.float64).ravel()
    Ka = affinity_block(z, ncols, sel, sel, hx, hy)
    Kab = affinity_block(z, ncols, sel, rest, hx, hy)
    perm = np.concatenate([sel, rest])
    return perm, Ka, Kab


# --------------------------------------------------------------------------------------------
# filter.cpp:204-228
def eigen_decomposition(M: np.ndarray, eps: float = EPS):
    """Symmetric eigensolve on the LOWER triangle, descending, prefix with lambda >= eps."""
    M = np.asarray(M, dtype=np.float64)
    if M.shape[0] == 0:
        return np.zeros((0, 0)), np.zeros(0)
    w, v = scipy.linalg.eigh(M, lower=True)
    D = w[::-1]
    U = v[:, ::-1]
    r = 0
    while r < D.size and D[r] >= eps:               # filter.cpp:213-214
        r += 1
    return np.ascontiguousarray(U[:, :r]), D[:r].copy()


# --------------------------------------------------------------------------------------------
# filter.cpp:257-280
def nystrom_approximation(Ka: np.ndarray, Kab: np.ndarray):
    U, lam = eigen_decomposition(Ka)
    inv, nnz = inplace_reciprocal(lam)              # :265-266 (all kept: lam >= eps already)
    U = U[:, :nnz]
    lam = lam[:nnz]
    phi = np.vstack([U, (Kab.T @ U) * inv[:nnz][None, :]])   # :275
    return lam, phi


# --------------------------------------------------------------------------------------------
# filter.cpp:230-254
def sinkhorn(phi: np.ndarray, eigvals: np.ndarray, max_iter: int = 10):
    n = phi.shape[0]
    r = np.ones(n)
    c = np.zeros(n)
    for _ in range(max_iter):                       # :238-245
        c, _ = inplace_reciprocal(phi @ (eigvals * (phi.T @ r)))
        r, _ = inplace_reciprocal(phi @ (eigvals * (phi.T @ c)))
    p = phi.shape[1]                                # :247  -- phi.cols(), i.e. the RANK, not #samples
    left = r[:p, None] * (phi[:p] * eigvals[None, :])            # R * (phi.topRows(p) * D)
    Wa = left @ (c[:p, None] * phi[:p]).T                        # :249
    Wab = left @ (c[p:, None] * phi[p:]).T                       # :250
    return Wa, Wab, r, c


# --------------------------------------------------------------------------------------------
# filter.cpp:282-331 (non-Spectra branch :313-316)
def orthogonalize(Wa: np.ndarray, Wab: np.ndarray, n_eig_vectors: int = 5, eps: float = EPS):
    Ua, la = eigen_decomposition(Wa)
    inv, _ = inplace_reciprocal(la)
    inv_root = np.sqrt(inv)
    inv_root_wa = (Ua * inv_root[None, :]) @ Ua.T               # :292
    Q = Wa + inv_root_wa @ (Wab @ Wab.T) @ inv_root_wa          # :296
    Vq, Sq = eigen_decomposition(Q)
    k = min(n_eig_vectors, Vq.shape[1])
    Vq = Vq[:, :k]
    Sq = Sq[:k]
    inv_sq, _ = inplace_reciprocal(Sq)
    inv_root_sq = np.sqrt(inv_sq)
    tmp = np.vstack([Wa, Wab.T])                                # :324-325
    V = ((tmp @ inv_root_wa) @ Vq) * inv_root_sq[None, :]       # :327 (left to right)
    return V, Sq, dict(la=la, Q=Q, inv_root_wa=inv_root_wa, Vq=Vq)


# --------------------------------------------------------------------------------------------
# filter.cpp:334-347
def transform_eigenvalues(eigvals: np.ndarray, weights) -> np.ndarray:
    eigvals = np.asarray(eigvals, dtype=np.float64)
    fS = np.full(eigvals.shape, float(weights[0]))
    for k in range(1, len(weights)):
        fS = fS + (weights[k] - weights[k - 1]) * np.power(eigvals, float(k))
    return fS


@dataclass
class TrainedFilter:
    """State of nle::NLEFilter after trainFilter (filter.hpp:52-53) plus stage intermediates."""
    rows: int
    cols: int
    eigvecs: np.ndarray              # N x k', pixel (raster) order   -- m_eigvecs
    eigvals: np.ndarray              # k'                            -- m_eigvals
    stages: dict = field(default_factory=dict)


# --------------------------------------------------------------------------------------------
# filter.cpp:480-502
def train_dense(lum: np.ndarray, n_row_samples: int, n_col_samples: int, hx: float, hy: float,
                n_sinkhorn_iter: int = 10, n_eigen_vectors: int = 5) -> TrainedFilter:
    lum = np.asarray(lum, dtype=np.float64)
    perm, Ka, Kab = compute_kernel(lum, n_row_samples, n_col_samples, hx, hy)
    p = Ka.shape[0]
    lam, phi = nystrom_approximation(Ka, Kab)
    del Kab
    Wa, Wab, rvec, cvec = sinkhorn(phi, lam, n_sinkhorn_iter)
    r = lam.size
    del phi
    V, Sq, aux = orthogonalize(Wa, Wab, n_eigen_vectors)
    del Wab
    Vf = np.empty_like(V)
    Vf[perm] = V                                     # :502  (P * V): row i -> pixel perm[i]
    stages = dict(perm=perm, p=p, r=r, r2=int(aux["la"].size), Ka=Ka, lam=lam, Wa=Wa,
                  rvec_head=rvec[:r].copy(), c=cvec, la=aux["la"], Q=aux["Q"], Sq=Sq)
    return TrainedFilter(lum.shape[0], lum.shape[1], Vf, Sq, stages)


# filter.cpp:445-458
def apply(flt: TrainedFilter, channel: np.ndarray, transformed: np.ndarray) -> np.ndarray:
    if channel.size != flt.eigvecs.shape[0]:
        raise RuntimeError("Number of values in channel must match that of training image.")
    z = np.ascontiguousarray(channel, dtype=np.float64).ravel()
    out = flt.eigvecs @ (transformed * (flt.eigvecs.T @ z))
    return out.reshape(channel.shape)


def clamp_round_u8(x: np.ndarray) -> np.ndarray:
    """filter.cpp:434-436: max(.,0), min(.,255), convertTo(CV_8U) == cvRound == half-to-even."""
    return np.rint(np.minimum(np.maximum(x, 0.0), 255.0)).astype(np.uint8)


def enhance_luminance(flt: TrainedFilter, lum_u8: np.ndarray, weights) -> np.ndarray:
    """filter.cpp:426-436 on the L channel only (colour conversion is done by the caller)."""
    fS = transform_eigenvalues(flt.eigvals, weights)
    return clamp_round_u8(apply(flt, lum_u8.astype(np.float64), fS))


# --------------------------------------------------------------------------------------------
# Image-level entry points (need cv2; filter.cpp:412-443, 460-469, 514-519)
def bgr_to_lab(image_bgr_u8):
    import cv2
    return cv2.cvtColor(image_bgr_u8, cv2.COLOR_BGR2Lab)


def lab_to_bgr(lab_u8):
    import cv2
    return cv2.cvtColor(lab_u8, cv2.COLOR_Lab2BGR)


def train_for_enhancement(image_bgr_u8, n_row_samples, n_col_samples, hx, hy,
                          n_sinkhorn_iter=10, n_eigen_vectors=5, streaming=False, **kw):
    lab = bgr_to_lab(image_bgr_u8)
    lum = lab[:, :, 0].astype(np.float64)
    fn = train_streaming if streaming else train_dense
    return fn(lum, n_row_samples, n_col_samples, hx, hy, n_sinkhorn_iter, n_eigen_vectors, **kw)


def enhance(flt: TrainedFilter, image_bgr_u8, weights):
    if image_bgr_u8.ndim != 3 or image_bgr_u8.shape[2] != 3:
        raise RuntimeError("Can only enhance RGB image.")                       # :415
    if image_bgr_u8.shape[0] * image_bgr_u8.shape[1] != flt.eigvecs.shape[0]:
        raise RuntimeError("Cannot apply filter on image with different size from the image "
                           "filter was trained on.")                             # :419
    lab = bgr_to_lab(image_bgr_u8)
    lab[:, :, 0] = enhance_luminance(flt, lab[:, :, 0], weights)
    return lab_to_bgr(lab)


# --------------------------------------------------------------------------------------------
# Streaming (factor-form) evaluation: SURVEY.md Appendix A.3-A.7.  Same mathematics as
# train_dense; no O(p*N) array is ever stored.  This is the formulation the CUDA path uses.
def _tiles(n, tile):
    for s in range(0, n, tile):
        yield s, min(n, s + tile)


def train_streaming(lum: np.ndarray, n_row_samples: int, n_col_samples: int, hx: float, hy: float,
                    n_sinkhorn_iter: int = 10, n_eigen_vectors: int = 5,
                    tile: int = 16384, slab=None, allreduce=None, block_fn=None, ka_fn=None,
                    gram_fn=None) -> TrainedFilter:
    """block_fn: affinity_block (default, NumPy) or affinity_block_c (same loop in C, threaded).
    ka_fn / gram_fn: hooks of scripts/precision_probe.py ONLY (Ka -> perturbed Ka; X = Kab_tile * c -> X X^T evaluated in a
    reduced-precision arithmetic); the oracle proper never passes them.
    slab=(row0,row1) restricts every pixel sum to the image rows owned by one rank and
    `allreduce(np.ndarray)` (in-place sum over ranks) completes them: the CPU model of the row-sharded
    multi-GPU path.  The returned eigvecs then cover only the slab's pixels."""
    lum = np.asarray(lum, dtype=np.float64)
    nrows, ncols = lum.shape
    if n_row_samples > nrows or n_col_samples > ncols:
        raise RuntimeError("Number of samples per row and col must be <= that of image.")
    z = lum.ravel()
    sel, rest = sample_pixels(nrows, ncols, n_row_samples, n_col_samples)
    if slab is not None:
        rest = rest[(rest >= slab[0] * ncols) & (rest < slab[1] * ncols)]
    if allreduce is None:
        def allreduce(a):
            return a
    perm = np.concatenate([sel, rest])
    p, nrest = sel.size, rest.size
    T = n_sinkhorn_iter

    blk = block_fn or affinity_block

    def kb(s, e):                       # p x (e-s) block of Kab for rest pixels s..e-1
        return blk(z, ncols, sel, rest[s:e], hx, hy)

    Ka = blk(z, ncols, sel, sel, hx, hy)
    if ka_fn is not None:
        Ka = ka_fn(Ka)
    U, lam = eigen_decomposition(Ka)                      # A.3
    r = lam.size
    inv_lam = 1.0 / lam

    def phiT_x(x_sel, s_vec):           # phi^T x = U^T x_sel + Lam^-1 U^T (Kab x_rest)     (A.4)
        return U.T @ x_sel + inv_lam * (U.T @ s_vec)

    # Sinkhorn in factor form.  x lives as (x_sel [p], x_rest [N-p]).
    def half_step(t, need_rest=True):   # returns recip(phi Lam t) split as (sel, rest) and Kab x_rest
        w = U @ t                       # for rest pixels (K~x)_j = k_j^T (U t)
        y_sel = U @ (lam * t)           # for samples     (K~x)_s = U[s,:] Lam t
        x_sel, _ = inplace_reciprocal(y_sel)
        if not need_rest:
            return x_sel, None, None
        x_rest = np.empty(nrest)
        s_vec = np.zeros(p)
        for s, e in _tiles(nrest, tile):
            K = kb(s, e)
            xr, _ = inplace_reciprocal(K.T @ w)
            x_rest[s:e] = xr
            s_vec += K @ xr
        allreduce(s_vec)
        return x_sel, x_rest, s_vec

    s0 = np.zeros(p)
    for s, e in _tiles(nrest, tile):
        s0 += kb(s, e).sum(axis=1)                        # Kab * 1
    allreduce(s0)
    t = phiT_x(np.ones(p), s0)
    c_sel = c_rest = r_sel = None
    if T < 1:
        raise RuntimeError("n_sinkhorn_iter must be >= 1")
    for it in range(T):
        c_sel, c_rest, s_vec = half_step(t)
        t = phiT_x(c_sel, s_vec)
        last = it == T - 1                                # the final rvec is only used on perm[0:r]
        r_sel, _r_rest, s_vec = half_step(t, need_rest=not last)
        if not last:
            t = phiT_x(r_sel, s_vec)

    # --- Wa, Gram (A.5).  "Landmarks" are the first r entries of perm (filter.cpp:247).
    phi_top = U[:r]                                       # r x r  (rows of phi for perm[0:r])
    L = r_sel[:r, None] * (phi_top * lam[None, :])        # diag(rvec) phi_top Lam
    Wa = L @ (c_sel[:r, None] * phi_top).T
    Gp = np.zeros((p, p))                                 # sum_j c_j^2 k_j k_j^T over rest pixels
    for s, e in _tiles(nrest, tile):
        K = kb(s, e) * c_rest[None, s:e]
        Gp += K @ K.T if gram_fn is None else gram_fn(K)
    allreduce(Gp)
    UL = U * inv_lam[None, :]                             # p x r : U Lam^-1
    G = UL.T @ Gp @ UL
    demoted = U[r:p] * c_sel[r:p, None]                   # samples r..p-1 are "rest" for Wab
    G += demoted.T @ demoted
    WabWabT = L @ G @ L.T

    # --- orthogonalise (A.6)
    Ua, la = eigen_decomposition(Wa)
    inv_root_wa = (Ua * (1.0 / np.sqrt(la))[None, :]) @ Ua.T
    Q = Wa + inv_root_wa @ WabWabT @ inv_root_wa
    Vq, Sq = eigen_decomposition(Q)
    k = min(n_eigen_vectors, Vq.shape[1])
    Vq, Sq = Vq[:, :k], Sq[:k]
    Mv = inv_root_wa @ Vq * (1.0 / np.sqrt(Sq))[None, :]  # r x k'

    # --- extension (A.6): V_pi = [Wa ; Wab^T] Mv
    Z_rest = L.T @ Mv                                     # r x k'
    Y = UL @ Z_rest                                       # p x k'
    V = np.zeros((nrows * ncols, k))
    V[perm[:r]] = Wa @ Mv                                 # top block: rows of Wa
    V[perm[r:p]] = c_sel[r:p, None] * (U[r:p] @ Z_rest)   # demoted samples: c_j phi_j Z_rest
    for s, e in _tiles(nrest, tile):
        K = kb(s, e)
        V[rest[s:e]] = c_rest[s:e, None] * (K.T @ Y)
    c_full = np.concatenate([c_sel, c_rest])
    if slab is not None:
        lo, hi = slab[0] * ncols, slab[1] * ncols
        V = V[lo:hi]
    stages = dict(perm=perm, p=p, r=r, r2=int(la.size), Ka=Ka, lam=lam, Wa=Wa,
                  rvec_head=r_sel[:r].copy(), c=c_full, la=la, Q=Q, Sq=Sq, G=G)
    return TrainedFilter(nrows, ncols, V, Sq, stages)
