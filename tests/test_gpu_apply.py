"""GPU: the TMA-streamed output path (csrc/apply_kernels.cu: vtz_tma_kernel, recompose_tma_kernel and their fused
BGR<->Lab epilogues) against V diag(fS) V^T z evaluated in NumPy from the filter's own eigenvectors (NLEFilter::apply,
filter.cpp:445-458; the clamp / round / colour conversion of NLEFilter::enhance, filter.cpp:422-440), over the shapes
that select different tilings: k odd and even, k <= 50 (128-row tiles), 50 < k <= 100 (64 rows, 2 threads per row),
k > 128 (several columns per thread), k > 400 (plain-load kernels), pixel counts that are not a multiple of the tile and
odd x odd (the bulk copy moves multiples of 16 bytes: the last double is carried separately)."""
import cv2
import numpy as np
import pytest

from nle_testlib import synth_lum

pytestmark = pytest.mark.gpu


def _rough(rows, cols, seed):
    rng = np.random.default_rng(seed)
    base = synth_lum(rows, cols, seed=seed).astype(np.int32)
    return np.clip(base + rng.integers(-40, 41, size=(rows, cols)), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("rows,cols,grid,k_req", [
    (63, 81, (7, 9), 1), (63, 81, (7, 9), 3), (63, 81, (7, 9), 27), (64, 80, (8, 10), 50), (63, 81, (9, 9), 51),
    (100, 130, (10, 13), 100), (97, 131, (12, 13), 129), (90, 120, (15, 18), 200), (91, 121, (17, 19), 257),
    (96, 128, (21, 21), 420)])
def test_apply_paths_match_numpy(nb, rows, cols, grid, k_req):
    L = _rough(rows, cols, seed=rows + k_req)
    f = nb.NLEFilter().trainFilter(L, grid[0], grid[1], 6.0, 12.0, 4, k_req)      # short hx, hy: Ka close to full rank
    k = f.info().k
    assert k >= min(k_req, 20), f"only {k} eigenvectors survived; pick a rougher test image"
    V, S = f.eigvecs, f.eigvals
    z = L.astype(np.float64).ravel()
    rng = np.random.default_rng(k)
    fS = rng.uniform(0.5, 2.0, k)
    ref = V @ (fS * (V.T @ z))
    scale = max(1.0, np.abs(ref).max())
    got = f.apply(L.astype(np.float64), fS).ravel()
    assert np.abs(got - ref).max() <= 1e-10 * scale
    w = [2.0, 3.0, 4.0, 1.0]
    tS = nb.transformEigenValues(S, w)
    exact = V @ (tS * (V.T @ z))
    out = f.enhanceLuminance(L, w).ravel().astype(int)
    want = np.rint(np.clip(exact, 0, 255)).astype(int)
    near_tie = np.abs(exact - np.floor(exact) - 0.5) < 1e-7               # rounding ties may fall either way
    assert np.all((out == want) | near_tie)


@pytest.mark.parametrize("rows,cols,k_req", [(63, 81, 9), (100, 130, 64), (97, 131, 140)])
def test_fused_bgr_enhance_equals_unfused_pipeline(nb, rows, cols, k_req):
    rng = np.random.default_rng(rows)
    img = np.stack([_rough(rows, cols, seed=c + 1) for c in range(3)], axis=2)
    img[:2] = rng.integers(0, 256, size=(2, cols, 3), dtype=np.uint8)           # saturated colours exercise the Lab tables
    f = nb.NLEFilter().trainForEnhancement(img, 9, 11, 6.0, 12.0, 4, k_req)
    w = [1.5, 2.5, 3.0, 1.0]
    got = f.enhance(img, w)                                                      # BGR2Lab / Lab2BGR fused into the two passes
    lab = cv2.cvtColor(img, cv2.COLOR_BGR2Lab)
    lab[:, :, 0] = f.enhanceLuminance(np.ascontiguousarray(lab[:, :, 0]), w)     # same filter, colour conversion by OpenCV
    assert np.array_equal(got, cv2.cvtColor(lab, cv2.COLOR_Lab2BGR))
