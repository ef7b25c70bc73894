"""Developer diagnostic: the top-k block eigensolver (csrc/eig_topk.cu) against LAPACK on synthetic PSD spectra, with wall
times next to the full solver's (both through the host-pointer C ABI, so both include the same upload).

  python scripts/gpu_topk.py
"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nonlocal_image_edit_b200 as nb

rng = np.random.default_rng(1)


def timed(fn, reps=3):
    fn()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); out = fn(); best = min(best, time.perf_counter() - t0)
    return out, best * 1e3


for n, k, decay in ((612, 50, 0.96), (1357, 50, 0.985), (1950, 100, 0.99), (800, 100, 0.97), (612, 50, 0.995), (400, 20, 0.9)):
    lam = np.concatenate([[1.0, 0.86, 0.72], 0.6 * decay ** np.arange(n - 3)]) + 1e-9
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    A = (Q * lam) @ Q.T
    A = (A + A.T) / 2
    (U, D, prod), ms = timed(lambda: nb.topkEigenDecomposition(A, k, assume_psd=True, return_products=True))
    (_, _), ms_full = timed(lambda: nb.eigenDecomposition(A, eps=-1e300))
    (_, _), ms_up = timed(lambda: nb.eigenDecomposition(A[:8, :8].copy(), eps=-1e300))
    w = np.linalg.eigvalsh(A)[::-1][:k]
    err = np.abs(D - w[:D.size]).max() / w[0] if D.size == k else float("nan")
    res = np.abs(A @ U - U * D).max()
    orth = np.abs(U.T @ U - np.eye(U.shape[1])).max()
    print(f"n={n:5d} k={k:3d} decay={decay}: products={prod:4d} got {D.size} pairs, eig err {err:.1e}, resid {res:.1e}, orth {orth:.1e}; "
          f"top-k {ms:7.2f} ms, full {ms_full:7.2f} ms (both incl. transfers)", flush=True)
