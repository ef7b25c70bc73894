#!/usr/bin/env python
"""NumPy prototype of the top-k eigensolver for the third eigensolve (eig of the r2 x r2 block of Q, filter.cpp:311-316):
Chebyshev-filtered block subspace iteration with Cholesky-QR and two or three Rayleigh-Ritz steps -- the structure
csrc/eig_topk.cu implements with DMMA GEMMs.  Run on the Q of a crop of the bench image (oracle) and on synthetic spectra.

  python scripts/proto_topk.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cholqr(X):
    G = X.T @ X
    L = np.linalg.cholesky(G)
    return np.linalg.solve(L, X.T).T


def cheb(A, X, d, lb, cut):
    c, e = 0.5 * (cut + lb), 0.5 * (cut - lb)
    T0 = X
    T1 = (A @ X - c * X) / e
    for _ in range(2, d + 1):
        T0, T1 = T1, 2.0 * (A @ T1 - c * T1) / e - T0
    return T1


def choose_degree(theta0, theta_k, lb, cut, max_ratio=1e6, dmax=6):
    c, e = 0.5 * (cut + lb), 0.5 * (cut - lb)
    x0, xk = (theta0 - c) / e, max((theta_k - c) / e, 1.0 + 1e-9)
    best = 2
    for d in range(2, dmax + 1):
        ratio = np.cosh(d * np.arccosh(x0)) / np.cosh(d * np.arccosh(xk))
        if ratio <= max_ratio:
            best = d
    return best


def topk(A, k, guard=None, tol=2e-13, max_blocks=3, rounds=8, seed=0, verbose=True):
    n = A.shape[0]
    m = k + (guard if guard is not None else max(14, k // 4))
    m = (m + 7) // 8 * 8
    rng = np.random.default_rng(seed)
    X = cholqr(cholqr(rng.standard_normal((n, m))))
    for _ in range(2):                                   # two plain power steps: a first look at the spectrum
        X = cholqr(A @ X)
    ngemm = 2
    for blk in range(max_blocks):
        Y = A @ X
        H = X.T @ Y
        th, W = np.linalg.eigh((H + H.T) / 2)
        th, W = th[::-1], W[:, ::-1]
        X, Y = X @ W, Y @ W
        res = np.linalg.norm(Y - X * th, axis=0)
        ngemm += 1
        if verbose:
            print(f"   RR {blk}: theta0={th[0]:.4f} theta_k={th[k-1]:.3e} theta_m={th[-1]:.3e} max res(top k)={res[:k].max():.2e} gemms={ngemm}")
        if res[:k].max() <= tol * th[0]:
            return th[:k], X[:, :k], dict(gemms=ngemm, rr=blk + 1, converged=True)
        lb, cut = -1e-3 * th[0], th[-1]
        d = choose_degree(th[0], th[k - 1], lb, cut)
        for _ in range(rounds):
            X = cholqr(cheb(A, X, d, lb, cut))
            ngemm += d
    return th[:k], X[:, :k], dict(gemms=ngemm, rr=max_blocks, converged=False)


def main():
    from oracle import nle_oracle as O
    import bench
    cases = []
    _, lum = bench.workload_images(1024, 1024)
    crop = lum[:320, :320].astype(np.float64)
    t0 = time.time()
    flt = O.train_streaming(crop, 40, 40, bench.HX, bench.HY, bench.T_SINK, bench.K_EIG, block_fn=O.affinity_block_c)
    Q = flt.stages["Q"]
    Qs = np.tril(Q) + np.tril(Q, -1).T
    w, v = np.linalg.eigh(Qs)
    keep = w >= 1e-10
    M = (v[:, keep].T @ Qs @ v[:, keep])
    print(f"bench-image crop 320x320: r={flt.stages['r']} r2={flt.stages['r2']} block n={M.shape[0]} ({time.time()-t0:.0f}s)")
    cases.append(("bench crop Q block", M, 50))
    rng = np.random.default_rng(1)
    for n, k, decay in ((612, 50, 0.96), (1357, 50, 0.985), (1950, 100, 0.99), (800, 100, 0.97)):
        lam = np.concatenate([[1.0, 0.86, 0.72], 0.6 * decay ** np.arange(n - 3)]) + 1e-9
        Qm, _ = np.linalg.qr(rng.standard_normal((n, n)))
        cases.append((f"synthetic n={n} decay={decay}", (Qm * lam) @ Qm.T, k))
    for name, A, k in cases:
        if A.shape[0] < 4 * (k + 16):
            print(name, "skipped: block too small for the top-k path")
            continue
        print(name, "n =", A.shape[0], "k =", k)
        th, X, info = topk(A, k)
        wref = np.linalg.eigvalsh((A + A.T) / 2)[::-1][:k]
        print(f"   -> converged={info['converged']} gemms={info['gemms']} rr={info['rr']} max rel eig err={np.abs(th-wref).max()/wref[0]:.2e} "
              f"rel on smallest={abs(th[-1]-wref[-1])/wref[-1]:.2e} orth={np.abs(X.T@X-np.eye(k)).max():.1e}")


if __name__ == "__main__":
    main()
