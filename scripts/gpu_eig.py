"""Developer diagnostic: eigensolver accuracy on a matrix zoo and timing on Ka matrices of the bench workload.

  NLE_B200_EIG=jacobi python scripts/gpu_eig.py      # old Jacobi path
  python scripts/gpu_eig.py [sizes...]               # direct path (tridiagonalisation + divide & conquer)
"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nonlocal_image_edit_b200 as nb
from oracle import nle_oracle as O
from bench import synth_luminance

os.environ.setdefault("NLE_B200_EIG_STRICT", "1")
rng = np.random.default_rng(0)


def check(A, name):
    n = A.shape[0]
    try:
        t0 = time.time(); U, D = nb.eigenDecomposition(A, eps=-1e300); dt = time.time() - t0
    except Exception as ex:
        print(f'{name:26s} n={n:5d} FAILED: {ex}', flush=True)
        return False
    w = np.linalg.eigvalsh(np.tril(A) + np.tril(A, -1).T)[::-1]
    As = np.tril(A) + np.tril(A, -1).T
    nrm = max(np.abs(w).max(), 1e-300)
    err = np.abs(D - w).max() / nrm
    orth = np.abs(U.T @ U - np.eye(n)).max()
    res = np.abs(As @ U - U * D).max() / nrm
    ok = err < 1e-13 * max(1, n / 50) and orth < 1e-11 and res < 1e-12 * max(1, n / 50)
    print(f'{name:26s} n={n:5d} {dt*1e3:8.1f} ms  eig rel err {err:.2e} orth {orth:.2e} resid {res:.2e} '
          f'count {(D >= 1e-10).sum()} vs {(w >= 1e-10).sum()} {"ok" if ok else "BAD"}', flush=True)
    return ok


bad = 0
if not sys.argv[1:]:
    for n in (1, 2, 3, 5, 17, 32, 33, 64, 65, 100, 257, 300, 513):
        A = rng.standard_normal((n, n)); A = (A + A.T) / 2
        bad += not check(A, 'random symmetric')
    bad += not check(np.array([[2., -1, 0], [-1, 2, -1], [0, -1, 2]]), 'test_filter.cpp 3x3')
    bad += not check(np.eye(70), 'identity')
    bad += not check(np.zeros((70, 70)), 'zeros')
    bad += not check(np.ones((70, 70)), 'ones (rank 1)')
    B = rng.standard_normal((90, 7)); bad += not check(B @ B.T, 'rank 7 PSD')
    Qr, _ = np.linalg.qr(rng.standard_normal((120, 120)))
    bad += not check(Qr @ np.diag(np.repeat([3.0, 1.0, -2.0, 0.0], 30)) @ Qr.T, '4 clusters x30')
    bad += not check(Qr @ np.diag(np.logspace(2, -17, 120)) @ Qr.T, 'graded 1e2..1e-17')
    T = np.diag(2.0 * np.ones(150)) - np.diag(np.ones(149), 1) - np.diag(np.ones(149), -1)
    bad += not check(T, '1-2-1 tridiagonal')
    W = np.diag(np.abs(np.arange(-40, 41)).astype(float)) + np.diag(np.ones(80), 1) + np.diag(np.ones(80), -1)
    bad += not check(W, 'Wilkinson W81+')
    for n in (700, 1000, 2000):
        A = rng.standard_normal((n, n)); A = (A + A.T) / 2
        bad += not check(A, 'random symmetric')
        bad += not check(100.0 * A, 'random symmetric x100')

lum = synth_luminance(1024, 1024).astype(np.float64)
sel, _ = O.sample_pixels(1024, 1024, 40, 40)
Kfull = O.affinity_block(lum.ravel(), 1024, sel, sel, 500.0, 30.0)
sizes = [int(x) for x in (sys.argv[1:] or [200, 600, 1031, 1600, 2500])]
for n in sizes:
    if n <= 1600:
        idx = np.linspace(0, 1599, n).astype(int)
        A = Kfull[np.ix_(idx, idx)]
    else:
        g = int(round(np.sqrt(n)))
        sel2, _ = O.sample_pixels(1024, 1024, g, g)
        A = O.affinity_block(lum.ravel(), 1024, sel2, sel2, 500.0, 30.0)
        n = A.shape[0]
    bad += not check(A, 'Ka (bench image)')
    bad += not check(A, 'Ka again (warm)')
print('BAD CASES:', bad)
sys.exit(1 if bad else 0)
