"""Developer probe: one tridiagonalisation-dominated eigensolve per size, with the per-phase cycle counters
(NLE_B200_TRD_PROF=1) and the wall-clock per phase (NLE_B200_EIG_PROF=1) on stderr.
   python scripts/gpu_trd_probe.py [n ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nonlocal_image_edit_b200 as nb  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [1600]
for n in sizes:
    rng = np.random.default_rng(n)
    B = rng.standard_normal((n, n // 2))
    A = B @ B.T / n + 1e-3 * np.eye(n)
    for rep in range(3):
        U, D = nb.eigenDecomposition(A, eps=-1e300)
    print("n", n, "ok", float(D[0]))
