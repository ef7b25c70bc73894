"""Developer timing: grid.sync tridiagonalisation vs the shared-memory-resident one (NLE_B200_TRD=resident).

  NLE_B200_EIG_PROF=1 python scripts/gpu_trd_resident.py [n ...]     # per-phase ms on stderr, default n = 612 1041 1600
"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nonlocal_image_edit_b200 as nb

os.environ.setdefault("NLE_B200_EIG_PROF", "1")
os.environ["NLE_B200_EIG_STRICT"] = "1"
sizes = [int(a) for a in sys.argv[1:]] or [612, 1041, 1600]
for n in sizes:
    rng = np.random.default_rng(n)
    B = rng.standard_normal((n, n // 2))
    A = B @ B.T / n + 1e-3 * np.eye(n)
    out = {}
    for mode in ("gridsync", "resident"):
        if mode == "resident":
            os.environ["NLE_B200_TRD"] = "resident"
        else:
            os.environ.pop("NLE_B200_TRD", None)
        nb.eigenDecomposition(A, eps=-1e300)                      # warm-up (attributes, pool)
        t0 = time.time()
        for _ in range(3):
            out[mode] = nb.eigenDecomposition(A, eps=-1e300)
        print(f"n={n} {mode}: {(time.time() - t0) / 3 * 1e3:.2f} ms per call (host clock, includes copies)", flush=True)
    same = np.array_equal(out["gridsync"][1], out["resident"][1]) and np.array_equal(out["gridsync"][0], out["resident"][0])
    print(f"n={n} bit-identical: {same}", flush=True)
