"""Lean profiling driver: N training calls (+ one enhance) on the bench workload, nothing else -- for `ncu -k regex:<kernel>`
captures that should not pay for the whole of bench.py.   python scripts/gpu_one_train.py [ntrain] [rows]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import nonlocal_image_edit_b200 as nb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rows = int(sys.argv[2]) if len(sys.argv) > 2 else bench.BASE_ROWS
nb.load().nle_b200_set_keep_stages(0)
_, lum = bench.workload_images(rows, bench.COLS)
for _ in range(n):
    f = nb.NLEFilter().trainFilter(lum, bench.GRID[0], bench.GRID[1], bench.HX, bench.HY, bench.T_SINK, bench.K_EIG)
out = f.enhanceLuminance(lum, bench.WEIGHTS)
inf = f.info()
print("ok", inf.p, inf.r, inf.r2, inf.k, int(out.sum()))
