"""Developer diagnostic: is the top-k block solver still the faster choice for eig(Q)?  Trains the weak-scaling images of 2 and 8
GPUs (2048 x 1024, 8192 x 1024; p = 1600, k = 50) and a 4096 x 1024 slab-shaped stand-in for the 16.7 MP target grid (p = 2500, k = 100)
UNSHARDED on one GPU and prints the stage times; run once as is and once with NLE_B200_TOPK=off.

  python scripts/gpu_topk_switch.py ; NLE_B200_TOPK=off python scripts/gpu_topk_switch.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import nonlocal_image_edit_b200 as nb  # noqa: E402

nb.load().nle_b200_set_keep_stages(0)
for rows, cols, grid, k in ((2048, 1024, (40, 40), 50), (8192, 1024, (40, 40), 50), (4096, 2048, (50, 50), 100)):
    _, lum = bench.workload_images(rows, cols)
    best = None
    for rep in range(3):
        f = nb.NLEFilter().trainFilter(lum, grid[0], grid[1], bench.HX, bench.HY, bench.T_SINK, k)
        m = f.stage(8)
        if best is None or m[4] < best[4]:
            best = m
    inf = f.info()
    print(f"{rows}x{cols} p={inf.p} r={inf.r} r2={inf.r2} k={inf.k} topk_products={inf.topk_products}: small_algebra_2eigs {best[4]:.2f} ms "
          f"(tridiag of the three solves {best[8]:.2f}, D&C {best[9]:.2f}, back-transformation {best[10]:.2f}), train_total {best[6]:.2f} ms", flush=True)
