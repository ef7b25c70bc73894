#!/bin/bash
# Round-2 check of the hybrid tridiagonalisation (tridiag_kernel partial mode -> tridiag_cluster_kernel hand-over):
# bit-equality test, per-phase times at n = 1800 / 2048 / 2500 with and without the hand-over, bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eig_variants.py -x -q > gpurun_out/trd_hybrid_test.log 2>&1; echo "variants test rc=$?"
tail -3 gpurun_out/trd_hybrid_test.log
for mode in hybrid nohybrid; do
  if [ $mode = nohybrid ]; then export NLE_B200_TRD=nohybrid; else unset NLE_B200_TRD; fi
  NLE_B200_EIG_PROF=1 NLE_B200_EIG_STRICT=1 timeout 300 python - > gpurun_out/trd_hybrid_$mode.log 2>&1 <<'PY'
import numpy as np, nonlocal_image_edit_b200 as nb
for n in (1800, 2048, 2500):
    rng = np.random.default_rng(n)
    B = rng.standard_normal((n, n // 2)); A = B @ B.T / n + 1e-3 * np.eye(n)
    for rep in range(3):
        nb.eigenDecomposition(A, eps=-1e300)
PY
  echo "timing $mode rc=$?"; grep "eig_dc" gpurun_out/trd_hybrid_$mode.log
done
unset NLE_B200_TRD
timeout 900 python bench.py > gpurun_out/bench_r2y.json 2> gpurun_out/bench_r2y.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/bench_r2y.json 2>/dev/null | head -40
