#!/bin/bash
# tridiagonalisation: bit-equality / eigensolver tests, per-phase times with and without the column cache of tridiag_kernel, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eig_variants.py tests/test_gpu_topk.py -x -q > gpurun_out/trd_variants.log 2>&1; echo "variants rc=$?"; tail -2 gpurun_out/trd_variants.log
for mode in cache nocache; do
  if [ $mode = nocache ]; then export NLE_B200_TRD_NOCACHE=1; else unset NLE_B200_TRD_NOCACHE; fi
  NLE_B200_EIG_PROF=1 NLE_B200_EIG_STRICT=1 timeout 300 python - > gpurun_out/trd_times_$mode.log 2>&1 <<'PY'
import numpy as np, nonlocal_image_edit_b200 as nb
for n in (1800, 2048, 2500, 3000):
    rng = np.random.default_rng(n)
    B = rng.standard_normal((n, n // 2)); A = B @ B.T / n + 1e-3 * np.eye(n)
    for rep in range(3):
        nb.eigenDecomposition(A, eps=-1e300)
PY
  echo "== $mode rc=$?"; grep -E "eig_dc" gpurun_out/trd_times_$mode.log | awk 'NR%3==0'
done
unset NLE_B200_TRD_NOCACHE
timeout 900 python bench.py > gpurun_out/bench_trd.json 2> gpurun_out/bench_trd.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/bench_trd.json 2>/dev/null
