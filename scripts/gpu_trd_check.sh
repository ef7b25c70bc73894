#!/bin/bash
# tridiagonalisation: bit-equality / eigensolver tests, per-phase cycles and times, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eig_variants.py tests/test_gpu_topk.py -x -q > gpurun_out/trd_variants.log 2>&1; echo "variants rc=$?"; tail -2 gpurun_out/trd_variants.log
NLE_B200_TRD_PROF=1 NLE_B200_EIG_PROF=1 NLE_B200_EIG_STRICT=1 timeout 300 python - > gpurun_out/trd_times.log 2>&1 <<'PY'
import numpy as np, nonlocal_image_edit_b200 as nb
for n in (612, 1041, 1600, 2500):
    rng = np.random.default_rng(n)
    B = rng.standard_normal((n, n // 2)); A = B @ B.T / n + 1e-3 * np.eye(n)
    for rep in range(3):
        nb.eigenDecomposition(A, eps=-1e300)
PY
echo "rc=$?"; grep -E "trd cluster|eig_dc" gpurun_out/trd_times.log | awk 'NR%6==5 || NR%6==0'
timeout 900 python bench.py --no-targets --no-cpu-baseline > gpurun_out/bench_trd.json 2> gpurun_out/bench_trd.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/bench_trd.json 2>/dev/null | head -2
