#!/bin/bash
# First GPU call of the next round (about 1 minute of box time):
#   1. bit-equality of every tridiagonalisation variant with the default kernel (opt-in pytest);
#   2. time per eigensolve of the variants at the three sizes of the bench step (612, 1041, 1600), with the
#      per-phase cycle counters of the resident kernel.
# Usage: gpurun --timeout 240 -- 'bash scripts/gpu_trd_sweep.sh'
mkdir -p gpurun_out
NLE_B200_TEST_TRD_RESIDENT=1 timeout 150 python -m pytest tests/test_gpu_trd_resident.py -x -q > gpurun_out/trd_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/trd_pytest.log
tail -3 gpurun_out/trd_pytest.log
TRD_QUICK=1 timeout 80 python scripts/gpu_trd_resident.py 612 1041 1600 \
    resident cluster:2 cluster:4 cluster:8 dyn:4 dyn:8 dyn:16 dyn:32 grid:37 grid:74 grid:111 resident+grid:111 > gpurun_out/trd_sweep.log 2>&1
echo "sweep rc=$?" >> gpurun_out/trd_sweep.log
grep -v "^\[eig n" gpurun_out/trd_sweep.log | sed 's/, divide&conquer.*//' | tail -150
