#!/bin/bash
# One gpurun call: GPU tests, bench line (with CPU baseline), reference arm, ncu launch list, ncu --set full of the named kernels.
# usage: scripts/gpu_profile_final.sh <tag> <kernel-regex-1> [kernel-regex-2 ...]
set -u
TAG=${1:-r1g}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_${TAG}.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.log 2>gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.log 2>gpurun_out/bench_ref_${TAG}.err; echo "reference arm rc=$?"
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "ncu launches rc=$?"
i=0
for KRE in "$@"; do
  i=$((i+1))
  ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_$i \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}_$i.log 2>&1
  echo "ncu full ${KRE} rc=$?"
done
