#!/bin/bash
# One gpurun call: bench line, ncu launch list of one bench run, optional ncu --set full of named kernels.
# usage: scripts/gpu_profile_launches.sh <tag> [kernel-regex ...]
set -u
TAG=${1:-r1j}; shift
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.log 2>gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-targets > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-targets > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "ncu launches rc=$?"
i=0
for KRE in "$@"; do
  i=$((i+1))
  ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 3 -c 1 -f -o gpurun_out/prof_${TAG}$i \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-targets > gpurun_out/ncu_full_${TAG}$i.log 2>&1
  echo "ncu full ${KRE} rc=$?"
done
