#!/bin/bash
# One gpurun call: GPU parity tests, bench line, ncu launch list of one bench run, ncu --set full of the top kernels.
# usage: scripts/gpu_profile.sh <tag> [kernel-regex]
set -u
TAG=${1:-r1}
KRE=${2:-gram_kernel}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?" 
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_${TAG}.log 2>gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_${TAG}.log
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 1 -c 1 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out
