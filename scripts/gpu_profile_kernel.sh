#!/bin/bash
# One gpurun call: ncu --set full capture of ONE kernel (regex) inside a short bench run, after the same run passed plain.
# usage: scripts/gpu_profile_kernel.sh <tag> <kernel-regex> [skip-count]
set -u
TAG=${1:-k1}
KRE=${2:-gram_kernel}
SKIP=${3:-1}
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KRE} -s ${SKIP} -c 1 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu full rc=$?"
python scripts/show_bench.py gpurun_out/plain_${TAG}.log
