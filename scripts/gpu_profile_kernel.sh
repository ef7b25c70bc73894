#!/bin/bash
# One gpurun call: ncu --set full (with source) of ONE launch of each named kernel inside scripts/gpu_one_train.py.
# usage: scripts/gpu_profile_kernel.sh <tag> <kernel-regex> [<kernel-regex> ...]      (launch-skip 1 = second training call)
set -u
TAG=$1; shift
mkdir -p gpurun_out
python scripts/gpu_one_train.py 2 > gpurun_out/plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG}.log; exit 1; }
i=0
for KRE in "$@"; do
  i=$((i+1))
  ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 1 -c 1 -f -o gpurun_out/prof_${TAG}$i \
      python scripts/gpu_one_train.py 2 > gpurun_out/ncu_full_${TAG}$i.log 2>&1
  echo "ncu full ${KRE} rc=$?"
done
