#!/bin/bash
# What the driver runs at round end, in one call: GPU tests, smoke(), the bench line, the reference arm.
TAG=${1:-final}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_${TAG}.log 2>&1; echo "gpu suite rc=$?"; tail -2 gpurun_out/gputest_${TAG}.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_${TAG}.log
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; echo "reference arm rc=$?"
python scripts/show_bench.py gpurun_out/bench_${TAG}.json
