#!/bin/bash
# parity tests + bench line (developer loop)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_quick.log 2>&1; echo "gpu suite rc=$?"; tail -3 gpurun_out/gputest_quick.log
timeout 900 python bench.py --no-targets > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/bench_quick.json 2>/dev/null | head -2
