#!/bin/bash
# developer: stage trace of the bench step (host wall-clock, stream drained at each label) + the bench line
NLE_B200_TRACE=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2> gpurun_out/trace.err > gpurun_out/trace.log
tail -40 gpurun_out/trace.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['stage_ms'], d['filter'], d['roofline']['frac'])
PY
