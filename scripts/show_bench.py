"""Print the interesting fields of a bench.py JSON line (developer helper)."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print({k: round(v, 2) for k, v in d['stage_ms'].items()})
print(d['filter'], 'roofline', {k: d['roofline'][k] for k in ('achieved', 'peak', 'frac', 'launch_ms')}, d.get('clocks'))
print('per-step ms', d.get('step_ms'))
