"""CPU: the JSON contract of `bench.py --impl reference` (the reference arm the driver runs beside the GPU arm).

The arm times the CPU restatement of the reference (oracle/nle_oracle.py) on a crop of the bench workload and must say so: its
`config` names the crop it ran (rows/cols of the crop, not of the GPU workload), `e2e` repeats the line's own value with zero
transfer bytes, `cpu_baseline` describes this run, no GPU kernel is claimed.  One step of the 384 x 384 crop (~15 s here)."""
import json
import os
import subprocess
import sys

from nle_testlib import ROOT


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ)
    env.pop("RANK", None); env.pop("WORLD_SIZE", None)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1                                     # ONE JSON line
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"] == "enhance MP/s (p=1600,k=50)" and d["unit"] == "MP/s" and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1
    assert d["dtype"] == "f64" and d["vs_baseline"] is None and d["gpu_launches"] == 0
    cfg = d["config"]
    assert cfg["rows"] == cfg["cols"] == 384 and cfg["p"] == 1600 and "crop" in cfg["workload"]      # the crop it ran, labelled as such
    assert d["e2e"] == {"value": d["value"], "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] >= 1 and "384x384" in cb["sample"]
    assert d["value"] > 0 and abs(d["value"] - 384 * 384 / 1e6 / (d["ms_per_step"] * 1e-3)) <= 1e-9 * d["value"]
