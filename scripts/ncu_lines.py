#!/usr/bin/env python
"""Per-source-line stall samples / executed instructions of an ncu report (developer helper).
   python scripts/ncu_lines.py gpurun_out/prof_k8.ncu-rep [min_pct]"""
import csv, subprocess, sys
rep = sys.argv[1]
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
iS, iE = hdr.index("# Samples"), hdr.index("Instructions Executed")
lines = {}
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or not r[0]:
        continue       # SASS rows have an empty line number; the line rows carry the per-line totals
    try:
        s, e = int(r[iS]), int(r[iE])
    except ValueError:
        continue
    k = int(r[0])
    old = lines.get(k, (0, 0, r[1]))
    lines[k] = (old[0] + s, old[1] + e, r[1])
totS = sum(v[0] for v in lines.values()) or 1
totE = sum(v[1] for v in lines.values()) or 1
print(f"total samples {totS}, warp instructions {totE}")
for k in sorted(lines):
    s, e, src = lines[k]
    if 100 * s / totS >= minpct or 100 * e / totE >= minpct:
        print(f"{k:>5} samp {100*s/totS:5.1f}% inst {100*e/totE:5.1f}%  {src.strip()[:110]}")
