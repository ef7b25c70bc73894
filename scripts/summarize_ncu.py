#!/usr/bin/env python
"""Summarise the ncu outputs of scripts/gpu_profile.sh into profiles/ (run here, no GPU needed).

  python scripts/summarize_ncu.py <tag>
reads  gpurun_out/launches_<tag>.csv   (ncu --metrics gpu__time_duration.sum launch list of one bench run)
       gpurun_out/prof_<tag>.ncu-rep   (ncu --set full capture of the top kernel)
       gpurun_out/bench_<tag>.log      (the bench JSON line of the same build, NOT run under ncu)
writes profiles/<tag>_launches.md, profiles/<tag>_<kernel>_full.md, profiles/<tag>_bench.json
"""
import collections
import csv
import os
import re
import shutil
import subprocess
import sys

tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go = os.path.join(root, "gpurun_out")
po = os.path.join(root, "profiles")
os.makedirs(po, exist_ok=True)

lp = os.path.join(go, f"launches_{tag}.csv")
if os.path.exists(lp):
    lines = [l for l in open(lp) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("unnamed>::", "").replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}[row["Metric Unit"]]
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    with open(os.path.join(po, f"{tag}_launches.md"), "w") as f:
        f.write(f"# ncu launch list, tag {tag}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over "
                f"`python bench.py --steps 1 --warmup 3 --no-cpu-baseline` (3 warm-up + 1 timed device-resident steps, "
                f"2 warm-up + 1 timed host-pointer steps = 7 train+enhance steps).  Times are cold-cache and serialised: "
                f"compare SHARES.\n\ntotal GPU time {tot / 1e6:.2f} ms over {sum(c for c, _ in agg.values())} launches\n\n"
                f"| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {c} | {v / 1e6:.3f} | {100 * v / tot:.2f}% |\n")
    print("wrote launches summary")

rp = os.path.join(go, f"prof_{tag}.ncu-rep")
if os.path.exists(rp):
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    keep = re.compile(r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
                      r"launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_dynamic|waves_per_multiprocessor|occupancy_limit_.*)|"
                      r"sm__warps_active\.avg\.pct_of_peak_sustained_active|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                      r"sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|sm__inst_executed_pipe_fp64\.avg\.pct_of_peak_sustained_active|"
                      r"sm__pipe_tensor.*cycles_active\.avg\.pct_of_peak_sustained_active|sm__ops_path_tensor_src_fp64\.avg\.pct_of_peak_sustained_elapsed|"
                      r"sm__issue_active\.avg\.pct_of_peak_sustained_elapsed|smsp__inst_executed\.sum|sm__cycles_elapsed\.avg|"
                      r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|lts__t_sector_hit_rate\.pct|"
                      r"lts__t_bytes\.sum|smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio)$")
    for vals in rows[2:]:
        kname = re.sub(r"\(.*", "", vals[hdr.index("Kernel Name")]).replace("void ", "").replace("nle::", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("unnamed::", "").replace("<", "_").replace(">", "")
        out = os.path.join(po, f"{tag}_{kname}_full.md")
        with open(out, "w") as f:
            f.write(f"# ncu --set full --clock-control none, kernel `{kname}`, tag {tag}\n\n"
                    f"one launch inside `python bench.py --steps 1 --warmup 3 --no-cpu-baseline`; report kept in gpurun_out/ (scratch)\n\n"
                    f"| metric | unit | value |\n|---|---|---:|\n")
            for i, h in enumerate(hdr):
                if keep.match(h):
                    f.write(f"| {h} | {units[i]} | {vals[i]} |\n")
        print("wrote", out)

bp = os.path.join(go, f"bench_{tag}.log")
if os.path.exists(bp):
    shutil.copy(bp, os.path.join(po, f"{tag}_bench.json"))
