"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python scripts/ncu_launch_summary.py file.csv"""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg, tot = collections.OrderedDict(), 0.0
for row in csv.DictReader(lines):
    try:
        t = float(row["Metric Value"].replace(",", ""))
    except (ValueError, KeyError):
        continue
    t = t / 1000 if row["Metric Unit"] == "ns" else t * 1000 if row["Metric Unit"] == "ms" else t
    a = agg.setdefault(row["Kernel Name"].split("(")[0], [0, 0.0])
    a[0] += 1; a[1] += t; tot += t
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:9.1f} us {v[0]:5d} x {v[1]/v[0]:8.1f}  {100*v[1]/tot:5.1f} %  {k[:100]}")
print(f"{tot:9.1f} us total")
