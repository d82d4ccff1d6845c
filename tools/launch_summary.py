#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time and share.
Usage: python tools/launch_summary.py gpurun_out/launches.csv"""
import collections
import csv
import sys

lines = open(sys.argv[1]).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(lines[start:]))
agg = collections.OrderedDict()
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    k = r["Kernel Name"].split("(")[0]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r["Metric Value"].replace(",", "")) / 1e3
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{v[1]:10.1f} us {v[0]:4d}x {100 * v[1] / tot:5.1f}%  {k}")
