#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total
time and share.  usage: summarize_launches.py launches.csv "title" > summary.txt"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith("==")]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows[1:]:
    if len(r) <= vi:
        continue
    scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[ui], 1.0)
    name = r[ki].split("(")[0]
    e = agg.setdefault(name, [0, 0.0])
    e[0] += 1
    e[1] += float(r[vi].replace(",", "")) * scale
tot = sum(v[1] for v in agg.values())
print(sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
print("kernel, launches, total_ns, share")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:120]}, {v[0]}, {v[1]:.0f}, {v[1] / tot:.4f}")
