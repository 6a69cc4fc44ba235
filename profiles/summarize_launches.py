#!/usr/bin/env python
"""Condenses an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals.
usage: summarize_launches.py launches.csv > summary.txt"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
agg = OrderedDict()
for r in rows[1:]:
    name = re.sub(r"<.*", "", r[ki].replace("void ", ""))[:60]
    if "nthb::" in r[ki]:
        name = re.sub(r"\(.*", "", r[ki].replace("void ", ""))
    a = agg.setdefault(name, [0, 0.0, r[gi], r[bi]])
    a[0] += 1
    a[1] += float(r[vi].replace(",", ""))
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':70s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}  grid block")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:70s} {a[0]:8d} {a[1]/1e3:12.1f} {a[1]/a[0]/1e3:10.1f} {a[1]/tot*100:6.1f}%  {a[2]} {a[3]}")
