#!/usr/bin/env python
"""Top stall sites (SASS level) of the first kernel in an .ncu-rep.  usage: ncu_top_stalls.py report.ncu-rep [N]"""
import csv
import io
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
si, ai = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for n, r in enumerate(rows[2:]):
    if len(r) <= ai:
        continue
    v = float(r[ai] or 0)
    top = sorted(((float(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    data.append((v, n, " ".join(r[si].split())[:70], ", ".join(f"{b}={a:.0f}" for a, b in top if a)))
tot = sum(d[0] for d in data)
print(f"total samples {tot:.0f}")
for v, n, s, t in sorted(data, reverse=True)[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{v:7.0f} {v / tot * 100:5.1f}%  #{n:<5d} {s:70s} {t}")
