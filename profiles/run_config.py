#!/usr/bin/env python
"""Launches one BASELINE config's hot path a few times (for ncu / compute-sanitizer): python profiles/run_config.py c5 [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import nthash_b200  # noqa: E402
from nthash_b200._lib import LIB  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = dict(bench.CONFIGS[name])
if len(sys.argv) > 3:
    cfg["n_reads"] = int(sys.argv[3])
w = bench.Workload(torch, nthash_b200, LIB, cfg, 0)
for _ in range(steps):
    w.step(False)
torch.cuda.synchronize()
print("done", name, steps)
