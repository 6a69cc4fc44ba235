#!/usr/bin/env python
"""Probe: how much of the ragged path's time is partial-sector stores at row boundaries?  Same read-length range, once
with arbitrary lengths and once with lengths whose window counts are multiples of 4 (every row starts and ends on a
32-byte boundary: no peel, no partial tail)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch

import bench
import nthash_b200

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
n, k, h = 10_000_000, 31, 1
for label, aligned in (("arbitrary lengths 100-150", False), ("lengths with 4 | windows (102..150)", True)):
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    if aligned:
        lens = 30 + 4 * torch.randint(18, 31, (n,), device="cuda", generator=g, dtype=torch.int64)
    else:
        lens = torch.randint(100, 151, (n,), device="cuda", generator=g, dtype=torch.int64)
    off = torch.zeros(n + 1, dtype=torch.int64, device="cuda"); off[1:] = torch.cumsum(lens, 0)
    nb = int(off[-1])
    bases = bench.splitmix_bases_torch(torch, (nb + 31) // 32 * 32, 95)[:nb]
    rows = int((lens - k + 1).sum())
    ab = nb + rows * h * 8
    out = torch.empty((rows, h), dtype=torch.int64, device="cuda")
    plan = nthash_b200.RaggedPlan(off, k)
    for env in ("0", "1"):
        os.environ["NTHASH_B200_FAST_DIRECT"] = env
        f = lambda: nthash_b200.kmer_hashes_planned(plan, bases, h, want_valid=False, out=out)  # noqa: E731
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"{label}, direct={env}: {ms:.3f} ms = {ab / ms / 1e6 / peak:.3f} of the HBM peak", flush=True)
    del bases, out, plan
