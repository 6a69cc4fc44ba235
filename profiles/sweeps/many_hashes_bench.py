#!/usr/bin/env python
"""NtHash with many hashes per k-mer (h = 5..255) and strand outputs on C2-shaped reads: device-resident time per call
and fraction of the HBM peak (algorithmic bytes = bases + 8*h per window [+16 with strands]).
usage: python profiles/sweeps/many_hashes_bench.py [h ...]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import nthash_b200  # noqa: E402

L, K = 150, 31
peak, _ = bench.load_peak()
hs = [int(x) for x in sys.argv[1:]] or [1, 4, 5, 8, 9, 16, 32, 64, 255]
buf = bench.splitmix_bases_torch(torch, 10_000_000 * L, 42)
for h in hs:
    for strands in (False, True):
        if strands and h not in (1, 4, 16):
            continue
        n = max(50_000, min(10_000_000, int(60e9 / ((L - K + 1) * 8 * (h + (2 if strands else 0)))) // 1000 * 1000))
        bases = buf[: n * L]
        rows = n * (L - K + 1)
        out = torch.empty((rows, h), dtype=torch.int64, device="cuda")
        f = lambda: nthash_b200.kmer_hashes_uniform(bases, n, L, K, h, want_valid=False, want_strands=strands, out=out)  # noqa: E731
        for _ in range(2):
            r = f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            r = f()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        by = n * L + rows * 8 * (h + (2 if strands else 0))
        print(json.dumps({"h": h, "strands": strands, "reads": n, "ms": round(ms, 3), "GBps": round(by / ms / 1e6, 1), "frac": round(by / ms / 1e6 / peak, 3)}), flush=True)
        del out, r
        torch.cuda.empty_cache()
