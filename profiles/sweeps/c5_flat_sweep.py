#!/usr/bin/env python
"""C5 (12.5 k x 50 kb, k=63, h=1): threads per CTA x windows per store x flat item length, best / median of 12 launches.
usage: python profiles/sweeps/c5_flat_sweep.py"""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch

import bench
import nthash_b200

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
cfg = bench.CONFIGS["c5"]
n, L, k, h = cfg["n_reads"], cfg["read_len"], cfg["k"], cfg["h"]
bases = bench.splitmix_bases_torch(torch, n * L, cfg["seed"])[: n * L]
out = torch.empty((n * (L - k + 1), h), dtype=torch.int64, device="cuda")
ab = bench.algorithmic_bytes(n, L, k, h)
ref = None
grid = [(0, 0, 168)] + [(nt, ws, seg) for ws, seg in ((2, 120), (2, 200), (2, 280), (2, 360), (2, 440), (1, 288), (1, 352), (0, 168), (0, 216)) for nt in (96, 128, 160)] + [(0, 0, 168)]
for nt, ws, seg in grid:
    for key, v in (("NTHASH_B200_FAST_NT", nt), ("NTHASH_B200_FAST_WS", ws), ("NTHASH_B200_FLAT_SEG", seg)):
        if nt == 0:
            os.environ.pop(key, None)
        else:
            os.environ[key] = str(v)
    try:
        for _ in range(3):
            nthash_b200.kmer_hashes_uniform(bases, n, L, k, h, want_valid=False, out=out)
        torch.cuda.synchronize()
    except Exception as e:
        print(f"nt={nt} ws={ws} seg={seg}: skipped ({str(e)[:60]})", flush=True)
        continue
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(13)]
    ev[0].record()
    for i in range(12):
        nthash_b200.kmer_hashes_uniform(bases, n, L, k, h, want_valid=False, out=out)
        ev[i + 1].record()
    torch.cuda.synchronize()
    t = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(12))
    cs = int(out[::9973].sum())
    ref = cs if ref is None else ref
    print(f"nt={nt or 'default'} ws={ws} seg={seg}: best {t[0]:.4f} median {statistics.median(t):.4f} ms  frac(best) {ab / t[0] / 1e6 / peak:.3f}"
          f"{'' if cs == ref else '  CHECKSUM DIFFERS'}", flush=True)
