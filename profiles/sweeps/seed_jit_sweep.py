#!/usr/bin/env python
"""Sweeps the build knobs of the NVRTC-specialised SeedNtHash kernel (threads per CTA, windows per tile row,
tile buffers per warp, 3-D box stores on/off) on C4 and prints kernel time / fraction of the measured HBM peak.
usage: python profiles/sweeps/seed_jit_sweep.py [reads]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch

import bench
import nthash_b200

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
cfg = bench.CONFIGS["c4"]
n = int(sys.argv[1]) if len(sys.argv) > 1 else cfg["n_reads"]
L, k, h, seeds = cfg["read_len"], cfg["k"], cfg["h"], cfg["seeds"]
H = h * len(seeds)
bases = bench.splitmix_bases_torch(torch, n * L, 42)[: n * L]
out = torch.empty((n * (L - k + 1), H), dtype=torch.int64, device="cuda")
ab = bench.algorithmic_bytes(n, L, k, H)
ref = None
grid = [(1, tw, nt, nb) for tw in (4, 8) for nt in (128, 192, 256, 320) for nb in (1, 2)] + [(0, 5, 256, 1), (0, 3, 256, 2)]
if os.environ.get("SEED_SWEEP_FINE"):
    grid = [(1, 4, nt, nb) for nt in (64, 96, 128, 160, 224) for nb in (2, 3)]
for box, tw, nt, nb in grid:
    os.environ["NTHASH_B200_SEED_JIT_TW"] = str(tw)
    os.environ["NTHASH_B200_SEED_JIT_NT"] = str(nt)
    os.environ["NTHASH_B200_SEED_JIT_NBUF"] = str(nb)
    if box:
        os.environ.pop("NTHASH_B200_SEED_JIT_NO_BOX", None)
    else:
        os.environ["NTHASH_B200_SEED_JIT_NO_BOX"] = "1"
    try:
        plan = nthash_b200.SeedPlan(seeds, h)
        for _ in range(2):
            nthash_b200.seed_hashes_uniform(plan, bases, n, L, want_valid=False, out=out)
        torch.cuda.synchronize()
    except Exception as e:
        print(f"box={box} tw={tw} nt={nt} nbuf={nb}: skipped ({str(e)[:70]})", flush=True)
        continue
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        nthash_b200.seed_hashes_uniform(plan, bases, n, L, want_valid=False, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    cs = int(out[:: 9973].sum())
    ref = cs if ref is None else ref
    print(f"box={box} tw={tw} nt={nt} nbuf={nb}: {ms:.3f} ms  {ab / ms / 1e6:.0f} GB/s  frac {ab / ms / 1e6 / peak:.3f}"
          f"{'' if cs == ref else '  CHECKSUM DIFFERS'}  [{plan.kernel_note() if hasattr(plan, 'kernel_note') else ''}]", flush=True)
