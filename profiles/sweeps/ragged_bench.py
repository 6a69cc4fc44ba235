#!/usr/bin/env python
"""Device-resident throughput of the ragged entry points (nthash_kmer_plan_dev + nthash_kmer_batch_dev) on trimmed-read
shaped batches: lengths uniform in [lo, hi].  GPU only.  usage: python profiles/sweeps/ragged_bench.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch

import bench
import nthash_b200

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
for n, lo, hi, k, h in ((10_000_000, 100, 150, 31, 1), (10_000_000, 36, 150, 31, 1), (5_000_000, 100, 250, 31, 2), (200_000, 1000, 20000, 63, 1)):
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    lens = torch.randint(lo, hi + 1, (n,), device="cuda", generator=g, dtype=torch.int64)
    off = torch.zeros(n + 1, dtype=torch.int64, device="cuda"); off[1:] = torch.cumsum(lens, 0)
    nb = int(off[-1])
    bases = bench.splitmix_bases_torch(torch, (nb + 31) // 32 * 32, 99)[:nb]
    rows = int(torch.clamp(lens - k + 1, min=0).sum())
    ab = nb + rows * h * 8
    res = nthash_b200.kmer_hashes(bases, off, k, h, want_valid=False)
    torch.cuda.synchronize()
    out = res.out
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        nthash_b200.kmer_hashes(bases, off, k, h, want_valid=False, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"ragged n={n} len[{lo},{hi}] k={k} h={h}: {ms:.3f} ms per call (layout scan + item planning + kernel), {rows / ms / 1e6:.1f} G k-mers/s, "
          f"{ab / ms / 1e6:.0f} GB/s = {ab / ms / 1e6 / peak:.3f} of the HBM peak", flush=True)
    del bases, out, res
    torch.cuda.empty_cache()

# SeedNtHash (configs[3] seeds, 3 hashes each) on ragged reads: the ragged variant of the specialised kernel vs the generic one
seeds = bench.CONFIGS["c4"]["seeds"]
plan = nthash_b200.SeedPlan(seeds, 3)
for n, lo, hi in ((5_000_000, 100, 150),):
    g = torch.Generator(device="cuda"); g.manual_seed(6)
    lens = torch.randint(lo, hi + 1, (n,), device="cuda", generator=g, dtype=torch.int64)
    off = torch.zeros(n + 1, dtype=torch.int64, device="cuda"); off[1:] = torch.cumsum(lens, 0)
    nb = int(off[-1])
    bases = bench.splitmix_bases_torch(torch, (nb + 31) // 32 * 32, 98)[:nb]
    rows = int(torch.clamp(lens - 31 + 1, min=0).sum())
    ab = nb + rows * 6 * 8
    for label, env in (("specialised (ragged variant)", None), ("generic interpreter", "1")):
        if env:
            os.environ["NTHASH_B200_DISABLE_SEED_JIT"] = env
        else:
            os.environ.pop("NTHASH_B200_DISABLE_SEED_JIT", None)
        res = nthash_b200.seed_hashes(plan, bases, off, want_valid=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            res = nthash_b200.seed_hashes(plan, bases, off, want_valid=False)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"ragged SeedNtHash n={n} len[{lo},{hi}] 2 seeds x 3, {label}: {ms:.3f} ms per call, {rows / ms / 1e6:.1f} G windows/s, "
              f"{ab / ms / 1e6:.0f} GB/s = {ab / ms / 1e6 / peak:.3f} of the HBM peak", flush=True)
        del res
