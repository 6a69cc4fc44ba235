#!/usr/bin/env python
"""Device-resident throughput of the ragged entry points on trimmed-read shaped batches (lengths uniform in [lo, hi]):
planned calls (nthash_ragged_plan_create once, then nthash_kmer_batch_planned_dev: kernels only) and unplanned ones
(nthash_kmer_plan_dev + nthash_kmer_batch_dev per call).  GPU only.  usage: python profiles/sweeps/ragged_bench.py [kmer|seed|all]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch

import bench
import nthash_b200

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
what = sys.argv[1] if len(sys.argv) > 1 else "all"


def timed(f, reps=5):
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def make(n, lo, hi, seed):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    lens = torch.randint(lo, hi + 1, (n,), device="cuda", generator=g, dtype=torch.int64)
    off = torch.zeros(n + 1, dtype=torch.int64, device="cuda"); off[1:] = torch.cumsum(lens, 0)
    nb = int(off[-1])
    bases = bench.splitmix_bases_torch(torch, (nb + 31) // 32 * 32, 90 + seed)[:nb]
    return lens, off, nb, bases


if what in ("kmer", "all"):
    for n, lo, hi, k, h in ((10_000_000, 100, 150, 31, 1), (10_000_000, 36, 150, 31, 1), (5_000_000, 100, 250, 31, 2), (200_000, 1000, 20000, 63, 1)):
        lens, off, nb, bases = make(n, lo, hi, 5)
        rows = int(torch.clamp(lens - k + 1, min=0).sum())
        ab = nb + rows * h * 8
        out = torch.empty((rows, h), dtype=torch.int64, device="cuda")
        plan = nthash_b200.RaggedPlan(off, k)
        for label, env in (("rows via shared memory", "0"), ("direct 32-byte stores", "1")):
            os.environ["NTHASH_B200_FAST_DIRECT"] = env
            ms_p = timed(lambda: nthash_b200.kmer_hashes_planned(plan, bases, h, want_valid=False, out=out))
            ms_u = timed(lambda: nthash_b200.kmer_hashes(bases, off, k, h, want_valid=False, out=out))
            print(f"ragged n={n} len[{lo},{hi}] k={k} h={h} [{label}]: planned {ms_p:.3f} ms = {ab / ms_p / 1e6 / peak:.3f} of the HBM peak "
                  f"({rows / ms_p / 1e6:.1f} G k-mers/s); unplanned (layout scan + planning + kernel) {ms_u:.3f} ms = {ab / ms_u / 1e6 / peak:.3f}", flush=True)
        del bases, out, plan
        torch.cuda.empty_cache()
    os.environ.pop("NTHASH_B200_FAST_DIRECT", None)

if what in ("seed", "all"):
    # SeedNtHash (configs[3] seeds, 3 hashes each) on ragged reads: the ragged variant of the specialised kernel vs the generic one
    seeds = bench.CONFIGS["c4"]["seeds"]
    sp = nthash_b200.SeedPlan(seeds, 3)
    for n, lo, hi in ((5_000_000, 100, 150),):
        lens, off, nb, bases = make(n, lo, hi, 6)
        rows = int(torch.clamp(lens - 31 + 1, min=0).sum())
        ab = nb + rows * 6 * 8
        plan = nthash_b200.RaggedPlan(off, 31)
        out = torch.empty((rows, 6), dtype=torch.int64, device="cuda")
        for label, env in (("specialised (ragged variant)", None), ("generic interpreter", "1")):
            if env:
                os.environ["NTHASH_B200_DISABLE_SEED_JIT"] = env
            else:
                os.environ.pop("NTHASH_B200_DISABLE_SEED_JIT", None)
            ms_p = timed(lambda: nthash_b200.seed_hashes_planned(sp, plan, bases, want_valid=False, out=out), 3)
            ms_u = timed(lambda: nthash_b200.seed_hashes(sp, bases, off, want_valid=False), 3)
            print(f"ragged SeedNtHash n={n} len[{lo},{hi}] 2 seeds x 3, {label}: planned {ms_p:.3f} ms = {ab / ms_p / 1e6 / peak:.3f} of the HBM peak "
                  f"({rows / ms_p / 1e6:.1f} G windows/s); unplanned {ms_u:.3f} ms = {ab / ms_u / 1e6 / peak:.3f}", flush=True)
