#!/usr/bin/env python
"""A/B timing helper: device-resident ms per launch (best / median of N) of the BASELINE configs and the reduce consumers
with whatever libnthash_b200.so is in place.  python profiles/ab_time.py [tag] [reps]"""
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import nthash_b200  # noqa: E402
from nthash_b200._lib import LIB  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "lib"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 12


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    t = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return t[0], statistics.median(t)


for name in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["c2", "c3", "c4", "c5"]):
    cfg = dict(bench.CONFIGS[name])
    w = bench.Workload(torch, nthash_b200, LIB, cfg, 0)
    best, med = timed(lambda: w.step(False))
    ab = bench.algorithmic_bytes(w.n, w.L, w.k, w.H)
    line = f"{tag} {name}: best {best:.4f} ms median {med:.4f} ms  frac(best) {ab / best / 1e6 / 6551.7:.3f} frac(median) {ab / med / 1e6 / 6551.7:.3f}"
    if w.seeds:
        b2, m2 = timed(lambda: nthash_b200.seed_reduce_uniform(w.plan, w.bases, w.n, w.L))
        line += f" | seed reduce best {b2:.3f} median {m2:.3f} ms"
    else:
        b2, m2 = timed(lambda: nthash_b200.kmer_reduce_uniform(w.bases, w.n, w.L, w.k, w.h))
        line += f" | reduce best {b2:.4f} median {m2:.4f} ms"
    if not w.seeds and w.L <= 250 and (w.nk * w.h) % 8 == 0 and w.h <= 4:
        # the same reads as 2-bit packed bytes, hashed directly (nthash_kmer_batch_packed2bit_uniform_dev)
        lut2 = torch.zeros(256, dtype=torch.uint8, device="cuda")
        lut2[torch.tensor(list(b"ACGT"), device="cuda").long()] = torch.arange(4, dtype=torch.uint8, device="cuda")
        codes = lut2[w.bases.long()]
        if codes.numel() % 4:
            codes = torch.cat([codes, torch.zeros(4 - codes.numel() % 4, dtype=torch.uint8, device="cuda")])
        c4 = codes.view(-1, 4)
        packed = torch.zeros(c4.shape[0] + 64, dtype=torch.uint8, device="cuda")
        packed[: c4.shape[0]] = c4[:, 0] | (c4[:, 1] << 2) | (c4[:, 2] << 4) | (c4[:, 3] << 6)
        del codes, c4, lut2
        ref_sum = int(w.out.sum())
        b3, m3 = timed(lambda: nthash_b200.kmer_hashes_packed2bit_uniform(packed, None, 0, w.n, w.L, w.k, w.h, want_valid=False, out=w.out))
        abp = w.n * w.L // 4 + w.rows * w.H * 8
        line += f" | packed2bit direct best {b3:.4f} median {m3:.4f} ms ({abp / b3 / 1e6:.0f} GB/s of its own {abp / 1e9:.2f} GB; same sum: {int(w.out.sum()) == ref_sum})"
        del packed
    print(line, flush=True)
    del w
    torch.cuda.empty_cache()
