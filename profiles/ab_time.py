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
    print(line, flush=True)
    del w
    torch.cuda.empty_cache()
