#!/usr/bin/env python
"""Launches the ragged NtHash path a few times (for ncu): python profiles/run_ragged.py [n lo hi k h steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import nthash_b200  # noqa: E402

a = [int(x) for x in sys.argv[1:]]
n, lo, hi, k, h, steps = (a + [10_000_000, 100, 150, 31, 1, 3][len(a):])[:6]
g = torch.Generator(device="cuda"); g.manual_seed(5)
lens = torch.randint(lo, hi + 1, (n,), device="cuda", generator=g, dtype=torch.int64)
off = torch.zeros(n + 1, dtype=torch.int64, device="cuda"); off[1:] = torch.cumsum(lens, 0)
nb = int(off[-1])
bases = bench.splitmix_bases_torch(torch, (nb + 31) // 32 * 32, 99)[:nb]
res = nthash_b200.kmer_hashes(bases, off, k, h, want_valid=False)
for _ in range(steps):
    nthash_b200.kmer_hashes(bases, off, k, h, want_valid=False, out=res.out)
torch.cuda.synchronize()
print("done", n, lo, hi, k, h)
