#!/usr/bin/env python
"""End-to-end time of nthash_kmer_batch_uniform (pinned host buffers, 10 M x 150 bp, k=31, h=1) as a function of the host
pipeline's chunk size (NTHASH_B200_HOST_CHUNK_VALUES, hash values per chunk).  GPU only."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch

import bench
import nthash_b200
from nthash_b200._lib import LIB, check

n, L, k, h = 10_000_000, 150, 31, 1
rows = n * (L - k + 1)
d = bench.splitmix_bases_torch(torch, n * L, 42)[: n * L]
hb = torch.empty(n * L, dtype=torch.uint8).pin_memory(); hb.copy_(d)
ho = torch.empty((rows, h), dtype=torch.int64).pin_memory()
hv = torch.empty(int(LIB.nthash_valid_words(rows)), dtype=torch.int32).pin_memory()
for chunk in (8, 16, 32, 48, 64, 96, 128, 256):
    os.environ["NTHASH_B200_HOST_CHUNK_VALUES"] = str(chunk << 20)
    for with_valid in (True, False):
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            check(LIB.nthash_kmer_batch_uniform(hb.data_ptr(), n, L, k, h, ho.data_ptr(), hv.data_ptr() if with_valid else None, None, None, 0))
            ts.append(time.perf_counter() - t0)
        print(f"chunk {chunk:4d} M values ({chunk * 8} MB), valid_bits={with_valid}: best {min(ts) * 1e3:.1f} ms  ({rows / min(ts) / 1e9:.2f} G k-mers/s)", flush=True)
