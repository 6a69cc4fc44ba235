#!/usr/bin/env python
"""FASTQ text resident on the GPU -> bases/read_off (nthash_fastq_extract_dev) -> NtHash k=31 over the result.  GPU only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch

import nthash_b200

n, L = 5_000_000, 150
g = torch.Generator(device="cuda"); g.manual_seed(1)
rec = torch.empty((n, 3 + L + 3 + L + 1), dtype=torch.uint8, device="cuda")
rec[:, 0] = ord("@"); rec[:, 1] = ord("r"); rec[:, 2] = 10
rec[:, 3:3 + L] = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")[torch.randint(0, 4, (n, L), device="cuda", generator=g)]
rec[:, 3 + L] = 10; rec[:, 4 + L] = ord("+"); rec[:, 5 + L] = 10
rec[:, 6 + L:6 + 2 * L] = 70
rec[:, 6 + 2 * L] = 10
text = rec.view(-1)
for it in range(3):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    bases, off = nthash_b200.fastq_extract(text)
    e1.record()
    res = nthash_b200.kmer_hashes(bases, off, 31, 1, want_valid=False)
    e2.record(); torch.cuda.synchronize()
print(f"FASTQ {text.numel() / 1e9:.2f} GB, {n} records: extract {e0.elapsed_time(e1):.2f} ms ({text.numel() / e0.elapsed_time(e1) / 1e6:.0f} GB/s of text), "
      f"ragged NtHash over the result {e1.elapsed_time(e2):.2f} ms; reads {off.numel() - 1}, bases {bases.numel()}")
