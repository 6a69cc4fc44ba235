#!/usr/bin/env python
"""Sweeps the launch knobs of kmer_fast_kernel (threads per CTA, windows per store, buffers per lane) on the
bench configurations and prints kernel time / fraction of the measured HBM peak.  GPU only.
usage: python profiles/sweeps/kmer_fast_sweep.py c2 [c3 c5 ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch

import bench
import nthash_b200

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
grid = [(nt, ws, nb) for nt in (96, 128, 160, 192, 256) for ws in (0, 1, 2) for nb in (1, 2)]
for name in sys.argv[1:] or ["c2"]:
    cfg = bench.CONFIGS[name]
    n, L, k, h = cfg["n_reads"], cfg["read_len"], cfg["k"], cfg["h"]
    bases = bench.splitmix_bases_torch(torch, n * L, cfg["seed"])[: n * L]
    rows = n * (L - k + 1)
    out = torch.empty((rows, h), dtype=torch.int64, device="cuda")
    ab = bench.algorithmic_bytes(n, L, k, h)
    ref = None
    for nt, ws, nb in (grid if L <= 400 else [(nt, ws, nb) for nt in (64, 96, 128, 192, 256) for ws in (0, 1) for nb in (1, 2)]):
        os.environ["NTHASH_B200_FAST_NT"] = str(nt)
        os.environ["NTHASH_B200_FAST_WS"] = str(ws)
        os.environ["NTHASH_B200_FAST_NBUF"] = str(nb)
        try:
            for _ in range(2):
                nthash_b200.kmer_hashes_uniform(bases, n, L, k, h, want_valid=False, out=out)
            torch.cuda.synchronize()
        except Exception as e:  # configuration does not fit shared memory
            print(f"{name} nt={nt} ws={ws} nbuf={nb}: skipped ({str(e)[:60]})", flush=True)
            continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            nthash_b200.kmer_hashes_uniform(bases, n, L, k, h, want_valid=False, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        cs = int(out[:: 9973].sum())
        ref = cs if ref is None else ref
        print(f"{name} nt={nt} ws={ws} nbuf={nb}: {ms:.3f} ms  {ab / ms / 1e6:.0f} GB/s  frac {ab / ms / 1e6 / peak:.3f}"
              f"{'' if cs == ref else '  CHECKSUM DIFFERS'}", flush=True)
    del bases, out
    torch.cuda.empty_cache()
