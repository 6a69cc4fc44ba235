"""Windows per store x threads per CTA for three read lengths (10 M reads, k=31, h=1).  NOTE: 54 configurations back to back heat the GPU;
the later rows run on lower clocks (L=150 default measured 1.80 ms alone, 2.16 ms at the end of this script): compare within a row block only."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, bench, nthash_b200
peak = 6551.7
for L in (190, 158, 150):
    n, k, h = 10_000_000, 31, 1
    bases = bench.splitmix_bases_torch(torch, n * L, 42)[: n * L]
    out = torch.empty((n * (L - k + 1), h), dtype=torch.int64, device="cuda")
    ab = bench.algorithmic_bytes(n, L, k, h)
    for ws in (0, 1, 2):
        for nt in (64, 96, 128, 160, 192, 256):
            os.environ["NTHASH_B200_FAST_WS"] = str(ws)
            os.environ["NTHASH_B200_FAST_NT"] = str(nt)
            try:
                for _ in range(2):
                    nthash_b200.kmer_hashes_uniform(bases, n, L, k, h, want_valid=False, out=out)
                torch.cuda.synchronize()
            except Exception as e:
                print(f"L={L} ws={ws} nt={nt}: skipped", flush=True); continue
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(9)]
            ev[0].record()
            for i in range(8):
                nthash_b200.kmer_hashes_uniform(bases, n, L, k, h, want_valid=False, out=out)
                ev[i + 1].record()
            torch.cuda.synchronize()
            t = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(8))
            print(f"L={L} nk={L-k+1} ws={ws} nt={nt}: best {t[0]:.4f} median {statistics.median(t):.4f} ms frac {ab / t[0] / 1e6 / peak:.3f}", flush=True)
    del bases, out
    torch.cuda.empty_cache()
