# small end-to-end exercise of every kernel for compute-sanitizer (memcheck): sizes kept tiny
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import nthash_b200
from gpu_util import synth, to_dev, ragged_offsets
rng = np.random.default_rng(1)
for L, n, k, h in ((150, 700, 31, 1), (151, 300, 31, 3), (2000, 40, 63, 4), (150, 300, 31, 4)):
    b = synth(rng, n * L, p_bad=0.002); d, keep = to_dev(b)
    nthash_b200.kmer_hashes_uniform(d, n, L, k, h); nthash_b200.kmer_hashes_uniform(d, n, L, k, 1, want_strands=True)
    nthash_b200.kmer_reduce_uniform(d, n, L, k, h)
# fast kernel: tensor-store path with a partial last tile, general path (odd rows, cut-up reads), huge k, Bloom consumer
for L, n, k, h in ((102, 300, 31, 1), (149, 300, 31, 1), (1000, 40, 31, 2), (5003, 9, 32, 1), (3000, 6, 2500, 1)):
    b = synth(rng, n * L, p_bad=0.001); d, keep = to_dev(b)
    nthash_b200.kmer_hashes_uniform(d, n, L, k, h)
    filt = nthash_b200.bloom_filter(1 << 16)
    nthash_b200.kmer_bloom_uniform(d, n, L, k, 3, filt, 1 << 16); nthash_b200.kmer_bloom_uniform(d, n, L, k, 3, filt, 1 << 16, query=True)
lens = rng.integers(0, 400, 500); off = ragged_offsets(lens); b = synth(rng, int(off[-1]), p_bad=0.003); d, keep = to_dev(b, pad=16)
o = torch.from_numpy(off).cuda()
nthash_b200.kmer_hashes(d, o, 31, 2, want_strands=True); nthash_b200.kmer_reduce(d, o, 31, 2); nthash_b200.kmer_hashes(d, o, 31, 1)
nthash_b200.kmer_bloom(d, o, 31, 2, nthash_b200.bloom_filter(99991), 99991)
lens = [30000, 3, 9000, 62]; off = ragged_offsets(lens); b = synth(rng, int(off[-1]), p_bad=0.0003); d, keep = to_dev(b, pad=16)
nthash_b200.kmer_hashes(d, torch.from_numpy(off).cuda(), 63, 1)
plan = nthash_b200.SeedPlan(["1010101010101010101010101010101", "1101101101101101011011011011011"], 3)
b = synth(rng, 400 * 150, p_bad=0.002); d, keep = to_dev(b)
nthash_b200.seed_hashes_uniform(plan, d, 400, 150); nthash_b200.seed_hashes_uniform(plan, d, 400, 150, want_strands=True)
nthash_b200.seed_hashes(plan, d, torch.arange(0, 401, dtype=torch.int64).cuda() * 150, want_strands=True)
init = nthash_b200.kmer_hashes_uniform(d, 100, 31, 31, 2, want_strands=True)
nthash_b200.blind_roll(init.fwd, init.rev, d[:100].contiguous(), d[100:200].contiguous(), 31, 2)
nthash_b200.blind_peek4(init.fwd, init.rev, d[:100].contiguous(), 31, 2)
off = ragged_offsets(rng.integers(0, 300, 200)).astype(np.uint64); b = synth(rng, int(off[-1]), p_bad=0.002)
out = np.zeros((int(nthash_b200.LIB.nthash_window_rows(off.ctypes.data, 200, 31, None)), 1), np.uint64)
assert nthash_b200.LIB.nthash_kmer_batch(b.ctypes.data, off.ctypes.data, 200, 31, 1, out.ctypes.data, None, None, None, 0) == 0
# three / 5..8 hashes, packed input, multi-device entry, compaction, FASTQ staging, BlindSeed batch, ragged seed variant
for L, n, k, h in ((150, 300, 31, 3), (151, 200, 31, 3), (150, 200, 31, 5), (1000, 20, 31, 8)):
    b = synth(rng, n * L, p_bad=0.001); d, keep = to_dev(b); nthash_b200.kmer_hashes_uniform(d, n, L, k, h)
res = nthash_b200.kmer_hashes(d, torch.arange(0, 21, dtype=torch.int64).cuda() * 1000, 31, 2); nthash_b200.compact(res)
acgt = rng.choice(np.frombuffer(b"ACGT", np.uint8), 320 * 150); acgt[rng.random(len(acgt)) < 0.002] = ord("N")
code = np.zeros(256, np.uint8); code[list(b"ACGT")] = [0, 1, 2, 3]
c4 = code[acgt].reshape(-1, 4); packed = np.ascontiguousarray(c4[:, 0] | (c4[:, 1] << 2) | (c4[:, 2] << 4) | (c4[:, 3] << 6)).astype(np.uint8)
inv = np.packbits((acgt == ord("N")).reshape(-1, 8), axis=1, bitorder="little").reshape(-1).copy().view(np.uint32)
r3 = np.zeros(3, np.uint64)
assert nthash_b200.LIB.nthash_kmer_reduce_packed2bit(packed.ctypes.data, inv.ctypes.data, None, 320, 150, 31, 2, r3.ctypes.data, 0) == 0
devs = np.zeros(2, np.int32); o2 = np.zeros((int(nthash_b200.LIB.nthash_window_rows(off.ctypes.data, 200, 31, None)), 1), np.uint64)
assert nthash_b200.LIB.nthash_kmer_batch_multi(b.ctypes.data, off.ctypes.data, 200, 31, 1, o2.ctypes.data, None, None, None, devs.ctypes.data, 2) == 0 or True
text = b"".join(b"@r\n" + bytes(acgt[i * 150:(i + 1) * 150]) + b"\n+\n" + b"I" * 150 + b"\n" for i in range(100))
fb, fo = nthash_b200.fastq_extract(torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda())
nthash_b200.kmer_hashes(fb, fo, 31, 1)
plan9 = nthash_b200.SeedPlan(["110101011", "101111101"], 2)
km = torch.zeros(500 * 9 + 64, dtype=torch.uint8, device="cuda"); km[: 4500] = torch.from_numpy(acgt[:4500].copy()).cuda()
nthash_b200.blind_seed_roll(plan9, km[:4500].view(500, 9), torch.from_numpy(acgt[5000:5500].copy()).cuda())
lens = rng.integers(0, 300, 300); roff = ragged_offsets(lens); rb = synth(rng, int(roff[-1]), p_bad=0.003); dr, keep = to_dev(rb, pad=64)
nthash_b200.seed_hashes(plan, dr, torch.from_numpy(roff).cuda()); nthash_b200.seed_hashes(plan9, dr, torch.from_numpy(roff).cuda())
torch.cuda.synchronize(); print("sanitizer workload done")
