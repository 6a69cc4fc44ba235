"""BASELINE.json configs[2], [3] and [4] at their FULL sizes on the GPU (configs[1] has its own test in
test_gpu_kmer.py).  At these sizes the oracle cannot hash everything in seconds, so each test combines
  * a bit-exact comparison of a deterministic sample of reads with the oracle,
  * a whole-prefix 64-bit checksum against the threaded oracle, and
  * size-independent properties of the domain over the WHOLE batch (strand symmetry, reference tests.cpp:119-133 /
    :349-377; the NTM64 extension formula, src/internal.hpp:104-118; agreement of independent output paths).
Inputs are the bench's own reads (splitmix64 stream of SURVEY.md 8d), so what is checked here is what bench.py times."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

import bench  # noqa: E402  (generator + config table only)
from gpu_util import u64  # noqa: E402
from oracle_lib import ORACLE  # noqa: E402

M64 = (1 << 64) - 1
MULTISEED = 0x90B45D39FB6DA1FA


def _s64(v):
    v &= M64
    return v - (1 << 64) if v >= (1 << 63) else v


def _revcomp(bases_2d):
    """Reverse complement of ASCII reads [n, L] on the device."""
    lut = torch.zeros(256, dtype=torch.uint8, device="cuda")
    for a, b in zip(b"ACGT", b"TGCA"):
        lut[a] = b
    return lut[bases_2d.flip(1).long()].contiguous()


def _reads(cfg):
    import nthash_b200  # noqa: F401
    n, L = cfg["n_reads"], cfg["read_len"]
    buf = bench.splitmix_bases_torch(torch, n * L, cfg["seed"])
    return buf, buf[: n * L]


def test_full_size_config3_h4():
    """configs[2]: 10 M x 150 bp, k=31, h=4 (38.4 GB of hashes)."""
    import nthash_b200
    cfg = bench.CONFIGS["c3"]
    n, L, k, h = cfg["n_reads"], cfg["read_len"], cfg["k"], cfg["h"]
    nk = L - k + 1
    buf, bases = _reads(cfg)
    out = nthash_b200.kmer_hashes_uniform(bases, n, L, k, h, want_valid=False).out
    torch.cuda.synchronize()
    # NTM64 over the whole batch: hash[j] = t ^ (t >> 27), t = hash[0] * (j ^ k*MULTISEED)   (internal.hpp:104-118)
    h0 = out[:, 0]
    for j in range(1, h):
        bad = 0
        for a in range(0, out.shape[0], 1 << 27):
            t = h0[a:a + (1 << 27)] * _s64(j ^ ((k * MULTISEED) & M64))
            e = t ^ ((t >> 27) & ((1 << 37) - 1))
            bad += int((e != out[a:a + (1 << 27), j]).sum())
        assert bad == 0, f"hash[{j}] is not the NTM64 extension of hash[0] on {bad} windows"
    # hash[0] is what h=1 produces (configs[1]'s kernel instance): two template instances agree on 1.2e9 windows
    h1 = nthash_b200.kmer_hashes_uniform(bases, n, L, k, 1, want_valid=False).out
    assert torch.equal(h1[:, 0], h0)
    del h1
    # the fused consumer reaches the same checksum without storing anything
    red = nthash_b200.kmer_reduce_uniform(bases, n, L, k, h)
    assert int(red[0]) == n * nk and (int(red[1]) & M64) == (int(out.sum()) & M64)
    # sampled reads, bit for bit
    idx = torch.arange(0, n, 9973, device="cuda")
    sample = bases.view(n, L)[idx].cpu().numpy()
    ora = ORACLE.kmer_batch(sample.reshape(-1), np.arange(len(idx) + 1, dtype=np.uint64) * L, k, h, threads=8)
    assert (u64(out.view(n, nk, h)[idx]).reshape(-1, h) == ora["out"]).all()
    # whole-prefix checksum against the threaded oracle
    m = 500_000
    pre = ORACLE.kmer_batch(bases[: m * L].cpu().numpy(), np.arange(m + 1, dtype=np.uint64) * L, k, h, want=(), threads=os.cpu_count() or 1)
    assert pre["n_emit"] == m * nk and pre["sum"] == (int(out[: m * nk].sum()) & M64)


def test_full_size_config4_spaced_seeds():
    """configs[3]: SeedNtHash, 10 M x 150 bp, two palindromic seeds x 3 hashes (57.6 GB of hashes)."""
    import nthash_b200
    cfg = bench.CONFIGS["c4"]
    n, L, k, h, seeds = cfg["n_reads"], cfg["read_len"], cfg["k"], cfg["h"], cfg["seeds"]
    nk, H = L - k + 1, h * len(seeds)
    buf, bases = _reads(cfg)
    plan = nthash_b200.SeedPlan(seeds, h)
    assert plan.symmetric
    res = nthash_b200.seed_hashes_uniform(plan, bases, n, L)
    out = res.out
    torch.cuda.synchronize()
    assert int(res.valid_bits[: (n * nk) // 32].ne(-1).sum()) == 0  # clean reads: every window is visited
    # strand symmetry of palindromic seeds (tests.cpp:349-377) on the whole batch, in slabs that fit next to `out`
    slab = 1_000_000
    for a in range(0, n, slab):
        b = min(n, a + slab)
        rc = torch.zeros((b - a) * L + 64, dtype=torch.uint8, device="cuda")
        rc[: (b - a) * L] = _revcomp(bases[a * L: b * L].view(b - a, L)).view(-1)
        o2 = nthash_b200.seed_hashes_uniform(plan, rc[: (b - a) * L], b - a, L, want_valid=False).out
        assert torch.equal(out[a * nk: b * nk].view(b - a, nk, H), o2.view(b - a, nk, H).flip(1)), f"strand symmetry broken in reads {a}..{b}"
        del o2, rc
    # NTM64 per seed over the whole batch
    for s in range(len(seeds)):
        for j in range(1, h):
            t = out[:, s * h] * _s64(j ^ ((k * MULTISEED) & M64))
            assert torch.equal(t ^ ((t >> 27) & ((1 << 37) - 1)), out[:, s * h + j])
            del t
    # sampled reads and a whole prefix against the oracle
    idx = torch.arange(0, n, 19997, device="cuda")
    sample = bases.view(n, L)[idx].cpu().numpy()
    ora = ORACLE.seed_batch(sample.reshape(-1), np.arange(len(idx) + 1, dtype=np.uint64) * L, seeds, h, threads=8)
    assert (u64(out.view(n, nk, H)[idx]).reshape(-1, H) == ora["out"]).all()
    m = 200_000
    pre = ORACLE.seed_batch(bases[: m * L].cpu().numpy(), np.arange(m + 1, dtype=np.uint64) * L, seeds, h, want=(), threads=os.cpu_count() or 1)
    assert pre["n_emit"] == m * nk and pre["sum"] == (int(out[: m * nk].sum()) & M64)
    # the consumer's checksum over everything equals the stored rows'
    red = nthash_b200.seed_reduce_uniform(plan, bases, n, L)
    assert int(red[0]) == n * nk and (int(red[1]) & M64) == (int(out.sum()) & M64)


@pytest.mark.parametrize("shard", [0, 7])
def test_full_size_config5_long_read_shard(shard):
    """configs[4]: 100 k x 50 kb, k=63, h=1 sharded over 8 GPUs — one rank's 12.5 k reads (6.2e8 k-mers), exactly as
    bench.py cuts the splitmix stream (rank r hashes reads [12500 r, 12500 (r+1)))."""
    import nthash_b200
    cfg = bench.CONFIGS["c5"]
    n, L, k = cfg["n_reads"], cfg["read_len"], cfg["k"]
    nk = L - k + 1
    buf = bench.splitmix_bases_torch(torch, n * L, cfg["seed"], first_base=shard * n * L)
    bases = buf[: n * L]
    res = nthash_b200.kmer_hashes_uniform(bases, n, L, k, 1, want_strands=False)
    out = res.out.view(n, nk)
    torch.cuda.synchronize()
    assert int(res.valid_bits.view(torch.int32)[: (n * nk) // 32].ne(-1).sum()) == 0  # every window visited
    # strand symmetry over the whole shard
    rc = torch.zeros(n * L + 64, dtype=torch.uint8, device="cuda")
    rc[: n * L] = _revcomp(bases.view(n, L)).view(-1)
    out_r = nthash_b200.kmer_hashes_uniform(rc[: n * L], n, L, k, 1, want_valid=False).out.view(n, nk)
    assert torch.equal(out, out_r.flip(1))
    del out_r, rc
    # strands + fused consumer agree with the stored canonical hashes
    st = nthash_b200.kmer_hashes_uniform(bases, n, L, k, 1, want_valid=False, want_strands=True)
    assert torch.equal(st.fwd + st.rev, out.reshape(-1)) and torch.equal(st.out.view(-1), out.reshape(-1))
    del st
    red = nthash_b200.kmer_reduce_uniform(bases, n, L, k, 1)
    assert int(red[0]) == n * nk and (int(red[1]) & M64) == (int(out.sum()) & M64)
    # reads against the oracle bit for bit (the first, the last, and a stride in between), and a prefix checksum
    idx = torch.tensor(sorted({0, 1, n - 1, *range(0, n, 997)}), device="cuda")
    sample = bases.view(n, L)[idx].cpu().numpy()
    ora = ORACLE.kmer_batch(sample.reshape(-1), np.arange(len(idx) + 1, dtype=np.uint64) * L, k, 1, threads=8)
    assert (u64(out[idx]).reshape(-1, 1) == ora["out"]).all()
    m = 2000
    pre = ORACLE.kmer_batch(bases[: m * L].cpu().numpy(), np.arange(m + 1, dtype=np.uint64) * L, k, 1, want=(), threads=os.cpu_count() or 1)
    assert pre["n_emit"] == m * nk and pre["sum"] == (int(out[:m].sum()) & M64)


@pytest.mark.parametrize("lo,hi,seeds", [(100, 150, None), (36, 150, None), (100, 150, "c4")])
def test_full_size_ragged_batches(lo, hi, seeds):
    """The ragged batches bench.py reports (`configs.ragged`): 10 M trimmed-read shaped reads through the planned entry points
    (the layout's precomputed deal, direct sector stores; the generated seed kernel's direct form).  Sampled blocks of reads
    bit-exact against the oracle — the first, the last (a partial block of 256 items) and a few in between — plus a 64-bit
    checksum of a long prefix against the threaded oracle, and validity = "every window" for clean reads."""
    import nthash_b200
    n, k = (10_000_000, 31) if seeds is None else (5_000_000, 31)
    h = 1 if seeds is None else 3
    seed_list = bench.CONFIGS["c4"]["seeds"] if seeds else None
    H = h * (len(seed_list) if seed_list else 1)
    g = torch.Generator(device="cuda")
    g.manual_seed(lo * 1000 + hi)
    lens = torch.randint(lo, hi + 1, (n - 3,), device="cuda", generator=g, dtype=torch.int64)
    lens = torch.cat([lens, torch.tensor([k - 1, k, 0], device="cuda")])  # the last, partial block ends in reads without / with one window
    off = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    off[1:] = torch.cumsum(lens, 0)
    nb = int(off[-1])
    bases = bench.splitmix_bases_torch(torch, (nb + 31) // 32 * 32, 77)[:nb]
    plan = nthash_b200.RaggedPlan(off, k)
    if seed_list:
        res = nthash_b200.seed_hashes_planned(nthash_b200.SeedPlan(seed_list, h), plan, bases)
    else:
        res = nthash_b200.kmer_hashes_planned(plan, bases, h)
    torch.cuda.synchronize()
    koff = plan.koff()
    assert plan.rows == int(koff[-1]) == int(torch.clamp(lens - k + 1, min=0).sum())
    assert bool(res.valid_mask().all())  # ACGT only: the reference visits every window
    out = res.out
    off_h, koff_h = off.cpu().numpy(), koff.cpu().numpy()

    def oracle_rows(r0, r1):
        b = bases[off_h[r0]: off_h[r1]].cpu().numpy()
        o = (off_h[r0: r1 + 1] - off_h[r0]).astype(np.uint64)
        return (ORACLE.seed_batch(b, o, seed_list, h, threads=8) if seed_list else ORACLE.kmer_batch(b, o, k, h, threads=8))

    for r0 in (0, 255, 256 * 1234 + 17, n // 2, n - 700, n - 256):
        r1 = min(n, r0 + 600)
        ora = oracle_rows(r0, r1)
        got = u64(out[koff_h[r0]: koff_h[r1]]).reshape(-1, H)
        assert got.shape == ora["out"].shape and (got == ora["out"]).all(), f"reads {r0}..{r1}"
    m = 150_000  # whole-prefix checksum
    ora = oracle_rows(0, m)
    assert int(out[: koff_h[m]].sum()) & M64 == ora["sum"] and int(koff_h[m]) == ora["n_emit"]
