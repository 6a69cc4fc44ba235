"""GPU parity of the NtHash batch kernel (csrc/kmer_kernel.cu) against the CPU oracle.

Everything goes through the C ABI (include/nthash_b200.h) via the ctypes binding.  Bar: bit-exact
uint64 hashes, identical set of emitted positions (reference: NtHash::roll, src/kmer.cpp:246-264).
"""
import ctypes as C

import numpy as np
import pytest
import torch

from gpu_util import assert_batch_equal, ragged_offsets, synth, to_dev, u64
from oracle_lib import ORACLE

pytestmark = pytest.mark.gpu

import nthash_b200  # noqa: E402  (fails loudly if the CUDA library is missing)


def run_ragged(bases, off, k, h, strands=False):
    d_b, _keep = to_dev(bases)
    d_off = torch.from_numpy(off).cuda()
    res = nthash_b200.kmer_hashes(d_b, d_off, k, h, want_valid=True, want_strands=strands)
    torch.cuda.synchronize()
    return res


def test_golden_vectors_through_gpu():
    # reference tests/tests.cpp:47-69 (k-mer hash values) and :181-208 (skipping Ns)
    seq = np.frombuffer(b"ACATGCATGCA", np.uint8)
    res = run_ragged(seq, ragged_offsets([len(seq)]), 5, 3)
    out = u64(res.out).reshape(-1, 3)
    assert [int(x) for x in out[1]] == [0x38CC00F940AEBDAE, 0xAB7E1B110E086FC6, 0x11A1818BCFDD553]
    assert [int(x) for x in out[2]] == [0x603A48C5A11C794A, 0xE66016E61816B9C4, 0xC5B13CB146996FFE]
    s = bytearray(b"ACGTACACTGGACTGAGTCT"); s[10] = s[11] = ord("N")
    res = run_ragged(np.frombuffer(bytes(s), np.uint8), ragged_offsets([20]), 8, 3)
    assert list(np.flatnonzero(res.valid_mask().cpu().numpy())) == [0, 1, 2, 12]


def test_config1_single_1kb_sequence():
    # BASELINE.json configs[0]; SURVEY.md Appendix C known answers
    seq = ORACLE.gen_bases(1000, 42)
    res = run_ragged(seq, ragged_offsets([1000]), 31, 1, strands=True)
    out = u64(res.out)
    assert out.shape == (970, 1)
    assert int(out[0, 0]) == 0xB6A7A648205E25D6 and int(u64(res.fwd)[0]) == 0x54C31B64E55CF218
    assert int(out.sum(dtype=np.uint64)) == 0x429E8D1795548BB7
    assert_batch_equal(res, ORACLE.kmer_batch(seq, ragged_offsets([1000]).astype(np.uint64), 31, 1), 1, True)
    # the uniform entry point gives the same rows
    d_b, _k = to_dev(seq)
    resu = nthash_b200.kmer_hashes_uniform(d_b, 1, 1000, 31, 1)
    assert (u64(resu.out) == out).all()


@pytest.mark.parametrize("k", [3, 4, 5, 31, 32, 33, 63, 64, 65, 127, 255])
def test_ragged_dirty_reads_all_k(k):
    rng = np.random.default_rng(k)
    lens = rng.integers(0, 3 * k + 120, 700)
    lens[:6] = [0, 1, k - 1, k, k + 1, 2 * k]
    off = ragged_offsets(lens)
    bases = synth(rng, int(off[-1]), p_bad=0.004, lower=0.1)
    for h in (1, 2, 3, 4):
        res = run_ragged(bases, off, k, h, strands=(h == 3))
        ora = ORACLE.kmer_batch(bases, off.astype(np.uint64), k, h)
        assert_batch_equal(res, ora, h, check_strands=(h == 3))


@pytest.mark.parametrize("h", [1, 2, 3, 4])
def test_ragged_deal_precomputed_and_in_kernel(h, monkeypatch):
    """Ragged batches: the CTAs deal their items out by length class.  The deal comes with the layout (KmerGeom::item_perm,
    blocks of 256 items, computed by item_perm_kernel) or, for CTAs of another size or with NTHASH_B200_NO_ITEM_PERM, from
    the counting sort inside the kernel.  Same rows either way, and the oracle's; one batch of short reads (reads are
    items) and one with reads long enough to be cut into items."""
    k = 31
    rng = np.random.default_rng(900 + h)
    for lens in (rng.integers(0, 260, 1500), np.concatenate([rng.integers(20, 400, 300), [30000, 5, 12000, 31, 30]])):
        off = ragged_offsets(lens)
        bases = synth(rng, int(off[-1]), p_bad=0.003, lower=0.1)
        ora = ORACLE.kmer_batch(bases, off.astype(np.uint64), k, h, threads=4)
        res = run_ragged(bases, off, k, h)
        assert_batch_equal(res, ora, h)
        for env, val in (("NTHASH_B200_NO_ITEM_PERM", "1"), ("NTHASH_B200_FAST_NT", "128")):
            monkeypatch.setenv(env, val)
            alt = run_ragged(bases, off, k, h)
            monkeypatch.delenv(env)
            assert torch.equal(alt.out, res.out) and torch.equal(alt.valid_bits, res.valid_bits), env


def test_many_hashes_and_clean_reads():
    rng = np.random.default_rng(99)
    off = ragged_offsets(rng.integers(40, 200, 300))
    bases = synth(rng, int(off[-1]))
    for h in (7, 255):
        assert_batch_equal(run_ragged(bases, off, 31, h), ORACLE.kmer_batch(bases, off.astype(np.uint64), 31, h), h)


@pytest.mark.parametrize("read_len,k,h", [(150, 31, 1), (150, 31, 4), (151, 31, 1), (100, 21, 2), (36, 31, 1),
                                          (128, 31, 1), (250, 63, 3), (2000, 63, 1), (5003, 31, 1), (31, 31, 1)])
def test_uniform_batches(read_len, k, h):
    rng = np.random.default_rng(read_len * 7 + k)
    n = 3000 if read_len <= 300 else 150
    bases = synth(rng, n * read_len, p_bad=0.0005)
    d_b, _keep = to_dev(bases)
    res = nthash_b200.kmer_hashes_uniform(d_b, n, read_len, k, h, want_strands=(h == 1))
    torch.cuda.synchronize()
    off = (np.arange(n + 1, dtype=np.uint64) * read_len)
    assert_batch_equal(res, ORACLE.kmer_batch(bases, off, k, h, threads=8), h, check_strands=(h == 1))


FAST_SHAPES = [
    # (reads, read_len, k, h)           output path the launch picks (csrc/kmer_fast_kernel.cu)
    (3000, 150, 31, 1), (1000, 150, 31, 2), (1000, 150, 31, 4),   # 3-D tensor stores, one item per read
    (33, 150, 31, 1), (1, 150, 31, 1), (257, 102, 31, 1),         # partly filled warps / CTAs
    (2000, 149, 31, 1), (2000, 152, 31, 1), (700, 151, 31, 2),    # rows not 64-byte multiples: per-lane bulk stores + peel
    (300, 1000, 31, 1), (300, 1000, 31, 4), (200, 1230, 31, 2),   # cut-up reads, 4-D tensor stores (overlapping / exact last item)
    (301, 401, 31, 2), (60, 5003, 32, 1), (9, 50000, 63, 1),      # two items per read; 21 items; configs[4] shape
    (150, 1001, 31, 1), (64, 3001, 21, 1),                        # odd row length: cut-up reads through per-lane stores
    (1500, 150, 31, 3), (700, 151, 31, 3), (120, 1000, 31, 3),    # three hashes: 24-byte windows straddle the 16-byte chunks
    (500, 150, 31, 5), (300, 151, 31, 8), (60, 1000, 31, 7), (400, 149, 21, 6),   # 5..8 hashes: runtime count in the general output path
]


@pytest.mark.parametrize("n,read_len,k,h", FAST_SHAPES)
@pytest.mark.parametrize("no_box", [False, True])
def test_fast_kernel_output_paths(n, read_len, k, h, no_box, monkeypatch):
    """Uniform batches without strand outputs take kmer_fast_kernel; every output path must match the oracle,
    including rows the reference skips (non-ACGTU bytes) and rows written twice by overlapping items."""
    if no_box:
        monkeypatch.setenv("NTHASH_B200_FAST_NO_BOX", "1")
    rng = np.random.default_rng(n * 131 + read_len * 7 + k + h)
    bases = synth(rng, n * read_len, p_bad=0.0004, lower=0.05)
    bases[-1] = ord("N")                      # the very last window is one the reference skips
    if read_len > 400:
        bases[read_len - k - 3] = ord("n")    # inside the overlap of the last two items of read 0
    d_b, _keep = to_dev(bases)
    res = nthash_b200.kmer_hashes_uniform(d_b, n, read_len, k, h)
    torch.cuda.synchronize()
    off = np.arange(n + 1, dtype=np.uint64) * read_len
    assert_batch_equal(res, ORACLE.kmer_batch(bases, off, k, h, threads=8), h)
    for ws, nt, nbuf in ((1, 96, 2), (2, 160, 1)):   # the other launch knobs (sweeps use them)
        monkeypatch.setenv("NTHASH_B200_FAST_WS", str(ws))
        monkeypatch.setenv("NTHASH_B200_FAST_NT", str(nt))
        monkeypatch.setenv("NTHASH_B200_FAST_NBUF", str(nbuf))
        res2 = nthash_b200.kmer_hashes_uniform(d_b, n, read_len, k, h)
        torch.cuda.synchronize()
        assert torch.equal(res2.out, res.out) and torch.equal(res2.valid_bits, res.valid_bits)


@pytest.mark.parametrize("n,read_len,k,h", [(37, 5000, 63, 1), (100, 1200, 31, 2), (50, 3000, 21, 4), (64, 2113, 31, 3), (3, 700, 31, 1),
                                            (1, 100000, 63, 1), (1000, 600, 33, 1), (700, 1057, 5, 1)])
@pytest.mark.parametrize("flat_seg", [0, 168, 264, 96])
def test_flat_items_long_uniform_reads(n, read_len, k, h, flat_seg, monkeypatch):
    """Uniform long reads run as FLAT items (fixed runs of dense windows that ignore read boundaries, 3-D tensor stores)
    plus a fix-up launch for the rows behind every read boundary and the partial last item.  Dirty bytes are planted on
    both sides of read boundaries, where the two kernels' responsibilities meet."""
    if flat_seg:  # 0: the defaults (h <= 2: 200-window items, 40-window stores, CTAs of 96 threads)
        monkeypatch.setenv("NTHASH_B200_FLAT_SEG", str(flat_seg))
    rng = np.random.default_rng(n * 17 + read_len + k + h)
    bases = synth(rng, n * read_len, p_bad=0.0003, lower=0.05)
    for r in range(1, n, max(1, n // 7)):
        b = r * read_len
        for d in (-1, 0, k - 1, -k, k + 5)[: 2 + (r % 4)]:
            bases[b + d] = ord("N")
    d_b, _keep = to_dev(bases)
    res = nthash_b200.kmer_hashes_uniform(d_b, n, read_len, k, h)
    torch.cuda.synchronize()
    off = np.arange(n + 1, dtype=np.uint64) * read_len
    assert_batch_equal(res, ORACLE.kmer_batch(bases, off, k, h, threads=8), h)
    monkeypatch.setenv("NTHASH_B200_NO_FLAT", "1")
    res2 = nthash_b200.kmer_hashes_uniform(d_b, n, read_len, k, h)
    torch.cuda.synchronize()
    assert torch.equal(res2.out, res.out) and torch.equal(res2.valid_bits, res.valid_bits)


@pytest.mark.parametrize("n,read_len,k,h", [(3000, 150, 31, 1), (1000, 150, 31, 2), (1500, 150, 31, 3), (1000, 150, 31, 4), (33, 150, 31, 1),
                                            (1, 150, 31, 1), (257, 102, 31, 1), (900, 250, 63, 4), (37, 5000, 63, 1), (100, 1200, 31, 2),
                                            (1, 100000, 63, 1), (64, 2113, 31, 3), (700, 1057, 5, 1)])
def test_nibble_strip_kernel(n, read_len, k, h, monkeypatch):
    """kmer_pack_kernel (opt-in, NTHASH_B200_PACK=1): warp-private strips of nibble-packed bases, validity per 16-byte chunk,
    exact scrub from global memory.  Whole-read items and flat items, dirty bytes at read boundaries, every CTA size."""
    monkeypatch.setenv("NTHASH_B200_PACK", "1")
    rng = np.random.default_rng(n * 31 + read_len + k + h)
    bases = synth(rng, n * read_len, p_bad=0.0004, lower=0.05)
    bases[-1] = ord("N")
    bases[0] = ord("n") if n > 40 else bases[0]
    for r in range(1, n, max(1, n // 5)):
        bases[r * read_len - 1] = ord("N")
        bases[r * read_len + k - 1] = ord("R")
    d_b, _keep = to_dev(bases)
    off = np.arange(n + 1, dtype=np.uint64) * read_len
    ora = ORACLE.kmer_batch(bases, off, k, h, threads=8)
    for nt in (0, 64, 160):
        if nt:
            monkeypatch.setenv("NTHASH_B200_PACK_NT", str(nt))
        res = nthash_b200.kmer_hashes_uniform(d_b, n, read_len, k, h)
        torch.cuda.synchronize()
        assert_batch_equal(res, ora, h)


@pytest.mark.parametrize("n,read_len,k,h", [(40, 6000, 2000, 1), (6, 70000, 65535, 2), (300, 1300, 1023, 4)])
def test_huge_k_uniform(n, read_len, k, h):
    # k up to uint16 max (the reference's k is a uint16_t, nthash.hpp:74); the general kernel's CTA tile no longer
    # fits shared memory there, the fast kernel's smaller CTAs do
    rng = np.random.default_rng(k)
    bases = synth(rng, n * read_len, p_bad=0.00002)
    d_b, _keep = to_dev(bases)
    res = nthash_b200.kmer_hashes_uniform(d_b, n, read_len, k, h)
    torch.cuda.synchronize()
    off = np.arange(n + 1, dtype=np.uint64) * read_len
    assert_batch_equal(res, ORACLE.kmer_batch(bases, off, k, h, threads=8), h)


def test_huge_k_ragged():
    rng = np.random.default_rng(77)
    lens = [5000, 100, 999, 1000, 1001, 20000, 0, 3000]
    off = ragged_offsets(lens)
    bases = synth(rng, int(off[-1]), p_bad=0.0001)
    assert_batch_equal(run_ragged(bases, off, 1000, 2), ORACLE.kmer_batch(bases, off.astype(np.uint64), 1000, 2, threads=4), 2)


def test_ragged_long_reads_use_item_table():
    rng = np.random.default_rng(5)
    lens = [70000, 10, 30000, 62, 63, 64, 12345, 0, 999]
    off = ragged_offsets(lens)
    bases = synth(rng, int(off[-1]), p_bad=0.0002)
    for k, h in ((63, 1), (31, 2)):
        assert_batch_equal(run_ragged(bases, off, k, h), ORACLE.kmer_batch(bases, off.astype(np.uint64), k, h, threads=4), h)


def test_host_pointer_entry():
    rng = np.random.default_rng(8)
    off = ragged_offsets(rng.integers(0, 400, 500)).astype(np.uint64)
    bases = synth(rng, int(off[-1]), p_bad=0.002)
    k, h = 31, 2
    ora = ORACLE.kmer_batch(bases, off, k, h)
    rows = ora["out"].shape[0]
    out = np.full((rows, h), 0xAB, np.uint64); fw = np.zeros(rows, np.uint64); rv = np.zeros(rows, np.uint64)
    vb = np.zeros((rows + 31) // 32, np.uint32)
    rc = nthash_b200.LIB.nthash_kmer_batch(bases.ctypes.data, off.ctypes.data, len(off) - 1, k, h, out.ctypes.data,
                                           vb.ctypes.data, fw.ctypes.data, rv.ctypes.data, 0)
    assert rc == 0, nthash_b200.LIB.nthash_last_error()
    bits = ((vb[:, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(-1)[:rows]
    assert (bits == ora["valid"]).all() and (out == ora["out"]).all() and (fw == ora["fwd"]).all() and (rv == ora["rev"]).all()


def test_strand_symmetry_and_rna_at_scale():
    # size-independent properties (reference tests.cpp:119-133, :210-226) on a batch the oracle would take long on
    rng = np.random.default_rng(3)
    n, L, k = 200_000, 150, 31
    fwd = synth(rng, n * L).reshape(n, L)
    comp = np.zeros(256, np.uint8); comp[list(b"ACGT")] = list(b"TGCA")
    rc = comp[fwd[:, ::-1]]
    rna = fwd.copy(); rna[rna == ord("T")] = ord("U")
    outs = []
    for arr in (fwd, rc, rna):
        d_b, _keep = to_dev(np.ascontiguousarray(arr).reshape(-1))
        outs.append(nthash_b200.kmer_hashes_uniform(d_b, n, L, k, 1).out.view(n, L - k + 1))
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1].flip(1))   # canonical: window p of a read == window nk-1-p of its revcomp
    assert torch.equal(outs[0], outs[2])           # U == T
    # checksum of checksums against the threaded oracle on the same 30 M bases
    ora = ORACLE.kmer_batch(fwd.reshape(-1), np.arange(n + 1, dtype=np.uint64) * L, k, 1, want=(), threads=8)
    got = int(u64(outs[0]).sum(dtype=np.uint64))
    assert got == ora["sum"] and ora["n_emit"] == n * (L - k + 1)


def test_host_pipeline_many_chunks(monkeypatch):
    # the host-buffer entry streams the batch in chunks over three CUDA streams; force tiny chunks
    monkeypatch.setenv("NTHASH_B200_HOST_CHUNK_VALUES", "3000")
    rng = np.random.default_rng(21)
    for lens in (rng.integers(0, 300, 400), np.full(300, 150)):          # ragged, then fixed-length (fast kernel per chunk)
        off = ragged_offsets(lens).astype(np.uint64)
        bases = synth(rng, int(off[-1]), p_bad=0.003)
        for k, h in ((31, 1), (21, 4)):
            ora = ORACLE.kmer_batch(bases, off, k, h)
            rows = ora["out"].shape[0]
            out = np.full((rows, h), 0xCD, np.uint64); fw = np.zeros(rows, np.uint64); rv = np.zeros(rows, np.uint64)
            vb = np.zeros((rows + 31) // 32, np.uint32)
            rc = nthash_b200.LIB.nthash_kmer_batch(bases.ctypes.data, off.ctypes.data, len(off) - 1, k, h, out.ctypes.data,
                                                   vb.ctypes.data, fw.ctypes.data, rv.ctypes.data, 0)
            assert rc == 0, nthash_b200.LIB.nthash_last_error()
            bits = ((vb[:, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(-1)[:rows]
            assert (bits == ora["valid"]).all() and (out == ora["out"]).all() and (fw == ora["fwd"]).all() and (rv == ora["rev"]).all()
            out2 = np.zeros((rows, h), np.uint64)   # without the optional outputs
            assert nthash_b200.LIB.nthash_kmer_batch(bases.ctypes.data, off.ctypes.data, len(off) - 1, k, h, out2.ctypes.data, None, None, None, 0) == 0
            assert (out2 == ora["out"]).all()


def test_fused_reduce_consumer():
    # count / sum / xor of every visited window's hashes without materialising them (reference examples/benchmark.cpp:34-39)
    rng = np.random.default_rng(31)
    for (n, L, k, h, p_bad) in ((5000, 150, 31, 1, 0.001), (3000, 150, 31, 4, 0.0), (2000, 151, 21, 3, 0.002), (40, 9000, 63, 2, 0.0005)):
        bases = synth(rng, n * L, p_bad=p_bad)
        d_b, _keep = to_dev(bases)
        got = u64(nthash_b200.kmer_reduce_uniform(d_b, n, L, k, h))
        ora = ORACLE.kmer_batch(bases, np.arange(n + 1, dtype=np.uint64) * L, k, h, want=(), threads=8)
        assert (int(got[0]), int(got[1]), int(got[2])) == (ora["n_emit"], ora["sum"], ora["xor"]), (n, L, k, h)
    lens = rng.integers(0, 400, 3000)
    off = ragged_offsets(lens)
    bases = synth(rng, int(off[-1]), p_bad=0.002)
    d_b, _keep = to_dev(bases)
    got = u64(nthash_b200.kmer_reduce(d_b, torch.from_numpy(off).cuda(), 31, 2))
    ora = ORACLE.kmer_batch(bases, off.astype(np.uint64), 31, 2, want=())
    assert (int(got[0]), int(got[1]), int(got[2])) == (ora["n_emit"], ora["sum"], ora["xor"])
    res = np.zeros(3, np.uint64)   # host entry: bases up, three words back
    offu = off.astype(np.uint64)
    assert nthash_b200.LIB.nthash_kmer_reduce(bases.ctypes.data, offu.ctypes.data, len(lens), 31, 2, res.ctypes.data, 0) == 0
    assert (int(res[0]), int(res[1]), int(res[2])) == (ora["n_emit"], ora["sum"], ora["xor"])


def _bloom_oracle(ora, h, bits):
    """numpy restatement of the Bloom consumer from the oracle's hashes: filter words + per-window positions."""
    vals = ora["out"][ora["valid"].astype(bool)]            # rows the reference visits, all h values
    pos = vals % np.uint64(bits)
    words = np.zeros((bits + 31) // 32, np.uint32)
    np.bitwise_or.at(words, (pos >> np.uint64(5)).astype(np.int64).reshape(-1),
                     (np.uint32(1) << (pos & np.uint64(31)).astype(np.uint32)).reshape(-1))
    return words, pos


@pytest.mark.parametrize("h,bits", [(1, 1 << 20), (3, 1 << 22), (5, 3_000_017), (4, 999_983)])
def test_fused_bloom_consumer(h, bits):
    # the caller the reference's header names (nthash.hpp:14-17): hashes() feeding a Bloom filter
    rng = np.random.default_rng(h * 1000 + bits % 97)
    n, L, k = 4000, 150, 31
    bases = synth(rng, n * L, p_bad=0.001)
    off = np.arange(n + 1, dtype=np.uint64) * L
    ora = ORACLE.kmer_batch(bases, off, k, h, threads=8)
    want_words, pos = _bloom_oracle(ora, h, bits)
    d_b, _keep = to_dev(bases)
    filt = nthash_b200.bloom_filter(bits)
    ins = u64(nthash_b200.kmer_bloom_uniform(d_b, n, L, k, h, filt, bits))
    assert int(ins[0]) == int(ora["n_emit"])
    assert (filt.cpu().numpy().view(np.uint32) == want_words).all(), "filter contents differ from the oracle's"
    # querying the same reads: every window hits; a second insert changes nothing and reports all windows as present
    q = u64(nthash_b200.kmer_bloom_uniform(d_b, n, L, k, h, filt, bits, query=True))
    assert (int(q[0]), int(q[1])) == (int(ora["n_emit"]), int(ora["n_emit"]))
    again = u64(nthash_b200.kmer_bloom_uniform(d_b, n, L, k, h, filt, bits))
    assert int(again[1]) == int(ora["n_emit"]) and (filt.cpu().numpy().view(np.uint32) == want_words).all()
    # other reads against that filter: the hit count is exact (false positives included)
    other = synth(rng, 1000 * L, p_bad=0.001)
    ora2 = ORACLE.kmer_batch(other, np.arange(1001, dtype=np.uint64) * L, k, h, threads=8)
    _, pos2 = _bloom_oracle(ora2, h, bits)
    isset = (want_words[(pos2 >> np.uint64(5)).astype(np.int64)] >> (pos2 & np.uint64(31)).astype(np.uint32)) & 1
    d_o, _keep2 = to_dev(other)
    q2 = u64(nthash_b200.kmer_bloom_uniform(d_o, 1000, L, k, h, filt, bits, query=True))
    assert (int(q2[0]), int(q2[1])) == (int(ora2["n_emit"]), int(isset.all(axis=1).sum()))
    # ragged reads into a fresh filter
    lens = rng.integers(0, 500, 800)
    roff = ragged_offsets(lens)
    rb = synth(rng, int(roff[-1]), p_bad=0.002)
    ora3 = ORACLE.kmer_batch(rb, roff.astype(np.uint64), k, h)
    words3, _ = _bloom_oracle(ora3, h, bits)
    d_r, _keep3 = to_dev(rb)
    f3 = nthash_b200.bloom_filter(bits)
    r3 = u64(nthash_b200.kmer_bloom(d_r, torch.from_numpy(roff).cuda(), k, h, f3, bits))
    assert int(r3[0]) == int(ora3["n_emit"]) and (f3.cpu().numpy().view(np.uint32) == words3).all()


def _pack2bit(ascii_bases):
    """ACGTN bytes -> (packed 2-bit array, invalid-base bitmap) in the layout of include/nthash_b200.h."""
    code = np.zeros(256, np.uint8); code[list(b"ACGT")] = [0, 1, 2, 3]
    n = len(ascii_bases)
    c = np.zeros((n + 3) // 4 * 4, np.uint8); c[:n] = code[ascii_bases]
    c = c.reshape(-1, 4)
    packed = (c[:, 0] | (c[:, 1] << 2) | (c[:, 2] << 4) | (c[:, 3] << 6)).astype(np.uint8)
    inv = np.zeros((n + 31) // 32 * 32, np.uint8); inv[:n] = ascii_bases == ord("N")
    bits = np.packbits(inv.reshape(-1, 8), axis=1, bitorder="little").reshape(-1).view(np.uint32)
    return np.ascontiguousarray(packed), np.ascontiguousarray(bits)


@pytest.mark.parametrize("n,read_len,k,h,p_n,first", [(3000, 150, 31, 1, 0.002, 0), (2000, 150, 31, 4, 0.0, 7), (1500, 150, 31, 3, 0.001, 31),
                                                       (1000, 100, 21, 2, 0.01, 64), (33, 150, 31, 1, 0.05, 3), (1, 150, 31, 1, 0.0, 1),
                                                       (700, 250, 63, 2, 0.001, 129)])
def test_packed_2bit_hashed_directly_on_the_device(n, read_len, k, h, p_n, first, monkeypatch):
    """Device-resident 2-bit packed reads (+ invalid-base bitmap) hashed by the nibble-strip kernel without an ASCII copy:
    same rows as the oracle on the equivalent ASCII batch, for every alignment of the first base in the packed stream."""
    rng = np.random.default_rng(n + read_len + h + first)
    bases = rng.choice(np.frombuffer(b"ACGT", np.uint8), n * read_len)
    if p_n:
        bases[rng.random(len(bases)) < p_n] = ord("N")
        bases[0] = ord("N"); bases[-1] = ord("N")
    lead = rng.choice(np.frombuffer(b"ACGTN", np.uint8), first)              # whatever precedes the batch in the stream
    packed, inv = _pack2bit(np.concatenate([lead, bases]))
    d_p = torch.zeros(len(packed) + 64, dtype=torch.uint8, device="cuda"); d_p[: len(packed)] = torch.from_numpy(packed)
    d_i = torch.zeros(len(inv) + 8, dtype=torch.int32, device="cuda"); d_i[: len(inv)] = torch.from_numpy(inv.view(np.int32))
    ora = ORACLE.kmer_batch(bases, np.arange(n + 1, dtype=np.uint64) * read_len, k, h, threads=4)
    res = nthash_b200.kmer_hashes_packed2bit_uniform(d_p, d_i if p_n else None, first, n, read_len, k, h)
    torch.cuda.synchronize()
    assert_batch_equal(res, ora, h)


def test_packed_2bit_direct_refuses_other_shapes():
    d_p = torch.zeros(1 << 16, dtype=torch.uint8, device="cuda")
    for n, L, k, h in ((10, 151, 31, 1), (4, 5000, 31, 1), (10, 150, 31, 5)):   # odd rows, long reads, five hashes
        with pytest.raises(nthash_b200.NtHashError):
            nthash_b200.kmer_hashes_packed2bit_uniform(d_p, None, 0, n, L, k, h)


@pytest.mark.parametrize("tiny_chunks", [False, True])
def test_packed_2bit_input(tiny_chunks, monkeypatch):
    # 2-bit packed bases + invalid-base bitmap through the host entries: same rows as the ASCII path / the oracle
    if tiny_chunks:
        monkeypatch.setenv("NTHASH_B200_HOST_CHUNK_VALUES", "2500")   # chunk boundaries at every alignment of the packed stream
    rng = np.random.default_rng(17)
    for lens, k, h in ((np.full(400, 150), 31, 1), (rng.integers(0, 300, 500), 21, 2), ([5000, 3, 2000, 31, 77], 31, 4)):
        off = ragged_offsets(lens).astype(np.uint64)
        bases = rng.choice(np.frombuffer(b"ACGT", np.uint8), int(off[-1]))
        bases[rng.random(len(bases)) < 0.002] = ord("N")
        packed, inv = _pack2bit(bases)
        ora = ORACLE.kmer_batch(bases, off, k, h)
        rows = ora["out"].shape[0]
        out = np.full((rows, h), 0xEE, np.uint64); vb = np.zeros((rows + 31) // 32, np.uint32)
        rc = nthash_b200.LIB.nthash_kmer_batch_packed2bit(packed.ctypes.data, inv.ctypes.data, off.ctypes.data, len(off) - 1, 0, k, h,
                                                          out.ctypes.data, vb.ctypes.data, 0)
        assert rc == 0, nthash_b200.LIB.nthash_last_error()
        bits = ((vb[:, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(-1)[:rows]
        assert (bits == ora["valid"]).all() and (out == ora["out"]).all()
        res = np.zeros(3, np.uint64)
        assert nthash_b200.LIB.nthash_kmer_reduce_packed2bit(packed.ctypes.data, inv.ctypes.data, off.ctypes.data, len(off) - 1, 0, k, h,
                                                             res.ctypes.data, 0) == 0
        assert (int(res[0]), int(res[1]), int(res[2])) == (ora["n_emit"], ora["sum"], ora["xor"])
        if len(set(np.diff(off.astype(np.int64)))) == 1:   # fixed-length batch: the entries without an offsets array
            L = int(off[1])
            out2 = np.zeros((rows, h), np.uint64); res2 = np.zeros(3, np.uint64); out3 = np.zeros((rows, h), np.uint64)
            assert nthash_b200.LIB.nthash_kmer_batch_packed2bit(packed.ctypes.data, inv.ctypes.data, None, len(off) - 1, L, k, h,
                                                                out2.ctypes.data, None, 0) == 0
            assert nthash_b200.LIB.nthash_kmer_reduce_packed2bit(packed.ctypes.data, inv.ctypes.data, None, len(off) - 1, L, k, h,
                                                                 res2.ctypes.data, 0) == 0
            assert nthash_b200.LIB.nthash_kmer_batch_uniform(bases.ctypes.data, len(off) - 1, L, k, h, out3.ctypes.data, None, None, None, 0) == 0
            assert (out2 == ora["out"]).all() and (out3 == ora["out"]).all() and (res2 == res).all()
    # without a bitmap every base is ACGT; and the device helper on an unaligned slice
    clean = rng.choice(np.frombuffer(b"ACGT", np.uint8), 10_001)
    packed, _ = _pack2bit(clean)
    offc = np.array([0, 10_001], np.uint64)
    orac = ORACLE.kmer_batch(clean, offc, 31, 1)
    res = np.zeros(3, np.uint64)
    assert nthash_b200.LIB.nthash_kmer_reduce_packed2bit(packed.ctypes.data, None, offc.ctypes.data, 1, 0, 31, 1, res.ctypes.data, 0) == 0
    assert (int(res[0]), int(res[1])) == (orac["n_emit"], orac["sum"])
    d_p = torch.from_numpy(packed).cuda()
    d_o = torch.zeros(10_016, dtype=torch.uint8, device="cuda")
    assert nthash_b200.LIB.nthash_unpack2bit_dev(d_p.data_ptr(), None, 37, 9000, d_o.data_ptr(), None) == 0
    torch.cuda.synchronize()
    assert (d_o[:9000].cpu().numpy() == clean[37:9037]).all()


def test_multi_device_host_entry():
    # nthash_kmer_batch_multi shards one host batch over the listed devices; with one GPU the list repeats it, which still
    # exercises the sharding, the per-shard pipelines in threads and the merge of the validity bitmaps at odd bit offsets
    rng = np.random.default_rng(23)
    ndev = torch.cuda.device_count()
    devs = np.array([i % ndev for i in range(3)], np.int32)
    for lens, k, h in ((rng.integers(0, 300, 1000), 31, 2), (np.full(777, 151), 21, 1), ([40, 5000, 7], 31, 1)):
        off = ragged_offsets(lens).astype(np.uint64)
        bases = synth(rng, int(off[-1]), p_bad=0.003)
        ora = ORACLE.kmer_batch(bases, off, k, h)
        rows = ora["out"].shape[0]
        out = np.full((rows, h), 0xAB, np.uint64); fw = np.zeros(rows, np.uint64); rv = np.zeros(rows, np.uint64)
        vb = np.full((rows + 31) // 32, 0xFFFFFFFF, np.uint32)
        rc = nthash_b200.LIB.nthash_kmer_batch_multi(bases.ctypes.data, off.ctypes.data, len(off) - 1, k, h, out.ctypes.data, vb.ctypes.data,
                                                     fw.ctypes.data, rv.ctypes.data, devs.ctypes.data, len(devs))
        assert rc == 0, nthash_b200.LIB.nthash_last_error()
        bits = ((vb[:, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(-1)[:rows]
        assert (bits == ora["valid"]).all() and (out == ora["out"]).all() and (fw == ora["fwd"]).all() and (rv == ora["rev"]).all()


def test_full_size_config2_properties():
    """BASELINE.json configs[1] at full size (10 M x 150 bp, k=31, h=1; 1.2e9 k-mers) through size-independent properties:
    strand symmetry of the canonical hash (reference tests.cpp:119-133), agreement of the three output paths' checksums
    (stored hashes, fused reduce consumer, 2-bit packed input), and a sampled comparison with the oracle."""
    n, L, k = 10_000_000, 150, 31
    nk = L - k + 1
    g = torch.Generator(device="cuda"); g.manual_seed(2024)
    codes = torch.randint(0, 4, (n, L), dtype=torch.uint8, device="cuda", generator=g)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")
    fwd = torch.zeros(n * L + 64, dtype=torch.uint8, device="cuda"); fwd[: n * L] = lut[codes.long()].view(-1)
    rc = torch.zeros(n * L + 64, dtype=torch.uint8, device="cuda"); rc[: n * L] = lut[(3 - codes.flip(1)).long()].view(-1)
    out_f = nthash_b200.kmer_hashes_uniform(fwd[: n * L], n, L, k, 1, want_valid=False).out.view(n, nk)
    out_r = nthash_b200.kmer_hashes_uniform(rc[: n * L], n, L, k, 1, want_valid=False).out.view(n, nk)
    torch.cuda.synchronize()
    assert torch.equal(out_f, out_r.flip(1))                      # window p of a read == window nk-1-p of its reverse complement
    del out_r, rc
    total = int(out_f.sum())                                      # wraps modulo 2^64 like the consumer's sum
    red = nthash_b200.kmer_reduce_uniform(fwd[: n * L], n, L, k, 1)
    assert int(red[0]) == n * nk and int(red[1]) == total
    # a sample of reads against the oracle, bit for bit
    idx = torch.arange(0, n, 9973, device="cuda")
    sample = fwd[: n * L].view(n, L)[idx].cpu().numpy()
    ora = ORACLE.kmer_batch(sample.reshape(-1), np.arange(len(idx) + 1, dtype=np.uint64) * L, k, 1, threads=8)
    assert (u64(out_f[idx]).reshape(-1, 1) == ora["out"]).all()
    # the packed-input host entry on the first million reads
    m = 1_000_000
    packed = (codes[:m].reshape(-1, 4).to(torch.int32) * torch.tensor([1, 4, 16, 64], device="cuda", dtype=torch.int32)).sum(1).to(torch.uint8).cpu().numpy()
    res = np.zeros(3, np.uint64)
    assert nthash_b200.LIB.nthash_kmer_reduce_packed2bit(packed.ctypes.data, None, None, m, L, k, 1, res.ctypes.data, 0) == 0
    assert int(res[0]) == m * nk and int(res[1]) == int(out_f[:m].sum()) & (2**64 - 1)


def test_compacted_output_is_the_reference_iteration_order():
    # nthash_compact_rows_dev keeps exactly what `while (h.roll())` yields, in order: per read, the reference iterator's
    # positions and hashes (ORACLE.kmer_read / seed_read restate NtHash / SeedNtHash as iterators)
    rng = np.random.default_rng(41)
    lens = rng.integers(0, 260, 300)
    off = ragged_offsets(lens)
    bases = synth(rng, int(off[-1]), p_bad=0.01, lower=0.1)
    k, h = 21, 2
    res = run_ragged(bases, off, k, h)
    comp, idx = nthash_b200.compact(res)
    koff = ORACLE.koff(off.astype(np.uint64), k)
    want_rows, want_h = [], []
    for r in range(len(lens)):
        it = ORACLE.kmer_read(bases[off[r]:off[r + 1]].tobytes(), k, h) if lens[r] >= k else None
        if it is not None and len(it[0]):
            want_rows.append(it[0].astype(np.uint64) + koff[r])
            want_h.append(it[1])
    want_rows = np.concatenate(want_rows); want_h = np.concatenate(want_h)
    assert (u64(idx) == want_rows).all() and (u64(comp).reshape(-1, h) == want_h).all()
    # SeedNtHash: its own visiting rule (jumps over invalid incoming bases, seed.cpp:524-530)
    seeds = ["110101011", "101111101"]
    plan = nthash_b200.SeedPlan(seeds, 2)
    d_b, _keep = to_dev(bases)
    sres = nthash_b200.seed_hashes(plan, d_b, torch.from_numpy(off).cuda())
    scomp, sidx = nthash_b200.compact(sres)
    koff9 = ORACLE.koff(off.astype(np.uint64), 9)
    rows_w, h_w = [], []
    for r in range(len(lens)):
        if lens[r] < 9:
            continue
        it = ORACLE.seed_read(bases[off[r]:off[r + 1]].tobytes(), seeds, 2)
        if it is not None and len(it[0]):
            rows_w.append(it[0].astype(np.uint64) + koff9[r]); h_w.append(it[1])
    assert (u64(sidx) == np.concatenate(rows_w)).all() and (u64(scomp).reshape(-1, 4) == np.concatenate(h_w)).all()
    # a large clean batch: nothing is dropped, order kept
    n, L = 100_000, 150
    clean = synth(rng, n * L)
    d_c, _k2 = to_dev(clean)
    big = nthash_b200.kmer_hashes_uniform(d_c, n, L, 31, 1)
    c2, i2 = nthash_b200.compact(big)
    assert c2.shape[0] == big.rows and torch.equal(c2, big.out) and torch.equal(i2, torch.arange(big.rows, device="cuda"))


@pytest.mark.parametrize("crlf,final_newline", [(False, True), (True, True), (False, False)])
def test_fastq_staging_on_the_device(crlf, final_newline):
    # FASTQ text -> bases/read_off on the GPU (nthash_fastq_extract_dev), then the ordinary ragged path
    rng = np.random.default_rng(3 + crlf)
    lens = rng.integers(1, 300, 2000)
    lens[:3] = [1, 30, 31]
    reads = [bytes(synth(rng, int(n), p_bad=0.01, lower=0.1)) for n in lens]
    eol = b"\r\n" if crlf else b"\n"
    rec = [b"@r%d some description" % i + eol + r + eol + b"+" + eol + bytes(rng.integers(33, 74, len(r), dtype=np.uint8)) + eol for i, r in enumerate(reads)]
    text = b"".join(rec)
    if not final_newline:
        text = text[: -len(eol)]
    d_text = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
    bases, read_off = nthash_b200.fastq_extract(d_text)
    torch.cuda.synchronize()
    want_off = ragged_offsets(lens)
    assert (read_off.cpu().numpy() == want_off).all()
    want_bases = np.frombuffer(b"".join(reads), np.uint8)
    assert (bases.cpu().numpy() == want_bases).all()
    res = nthash_b200.kmer_hashes(bases, read_off, 31, 2)
    assert_batch_equal(res, ORACLE.kmer_batch(want_bases, want_off.astype(np.uint64), 31, 2), 2)


def test_concurrent_host_calls_from_several_threads():
    # the ABI is thread-safe (per-thread error text, mutex-protected per-device caches): four host threads hash different
    # batches on the same GPU at once, plus a spaced-seed plan compiled concurrently
    import threading
    rng = np.random.default_rng(77)
    jobs = []
    for t in range(4):
        lens = rng.integers(0, 400, 600) if t % 2 else np.full(500, 150)
        off = ragged_offsets(lens).astype(np.uint64)
        bases = synth(rng, int(off[-1]), p_bad=0.002)
        k, h = (31, 1 + t)
        jobs.append((bases, off, k, h, ORACLE.kmer_batch(bases, off, k, h)))
    results, errors = [None] * 4, []

    def work(i):
        bases, off, k, h, ora = jobs[i]
        try:
            for _ in range(5):
                out = np.zeros_like(ora["out"])
                rc = nthash_b200.LIB.nthash_kmer_batch(bases.ctypes.data, off.ctypes.data, len(off) - 1, k, h, out.ctypes.data, None, None, None, 0)
                assert rc == 0, nthash_b200.LIB.nthash_last_error()
                assert (out == ora["out"]).all()
            if i == 0:
                nthash_b200.SeedPlan(["110011", "101101"], 2)
            results[i] = True
        except Exception as e:  # surfaced below
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors and all(results), errors
