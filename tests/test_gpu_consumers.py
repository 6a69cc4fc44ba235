"""The consumers added in round 2 next to the hash path: minimizer selection and the ntCard-style
cardinality sketch, each against a numpy restatement on top of the oracle's hash rows (tests/oracle_lib.py)."""
import ctypes as C

import numpy as np
import pytest
import torch

import nthash_b200
from gpu_util import ragged_offsets, synth, to_dev, u64
from nthash_b200._lib import LIB, check
from oracle_lib import ORACLE, minimizer_bits, sketch_counts

pytestmark = pytest.mark.gpu


def _bits(words, rows):
    w = words.cpu().numpy().view(np.uint32)
    return ((w[:, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(-1)[:rows].astype(bool)


@pytest.mark.parametrize("n,L,k,w", [(3000, 150, 31, 10), (500, 150, 21, 1), (700, 151, 31, 19), (64, 2000, 15, 64), (33, 60, 31, 31), (5, 40, 31, 12)])
def test_minimizers_uniform_device_entry(n, L, k, w, monkeypatch):
    rng = np.random.default_rng(n + L + w)
    bases = synth(rng, n * L, p_bad=0.003, lower=0.05)
    bases[: 3 * L] = ord("A")  # identical k-mers: ties must go to the leftmost
    d_b, _keep = to_dev(bases)
    off = np.arange(n + 1, dtype=np.uint64) * L
    ora = ORACLE.kmer_batch(bases, off, k, 1, threads=8)
    want = minimizer_bits(ora["out"][:, 0], ora["valid"], ORACLE.koff(off, k), w)
    for chunk in (None, "64"):  # default chunking, and many small chunks
        if chunk:
            monkeypatch.setenv("NTHASH_B200_MINIMIZER_CHUNK_READS", chunk)
        bits, mh, mr, cnt = nthash_b200.kmer_minimizers_uniform(d_b, n, L, k, w)
        got = _bits(bits, len(want))
        assert (got == want).all(), np.argwhere(got != want)[:8].tolist()
        assert cnt == int(want.sum())
        rows = np.flatnonzero(want)
        assert (u64(mr).astype(np.int64) == rows).all() and (u64(mh) == ora["out"][rows, 0]).all()
    bits2, _, _, cnt2 = nthash_b200.kmer_minimizers_uniform(d_b, n, L, k, w, want_lists=False)
    assert cnt2 == cnt and torch.equal(bits2, bits)
    # a capacity smaller than the count: the count is still exact, the lists hold the first `capacity` entries
    if cnt > 10:
        _, mh3, mr3, cnt3 = nthash_b200.kmer_minimizers_uniform(d_b, n, L, k, w, capacity=10)
        assert cnt3 == cnt and (u64(mr3).astype(np.int64) == np.flatnonzero(want)[:10]).all()


@pytest.mark.parametrize("ragged", [False, True])
def test_minimizers_host_entry(ragged):
    rng = np.random.default_rng(9 + ragged)
    k, w = 31, 10
    lens = rng.integers(0, 400, 6000) if ragged else np.full(4000, 150)
    off = ragged_offsets(lens)
    bases = synth(rng, int(off[-1]), p_bad=0.002)
    ora = ORACLE.kmer_batch(bases, off.astype(np.uint64), k, 1, threads=8)
    want = minimizer_bits(ora["out"][:, 0], ora["valid"], ORACLE.koff(off.astype(np.uint64), k), w)
    rows = len(want)
    bits = np.zeros((rows + 31) // 32, np.uint32)
    cap = rows
    mh = np.zeros(cap, np.uint64); mr = np.zeros(cap, np.uint64)
    cnt = C.c_uint64(0)
    off64 = off.astype(np.uint64)
    check(LIB.nthash_kmer_minimizers(bases.ctypes.data, off64.ctypes.data if ragged else None, len(lens), 0 if ragged else 150, k, w,
                                     bits.ctypes.data, mh.ctypes.data, mr.ctypes.data, cap, C.byref(cnt), 0))
    got = ((bits[:, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(-1)[:rows].astype(bool)
    assert (got == want).all() and cnt.value == int(want.sum())
    sel = np.flatnonzero(want)
    assert (mr[: cnt.value].astype(np.int64) == sel).all() and (mh[: cnt.value] == ora["out"][sel, 0]).all()


@pytest.mark.parametrize("ragged", [False, True])
def test_cardinality_sketch(ragged):
    rng = np.random.default_rng(21 + ragged)
    k, s, r = 25, 4, 12   # a 1/16 sample so that a small batch fills the table
    lens = rng.integers(0, 300, 5000) if ragged else np.full(6000, 100)
    off = ragged_offsets(lens)
    bases = synth(rng, int(off[-1]), p_bad=0.002, lower=0.1)
    bases[: 2000] = bases[2000: 4000]  # repeated sequence: multiplicities above one
    ora = ORACLE.kmer_batch(bases, off.astype(np.uint64), k, 1, threads=8)
    want, n_s = sketch_counts(ora["out"][:, 0], ora["valid"], s, r)
    d_b, _keep = to_dev(bases)
    if ragged:
        counters, res = nthash_b200.kmer_sketch(d_b, torch.from_numpy(off).cuda(), k, s, r)
    else:
        counters, res = nthash_b200.kmer_sketch_uniform(d_b, len(lens), 100, k, s, r)
    torch.cuda.synchronize()
    assert (counters.cpu().numpy().view(np.uint32) == want).all()
    assert int(res[0]) == int(ora["n_emit"]) and int(res[1]) == n_s
    # accumulation over batches: the same batch again doubles every counter
    if not ragged:
        counters, res = nthash_b200.kmer_sketch_uniform(d_b, len(lens), 100, k, s, r, counters=counters)
        assert (counters.cpu().numpy().view(np.uint32) == 2 * want).all()


def test_consumer_argument_checks():
    cnt = np.zeros(3, np.uint64)
    assert LIB.nthash_kmer_sketch_uniform_dev(None, 0, 1, 150, 31, 0, 20, cnt.ctypes.data, cnt.ctypes.data, None) == -1
    assert LIB.nthash_kmer_sketch_uniform_dev(None, 0, 1, 150, 31, 40, 30, cnt.ctypes.data, cnt.ctypes.data, None) == -1
    assert LIB.nthash_kmer_minimizer_uniform_dev(None, 0, 1, 150, 31, 65, cnt.ctypes.data, None, None, 0, cnt.ctypes.data, None) == -1
    assert LIB.nthash_kmer_minimizers(None, None, 1, 150, 31, 0, cnt.ctypes.data, None, None, 0, cnt.ctypes.data, 0) == -1
