"""Ragged batches: the plan handle (nthash_ragged_plan_*), the `_dev` contract "only enqueue, never synchronise" (checked
by capturing the calls into a CUDA graph, which fails on any host synchronisation), the batch shape that used to
overflow a CTA's staged tile (one long read among many reads shorter than k), and the DIRECT store form of the kernel."""
import ctypes as C

import numpy as np
import pytest
import torch

import nthash_b200
from gpu_util import assert_batch_equal, ragged_offsets, synth, to_dev, u64
from nthash_b200._lib import LIB, check
from oracle_lib import ORACLE

pytestmark = pytest.mark.gpu


def _batch(rng, lens, p_bad=0.001):
    off = ragged_offsets(lens)
    bases = synth(rng, int(off[-1]), p_bad=p_bad, lower=0.05)
    d_b, keep = to_dev(bases, pad=96)
    return bases, off, d_b, keep


@pytest.mark.parametrize("shape", ["short", "long", "mixed"])
@pytest.mark.parametrize("h,strands", [(1, False), (2, True), (7, False)])
def test_planned_calls_match_the_oracle_and_the_unplanned_path(shape, h, strands):
    rng = np.random.default_rng(len(shape) * 100 + h)
    lens = {"short": rng.integers(0, 151, 4000), "long": rng.integers(200, 9000, 120),
            "mixed": np.concatenate([rng.integers(0, 60, 900), rng.integers(1000, 5000, 30), rng.integers(20, 40, 500)])}[shape]
    k = 31
    bases, off, d_b, _keep = _batch(rng, lens)
    d_off = torch.from_numpy(off).cuda()
    plan = nthash_b200.RaggedPlan(d_off, k)
    ora = ORACLE.kmer_batch(bases, off.astype(np.uint64), k, h, threads=8)
    assert plan.rows == ora["out"].shape[0] and plan.max_read_len == int(lens.max())
    assert (plan.koff().cpu().numpy() == ORACLE.koff(off.astype(np.uint64), k).astype(np.int64)).all()
    res = nthash_b200.kmer_hashes_planned(plan, d_b, h, want_strands=strands)
    torch.cuda.synchronize()
    assert_batch_equal(res, ora, h, check_strands=strands)
    # the same plan serves a second batch with the same layout and other bases
    bases2 = synth(rng, len(bases), p_bad=0.002)
    d_b2, _k2 = to_dev(bases2, pad=96)
    res2 = nthash_b200.kmer_hashes_planned(plan, d_b2, h)
    torch.cuda.synchronize()
    assert_batch_equal(res2, ORACLE.kmer_batch(bases2, off.astype(np.uint64), k, h, threads=8), h)
    # fused consumer through the plan
    red = torch.empty(3, dtype=torch.int64, device="cuda")
    check(LIB.nthash_kmer_reduce_planned_dev(plan._h, d_b.data_ptr(), d_b.numel(), h, red.data_ptr(), None))
    r = u64(red)
    assert (int(r[0]), int(r[1]), int(r[2])) == (int(ora["n_emit"]), int(ora["sum"]), int(ora["xor"]))


def test_planned_seed_batch():
    rng = np.random.default_rng(77)
    seeds = ["1101101101101101011011011011011", "1010101010101010101010101010101"]
    for lens in (rng.integers(0, 151, 3000), rng.integers(100, 4000, 200)):
        bases, off, d_b, _keep = _batch(rng, lens)
        plan = nthash_b200.RaggedPlan(torch.from_numpy(off).cuda(), 31)
        sp = nthash_b200.SeedPlan(seeds, 2)
        for strands in (False, True):
            res = nthash_b200.seed_hashes_planned(sp, plan, d_b, want_strands=strands)
            torch.cuda.synchronize()
            assert_batch_equal(res, ORACLE.seed_batch(bases, off.astype(np.uint64), seeds, 2, threads=8), 4, check_strands=strands)


@pytest.mark.parametrize("planned", [True, False])
@pytest.mark.parametrize("shape", ["short", "long"])
def test_dev_entries_only_enqueue__cuda_graph_capture(planned, shape):
    """A stream capture fails on any host synchronisation, blocking allocation or read-back inside the captured calls:
    the ragged `_dev` entries (planned and unplanned, with a validity bitmap, one-item-per-read and cut-up layouts) must
    be capturable, and replaying the graph on other bases must give that batch's hashes."""
    rng = np.random.default_rng(5 + planned)
    lens = rng.integers(0, 151, 5000) if shape == "short" else rng.integers(300, 7000, 150)
    k, h = 31, 2
    bases, off, d_b, _keep = _batch(rng, lens)
    d_off = torch.from_numpy(off).cuda()
    plan = nthash_b200.RaggedPlan(d_off, k)
    koff = plan.koff()
    rows = plan.rows
    out = torch.zeros((rows, h), dtype=torch.int64, device="cuda")
    valid = torch.zeros(int(LIB.nthash_valid_words(rows)), dtype=torch.int32, device="cuda")
    st = torch.cuda.Stream()

    def call():
        if planned:
            check(LIB.nthash_kmer_batch_planned_dev(plan._h, d_b.data_ptr(), d_b.numel(), h, out.data_ptr(), valid.data_ptr(), None, None, C.c_void_p(st.cuda_stream)))
        else:
            check(LIB.nthash_kmer_batch_dev(d_b.data_ptr(), d_b.numel(), d_off.data_ptr(), koff.data_ptr(), len(lens), int(lens.max()), k, h,
                                            out.data_ptr(), valid.data_ptr(), None, None, C.c_void_p(st.cuda_stream)))

    with torch.cuda.stream(st):
        call()  # warm-up outside the capture (function attributes, module load)
    st.synchronize()
    out.zero_(); valid.zero_()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        call()
    # new bases, same layout: replay
    bases2 = synth(rng, len(bases), p_bad=0.002)
    d_b[:] = torch.from_numpy(bases2).cuda()
    g.replay()
    torch.cuda.synchronize()
    res = nthash_b200.HashBatch(out, valid, None, rows)
    assert_batch_equal(res, ORACLE.kmer_batch(bases2, off.astype(np.uint64), k, h, threads=8), h)


@pytest.mark.parametrize("seeded", [False, True])
def test_one_long_read_among_many_reads_shorter_than_k(seeded):
    """k=31: a 400-base read forces the cut-up layout; the repeating 9 x 30-base + 1 x 31-base reads behind it own almost
    no windows, and their bytes used to be missing from the bound of a CTA's staged span (a device trap)."""
    rng = np.random.default_rng(3)
    lens = np.concatenate([[400], np.tile([30] * 9 + [31], 700), [2000], np.tile([30] * 9 + [31], 300)])
    bases, off, d_b, _keep = _batch(rng, lens, p_bad=0.0005)
    d_off = torch.from_numpy(off).cuda()
    if seeded:
        seeds = ["1101101101101101011011011011011"]
        plan = nthash_b200.SeedPlan(seeds, 2)
        res = nthash_b200.seed_hashes(plan, d_b, d_off)
        torch.cuda.synchronize()
        assert_batch_equal(res, ORACLE.seed_batch(bases, off.astype(np.uint64), seeds, 2, threads=8), 2)
        return
    for h, strands in ((1, False), (3, True), (12, False)):
        res = nthash_b200.kmer_hashes(d_b, d_off, 31, h, want_strands=strands)
        torch.cuda.synchronize()
        assert_batch_equal(res, ORACLE.kmer_batch(bases, off.astype(np.uint64), 31, h, threads=8), h, check_strands=strands)
    red = nthash_b200.kmer_reduce(d_b, d_off, 31, 2)
    ora = ORACLE.kmer_batch(bases, off.astype(np.uint64), 31, 2, threads=8)
    r = u64(red)
    assert (int(r[0]), int(r[1]), int(r[2])) == (int(ora["n_emit"]), int(ora["sum"]), int(ora["xor"]))


@pytest.mark.parametrize("h", [1, 2, 3, 4])
def test_direct_store_form(h, monkeypatch):
    """NTHASH_B200_FAST_DIRECT=1: the general output path without shared-memory rows (32-byte stores from registers)."""
    monkeypatch.setenv("NTHASH_B200_FAST_DIRECT", "1")
    rng = np.random.default_rng(40 + h)
    for lens in (rng.integers(0, 151, 3000), rng.integers(200, 3000, 100), np.full(500, 151)):
        bases, off, d_b, _keep = _batch(rng, lens)
        res = nthash_b200.kmer_hashes(d_b, torch.from_numpy(off).cuda(), 31, h, want_strands=(h <= 2))
        torch.cuda.synchronize()
        assert_batch_equal(res, ORACLE.kmer_batch(bases, off.astype(np.uint64), 31, h, threads=8), h, check_strands=(h <= 2))
    n, L = 700, 151  # uniform batch whose rows are not 64-byte multiples
    bases = synth(rng, n * L, p_bad=0.001)
    d_b, _keep = to_dev(bases)
    res = nthash_b200.kmer_hashes_uniform(d_b, n, L, 31, h, want_strands=True)
    torch.cuda.synchronize()
    assert_batch_equal(res, ORACLE.kmer_batch(bases, np.arange(n + 1, dtype=np.uint64) * L, 31, h, threads=8), h, check_strands=True)
