"""GPU parity of the SeedNtHash batch kernels (csrc/seed_kernel.cu) and the BlindNtHash batch kernel
(csrc/blind_kernel.cu) against the CPU oracle, through the C ABI.

Bar: bit-exact hashes and the reference's exact set of visited windows, including its quirks
(non-ACGTU bytes hashed inside a window, k-jump on an invalid incoming base, NUL rule:
reference src/seed.cpp:151, :493-544)."""
import ctypes as C

import numpy as np
import pytest
import torch

from gpu_util import assert_batch_equal, ragged_offsets, synth, to_dev, u64
from oracle_lib import ORACLE

pytestmark = pytest.mark.gpu

import nthash_b200  # noqa: E402

SEED_A31 = "1010101010101010101010101010101"
SEED_B31 = "1101101101101101011011011011011"

SEED_SETS = [
    ([SEED_A31, SEED_B31], 3),
    (["11100111"], 3),
    (["110011", "101101"], 2),
    (["11111"], 1),
    (["111110000000011111", "111111100001111111"], 2),
    (["111111111101111111111", "110111010010010111011"], 4),
    (["1101100", "0011011", "1000001"], 1),      # asymmetric
    (["0110", "1001", "0101"], 2),               # leading / trailing don't-cares
    (["1" * 40 + "0" * 23 + "1" * 40], 1),        # long runs, k = 103
    (["11011000001100101101011000011010110100110000011011", "01010000101001110100111011011100101110010100001010",
      "11100000100111010111000100100011101011100100000111", "01111000011000111101000011000010111100011000011110",
      "00111000011000111101000011000010111100011000011100", "00000000000000000000000011000000000000000000000000",
      "11111111111111111111111100111111111111111111111111", "11111111111111111111111111111111111111111111111111"], 4),
]


def run_ragged(plan, bases, off, strands=True):
    d_b, _keep = to_dev(bases)
    res = nthash_b200.seed_hashes(plan, d_b, torch.from_numpy(off).cuda(), want_strands=strands)
    torch.cuda.synchronize()
    return res


def test_golden_spaced_seed_values_through_gpu():
    # reference tests/tests.cpp:228-248
    plan = nthash_b200.SeedPlan(["11100111"], 3)
    seq = np.frombuffer(b"ACATGCATGCA", np.uint8)
    out = u64(run_ragged(plan, seq, ragged_offsets([11])).out).reshape(-1, 3)
    want = [[0x10BE4904AD8DE5D, 0x3E29E4F4C991628C, 0x3F35C984B13FEB20], [0x8200A7AA3EAF17C8, 0x344198402F4C2A9C, 0xB6423FE62E69C40C],
            [0x3CE8ADCBEAA56532, 0x162E91A4DBEDBF11, 0x53173F786A031F45]]
    assert [[int(x) for x in r] for r in out[:3]] == want
    assert plan.symmetric and not nthash_b200.SeedPlan(["1101100"], 1).symmetric


def test_config_seeds_known_answers():
    # seeds A,B on the 1 kb sequence, and the 200-base dirty read (109 visited windows)
    plan = nthash_b200.SeedPlan([SEED_A31, SEED_B31], 3)
    seq = ORACLE.gen_bases(1000, 42)
    out = u64(run_ragged(plan, seq, ragged_offsets([1000])).out)
    assert [int(x) for x in out[0]] == [0xE09BC50CEA32E5AA, 0xF206E6A6EBAE2333, 0x5033979BEAA8B115,
                                       0x194F0B34C1432F4C, 0x74181E1C17687846, 0x282AFC7155FDA2BA]
    assert int(out.sum(dtype=np.uint64)) == 0x6C20E8349FC9CF61
    dirty = seq[:200].copy(); dirty[50] = dirty[51] = ord("N"); dirty[120] = ord("n"); dirty[199] = ord("N")
    plan1 = nthash_b200.SeedPlan([SEED_A31, SEED_B31], 1)
    res = run_ragged(plan1, dirty, ragged_offsets([200]))
    assert int(res.valid_mask().sum()) == 109 and int(u64(res.out).sum(dtype=np.uint64)) == 0x69A56848390E3347


@pytest.mark.parametrize("seeds,h", SEED_SETS, ids=lambda v: v[0][:10] if isinstance(v, list) else str(v))
def test_ragged_dirty_reads_all_seed_sets(seeds, h, capfd):
    k = len(seeds[0])
    rng = np.random.default_rng(len(seeds) * 1000 + k)
    lens = rng.integers(0, 3 * k + 100, 400)
    lens[:6] = [0, 1, k - 1, k, k + 1, 2 * k]
    off = ragged_offsets(lens)
    bases = synth(rng, int(off[-1]), p_bad=0.006, lower=0.1)
    # a cluster of invalid bytes closer than k, and NUL bytes: the jump rule and init's NUL rule
    for r in (10, 20, 30):
        s, e = int(off[r]), int(off[r + 1])
        if e - s > 2 * k:
            bases[s + k:s + k + 3] = [ord("N"), ord("A"), ord("R")]
    bases[rng.integers(0, len(bases), 6)] = 0
    plan = nthash_b200.SeedPlan(seeds, h)
    res = run_ragged(plan, bases, off)
    ora = ORACLE.seed_batch(bases, off.astype(np.uint64), seeds, h)
    assert_batch_equal(res, ora, len(seeds) * h, check_strands=True)
    capfd.readouterr()


@pytest.mark.parametrize("jit", [True, False])
def test_raw_control_bytes_hash_like_the_reference(jit, monkeypatch):
    """Raw bytes 1, 3, 4, 5, 7 are SEED_TAB's complement slots (src/internal.hpp:133): SeedNtHash hashes them with those
    seeds and never jumps on them.  Uniform (specialised kernel, with and without strands) and ragged batches."""
    if not jit:
        monkeypatch.setenv("NTHASH_B200_DISABLE_SEED_JIT", "1")
    rng = np.random.default_rng(99)
    seeds = [SEED_A31, SEED_B31]
    plan = nthash_b200.SeedPlan(seeds, 2)
    n, L = 600, 160
    bases = synth(rng, n * L, p_bad=0.001)
    at = rng.integers(0, n * L, 400)
    bases[at] = rng.choice(np.array([1, 3, 4, 5, 7, 2, 6], np.uint8), 400)
    d_b, _keep = to_dev(bases)
    off = np.arange(n + 1, dtype=np.uint64) * L
    ora = ORACLE.seed_batch(bases, off, seeds, 2, threads=4)
    for strands in (True, False):
        res = nthash_b200.seed_hashes_uniform(plan, d_b, n, L, want_strands=strands)
        torch.cuda.synchronize()
        assert_batch_equal(res, ora, 4, check_strands=strands)
    lens = rng.integers(20, 300, 500)
    roff = ragged_offsets(lens)
    rb = bases[: int(roff[-1])]
    res = run_ragged(plan, rb, roff, strands=False)
    assert_batch_equal(res, ORACLE.seed_batch(rb, roff.astype(np.uint64), seeds, 2, threads=4), 4)


@pytest.mark.parametrize("read_len,n,p_bad", [(150, 3000, 0.0), (150, 3000, 0.002), (151, 500, 0.001), (40, 1000, 0.0), (3000, 60, 0.0005)])
def test_uniform_config4_shape(read_len, n, p_bad):
    rng = np.random.default_rng(read_len + n)
    bases = synth(rng, n * read_len, p_bad=p_bad)
    d_b, _keep = to_dev(bases)
    plan = nthash_b200.SeedPlan([SEED_A31, SEED_B31], 3)
    res = nthash_b200.seed_hashes_uniform(plan, d_b, n, read_len, want_strands=True)
    torch.cuda.synchronize()
    ora = ORACLE.seed_batch(bases, np.arange(n + 1, dtype=np.uint64) * read_len, [SEED_A31, SEED_B31], 3, threads=8)
    assert_batch_equal(res, ora, 6, check_strands=True)


@pytest.mark.parametrize("seeds,h", SEED_SETS, ids=lambda v: v[0][:10] if isinstance(v, list) else str(v))
def test_ragged_specialised_kernel(seeds, h, capfd, monkeypatch):
    # ragged batches without strand outputs take the ragged variant of the NVRTC-specialised kernel (per-lane rows,
    # coalesced stores); same rows as the oracle and as the generic interpreter, dirty bytes and NULs included
    k = len(seeds[0])
    rng = np.random.default_rng(len(seeds) * 77 + k)
    for lens in (rng.integers(0, 3 * k + 100, 500), [40000, 3, k, 9000, k + 1, 0, 777]):
        lens = np.asarray(lens)
        off = ragged_offsets(lens)
        bases = synth(rng, int(off[-1]), p_bad=0.004, lower=0.1)
        bases[rng.integers(0, len(bases), 4)] = 0
        plan = nthash_b200.SeedPlan(seeds, h)
        res = run_ragged(plan, bases, off, strands=False)
        ora = ORACLE.seed_batch(bases, off.astype(np.uint64), seeds, h, threads=4)
        assert_batch_equal(res, ora, len(seeds) * h)
        monkeypatch.setenv("NTHASH_B200_DISABLE_SEED_JIT", "1")
        slow = run_ragged(plan, bases, off, strands=False)
        monkeypatch.delenv("NTHASH_B200_DISABLE_SEED_JIT")
        assert torch.equal(res.out, slow.out) and torch.equal(res.valid_bits, slow.valid_bits)
    capfd.readouterr()


@pytest.mark.parametrize("seeds,h", [([SEED_A31, SEED_B31], 3), (["11100111"], 3), (["110011", "101101"], 2), (["11111"], 1),
                                     (["111111111101111111111", "110111010010010111011"], 4), (["1101100", "0011011", "1000001"], 1)],
                         ids=lambda v: v[0][:10] if isinstance(v, list) else str(v))
def test_ragged_direct_and_row_forms_agree(seeds, h, monkeypatch):
    """The ragged variant of the generated kernel has two output forms: whole-sector stores straight from registers (groups of
    1, 2 or 4 windows, lanes started on a sector-aligned row) and the older private rows in shared memory
    (NTHASH_B200_SEED_JIT_RAGGED_ROWS=1, also what hash counts that do not fit the registers take).  Rows of every length and
    alignment, dirty bytes and NULs: same values, same validity, and the oracle's."""
    k = len(seeds[0])
    rng = np.random.default_rng(4242 + k + h)
    lens = np.concatenate([rng.integers(0, 3 * k + 150, 900), [k, k + 1, k + 2, k + 3, 0, 5000]])
    off = ragged_offsets(lens)
    bases = synth(rng, int(off[-1]), p_bad=0.003, lower=0.1)
    bases[rng.integers(0, len(bases), 5)] = 0
    ora = ORACLE.seed_batch(bases, off.astype(np.uint64), seeds, h, threads=4)
    direct = run_ragged(nthash_b200.SeedPlan(seeds, h), bases, off, strands=False)
    assert_batch_equal(direct, ora, len(seeds) * h)
    monkeypatch.setenv("NTHASH_B200_SEED_JIT_RAGGED_ROWS", "1")  # read when a plan's ragged variant is generated: a new plan
    rows = run_ragged(nthash_b200.SeedPlan(seeds, h), bases, off, strands=False)
    assert torch.equal(rows.out, direct.out) and torch.equal(rows.valid_bits, direct.valid_bits)


def _long_seed(k, rng):
    half = "".join(rng.choice(list("0111"), k // 2))
    s = half + ("1" if k % 2 else "") + half[::-1]
    return "1" + s[1:-1] + "1"


@pytest.mark.parametrize("case", list(range(len(SEED_SETS))) + ["k200", "k255x2"])
def test_uniform_strand_outputs_specialised_kernel(case, monkeypatch):
    """get_forward_hash() / get_reverse_hash() arrays (nthash.hpp:489-500) from the NVRTC-specialised kernel (what the C++
    SeedNtHash shim asks for on every call), and seeds up to k = 256 there: equal to the oracle and to the interpreter,
    on rows that are and are not 32-byte multiples, with dirty bytes and NULs."""
    rng = np.random.default_rng(1000 + len(str(case)))
    if case == "k200":
        seeds, h = [_long_seed(200, rng)], 2
    elif case == "k255x2":
        seeds, h = [_long_seed(255, rng), "1" * 255], 1
    else:
        seeds, h = SEED_SETS[case]
    k, m = len(seeds[0]), len(seeds)
    plan = nthash_b200.SeedPlan(seeds, h)
    for n, L in ((600, k + 119), (301, k + 120), (23, 3 * k + 1501)):
        bases = synth(rng, n * L, p_bad=0.002, lower=0.1)
        bases[rng.integers(0, len(bases), 3)] = 0
        d_b, _keep = to_dev(bases)
        res = nthash_b200.seed_hashes_uniform(plan, d_b, n, L, want_strands=True)
        torch.cuda.synchronize()
        ora = ORACLE.seed_batch(bases, np.arange(n + 1, dtype=np.uint64) * L, seeds, h, threads=8)
        assert_batch_equal(res, ora, m * h, check_strands=True)
        monkeypatch.setenv("NTHASH_B200_SEED_JIT_NO_STRANDS", "1")
        slow = nthash_b200.seed_hashes_uniform(plan, d_b, n, L, want_strands=True)
        monkeypatch.delenv("NTHASH_B200_SEED_JIT_NO_STRANDS")
        torch.cuda.synchronize()
        assert torch.equal(res.out, slow.out) and torch.equal(res.fwd, slow.fwd) and torch.equal(res.rev, slow.rev) and torch.equal(res.valid_bits, slow.valid_bits)


def test_generic_and_specialised_kernels_agree(monkeypatch):
    # uniform batches take the NVRTC-specialised kernel; the generic interpreter must give the same rows
    rng = np.random.default_rng(12)
    n, L = 2000, 150
    bases = synth(rng, n * L, p_bad=0.001)
    d_b, _keep = to_dev(bases)
    for seeds, h in ((["1010101010101010101010101010101", "1101101101101101011011011011011"], 3), (["110011", "101101"], 2),
                     (["1" * 40 + "0" * 23 + "1" * 40], 1), (["1101100", "0011011", "1000001"], 1)):
        plan = nthash_b200.SeedPlan(seeds, h)
        assert b"NVRTC" in nthash_b200.LIB.nthash_seed_plan_kernel_note(plan._h), nthash_b200.LIB.nthash_seed_plan_kernel_note(plan._h)
        fast = nthash_b200.seed_hashes_uniform(plan, d_b, n, L)
        monkeypatch.setenv("NTHASH_B200_DISABLE_SEED_JIT", "1")
        slow = nthash_b200.seed_hashes_uniform(plan, d_b, n, L)
        monkeypatch.delenv("NTHASH_B200_DISABLE_SEED_JIT")
        torch.cuda.synchronize()
        assert torch.equal(fast.out, slow.out) and torch.equal(fast.valid_bits, slow.valid_bits)
        ora = ORACLE.seed_batch(bases, np.arange(n + 1, dtype=np.uint64) * L, seeds, h, threads=8)
        assert_batch_equal(fast, ora, len(seeds) * h)


def test_ragged_long_reads_and_host_entry():
    rng = np.random.default_rng(77)
    lens = [40000, 5, 9000, 31, 30, 777]
    off = ragged_offsets(lens)
    bases = synth(rng, int(off[-1]), p_bad=0.0003)
    seeds, h = [SEED_A31, SEED_B31], 2
    plan = nthash_b200.SeedPlan(seeds, h)
    ora = ORACLE.seed_batch(bases, off.astype(np.uint64), seeds, h, threads=4)
    assert_batch_equal(run_ragged(plan, bases, off), ora, 4, check_strands=True)
    # host-pointer entry (nthash_seed_batch)
    rows = ora["out"].shape[0]
    out = np.zeros((rows, 4), np.uint64); fw = np.zeros((rows, 2), np.uint64); rv = np.zeros((rows, 2), np.uint64)
    vb = np.zeros((rows + 31) // 32, np.uint32)
    arr = (C.c_char_p * 2)(*[s.encode() for s in seeds])
    offu = off.astype(np.uint64)
    rc = nthash_b200.LIB.nthash_seed_batch(bases.ctypes.data, offu.ctypes.data, len(lens), arr, 2, 31, h, out.ctypes.data,
                                           vb.ctypes.data, fw.ctypes.data, rv.ctypes.data, 0)
    assert rc == 0, nthash_b200.LIB.nthash_last_error()
    bits = ((vb[:, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(-1)[:rows]
    assert (bits == ora["valid"]).all() and (out == ora["out"]).all() and (fw == ora["fwd"]).all() and (rv == ora["rev"]).all()


def test_seed_properties_at_scale():
    # reference tests.cpp:349-377 (strand symmetry of palindromic seeds) and :447-463 (all-ones seed == NtHash)
    rng = np.random.default_rng(4)
    n, L, k = 100_000, 150, 31
    fwd = synth(rng, n * L).reshape(n, L)
    comp = np.zeros(256, np.uint8); comp[list(b"ACGT")] = list(b"TGCA")
    rc = np.ascontiguousarray(comp[fwd[:, ::-1]])
    plan = nthash_b200.SeedPlan([SEED_A31, SEED_B31, "1" * 31], 2)
    outs = []
    for arr in (fwd, rc):
        d_b, _keep = to_dev(arr.reshape(-1))
        outs.append(nthash_b200.seed_hashes_uniform(plan, d_b, n, L).out.view(n, L - k + 1, 6))
    d_b, _keep = to_dev(fwd.reshape(-1))
    kmer = nthash_b200.kmer_hashes_uniform(d_b, n, L, k, 2).out.view(n, L - k + 1, 2)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1].flip(1))
    assert torch.equal(outs[0][:, :, 4:6], kmer)


def test_seed_plan_errors():
    with pytest.raises(nthash_b200.NtHashError, match="not equal to k"):
        nthash_b200.SeedPlan(["1101", "101"], 1)   # seed.cpp:90-95
    with pytest.raises(nthash_b200.NtHashError):
        nthash_b200.SeedPlan(["11x11"], 1)


def test_blind_roll_and_peek_batches():
    # reference BlindNtHash::roll / peek (src/kmer.cpp:355-393); initial states = k-mer batch with read_len == k
    rng = np.random.default_rng(6)
    for k, h in ((5, 3), (31, 1), (64, 2)):
        n, steps = 2000, 12
        kmers = synth(rng, n * k).reshape(n, k)
        ins = rng.choice(np.frombuffer(b"ACGTNacgu", np.uint8), (n, steps))
        d_k, _keep = to_dev(kmers.reshape(-1))
        init = nthash_b200.kmer_hashes_uniform(d_k, n, k, k, h, want_strands=True)
        fwd, rev = init.fwd.clone(), init.rev.clone()
        window = np.concatenate([kmers, ins], axis=1)
        want = [ORACLE.blind_read(kmers[i].tobytes(), h, ins[i].tobytes()) for i in range(n)]
        assert (u64(init.out) == np.stack([w[0] for w in want])).all()
        for t in range(steps):
            out_base = torch.from_numpy(np.ascontiguousarray(window[:, t])).cuda()
            if t == 3:  # peek4 == what roll would give for each of ACGT, states untouched
                pk = u64(nthash_b200.blind_peek4(fwd, rev, out_base, k, h)).reshape(n, 4, h)
                for e, ch in enumerate(b"ACGT"):
                    f2, r2 = fwd.clone(), rev.clone()
                    got = nthash_b200.blind_roll(f2, r2, out_base, torch.full((n,), ch, dtype=torch.uint8, device="cuda"), k, h)
                    assert (u64(got) == pk[:, e]).all()
            in_base = torch.from_numpy(np.ascontiguousarray(ins[:, t])).cuda()
            got = u64(nthash_b200.blind_roll(fwd, rev, out_base, in_base, k, h))
            assert (got == np.stack([w[1][t] for w in want])).all()
            assert (u64(fwd) == np.array([w[2][t] for w in want], np.uint64)).all()
            assert (u64(rev) == np.array([w[3][t] for w in want], np.uint64)).all()


def test_seed_host_pipeline_many_chunks(monkeypatch):
    monkeypatch.setenv("NTHASH_B200_HOST_CHUNK_VALUES", "5000")
    rng = np.random.default_rng(22)
    seeds, h = [SEED_A31, SEED_B31], 2
    arr = (C.c_char_p * 2)(*[s.encode() for s in seeds])
    for lens in (rng.integers(0, 300, 300), np.full(200, 150)):
        off = ragged_offsets(lens).astype(np.uint64)
        bases = synth(rng, int(off[-1]), p_bad=0.004)
        ora = ORACLE.seed_batch(bases, off, seeds, h)
        rows = ora["out"].shape[0]
        out = np.zeros((rows, 4), np.uint64); fw = np.zeros((rows, 2), np.uint64); rv = np.zeros((rows, 2), np.uint64)
        vb = np.zeros((rows + 31) // 32, np.uint32)
        rc = nthash_b200.LIB.nthash_seed_batch(bases.ctypes.data, off.ctypes.data, len(off) - 1, arr, 2, 31, h, out.ctypes.data,
                                               vb.ctypes.data, fw.ctypes.data, rv.ctypes.data, 0)
        assert rc == 0, nthash_b200.LIB.nthash_last_error()
        bits = ((vb[:, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(-1)[:rows]
        assert (bits == ora["valid"]).all() and (out == ora["out"]).all() and (fw == ora["fwd"]).all() and (rv == ora["rev"]).all()
        out2 = np.zeros((rows, 4), np.uint64); vb2 = np.zeros_like(vb)   # no strands: fixed-length chunks take the specialised kernel
        assert nthash_b200.LIB.nthash_seed_batch(bases.ctypes.data, off.ctypes.data, len(off) - 1, arr, 2, 31, h, out2.ctypes.data, vb2.ctypes.data, None, None, 0) == 0
        assert (out2 == ora["out"]).all() and (vb2 == vb).all()


def test_blind_seed_roll_batch():
    # BlindSeedNtHash::roll(char) on many states (reference tests.cpp "Testing BlindSeedNtHash": after each roll the object
    # agrees with a SeedNtHash on the same window); invalid letters are hashed like any byte
    rng = np.random.default_rng(5)
    seeds, h = ["110101011", "101111101"], 3
    k, n = 9, 5000
    plan = nthash_b200.SeedPlan(seeds, h)
    windows = synth(rng, n * k, p_bad=0.01, lower=0.1).reshape(n, k)
    d_w = torch.zeros(n * k + 64, dtype=torch.uint8, device="cuda")
    d_w[: n * k] = torch.from_numpy(windows.reshape(-1).copy())
    kmers = d_w[: n * k].view(n, k)
    cur = windows.copy()
    for step in range(4):
        feed = synth(rng, n, p_bad=0.02, lower=0.1)
        out, fwd, rev = nthash_b200.blind_seed_roll(plan, kmers, torch.from_numpy(feed.copy()).cuda(), want_strands=True)
        torch.cuda.synchronize()
        cur = np.concatenate([cur[:, 1:], feed[:, None]], axis=1)
        assert (kmers.cpu().numpy() == cur).all()
        ora = ORACLE.seed_batch(cur.reshape(-1), np.arange(n + 1, dtype=np.uint64) * k, seeds, h)
        assert ora["valid"].all()
        assert (u64(out) == ora["out"]).all() and (u64(fwd) == ora["fwd"]).all() and (u64(rev) == ora["rev"]).all()


@pytest.mark.parametrize("seeds,h,n,L,p_bad", [([SEED_A31, SEED_B31], 3, 4000, 150, 0.002), ([SEED_A31, SEED_B31], 1, 40, 5000, 0.0008),
                                               (["110011", "101101"], 2, 3000, 64, 0.01), (["11100111"], 3, 1000, 151, 0.0),
                                               (["1" * 40 + "0" * 23 + "1" * 40], 1, 500, 300, 0.003)])
def test_seed_reduce_fused_in_the_specialised_kernel(seeds, h, n, L, p_bad, monkeypatch):
    """Uniform batches: count / sum / xor accumulated inside the generated kernel (nothing stored); items holding a byte for
    the exact path are redone by seed_reduce_dirty_kernel with the reference's visiting rule (jump on an invalid incoming
    base, NUL rule).  Same result as the oracle and as the two-pass form, short reads and reads cut into several items."""
    rng = np.random.default_rng(n + L)
    bases = synth(rng, n * L, p_bad=p_bad, lower=0.05)
    if p_bad:
        bases[rng.integers(0, n * L, 5)] = 0
        bases[rng.integers(0, n * L, 20)] = rng.choice(np.array([1, 3, 4, 5, 7], np.uint8), 20)
    d_b, _keep = to_dev(bases)
    plan = nthash_b200.SeedPlan(seeds, h)
    ora = ORACLE.seed_batch(bases, np.arange(n + 1, dtype=np.uint64) * L, seeds, h, want=(), threads=8)
    got = u64(nthash_b200.seed_reduce_uniform(plan, d_b, n, L))
    assert (int(got[0]), int(got[1]), int(got[2])) == (ora["n_emit"], ora["sum"], ora["xor"])
    monkeypatch.setenv("NTHASH_B200_SEED_REDUCE_TWO_PASS", "1")
    two = u64(nthash_b200.seed_reduce_uniform(plan, d_b, n, L))
    assert (got == two).all()


def test_seed_reduce_consumer(monkeypatch):
    # count / sum / xor of every visited window's hashes for SeedNtHash, computed on the device; dirty reads
    # follow the reference's own visiting rule
    rng = np.random.default_rng(15)
    seeds, h = ["1010101010101010101010101010101", "1101101101101101011011011011011"], 3
    plan = nthash_b200.SeedPlan(seeds, h)
    n, L = 6000, 150
    bases = synth(rng, n * L, p_bad=0.001)
    d_b, _keep = to_dev(bases)
    ora = ORACLE.seed_batch(bases, np.arange(n + 1, dtype=np.uint64) * L, seeds, h, want=(), threads=8)
    got = u64(nthash_b200.seed_reduce_uniform(plan, d_b, n, L))
    assert (int(got[0]), int(got[1]), int(got[2])) == (ora["n_emit"], ora["sum"], ora["xor"])
    # host entry, ragged, several pipeline chunks
    monkeypatch.setenv("NTHASH_B200_HOST_CHUNK_VALUES", "50000")
    lens = rng.integers(0, 300, 700)
    off = ragged_offsets(lens).astype(np.uint64)
    rb = synth(rng, int(off[-1]), p_bad=0.003)
    ora2 = ORACLE.seed_batch(rb, off, seeds, h, want=())
    res = np.zeros(3, np.uint64)
    arr = (C.c_char_p * 2)(*[s.encode() for s in seeds])
    assert nthash_b200.LIB.nthash_seed_reduce(rb.ctypes.data, off.ctypes.data, len(lens), arr, 2, 31, h, res.ctypes.data, 0) == 0, nthash_b200.LIB.nthash_last_error()
    assert (int(res[0]), int(res[1]), int(res[2])) == (ora2["n_emit"], ora2["sum"], ora2["xor"])
    # one seed, one hash per window (rows of a single value) through the two-pass form
    ora3 = ORACLE.seed_batch(rb, off, ["11100111"], 1, want=())
    arr1 = (C.c_char_p * 1)(b"11100111")
    assert nthash_b200.LIB.nthash_seed_reduce(rb.ctypes.data, off.ctypes.data, len(lens), arr1, 1, 8, 1, res.ctypes.data, 0) == 0, nthash_b200.LIB.nthash_last_error()
    assert (int(res[0]), int(res[1]), int(res[2])) == (ora3["n_emit"], ora3["sum"], ora3["xor"])
