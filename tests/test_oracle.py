"""Pins the CPU oracle (oracle/nthash_oracle.c) before anything is checked against it.

1. every golden vector the reference's own tests hold for the hot path
   (reference tests/tests.cpp:54-57, :193-200, :236-240) and the known answers
   SURVEY.md Appendix C records from the compiled reference;
2. the properties the reference's tests assert (tests.cpp blocks 2-4, 8, 10, 12, 17);
3. a randomized differential against the UNMODIFIED reference compiled into
   oracle/_ref (skipped only if neither /root/reference nor a prebuilt oracle/_ref exists).
"""
import numpy as np
import pytest

from oracle_lib import ORACLE, REF

LIBS = [pytest.param(ORACLE, id="port")] + ([pytest.param(REF, id="reference")] if REF else [])
needs_ref = pytest.mark.skipif(REF is None, reason="oracle/_ref not built and /root/reference absent")

SEED_A31 = "1010101010101010101010101010101"
SEED_B31 = "1101101101101101011011011011011"


def revcomp(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


# ---------------------------------------------------------------- goldens --
@pytest.mark.parametrize("lib", LIBS)
def test_golden_kmer_hash_values(lib):
    # tests.cpp:47-69 — positions 1 and 2 of ACATGCATGCA, k=5, h=3; BlindNtHash agrees
    pos, hv, _, _ = lib.kmer_read("ACATGCATGCA", 5, 3)
    want = np.array([[0x38CC00F940AEBDAE, 0xAB7E1B110E086FC6, 0x11A1818BCFDD553],
                     [0x603A48C5A11C794A, 0xE66016E61816B9C4, 0xC5B13CB146996FFE]], np.uint64)
    assert list(pos) == list(range(7))
    assert (hv[1:3] == want).all()
    h0, bh, _, _ = lib.blind_read("ACATG", 3, "CATGCA")
    assert (h0 == hv[0]).all() and (bh == hv[1:]).all()


@pytest.mark.parametrize("lib", LIBS)
def test_golden_spaced_seed_hash_values(lib):
    # tests.cpp:228-248
    pos, hv, _, _ = lib.seed_read("ACATGCATGCA", ["11100111"], 3)
    want = np.array([[0x10BE4904AD8DE5D, 0x3E29E4F4C991628C, 0x3F35C984B13FEB20],
                     [0x8200A7AA3EAF17C8, 0x344198402F4C2A9C, 0xB6423FE62E69C40C],
                     [0x3CE8ADCBEAA56532, 0x162E91A4DBEDBF11, 0x53173F786A031F45]], np.uint64)
    assert (hv[:3] == want).all() and list(pos) == [0, 1, 2, 3]


@pytest.mark.parametrize("lib", LIBS)
def test_golden_skipping_ns(lib):
    # tests.cpp:181-208
    seq = list("ACGTACACTGGACTGAGTCT")
    seq[10] = seq[11] = "N"
    k = (20 - 2) // 2 - 1
    pos, _, _, _ = lib.kmer_read("".join(seq), k, 3)
    assert list(pos) == list(range(0, 10 - k + 1)) + list(range(12, 20 - k + 1))


@pytest.mark.parametrize("lib", LIBS)
def test_survey_appendix_c_known_answers(lib):
    seq = ORACLE.gen_bases(1000, 42).tobytes()
    assert seq[:20] == b"CCCGGTGCTGGTTTGAGCGA"
    pos, hv, fw, rv = lib.kmer_read(seq, 31, 1)
    assert len(pos) == 970
    assert (int(fw[0]), int(rv[0]), int(hv[0, 0])) == (0x54C31B64E55CF218, 0x61E48AE33B0133BE, 0xB6A7A648205E25D6)
    assert int(hv[1, 0]) == 0x704A84790145DAEB and int(hv[2, 0]) == 0xBAD6DB28BFB4AC98
    assert int(hv.sum(dtype=np.uint64)) == 0x429E8D1795548BB7
    assert int(np.bitwise_xor.reduce(hv[:, 0])) == 0x23372BA181F010EB
    _, h4, _, _ = lib.kmer_read(seq, 31, 4)
    assert [int(x) for x in h4[0]] == [0xB6A7A648205E25D6, 0xDD1CB9BBF8438851, 0xB925C6D0DE029227, 0x6FCD6D1DE2B5D2AD]
    assert int(h4.sum(dtype=np.uint64)) == 0xA60EE73A342610F3
    p63, h63, f63, r63 = lib.kmer_read(seq, 63, 1)
    assert len(p63) == 938 and int(h63[0, 0]) == 0xE41A99C0E9B9DEA4
    assert (int(f63[0]), int(r63[0])) == (0x7A6A61727EE36902, 0x69B0384E6AD675A2)
    assert int(h63.sum(dtype=np.uint64)) == 0x085DA314475757B6
    ps, hs, _, _ = lib.seed_read(seq, [SEED_A31, SEED_B31], 3)
    assert len(ps) == 970
    assert [int(x) for x in hs[0]] == [0xE09BC50CEA32E5AA, 0xF206E6A6EBAE2333, 0x5033979BEAA8B115,
                                      0x194F0B34C1432F4C, 0x74181E1C17687846, 0x282AFC7155FDA2BA]
    assert int(hs.sum(dtype=np.uint64)) == 0x6C20E8349FC9CF61
    dirty = bytearray(seq[:200])
    dirty[50] = dirty[51] = ord("N"); dirty[120] = ord("n"); dirty[199] = ord("N")
    pd, hd, _, _ = lib.kmer_read(bytes(dirty), 31, 1)
    assert len(pd) == 106 and int(hd.sum(dtype=np.uint64)) == 0x05C7650B5A997843
    pq, hq, _, _ = lib.seed_read(bytes(dirty), [SEED_A31, SEED_B31], 1)
    assert len(pq) == 109 and int(hq.sum(dtype=np.uint64)) == 0x69A56848390E3347


def test_get_blocks_examples():
    # SURVEY.md A.5 (dumped from the reference's get_blocks, seed.cpp:19-66)
    assert ORACLE.get_blocks(SEED_A31) == ([], list(range(0, 31, 2)))
    assert ORACLE.get_blocks(SEED_B31) == ([(0, 31)], [2, 5, 8, 11, 14, 16, 19, 22, 25, 28])
    assert ORACLE.get_blocks("11100111") == ([(0, 3), (5, 8)], [])
    assert ORACLE.get_blocks("110011") == ([(0, 2), (4, 6)], [])
    assert ORACLE.get_blocks("101101") == ([(2, 4)], [0, 5])


# ------------------------------------------------------------- primitives --
def test_primitives():
    rng = np.random.default_rng(1)
    for x in [int(v) for v in rng.integers(0, 2**64, 200, dtype=np.uint64)] + [0, 2**64 - 1, 1 << 63, 1 << 32, 1 << 33, 1]:
        assert ORACLE._sror(ORACLE._srol(x)) == x
        y = x
        for d in range(0, 70):
            assert ORACLE._srol_n(x, d) == y
            y = ORACLE._srol(y)
        assert ORACLE._srol_n(x, 1023) == x  # lcm(31,33)
    for c in b"ACGTUacgtu":
        assert ORACLE._seed(c) != 0
    for c in b"NnRYKMSWBDHV-*. \x00\xff":
        assert ORACLE._seed(c) == 0
    assert ORACLE._seed(ord("U")) == ORACLE._seed(ord("T")) == ORACLE._seed(ord("t"))


@needs_ref
def test_strand_hashes_match_reference_init_path():
    # closed-form base hashes vs the reference's tetramer-table init (kmer.cpp:43-73,123-152)
    rng = np.random.default_rng(2)
    for k in [3, 4, 5, 6, 7, 8, 31, 32, 33, 63, 64, 65, 127, 200, 255, 1023, 1024, 2000]:
        s = bytes(rng.choice(list(b"ACGTacgtUu"), k).astype(np.uint8))
        _, _, fw, rv = ORACLE.kmer_read(s, k, 1)
        assert int(fw[0]) == REF._kmer_strand(s, k, 0) and int(rv[0]) == REF._kmer_strand(s, k, 1), k


# -------------------------------------------------------------- properties --
@pytest.mark.parametrize("lib", LIBS)
def test_properties_from_reference_tests(lib):
    # block 2: count and identical first/last 4-mer
    pos, hv, _, _ = lib.kmer_read("AGTCAGTC", 4, 3)
    assert len(pos) == 5 and (hv[0] == hv[-1]).all()
    # block 3: rolled == freshly initialised
    seq = "ACGTACACTGGACTGAGTCT"
    _, hv, _, _ = lib.kmer_read(seq, 18, 3)
    for i in range(3):
        assert (lib.kmer_read(seq[i:i + 18], 18, 3)[1][0] == hv[i]).all()
    # block 4: canonical
    assert (lib.kmer_read(seq, 20, 3)[1] == lib.kmer_read(revcomp(seq), 20, 3)[1]).all()
    # block 8: RNA
    d = "ACGTACACTGGACTGAGTCTACGG"
    assert (lib.kmer_read(d, 20, 3)[1] == lib.kmer_read(d.replace("T", "U"), 20, 3)[1]).all()
    # block 10: mutations at don't-care positions; roll == base
    seeds = ["111110000000011111", "111111100001111111"]
    ref_h = lib.seed_read(seq, seeds, 2)[1]
    assert ref_h.shape == (3, 4)
    for mut in ["ACGTACACTTGACTGAGTCT", "ACGTACACTGTACTGAGTCT", "ACGTACACTGCACTGAGTCT"]:
        assert (lib.seed_read(mut, seeds, 2)[1] == ref_h).all()
    for i in range(3):
        assert (lib.seed_read(seq[i:i + 18], seeds, 2)[1][0] == ref_h[i]).all()
    # block 12: strand symmetry of palindromic seeds, k=50, h=4 (first seed's values, as the reference compares)
    f = "CACTCGGCCACACACACACACACACACCCTCACACACACAAAACGCACAC"
    seeds50 = ["11011000001100101101011000011010110100110000011011",
               "01010000101001110100111011011100101110010100001010",
               "11100000100111010111000100100011101011100100000111",
               "01111000011000111101000011000010111100011000011110",
               "00111000011000111101000011000010111100011000011100",
               "00000000000000000000000011000000000000000000000000",
               "11111111111111111111111100111111111111111111111111",
               "11111111111111111111111111111111111111111111111111"]
    assert (lib.seed_read(f, seeds50, 4)[1] == lib.seed_read(revcomp(f), seeds50, 4)[1]).all()
    # block 17: k-mer == all-ones seed
    s = "ATGCTAGTAGCTGAC"
    assert (lib.kmer_read(s, 5, 3)[1] == lib.seed_read(s, ["11111"], 3)[1]).all()


@pytest.mark.parametrize("lib", LIBS)
def test_ctor_errors(lib):
    assert lib.kmer_read("ACGT", 5, 1) is None          # len < k, kmer.cpp:215-220
    assert lib.kmer_read("ACGTACGT", 4, 1, pos0=5) is None  # pos > len-k, kmer.cpp:221-225
    assert lib.seed_read("ACGTACGT", ["1101", "101"], 1) is None  # seed.cpp:90-95


# ------------------------------------------------- differential vs reference --
def _dirty(rng, n, p_bad):
    a = rng.choice(np.frombuffer(b"ACGT", np.uint8), n)
    bad = rng.random(n) < p_bad
    a[bad] = rng.choice(np.frombuffer(b"NnRYKMSWBDHV-acgtuU", np.uint8), int(bad.sum()))
    return a


@needs_ref
@pytest.mark.parametrize("k", [3, 4, 5, 16, 31, 32, 33, 63, 64, 65, 127, 255])
def test_kmer_port_vs_reference_random(k):
    rng = np.random.default_rng(k)
    for trial in range(6):
        n = int(rng.integers(k, 4 * k + 200))
        a = _dirty(rng, n, [0.0, 0.002, 0.02, 0.2][trial % 4])
        h = [1, 2, 4, 7][trial % 4]
        po, ho, fo, ro = ORACLE.kmer_read(a.tobytes(), k, h)
        pr, hr, fr, rr = REF.kmer_read(a.tobytes(), k, h)
        assert (po == pr).all() and (ho == hr).all() and (fo == fr).all() and (ro == rr).all()


@needs_ref
def test_kmer_h255_and_pos0():
    a = ORACLE.gen_bases(300, 7).tobytes()
    for pos0 in (0, 1, 17, 300 - 31):
        o = ORACLE.kmer_read(a, 31, 255, pos0)
        r = REF.kmer_read(a, 31, 255, pos0)
        assert all((x == y).all() for x, y in zip(o, r))


SEED_SETS = [
    [SEED_A31, SEED_B31],
    ["11100111"], ["110011", "101101"], ["11111"], ["1101011"], ["1010101"],
    ["111110000000011111", "111111100001111111"],
    ["111111111101111111111", "110111010010010111011"],
    ["1101100", "0011011", "1000001"],                      # asymmetric
    ["0110", "1001", "0101"],                               # leading/trailing zeros
    ["1" * 40 + "0" * 23 + "1" * 40],
]


@needs_ref
@pytest.mark.parametrize("seeds", SEED_SETS, ids=lambda s: s[0][:12])
def test_seed_port_vs_reference_random(seeds, capfd):
    k = len(seeds[0])
    rng = np.random.default_rng(len(seeds) * 100 + k)
    for trial in range(6):
        n = int(rng.integers(k, 4 * k + 150))
        a = _dirty(rng, n, [0.0, 0.01, 0.05, 0.3][trial % 4])
        if trial == 5:
            a[rng.integers(0, n, 3)] = 0  # NUL bytes: the only thing SeedNtHash::init rejects (seed.cpp:151)
        h = [1, 3, 2, 5][trial % 4]
        o = ORACLE.seed_read(a.tobytes(), seeds, h)
        r = REF.seed_read(a.tobytes(), seeds, h)
        assert len(o[0]) == len(r[0]) and all((x == y).all() for x, y in zip(o, r))
    capfd.readouterr()  # swallow the reference's "not symmetric" warnings


@needs_ref
def test_blind_port_vs_reference():
    rng = np.random.default_rng(5)
    for k in (3, 5, 31, 64, 100):
        kmer = bytes(rng.choice(list(b"ACGT"), k).astype(np.uint8))
        ins = bytes(rng.choice(list(b"ACGTNacgu"), 50).astype(np.uint8))  # no validity check in BlindNtHash
        o = ORACLE.blind_read(kmer, 3, ins)
        r = REF.blind_read(kmer, 3, ins)
        assert all((x == y).all() for x, y in zip(o, r))


@needs_ref
def test_batches_port_vs_reference():
    rng = np.random.default_rng(11)
    lens = rng.integers(0, 90, 200)
    lens[:5] = [0, 1, 30, 31, 32]
    read_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    bases = _dirty(rng, int(read_off[-1]), 0.01)
    for threads in (1, 3):
        o = ORACLE.kmer_batch(bases, read_off, 31, 2, threads=threads)
        r = REF.kmer_batch(bases, read_off, 31, 2, threads=threads)
        for key in ("n_emit", "sum", "xor"):
            assert o[key] == r[key]
        for key in ("out", "valid", "fwd", "rev"):
            assert (o[key] == r[key]).all()
        o = ORACLE.seed_batch(bases, read_off, [SEED_A31, SEED_B31], 3, threads=threads)
        r = REF.seed_batch(bases, read_off, [SEED_A31, SEED_B31], 3, threads=threads)
        for key in ("n_emit", "sum", "xor"):
            assert o[key] == r[key]
        for key in ("out", "valid", "fwd", "rev"):
            assert (o[key] == r[key]).all()
    # dense layout agrees with the per-read iterator
    koff = ORACLE.koff(read_off, 31)
    o = ORACLE.kmer_batch(bases, read_off, 31, 2)
    for r_i in range(len(lens)):
        if lens[r_i] < 31:
            continue
        seq = bases[int(read_off[r_i]):int(read_off[r_i + 1])].tobytes()
        pos, hv, _, _ = ORACLE.kmer_read(seq, 31, 2)
        rows = int(koff[r_i]) + pos.astype(np.int64)
        assert (o["out"][rows] == hv).all()
        assert int(o["valid"][int(koff[r_i]):int(koff[r_i + 1])].sum()) == len(pos)
