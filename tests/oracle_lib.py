"""ctypes bindings for the CPU oracle (oracle/) — TEST INFRASTRUCTURE ONLY.

`ORACLE` is the plain-C restatement (oracle/nthash_oracle.c, symbols nto_*);
`REF` is the unmodified reference compiled with a C harness (oracle/_ref, symbols
ntr_*) or None when neither /root/reference nor a prebuilt oracle/_ref exists.
Both expose the same Python methods so a test can be parametrized over them.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")

u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)


def build():
    """Compile the oracle (and oracle/_ref when the reference tree is present)."""
    subprocess.run(["make", "-s", "-C", ODIR], check=True, stdout=subprocess.DEVNULL)


def _ptr(a, typ):
    return a.ctypes.data_as(typ) if a is not None else None


class OracleLib:
    def __init__(self, path, prefix):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self.kind = "reference" if prefix == "ntr_" else "port"
        f = self._fn
        f("kmer_read", C.c_size_t, [C.c_char_p, C.c_size_t, C.c_uint, C.c_uint, C.c_size_t, u64p, u64p, u64p, u64p, C.c_size_t])
        f("seed_read", C.c_size_t, [C.c_char_p, C.c_size_t, C.POINTER(C.c_char_p), C.c_uint, C.c_uint, C.c_uint, C.c_size_t, u64p, u64p, u64p, u64p, C.c_size_t])
        f("blind_read", None, [C.c_char_p, C.c_uint, C.c_uint, C.c_char_p, C.c_size_t, u64p, u64p, u64p, u64p])
        f("kmer_batch", C.c_uint64, [C.c_void_p, u64p, C.c_uint64, C.c_uint, C.c_uint, u64p, u8p, u64p, u64p, C.c_int, u64p, u64p])
        f("seed_batch", C.c_uint64, [C.c_void_p, u64p, C.c_uint64, C.POINTER(C.c_char_p), C.c_uint, C.c_uint, C.c_uint, u64p, u8p, u64p, u64p, C.c_int, u64p, u64p])
        if prefix == "nto_":
            f("srol", C.c_uint64, [C.c_uint64])
            f("sror", C.c_uint64, [C.c_uint64])
            f("srol_n", C.c_uint64, [C.c_uint64, C.c_uint])
            f("seed", C.c_uint64, [C.c_ubyte])
            f("srol_table", C.c_uint64, [C.c_ubyte, C.c_uint])
            f("get_blocks", C.c_int, [C.c_char_p, C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.c_uint])
            f("gen_bases", None, [C.c_void_p, C.c_uint64, C.c_uint64])
        else:
            f("kmer_strand", C.c_uint64, [C.c_char_p, C.c_uint, C.c_int])

    def _fn(self, name, res, args):
        fn = getattr(self.lib, self.prefix + name)
        fn.restype = res
        fn.argtypes = args
        setattr(self, "_" + name, fn)

    # -- per-read iterators -------------------------------------------------
    def kmer_read(self, seq, k, h, pos0=0):
        """-> (positions, hashes[n,h], fwd[n], rev[n]) or None where the reference ctor errors."""
        seq = seq.encode() if isinstance(seq, str) else bytes(seq)
        cap = max(len(seq), 1)
        pos = np.zeros(cap, np.uint64); hv = np.zeros((cap, h), np.uint64)
        fw = np.zeros(cap, np.uint64); rv = np.zeros(cap, np.uint64)
        n = self._kmer_read(seq, len(seq), k, h, pos0, _ptr(pos, u64p), _ptr(hv, u64p), _ptr(fw, u64p), _ptr(rv, u64p), cap)
        if n == C.c_size_t(-1).value:
            return None
        return pos[:n], hv[:n], fw[:n], rv[:n]

    def seed_read(self, seq, seeds, h, pos0=0):
        seq = seq.encode() if isinstance(seq, str) else bytes(seq)
        k = len(seeds[0]); m = len(seeds)
        arr = (C.c_char_p * m)(*[s.encode() for s in seeds])
        cap = max(len(seq), 1)
        pos = np.zeros(cap, np.uint64); hv = np.zeros((cap, m * h), np.uint64)
        fw = np.zeros((cap, m), np.uint64); rv = np.zeros((cap, m), np.uint64)
        n = self._seed_read(seq, len(seq), arr, m, h, k, pos0, _ptr(pos, u64p), _ptr(hv, u64p), _ptr(fw, u64p), _ptr(rv, u64p), cap)
        if n == C.c_size_t(-1).value:
            return None
        return pos[:n], hv[:n], fw[:n], rv[:n]

    def blind_read(self, kmer, h, chars_in):
        kmer = kmer.encode() if isinstance(kmer, str) else bytes(kmer)
        chars_in = chars_in.encode() if isinstance(chars_in, str) else bytes(chars_in)
        n = len(chars_in)
        h0 = np.zeros(h, np.uint64); hv = np.zeros((n, h), np.uint64)
        fw = np.zeros(n, np.uint64); rv = np.zeros(n, np.uint64)
        self._blind_read(kmer, len(kmer), h, chars_in, n, _ptr(h0, u64p), _ptr(hv, u64p), _ptr(fw, u64p), _ptr(rv, u64p))
        return h0, hv, fw, rv

    # -- batches in the engine's dense layout ---------------------------------
    @staticmethod
    def koff(read_off, k):
        lens = np.diff(read_off.astype(np.int64))
        nk = np.maximum(lens - k + 1, 0)
        return np.concatenate([[0], np.cumsum(nk)]).astype(np.uint64)

    def kmer_batch(self, bases, read_off, k, h, want=("out", "valid", "fwd", "rev"), threads=1):
        """-> dict(n_emit, sum, xor, out[tot,h], valid[tot] (uint8), fwd[tot], rev[tot])."""
        bases = np.ascontiguousarray(bases, np.uint8); read_off = np.ascontiguousarray(read_off, np.uint64)
        n = len(read_off) - 1
        tot = int(self.koff(read_off, k)[-1])
        o = np.empty((tot, h), np.uint64) if "out" in want else None
        v = np.empty(tot, np.uint8) if "valid" in want else None
        fw = np.empty(tot, np.uint64) if "fwd" in want else None
        rv = np.empty(tot, np.uint64) if "rev" in want else None
        s = C.c_uint64(0); x = C.c_uint64(0)
        ne = self._kmer_batch(bases.ctypes.data, _ptr(read_off, u64p), n, k, h, _ptr(o, u64p), _ptr(v, u8p), _ptr(fw, u64p), _ptr(rv, u64p), threads, C.byref(s), C.byref(x))
        return dict(n_emit=ne, sum=s.value, xor=x.value, out=o, valid=v, fwd=fw, rev=rv)

    def seed_batch(self, bases, read_off, seeds, h, want=("out", "valid", "fwd", "rev"), threads=1):
        bases = np.ascontiguousarray(bases, np.uint8); read_off = np.ascontiguousarray(read_off, np.uint64)
        n = len(read_off) - 1
        k = len(seeds[0]); m = len(seeds)
        arr = (C.c_char_p * m)(*[s.encode() for s in seeds])
        tot = int(self.koff(read_off, k)[-1])
        o = np.empty((tot, m * h), np.uint64) if "out" in want else None
        v = np.empty(tot, np.uint8) if "valid" in want else None
        fw = np.empty((tot, m), np.uint64) if "fwd" in want else None
        rv = np.empty((tot, m), np.uint64) if "rev" in want else None
        s = C.c_uint64(0); x = C.c_uint64(0)
        ne = self._seed_batch(bases.ctypes.data, _ptr(read_off, u64p), n, arr, m, h, k, _ptr(o, u64p), _ptr(v, u8p), _ptr(fw, u64p), _ptr(rv, u64p), threads, C.byref(s), C.byref(x))
        return dict(n_emit=ne, sum=s.value, xor=x.value, out=o, valid=v, fwd=fw, rev=rv)

    # -- port-only helpers ------------------------------------------------------
    def gen_bases(self, n, seed):
        a = np.empty(n, np.uint8)
        self._gen_bases(a.ctypes.data, n, seed)
        return a

    def get_blocks(self, seed):
        cap = len(seed) + 2
        b = (C.c_uint * (2 * cap))(); m = (C.c_uint * cap)()
        nb = C.c_uint(0); nm = C.c_uint(0)
        rc = self._get_blocks(seed.encode(), b, C.byref(nb), m, C.byref(nm), cap)
        assert rc == 0
        return [(b[2 * i], b[2 * i + 1]) for i in range(nb.value)], [m[i] for i in range(nm.value)]


def _load():
    opath = os.path.join(ODIR, "libnthash_oracle.so")
    rpath = os.path.join(ODIR, "_ref", "libnthash_ref.so")
    if not os.path.exists(opath) or (not os.path.exists(rpath) and os.path.isdir("/root/reference/src")):
        build()
    oracle = OracleLib(opath, "nto_")
    ref = OracleLib(rpath, "ntr_") if os.path.exists(rpath) else None
    return oracle, ref


ORACLE, REF = _load()


# ---- consumers restated on top of the oracle's hash rows (numpy; what the fused GPU consumers must reproduce) -------------
def minimizer_bits(out0, valid, koff, w):
    """Boolean [rows]: row j is the leftmost minimum of hashes()[0] over the visited k-mers of at least one window of w
    consecutive k-mers of its read (include/nthash_b200.h: nthash_kmer_minimizer_*).  out0: uint64 [rows], valid: bool
    [rows], koff: dense row offsets of the reads."""
    from numpy.lib.stride_tricks import sliding_window_view
    out0 = np.asarray(out0, np.uint64)
    valid = np.asarray(valid).astype(bool)
    sel = np.zeros(len(out0), bool)
    big = np.uint64(0xFFFFFFFFFFFFFFFF)
    for r in range(len(koff) - 1):
        a, b = int(koff[r]), int(koff[r + 1])
        if b - a < w:
            continue
        key = np.where(valid[a:b], out0[a:b], big)      # an unvisited k-mer never wins (a visited one equal to 2^64-1: 2^-64)
        kw = sliding_window_view(key, w)
        m = kw.argmin(axis=1)                           # first occurrence = leftmost
        anyv = sliding_window_view(valid[a:b], w).any(axis=1)
        pos = (np.arange(len(m)) + m)[anyv]
        sel[a + pos] = True
    return sel


def sketch_counts(out0, valid, sample_bits, index_bits):
    """ntCard-style table: multiplicity of the index_bits below the top sample_bits of every visited k-mer's canonical hash
    whose top sample_bits are zero -> (counters uint32 [2^index_bits], number sampled)."""
    h = np.asarray(out0, np.uint64)[np.asarray(valid).astype(bool)]
    s = h[(h >> np.uint64(64 - sample_bits)) == 0]
    idx = ((s >> np.uint64(64 - sample_bits - index_bits)) & np.uint64((1 << index_bits) - 1)).astype(np.int64)
    return np.bincount(idx, minlength=1 << index_bits).astype(np.uint32), len(s)
