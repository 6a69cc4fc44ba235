"""Device-resident Python front end over the C ABI (include/nthash_b200.h).

PyTorch is used only for what it is good at here: owning HBM buffers and CUDA streams.  Every
hash is produced by the hand-written kernels in csrc/ through libnthash_b200.so; if that library
or a B200 is missing these functions raise — there is no CPU path in this package.

Shapes follow the C ABI's dense window rows: row w = koff[r] + p is the window of read r that
starts at base p (the reference's get_pos(), include/nthash/nthash.hpp:170).
"""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from ._lib import LIB, check


@dataclass
class HashBatch:
    """Result of one batch call; all tensors live on the GPU that hashed them."""
    out: torch.Tensor                    # int64 view of uint64 [rows, H]
    valid_bits: Optional[torch.Tensor]   # int32 words, bit (w & 31) of word (w >> 5)
    koff: Optional[torch.Tensor]         # int64 [n_reads + 1] (None for uniform batches: koff[r] = r * nk)
    rows: int
    fwd: Optional[torch.Tensor] = None
    rev: Optional[torch.Tensor] = None

    def valid_mask(self) -> torch.Tensor:
        """Boolean [rows] expansion of valid_bits (a convenience for tests, not a hot path)."""
        words = self.valid_bits.view(torch.int32)
        bits = (words.unsqueeze(1) >> torch.arange(32, device=words.device, dtype=torch.int32)) & 1
        return bits.reshape(-1)[: self.rows].bool()


def _stream_ptr(stream):
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _check_bases(bases):
    if not (bases.is_cuda and bases.dtype == torch.uint8 and bases.is_contiguous() and bases.dim() == 1):
        raise ValueError("bases must be a contiguous 1-D uint8 CUDA tensor")


def kmer_hashes_uniform(bases, n_reads, read_len, k, num_hashes=1, want_valid=True, want_strands=False,
                        out=None, valid_bits=None, stream=None) -> HashBatch:
    """NtHash over n_reads fixed-length reads stored back to back (C ABI: nthash_kmer_batch_uniform_dev)."""
    _check_bases(bases)
    nk = max(read_len - k + 1, 0)
    rows = n_reads * nk
    dev = bases.device
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty((rows, num_hashes), dtype=torch.int64, device=dev)
        if want_valid and valid_bits is None:
            valid_bits = torch.empty(int(LIB.nthash_valid_words(rows)), dtype=torch.int32, device=dev)
        fwd = torch.empty(rows, dtype=torch.int64, device=dev) if want_strands else None
        rev = torch.empty(rows, dtype=torch.int64, device=dev) if want_strands else None
        check(LIB.nthash_kmer_batch_uniform_dev(_ptr(bases), bases.numel(), n_reads, read_len, k, num_hashes,
                                                _ptr(out), _ptr(valid_bits), _ptr(fwd), _ptr(rev), _stream_ptr(stream)))
    return HashBatch(out, valid_bits, None, rows, fwd, rev)


def kmer_hashes_packed2bit_uniform(packed, invalid_bits, first_base, n_reads, read_len, k, num_hashes=1, want_valid=True, out=None,
                                   valid_bits=None, stream=None) -> HashBatch:
    """NtHash straight from device-resident 2-bit packed bases (nthash_kmer_batch_packed2bit_uniform_dev; raises NtHashError
    with NTHASH_ERR_UNSUPPORTED for shapes that need the ASCII expansion).  `packed`: uint8 CUDA tensor, readable to the next
    16-byte multiple; `invalid_bits`: int32 CUDA tensor or None."""
    rows = n_reads * max(read_len - k + 1, 0)
    dev = packed.device
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty((rows, num_hashes), dtype=torch.int64, device=dev)
        if want_valid and valid_bits is None:
            valid_bits = torch.empty(int(LIB.nthash_valid_words(rows)), dtype=torch.int32, device=dev)
        check(LIB.nthash_kmer_batch_packed2bit_uniform_dev(_ptr(packed), _ptr(invalid_bits), first_base, n_reads, read_len, k, num_hashes,
                                                           _ptr(out), _ptr(valid_bits) if want_valid else None, _stream_ptr(stream)))
    return HashBatch(out, valid_bits if want_valid else None, None, rows, None, None)


def kmer_hashes(bases, read_off, k, num_hashes=1, want_valid=True, want_strands=False, stream=None, out=None) -> HashBatch:
    """NtHash over ragged reads: read r is bases[read_off[r]:read_off[r+1]] (nthash_kmer_plan_dev + nthash_kmer_batch_dev).
    `out`: optional preallocated int64 CUDA tensor with at least rows * num_hashes elements (reused across calls)."""
    _check_bases(bases)
    if not (read_off.is_cuda and read_off.dtype == torch.int64 and read_off.is_contiguous()):
        raise ValueError("read_off must be a contiguous int64 CUDA tensor")
    n_reads = read_off.numel() - 1
    dev = bases.device
    with torch.cuda.device(dev):
        koff = torch.empty(n_reads + 1, dtype=torch.int64, device=dev)
        rows = C.c_uint64(0)
        max_len = C.c_uint64(0)
        check(LIB.nthash_kmer_plan_dev(_ptr(read_off), n_reads, k, _ptr(koff), C.byref(rows), C.byref(max_len), _stream_ptr(stream)))
        rows = rows.value
        if out is None:
            out = torch.empty((rows, num_hashes), dtype=torch.int64, device=dev)
        else:
            if not (out.is_cuda and out.dtype == torch.int64 and out.is_contiguous() and out.numel() >= rows * num_hashes):
                raise ValueError("out must be a contiguous int64 CUDA tensor with at least rows * num_hashes elements")
            out = out.view(-1)[: rows * num_hashes].view(rows, num_hashes)
        valid_bits = torch.empty(int(LIB.nthash_valid_words(rows)), dtype=torch.int32, device=dev) if want_valid else None
        fwd = torch.empty(rows, dtype=torch.int64, device=dev) if want_strands else None
        rev = torch.empty(rows, dtype=torch.int64, device=dev) if want_strands else None
        check(LIB.nthash_kmer_batch_dev(_ptr(bases), bases.numel(), _ptr(read_off), _ptr(koff), n_reads, max_len.value, k,
                                        num_hashes, _ptr(out), _ptr(valid_bits), _ptr(fwd), _ptr(rev), _stream_ptr(stream)))
    return HashBatch(out, valid_bits, koff, rows, fwd, rev)


class RaggedPlan:
    """Layout of one ragged batch, planned once on the current device (C ABI: nthash_ragged_plan_create): koff, the row
    total and the kernels' item tables.  `read_off` (int64 CUDA tensor, n_reads + 1) is kept alive by this object.
    Calls made with a plan only enqueue kernels (no host synchronisation; CUDA-graph capturable)."""

    def __init__(self, read_off, k, stream=None):
        if not (read_off.is_cuda and read_off.dtype == torch.int64 and read_off.is_contiguous()):
            raise ValueError("read_off must be a contiguous int64 CUDA tensor")
        self.read_off, self.k, self.n_reads = read_off, k, read_off.numel() - 1
        self._h = C.c_void_p()
        with torch.cuda.device(read_off.device):
            check(LIB.nthash_ragged_plan_create(_ptr(read_off), self.n_reads, k, _stream_ptr(stream), C.byref(self._h)))
        self.rows = int(LIB.nthash_ragged_plan_rows(self._h))
        self.max_read_len = int(LIB.nthash_ragged_plan_max_read_len(self._h))

    def koff(self):
        """int64 CUDA tensor [n_reads + 1]: the dense row offset of every read (recomputed into a tensor of the caller's;
        the plan's own copy stays inside the library: nthash_ragged_plan_koff)."""
        out = torch.empty(self.n_reads + 1, dtype=torch.int64, device=self.read_off.device)
        rows, max_len = C.c_uint64(0), C.c_uint64(0)
        with torch.cuda.device(self.read_off.device):
            check(LIB.nthash_kmer_plan_dev(_ptr(self.read_off), self.n_reads, self.k, _ptr(out), C.byref(rows), C.byref(max_len), _stream_ptr(None)))
        return out

    def close(self):
        if self._h:
            LIB.nthash_ragged_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def kmer_hashes_planned(plan, bases, num_hashes=1, want_valid=True, want_strands=False, out=None, valid_bits=None, stream=None) -> HashBatch:
    """NtHash over a planned ragged batch (nthash_kmer_batch_planned_dev): enqueue only."""
    _check_bases(bases)
    dev = bases.device
    rows = plan.rows
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty((rows, num_hashes), dtype=torch.int64, device=dev)
        if want_valid and valid_bits is None:
            valid_bits = torch.empty(int(LIB.nthash_valid_words(rows)), dtype=torch.int32, device=dev)
        fwd = torch.empty(rows, dtype=torch.int64, device=dev) if want_strands else None
        rev = torch.empty(rows, dtype=torch.int64, device=dev) if want_strands else None
        check(LIB.nthash_kmer_batch_planned_dev(plan._h, _ptr(bases), bases.numel(), num_hashes, _ptr(out), _ptr(valid_bits if want_valid else None),
                                                _ptr(fwd), _ptr(rev), _stream_ptr(stream)))
    return HashBatch(out, valid_bits if want_valid else None, None, rows, fwd, rev)


def seed_hashes_planned(seed_plan, plan, bases, want_valid=True, want_strands=False, out=None, valid_bits=None, stream=None) -> HashBatch:
    """SeedNtHash over a planned ragged batch (nthash_seed_batch_planned_dev): enqueue only."""
    _check_bases(bases)
    m, H = len(seed_plan.seeds), len(seed_plan.seeds) * seed_plan.h
    dev = bases.device
    rows = plan.rows
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty((rows, H), dtype=torch.int64, device=dev)
        if want_valid and valid_bits is None:
            valid_bits = torch.empty(int(LIB.nthash_valid_words(rows)), dtype=torch.int32, device=dev)
        fwd = torch.empty((rows, m), dtype=torch.int64, device=dev) if want_strands else None
        rev = torch.empty((rows, m), dtype=torch.int64, device=dev) if want_strands else None
        check(LIB.nthash_seed_batch_planned_dev(seed_plan._h, plan._h, _ptr(bases), bases.numel(), _ptr(out), _ptr(valid_bits if want_valid else None),
                                                _ptr(fwd), _ptr(rev), _stream_ptr(stream)))
    return HashBatch(out, valid_bits if want_valid else None, None, rows, fwd, rev)


class SeedPlan:
    """Compiled spaced-seed set on the current device (C ABI: nthash_seed_plan_create / _destroy).

    Mirrors what the SeedNtHash constructors do with their `seeds` argument
    (reference src/seed.cpp:449-491): strings of '1' (care) and '0' (don't care), all of length k."""

    def __init__(self, seeds, num_hashes_per_seed=1):
        self.seeds = list(seeds)
        self.k = len(self.seeds[0])
        self.h = num_hashes_per_seed
        arr = (C.c_char_p * len(self.seeds))(*[s.encode() for s in self.seeds])
        self._h = C.c_void_p()
        check(LIB.nthash_seed_plan_create(arr, len(self.seeds), self.k, num_hashes_per_seed, C.byref(self._h)))

    @property
    def symmetric(self):
        return bool(LIB.nthash_seed_plan_symmetric(self._h))

    def close(self):
        if self._h:
            LIB.nthash_seed_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def seed_hashes_uniform(plan, bases, n_reads, read_len, want_valid=True, want_strands=False, out=None,
                        valid_bits=None, stream=None) -> HashBatch:
    """SeedNtHash over fixed-length reads (nthash_seed_batch_uniform_dev); rows hold n_seeds*h values, seed-major."""
    _check_bases(bases)
    m, H = len(plan.seeds), len(plan.seeds) * plan.h
    rows = n_reads * max(read_len - plan.k + 1, 0)
    dev = bases.device
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty((rows, H), dtype=torch.int64, device=dev)
        if want_valid and valid_bits is None:
            valid_bits = torch.empty(int(LIB.nthash_valid_words(rows)), dtype=torch.int32, device=dev)
        fwd = torch.empty((rows, m), dtype=torch.int64, device=dev) if want_strands else None
        rev = torch.empty((rows, m), dtype=torch.int64, device=dev) if want_strands else None
        check(LIB.nthash_seed_batch_uniform_dev(plan._h, _ptr(bases), bases.numel(), n_reads, read_len, _ptr(out),
                                                _ptr(valid_bits), _ptr(fwd), _ptr(rev), _stream_ptr(stream)))
    return HashBatch(out, valid_bits, None, rows, fwd, rev)


def seed_hashes(plan, bases, read_off, want_valid=True, want_strands=False, stream=None) -> HashBatch:
    """SeedNtHash over ragged reads (nthash_kmer_plan_dev for the layout + nthash_seed_batch_dev)."""
    _check_bases(bases)
    n_reads = read_off.numel() - 1
    m, H = len(plan.seeds), len(plan.seeds) * plan.h
    dev = bases.device
    with torch.cuda.device(dev):
        koff = torch.empty(n_reads + 1, dtype=torch.int64, device=dev)
        rows = C.c_uint64(0)
        max_len = C.c_uint64(0)
        check(LIB.nthash_kmer_plan_dev(_ptr(read_off), n_reads, plan.k, _ptr(koff), C.byref(rows), C.byref(max_len), _stream_ptr(stream)))
        rows = rows.value
        out = torch.empty((rows, H), dtype=torch.int64, device=dev)
        valid_bits = torch.empty(int(LIB.nthash_valid_words(rows)), dtype=torch.int32, device=dev) if want_valid else None
        fwd = torch.empty((rows, m), dtype=torch.int64, device=dev) if want_strands else None
        rev = torch.empty((rows, m), dtype=torch.int64, device=dev) if want_strands else None
        check(LIB.nthash_seed_batch_dev(plan._h, _ptr(bases), bases.numel(), _ptr(read_off), _ptr(koff), n_reads, max_len.value,
                                        _ptr(out), _ptr(valid_bits), _ptr(fwd), _ptr(rev), _stream_ptr(stream)))
    return HashBatch(out, valid_bits, koff, rows, fwd, rev)


def blind_roll(fwd, rev, out_base, in_base, k, num_hashes=1, stream=None) -> torch.Tensor:
    """BlindNtHash::roll(char_in) on n states at once; fwd/rev are updated in place, returns hashes [n, h]."""
    n = fwd.numel()
    out = torch.empty((n, num_hashes), dtype=torch.int64, device=fwd.device)
    with torch.cuda.device(fwd.device):
        check(LIB.nthash_blind_roll_batch_dev(_ptr(fwd), _ptr(rev), _ptr(out_base), _ptr(in_base), n, k, num_hashes,
                                              _ptr(out), _stream_ptr(stream)))
    return out


def blind_peek4(fwd, rev, out_base, k, num_hashes=1, stream=None) -> torch.Tensor:
    """BlindNtHash::peek('A'|'C'|'G'|'T') on n states: hashes [n, 4, h]; states untouched."""
    n = fwd.numel()
    out = torch.empty((n, 4, num_hashes), dtype=torch.int64, device=fwd.device)
    with torch.cuda.device(fwd.device):
        check(LIB.nthash_blind_peek4_batch_dev(_ptr(fwd), _ptr(rev), _ptr(out_base), n, k, num_hashes, _ptr(out),
                                               _stream_ptr(stream)))
    return out


def kmer_reduce_uniform(bases, n_reads, read_len, k, num_hashes=1, stream=None) -> torch.Tensor:
    """Fused consumer (nthash_kmer_reduce_uniform_dev): int64 tensor [windows visited, sum, xor] on the GPU."""
    _check_bases(bases)
    res = torch.empty(3, dtype=torch.int64, device=bases.device)
    with torch.cuda.device(bases.device):
        check(LIB.nthash_kmer_reduce_uniform_dev(_ptr(bases), bases.numel(), n_reads, read_len, k, num_hashes, _ptr(res), _stream_ptr(stream)))
    return res


def kmer_reduce(bases, read_off, k, num_hashes=1, stream=None) -> torch.Tensor:
    """Fused consumer over ragged reads (nthash_kmer_plan_dev + nthash_kmer_reduce_dev)."""
    _check_bases(bases)
    n_reads = read_off.numel() - 1
    res = torch.empty(3, dtype=torch.int64, device=bases.device)
    with torch.cuda.device(bases.device):
        koff = torch.empty(n_reads + 1, dtype=torch.int64, device=bases.device)
        rows, max_len = C.c_uint64(0), C.c_uint64(0)
        check(LIB.nthash_kmer_plan_dev(_ptr(read_off), n_reads, k, _ptr(koff), C.byref(rows), C.byref(max_len), _stream_ptr(stream)))
        check(LIB.nthash_kmer_reduce_dev(_ptr(bases), bases.numel(), _ptr(read_off), _ptr(koff), n_reads, max_len.value, k, num_hashes,
                                         _ptr(res), _stream_ptr(stream)))
    return res


def seed_reduce_uniform(plan, bases, n_reads, read_len, stream=None) -> torch.Tensor:
    """SeedNtHash consumer (nthash_seed_reduce_uniform_dev): int64 [windows visited, sum, xor] on the GPU."""
    _check_bases(bases)
    res = torch.empty(3, dtype=torch.int64, device=bases.device)
    with torch.cuda.device(bases.device):
        check(LIB.nthash_seed_reduce_uniform_dev(plan._h, _ptr(bases), bases.numel(), n_reads, read_len, _ptr(res), _stream_ptr(stream)))
    return res


def blind_seed_roll(plan, kmers, in_base, want_strands=False, stream=None):
    """BlindSeedNtHash::roll(char_in) on n states (nthash_blind_seed_roll_batch_dev): `kmers` (uint8 [n, k], contiguous,
    updated in place) each drop their first base and take in_base[i]; returns hashes [n, n_seeds * h] (and fwd, rev [n, n_seeds])."""
    n, k = kmers.shape
    assert k == plan.k and kmers.is_contiguous() and kmers.dtype == torch.uint8
    H, m = len(plan.seeds) * plan.h, len(plan.seeds)
    dev = kmers.device
    with torch.cuda.device(dev):
        out = torch.empty((n, H), dtype=torch.int64, device=dev)
        fwd = torch.empty((n, m), dtype=torch.int64, device=dev) if want_strands else None
        rev = torch.empty((n, m), dtype=torch.int64, device=dev) if want_strands else None
        check(LIB.nthash_blind_seed_roll_batch_dev(plan._h, _ptr(kmers), kmers.numel(), _ptr(in_base), n, _ptr(out), _ptr(fwd), _ptr(rev),
                                                   _stream_ptr(stream)))
    return (out, fwd, rev) if want_strands else out


def fastq_extract(text, stream=None):
    """FASTQ text (uint8 CUDA tensor) -> (bases, read_off) in the layout of kmer_hashes / seed_hashes (nthash_fastq_extract_dev).
    `bases` keeps 64 readable bytes of slack past the last base."""
    if not (text.is_cuda and text.dtype == torch.uint8 and text.is_contiguous()):
        raise ValueError("text must be a contiguous uint8 CUDA tensor")
    nb = text.numel()
    with torch.cuda.device(text.device):
        bases = torch.empty(nb // 2 + 80, dtype=torch.uint8, device=text.device)
        read_off = torch.empty(nb // 8 + 2, dtype=torch.int64, device=text.device)   # a record is at least 8 bytes
        n_reads, n_bases = C.c_uint64(0), C.c_uint64(0)
        check(LIB.nthash_fastq_extract_dev(_ptr(text), nb, _ptr(bases), nb // 2, _ptr(read_off), nb // 8 + 1, C.byref(n_reads), C.byref(n_bases),
                                           _stream_ptr(stream)))
        bases[n_bases.value: n_bases.value + 64] = 0
    return bases[: n_bases.value], read_off[: n_reads.value + 1]


def compact(batch, stream=None):
    """HashBatch -> (hashes [n, H], dense row numbers [n]) of the windows the reference's loop visits, in its order
    (nthash_compact_rows_dev).  Synchronises to learn n."""
    if batch.valid_bits is None:
        raise ValueError("the batch was computed without a validity bitmap")
    out = batch.out
    H = out.shape[1]
    dev = out.device
    with torch.cuda.device(dev):
        comp = torch.empty_like(out)
        idx = torch.empty(batch.rows, dtype=torch.int64, device=dev)
        cnt = torch.zeros(1, dtype=torch.int64, device=dev)
        check(LIB.nthash_compact_rows_dev(_ptr(out), _ptr(batch.valid_bits), batch.rows, H, _ptr(comp), _ptr(idx), _ptr(cnt), _stream_ptr(stream)))
        n = int(cnt.item())
    return comp[:n], idx[:n]


def bloom_filter(bits, device="cuda") -> torch.Tensor:
    """An empty device-resident Bloom filter of `bits` bits (int32 words; bit b = bit b & 31 of word b >> 5)."""
    return torch.zeros((int(bits) + 31) // 32, dtype=torch.int32, device=device)


def kmer_bloom_uniform(bases, n_reads, read_len, k, num_hashes, filt, bits, query=False, stream=None) -> torch.Tensor:
    """Fused Bloom-filter consumer (nthash_kmer_bloom_uniform_dev): every visited window's num_hashes values set
    (query=False) or test (query=True) bits `hash % bits` of `filt`.  Returns int64 [windows visited, windows whose
    bits were all set already, 0] on the GPU."""
    _check_bases(bases)
    res = torch.empty(3, dtype=torch.int64, device=bases.device)
    with torch.cuda.device(bases.device):
        check(LIB.nthash_kmer_bloom_uniform_dev(_ptr(bases), bases.numel(), n_reads, read_len, k, num_hashes, _ptr(filt), int(bits),
                                                1 if query else 0, _ptr(res), _stream_ptr(stream)))
    return res


def kmer_bloom(bases, read_off, k, num_hashes, filt, bits, query=False, stream=None) -> torch.Tensor:
    """Fused Bloom-filter consumer over ragged reads (nthash_kmer_plan_dev + nthash_kmer_bloom_dev)."""
    _check_bases(bases)
    n_reads = read_off.numel() - 1
    res = torch.empty(3, dtype=torch.int64, device=bases.device)
    with torch.cuda.device(bases.device):
        koff = torch.empty(n_reads + 1, dtype=torch.int64, device=bases.device)
        rows, max_len = C.c_uint64(0), C.c_uint64(0)
        check(LIB.nthash_kmer_plan_dev(_ptr(read_off), n_reads, k, _ptr(koff), C.byref(rows), C.byref(max_len), _stream_ptr(stream)))
        check(LIB.nthash_kmer_bloom_dev(_ptr(bases), bases.numel(), _ptr(read_off), _ptr(koff), n_reads, max_len.value, k, num_hashes,
                                        _ptr(filt), int(bits), 1 if query else 0, _ptr(res), _stream_ptr(stream)))
    return res


def kmer_sketch_uniform(bases, n_reads, read_len, k, sample_bits=11, index_bits=20, counters=None, stream=None):
    """ntCard-style cardinality sketch (nthash_kmer_sketch_uniform_dev): returns (counters int32 [2^index_bits] on the GPU,
    int64 [windows visited, windows sampled, 0]).  Pass `counters` to accumulate over batches."""
    _check_bases(bases)
    dev = bases.device
    with torch.cuda.device(dev):
        if counters is None:
            counters = torch.zeros(1 << index_bits, dtype=torch.int32, device=dev)
        res = torch.empty(3, dtype=torch.int64, device=dev)
        check(LIB.nthash_kmer_sketch_uniform_dev(_ptr(bases), bases.numel(), n_reads, read_len, k, sample_bits, index_bits, _ptr(counters), _ptr(res),
                                                 _stream_ptr(stream)))
    return counters, res


def kmer_sketch(bases, read_off, k, sample_bits=11, index_bits=20, counters=None, stream=None):
    """The same over ragged reads (nthash_kmer_plan_dev + nthash_kmer_sketch_dev)."""
    _check_bases(bases)
    n_reads = read_off.numel() - 1
    dev = bases.device
    with torch.cuda.device(dev):
        if counters is None:
            counters = torch.zeros(1 << index_bits, dtype=torch.int32, device=dev)
        res = torch.empty(3, dtype=torch.int64, device=dev)
        koff = torch.empty(n_reads + 1, dtype=torch.int64, device=dev)
        rows, max_len = C.c_uint64(0), C.c_uint64(0)
        check(LIB.nthash_kmer_plan_dev(_ptr(read_off), n_reads, k, _ptr(koff), C.byref(rows), C.byref(max_len), _stream_ptr(stream)))
        check(LIB.nthash_kmer_sketch_dev(_ptr(bases), bases.numel(), _ptr(read_off), _ptr(koff), n_reads, max_len.value, k, sample_bits, index_bits,
                                         _ptr(counters), _ptr(res), _stream_ptr(stream)))
    return counters, res


def kmer_minimizers_uniform(bases, n_reads, read_len, k, window, want_lists=True, capacity=None, stream=None):
    """Minimizer selection (nthash_kmer_minimizer_uniform_dev): returns (min_bits int32 words, hashes int64 [n] or None,
    rows int64 [n] or None, count).  Synchronises to learn the count."""
    _check_bases(bases)
    dev = bases.device
    rows = n_reads * max(read_len - k + 1, 0)
    with torch.cuda.device(dev):
        bits = torch.empty(int(LIB.nthash_valid_words(rows)), dtype=torch.int32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int64, device=dev)
        if capacity is None:
            capacity = rows
        mh = torch.empty(capacity, dtype=torch.int64, device=dev) if want_lists else None
        mr = torch.empty(capacity, dtype=torch.int64, device=dev) if want_lists else None
        check(LIB.nthash_kmer_minimizer_uniform_dev(_ptr(bases), bases.numel(), n_reads, read_len, k, window, _ptr(bits), _ptr(mh), _ptr(mr),
                                                    capacity if want_lists else 0, _ptr(cnt), _stream_ptr(stream)))
        n = int(cnt.item())
    return bits, (mh[: min(n, capacity)] if want_lists else None), (mr[: min(n, capacity)] if want_lists else None), n
