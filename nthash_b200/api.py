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


def kmer_hashes(bases, read_off, k, num_hashes=1, want_valid=True, want_strands=False, stream=None) -> HashBatch:
    """NtHash over ragged reads: read r is bases[read_off[r]:read_off[r+1]] (nthash_kmer_plan_dev + nthash_kmer_batch_dev)."""
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
        out = torch.empty((rows, num_hashes), dtype=torch.int64, device=dev)
        valid_bits = torch.empty(int(LIB.nthash_valid_words(rows)), dtype=torch.int32, device=dev) if want_valid else None
        fwd = torch.empty(rows, dtype=torch.int64, device=dev) if want_strands else None
        rev = torch.empty(rows, dtype=torch.int64, device=dev) if want_strands else None
        check(LIB.nthash_kmer_batch_dev(_ptr(bases), bases.numel(), _ptr(read_off), _ptr(koff), n_reads, max_len.value, k,
                                        num_hashes, _ptr(out), _ptr(valid_bits), _ptr(fwd), _ptr(rev), _stream_ptr(stream)))
    return HashBatch(out, valid_bits, koff, rows, fwd, rev)
