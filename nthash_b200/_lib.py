"""ctypes binding of libnthash_b200.so — the C ABI declared in include/nthash_b200.h.

There is deliberately no fallback: if the library is missing and cannot be built, or a call
fails, this raises.  Nothing in this package touches oracle/.
"""
import ctypes as C
import os

from . import build as _build

u8p, u32p, u64p = C.c_void_p, C.c_void_p, C.c_void_p  # raw addresses (host or device)

_SIGS = {
    "nthash_fn_name": (C.c_char_p, []),
    "nthash_last_error": (C.c_char_p, []),
    "nthash_b200_abi_version": (C.c_int, []),
    "nthash_device_count": (C.c_int, []),
    "nthash_window_rows": (C.c_uint64, [u64p, C.c_uint64, C.c_uint32, u64p]),
    "nthash_valid_words": (C.c_uint64, [C.c_uint64]),
    "nthash_kmer_batch_uniform_dev": (C.c_int, [u8p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, u64p, u32p, u64p, u64p, C.c_void_p]),
    "nthash_kmer_plan_dev": (C.c_int, [u64p, C.c_uint64, C.c_uint32, u64p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_void_p]),
    "nthash_kmer_batch_dev": (C.c_int, [u8p, C.c_uint64, u64p, u64p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, u64p, u32p, u64p, u64p, C.c_void_p]),
    "nthash_ragged_plan_create": (C.c_int, [u64p, C.c_uint64, C.c_uint32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "nthash_ragged_plan_destroy": (None, [C.c_void_p]),
    "nthash_ragged_plan_rows": (C.c_uint64, [C.c_void_p]),
    "nthash_ragged_plan_max_read_len": (C.c_uint64, [C.c_void_p]),
    "nthash_ragged_plan_koff": (C.c_void_p, [C.c_void_p]),
    "nthash_kmer_batch_planned_dev": (C.c_int, [C.c_void_p, u8p, C.c_uint64, C.c_uint32, u64p, u32p, u64p, u64p, C.c_void_p]),
    "nthash_kmer_reduce_planned_dev": (C.c_int, [C.c_void_p, u8p, C.c_uint64, C.c_uint32, u64p, C.c_void_p]),
    "nthash_seed_batch_planned_dev": (C.c_int, [C.c_void_p, C.c_void_p, u8p, C.c_uint64, u64p, u32p, u64p, u64p, C.c_void_p]),
    "nthash_kmer_batch": (C.c_int, [u8p, u64p, C.c_uint64, C.c_uint32, C.c_uint32, u64p, u32p, u64p, u64p, C.c_int]),
    "nthash_kmer_reduce_uniform_dev": (C.c_int, [u8p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, u64p, C.c_void_p]),
    "nthash_kmer_reduce_dev": (C.c_int, [u8p, C.c_uint64, u64p, u64p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, u64p, C.c_void_p]),
    "nthash_kmer_reduce": (C.c_int, [u8p, u64p, C.c_uint64, C.c_uint32, C.c_uint32, u64p, C.c_int]),
    "nthash_kmer_batch_packed2bit": (C.c_int, [u8p, u32p, u64p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, u64p, u32p, C.c_int]),
    "nthash_kmer_reduce_packed2bit": (C.c_int, [u8p, u32p, u64p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, u64p, C.c_int]),
    "nthash_kmer_batch_multi": (C.c_int, [u8p, u64p, C.c_uint64, C.c_uint32, C.c_uint32, u64p, u32p, u64p, u64p, C.c_void_p, C.c_int]),
    "nthash_kmer_batch_uniform": (C.c_int, [u8p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, u64p, u32p, u64p, u64p, C.c_int]),
    "nthash_unpack2bit_dev": (C.c_int, [u8p, u32p, C.c_uint64, C.c_uint64, u8p, C.c_void_p]),
    "nthash_kmer_batch_packed2bit_uniform_dev": (C.c_int, [u8p, u32p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, u64p, u32p, C.c_void_p]),
    "nthash_seed_reduce": (C.c_int, [u8p, u64p, C.c_uint64, C.POINTER(C.c_char_p), C.c_uint32, C.c_uint32, C.c_uint32, u64p, C.c_int]),
    "nthash_seed_reduce_uniform_dev": (C.c_int, [C.c_void_p, u8p, C.c_uint64, C.c_uint64, C.c_uint32, u64p, C.c_void_p]),
    "nthash_blind_seed_roll_batch_dev": (C.c_int, [C.c_void_p, u8p, C.c_uint64, u8p, C.c_uint64, u64p, u64p, u64p, C.c_void_p]),
    "nthash_fastq_extract_dev": (C.c_int, [u8p, C.c_uint64, u8p, C.c_uint64, u64p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_void_p]),
    "nthash_compact_rows_dev": (C.c_int, [u64p, u32p, C.c_uint64, C.c_uint32, u64p, u64p, u64p, C.c_void_p]),
    "nthash_kmer_bloom_uniform_dev": (C.c_int, [u8p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64,
                                                C.c_int, u64p, C.c_void_p]),
    "nthash_kmer_bloom_dev": (C.c_int, [u8p, C.c_uint64, u64p, u64p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p,
                                        C.c_uint64, C.c_int, u64p, C.c_void_p]),
    "nthash_kmer_minimizer_uniform_dev": (C.c_int, [u8p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, u32p, u64p, u64p, C.c_uint64, u64p, C.c_void_p]),
    "nthash_kmer_minimizers": (C.c_int, [u8p, u64p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, u32p, u64p, u64p, C.c_uint64, u64p, C.c_int]),
    "nthash_kmer_sketch_uniform_dev": (C.c_int, [u8p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, u64p, C.c_void_p]),
    "nthash_kmer_sketch_dev": (C.c_int, [u8p, C.c_uint64, u64p, u64p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, u64p, C.c_void_p]),
    "nthash_seed_plan_create": (C.c_int, [C.POINTER(C.c_char_p), C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]),
    "nthash_seed_plan_destroy": (None, [C.c_void_p]),
    "nthash_seed_plan_symmetric": (C.c_int, [C.c_void_p]),
    "nthash_seed_plan_kernel_note": (C.c_char_p, [C.c_void_p]),
    "nthash_seed_jit_selftest": (C.c_int, [C.POINTER(C.c_char_p), C.c_uint32, C.c_uint32, C.c_uint32]),
    "nthash_seed_batch_uniform_dev": (C.c_int, [C.c_void_p, u8p, C.c_uint64, C.c_uint64, C.c_uint32, u64p, u32p, u64p, u64p, C.c_void_p]),
    "nthash_seed_batch_dev": (C.c_int, [C.c_void_p, u8p, C.c_uint64, u64p, u64p, C.c_uint64, C.c_uint64, u64p, u32p, u64p, u64p, C.c_void_p]),
    "nthash_seed_batch": (C.c_int, [u8p, u64p, C.c_uint64, C.POINTER(C.c_char_p), C.c_uint32, C.c_uint32, C.c_uint32, u64p, u32p, u64p, u64p, C.c_int]),
    "nthash_blind_roll_batch_dev": (C.c_int, [u64p, u64p, u8p, u8p, C.c_uint64, C.c_uint32, C.c_uint32, u64p, C.c_void_p]),
    "nthash_blind_peek4_batch_dev": (C.c_int, [u64p, u64p, u8p, C.c_uint64, C.c_uint32, C.c_uint32, u64p, C.c_void_p]),
    "nthash_blind_roll_batch": (C.c_int, [u64p, u64p, u8p, u8p, C.c_uint64, C.c_uint32, C.c_uint32, u64p, C.c_int]),
}


class NtHashError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"nthash_b200 error {code}: {msg}")
        self.code = code


def _load():
    path = _build.LIB
    if not os.path.exists(path) or (_build.stale() and os.path.exists(_build.NVCC)):
        path = _build.build()  # raises if nvcc is missing or compilation fails
    lib = C.CDLL(path)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


LIB = _load()
LIB_PATH = _build.LIB


def check(rc):
    if rc != 0:
        raise NtHashError(rc, LIB.nthash_last_error().decode())
