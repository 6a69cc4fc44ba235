"""CPU-side checks of the drop-in boundary: the shared library loads and exports exactly what
include/nthash_b200.h declares, and the host-only helpers compute the dense row layout."""
import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "nthash_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nthash_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import nthash_b200
    lib = C.CDLL(nthash_b200.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 8
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/nthash_b200.h but not exported"
    # and the Python binding covers all of them
    from nthash_b200 import _lib
    assert sorted(_lib._SIGS) == names


def test_fn_name_and_version():
    import nthash_b200
    assert nthash_b200.FN_NAME == "ntHash_v2"  # reference include/nthash/nthash.hpp:18
    assert nthash_b200.LIB.nthash_b200_abi_version() >= 2


def test_window_rows_helper():
    import nthash_b200
    lens = np.array([0, 5, 30, 31, 32, 150, 7], np.uint64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    koff = np.zeros(len(lens) + 1, np.uint64)
    tot = nthash_b200.LIB.nthash_window_rows(off.ctypes.data, len(lens), 31, koff.ctypes.data)
    assert tot == 0 + 0 + 0 + 1 + 2 + 120 + 0
    assert list(koff) == [0, 0, 0, 0, 1, 3, 123, 123]
    assert nthash_b200.LIB.nthash_valid_words(0) == 0 and nthash_b200.LIB.nthash_valid_words(33) == 2


def test_invalid_arguments_are_reported_not_fatal():
    # the reference exits the process on these (src/kmer.cpp:212-225); the C ABI returns codes
    import nthash_b200
    L = nthash_b200.LIB
    rc = L.nthash_kmer_batch(None, None, 1, 2, 1, None, None, None, None, 0)
    assert rc == -1 and b"k=2" in L.nthash_last_error()
    rc = L.nthash_kmer_batch(None, None, 1, 31, 0, None, None, None, None, 0)
    assert rc == -1 and b"num_hashes" in L.nthash_last_error()
    rc = L.nthash_kmer_batch(None, None, 1, 31, 1, None, None, None, None, 0)
    assert rc == -1


def test_new_entry_points_validate_before_touching_a_device():
    # every entry point added for the rows around the path (consumers, packed input, FASTQ staging, compaction, multi-GPU,
    # BlindSeed) rejects bad arguments with a code and a message - no exit(), no crash, no GPU needed to find out
    import nthash_b200
    L = nthash_b200.LIB
    cnt = np.zeros(3, np.uint64)
    n64 = C.c_uint64(0)
    cases = [
        (L.nthash_kmer_batch_uniform, (None, 1, 150, 2, 1, None, None, None, None, 0), b"k=2"),
        (L.nthash_kmer_batch_multi, (None, None, 1, 31, 1, None, None, None, None, None, 0), b"device"),
        (L.nthash_kmer_batch_packed2bit, (None, None, None, 1, 0, 31, 1, cnt.ctypes.data, None, 0), b"packed"),
        (L.nthash_kmer_reduce_packed2bit, (cnt.ctypes.data, None, None, 1, 0, 31, 1, cnt.ctypes.data, 0), b"uniform_read_len"),
        (L.nthash_kmer_reduce_packed2bit, (cnt.ctypes.data, None, None, 1, 150, 31, 1, None, 0), b"result"),
        (L.nthash_kmer_bloom_uniform_dev, (None, 0, 1, 150, 31, 3, None, 0, 0, cnt.ctypes.data, None), b"filter"),
        (L.nthash_kmer_bloom_uniform_dev, (None, 0, 1, 150, 31, 0, cnt.ctypes.data, 64, 0, cnt.ctypes.data, None), b"num_hashes"),
        (L.nthash_compact_rows_dev, (None, None, 10, 1, None, None, None, None), b"d_count"),
        (L.nthash_fastq_extract_dev, (None, 10, None, 0, None, 0, C.byref(n64), C.byref(n64), None), b"d_read_off"),
        (L.nthash_seed_reduce_uniform_dev, (None, None, 0, 1, 150, cnt.ctypes.data, None), b"plan"),
        (L.nthash_blind_seed_roll_batch_dev, (None, None, 0, None, 1, None, None, None, None), b"plan"),
        (L.nthash_unpack2bit_dev, (None, None, 0, 16, None, None), b"d_packed"),
    ]
    for fn, args, needle in cases:
        rc = fn(*args)
        assert rc < 0, (fn.__name__, rc)
        assert needle in L.nthash_last_error(), (fn.__name__, L.nthash_last_error())
    # zero-sized requests are fine and touch nothing
    assert L.nthash_kmer_batch_uniform(None, 0, 150, 31, 1, cnt.ctypes.data, None, None, None, 0) == 0
    assert L.nthash_kmer_reduce_packed2bit(None, None, None, 0, 150, 31, 1, cnt.ctypes.data, 0) == 0


def test_specialised_seed_kernel_compiles_without_a_gpu():
    # NVRTC generates sm_100a code for a seed set; no device needed (what __graft_entry__.build() also checks)
    import nthash_b200
    seeds = ["1010101010101010101010101010101", "1101101101101101011011011011011"]
    arr = (C.c_char_p * 2)(*[s.encode() for s in seeds])
    rc = nthash_b200.LIB.nthash_seed_jit_selftest(arr, 2, 31, 3)
    assert rc == 0, nthash_b200.LIB.nthash_last_error()
    arr = (C.c_char_p * 1)(b"1101")
    assert nthash_b200.LIB.nthash_seed_jit_selftest(arr, 1, 5, 1) == -1  # seed length != k (reference seed.cpp:90-95)
