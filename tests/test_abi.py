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
    assert nthash_b200.LIB.nthash_b200_abi_version() >= 1


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


def test_specialised_seed_kernel_compiles_without_a_gpu():
    # NVRTC generates sm_100a code for a seed set; no device needed (what __graft_entry__.build() also checks)
    import nthash_b200
    seeds = ["1010101010101010101010101010101", "1101101101101101011011011011011"]
    arr = (C.c_char_p * 2)(*[s.encode() for s in seeds])
    rc = nthash_b200.LIB.nthash_seed_jit_selftest(arr, 2, 31, 3)
    assert rc == 0, nthash_b200.LIB.nthash_last_error()
    arr = (C.c_char_p * 1)(b"1101")
    assert nthash_b200.LIB.nthash_seed_jit_selftest(arr, 1, 5, 1) == -1  # seed length != k (reference seed.cpp:90-95)
