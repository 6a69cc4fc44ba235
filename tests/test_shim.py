"""The drop-in C++ header (include/nthash/nthash.hpp) compiled against libnthash_b200.so.

CPU side: it compiles warning-free as plain C++17 and the caller-fed (Blind*) classes reproduce the
reference's golden values.  GPU side (-m gpu): every scenario of the reference's own tests/tests.cpp
that touches NtHash / SeedNtHash, now served by the CUDA kernels through the C ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "shim_tests")


def build_shim_tests():
    import nthash_b200  # builds the library if needed
    libdir = os.path.dirname(nthash_b200.LIB_PATH)
    src = os.path.join(ROOT, "tests", "cpp", "shim_tests.cpp")
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < max(os.path.getmtime(src), os.path.getmtime(nthash_b200.LIB_PATH),
                                                               os.path.getmtime(os.path.join(ROOT, "include", "nthash", "nthash.hpp"))):
        subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), src,
                        "-L" + libdir, "-lnthash_b200", "-Wl,-rpath," + libdir, "-o", BIN], check=True)
    return BIN


def test_header_compiles_and_host_side_classes_match_goldens():
    out = subprocess.run([build_shim_tests(), "--host-only"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


def test_constructor_errors_exit_like_the_reference():
    # raise_error(): message on stderr, exit status 1 (reference src/internal.hpp:16-22, src/kmer.cpp:212-225)
    build_shim_tests()
    src = os.path.join(ROOT, "tests", "cpp", "_die.cpp")
    exe = os.path.join(ROOT, "tests", "cpp", "_die")
    open(src, "w").write('#include <nthash/nthash.hpp>\nint main(){ nthash::NtHash h("ACGT", 4, 1, 5); return 0; }\n')
    import nthash_b200
    libdir = os.path.dirname(nthash_b200.LIB_PATH)
    subprocess.run(["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"), src, "-L" + libdir, "-lnthash_b200",
                    "-Wl,-rpath," + libdir, "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    os.remove(src); os.remove(exe)
    assert out.returncode == 1 and "smaller than k" in out.stderr and "[ntHash::NtHash]" in out.stderr


@pytest.mark.gpu
def test_reference_test_scenarios_through_the_gpu():
    out = subprocess.run([build_shim_tests()], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all blocks: 0 failure(s)" in out.stdout


@pytest.mark.gpu
def test_unmodified_reference_tests_pass_against_the_shim():
    """oracle/_ref/ref_tests_shim is the reference's own tests/tests.cpp, untouched, compiled against the drop-in header
    and linked with the CUDA engine (oracle/Makefile builds it where /root/reference exists; the binary travels)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_tests_shim")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_tests_shim not built (needs the reference tree at build time)")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert (out.stdout + out.stderr).count("Testing") == 17
