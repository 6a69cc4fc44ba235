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
        subprocess.run(["g++", "-std=c++17", "-O2", "-pthread", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), src,
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


REF_DIR = os.path.join(ROOT, "oracle", "_ref")
GPU_ENV = dict(os.environ, NTHASH_B200_HOST_CUTOFF="0")  # every sequence, however short, goes through the CUDA engine


def _class_bench(which, args, env=None, dirty=None):
    import json
    exe = os.path.join(REF_DIR, "class_bench_" + which)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/class_bench_* not built (needs the reference tree at build time)")
    e = dict(env or os.environ)
    if dirty:
        e["NTHASH_BENCH_DIRTY"] = str(dirty)
    out = subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True, timeout=600, env=e)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_short_sequences_are_rolled_on_the_host_like_the_reference():
    """Sequences of at most NTHASH_B200_HOST_CUTOFF windows never leave the host (one object per 150 bp read must not cost
    a kernel launch): every scenario of the restated reference tests passes without a GPU, and so does the reference's
    own unmodified tests/tests.cpp compiled against the drop-in header."""
    out = subprocess.run([build_shim_tests(), "--short-only"], capture_output=True, text=True)
    assert out.returncode == 0 and "all blocks: 0 failure(s)" in out.stdout, out.stdout + out.stderr
    exe = os.path.join(REF_DIR, "ref_tests_shim")
    if os.path.exists(exe):
        out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        assert (out.stdout + out.stderr).count("Testing") == 17


@pytest.mark.parametrize("args,dirty", [(("kmer", 20000, 150, 31, 1), None), (("kmer", 20000, 150, 31, 3), 20000),
                                        (("kmer", 5000, 100, 64, 3), 5000), (("seed", 8000, 150, 3), None),
                                        (("seed", 8000, 150, 2), 20000), (("kmer", 300, 400, 200, 2), 3000)])
def test_user_loop_on_short_reads_matches_the_compiled_reference(args, dirty):
    """`Hash h(read, ...); while (h.roll()) ...` over many short reads: same visited windows, hashes and strand hashes as the
    unmodified reference compiled from the same file (tests/cpp/class_bench.cpp), clean and dirty reads."""
    ref, shim = _class_bench("ref", args, dirty=dirty), _class_bench("shim", args, dirty=dirty)
    assert (ref["windows"], ref["sum"], ref["xor"]) == (shim["windows"], shim["sum"], shim["xor"])


@pytest.mark.gpu
def test_reference_test_scenarios_through_the_gpu():
    out = subprocess.run([build_shim_tests()], capture_output=True, text=True, env=GPU_ENV)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all blocks: 0 failure(s)" in out.stdout


@pytest.mark.gpu
def test_unmodified_reference_tests_pass_against_the_shim():
    """oracle/_ref/ref_tests_shim is the reference's own tests/tests.cpp, untouched, compiled against the drop-in header
    and linked with the CUDA engine (oracle/Makefile builds it where /root/reference exists; the binary travels)."""
    exe = os.path.join(REF_DIR, "ref_tests_shim")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_tests_shim not built (needs the reference tree at build time)")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=GPU_ENV)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert (out.stdout + out.stderr).count("Testing") == 17


@pytest.mark.gpu
@pytest.mark.parametrize("args,dirty", [(("kmer", 3000, 150, 31, 2), 10000), (("seed", 2000, 150, 2), 10000)])
def test_user_loop_on_short_reads_through_the_gpu(args, dirty):
    ref, shim = _class_bench("ref", args, dirty=dirty), _class_bench("shim", args, env=GPU_ENV, dirty=dirty)
    assert (ref["windows"], ref["sum"], ref["xor"]) == (shim["windows"], shim["sum"], shim["xor"])


@pytest.mark.gpu
@pytest.mark.parametrize("args,dirty", [(("kmer", 2, 30_000_000, 31, 1), None), (("kmer", 3, 5_000_000, 63, 2), 300),
                                        (("seed", 2, 8_000_000, 3), None), (("seed", 3, 3_000_000, 2), 300)])
def test_user_loop_on_long_sequences_is_chunked_and_bit_exact(args, dirty):
    """Chromosome-sized sequences go through the CUDA engine in bounded chunks: same checksums as the compiled reference,
    and the resident set stays within a few hundred MB of the sequence itself whatever its length."""
    ref, shim = _class_bench("ref", args, dirty=dirty), _class_bench("shim", args, dirty=dirty)
    assert (ref["windows"], ref["sum"], ref["xor"]) == (shim["windows"], shim["sum"], shim["xor"])
    seq_mb = args[1] * args[2] / 2**20
    assert shim["max_rss_mb"] < seq_mb + 1500, shim  # CUDA context + pinned staging + one chunk of windows, not O(sequence) hashes
