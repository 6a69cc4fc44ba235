"""Builds nthash_b200/libnthash_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc."""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnthash_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "-ldl", "--threads", "0",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.hpp"))
    deps.append(os.path.join(os.path.dirname(HERE), "include", "nthash_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library.  nvcc cross-compiles without a GPU."""
    if not force and not stale():
        return LIB
    # NTHASH_B200_NVCC_FLAGS / NTHASH_B200_LIB_OUT: experiments (e.g. an A/B build with -DNTH_ROLL_V1 next to the product library)
    extra = os.environ.get("NTHASH_B200_NVCC_FLAGS", "").split()
    cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", os.environ.get("NTHASH_B200_LIB_OUT", LIB)] + sources()
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
