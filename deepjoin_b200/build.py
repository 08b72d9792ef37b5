"""Compile the CUDA library in-tree: deepjoin_b200/lib/libdeepjoin_b200.so (sm_100a only)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libdeepjoin_b200.so")
LIB_DEBUG = os.path.join(HERE, "lib", "libdeepjoin_b200_debug.so")  # -DDJ_DEBUG: + the diagnostics tools/ use
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-shared", "-Xcompiler", "-fPIC"]


def sources():
    return sorted(os.path.join(SRC, f) for f in os.listdir(SRC)) + [os.path.join(HERE, "..", "include", "deepjoin_b200.h")]


def up_to_date(lib=LIB):
    return os.path.exists(lib) and all(os.path.getmtime(lib) >= os.path.getmtime(s) for s in sources())


def build(force=False, verbose=False, debug=False):
    """The product library; debug=True builds the diagnostics variant next to it (dj_debug_* entry points, the
    DEEPJOIN_B200_DEBUG knobs and cycle counters of the forward kernel), which tools/ load through
    DEEPJOIN_B200_LIB -- it never replaces the product library."""
    lib = LIB_DEBUG if debug else LIB
    if not force and up_to_date(lib):
        return lib
    os.makedirs(os.path.dirname(lib), exist_ok=True)
    # dj_abi.cu: kernels + the C ABI; host_sampler.cpp: host-only C++ (numpy's MT19937 stream), compiled by the host
    # compiler nvcc drives
    cmd = [NVCC] + FLAGS + (["-DDJ_DEBUG"] if debug else []) + (["-Xptxas", "-v"] if verbose else []) + [
        "-o", lib, os.path.join(SRC, "dj_abi.cu"), os.path.join(SRC, "host_sampler.cpp")]
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--debug" in sys.argv))
