"""Compile the CUDA library in-tree: deepjoin_b200/lib/libdeepjoin_b200.so (sm_100a only)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libdeepjoin_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-shared", "-Xcompiler", "-fPIC"]


def sources():
    return sorted(os.path.join(SRC, f) for f in os.listdir(SRC)) + [os.path.join(HERE, "..", "include", "deepjoin_b200.h")]


def up_to_date():
    return os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in sources())


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    # dj_abi.cu: kernels + the C ABI; host_sampler.cpp: host-only C++ (numpy's MT19937 stream), compiled by the host
    # compiler nvcc drives
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, os.path.join(SRC, "dj_abi.cu"),
                                                                     os.path.join(SRC, "host_sampler.cpp")]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
