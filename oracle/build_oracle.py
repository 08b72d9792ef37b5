"""Build the C part of the oracle (TEST INFRASTRUCTURE ONLY): oracle/_build/libmc_oracle.so.
The reference has no C/C++ sources of its own, so there is no oracle/_ref build."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libmc_oracle.so")


def build(force=False):
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(HERE, "mc_oracle.c")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return LIB
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB, srcs[0], "-lm"])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
