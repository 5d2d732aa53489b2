"""Builds the C oracle (test infrastructure) into oracle/liblattice_oracle.so.

The reference (/root/reference) holds no native sources for this path -- the lattice arithmetic
lives in an un-vendored dependency -- so there is nothing to compile into
oracle/_ref; DESIGN.md records that.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "lattice_oracle.c")
OUT = os.path.join(HERE, "liblattice_oracle.so")


def build(force: bool = False) -> str:
    if (not force) and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = ["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
           "-o", OUT, SRC, "-lm"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
