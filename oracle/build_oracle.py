"""Compile the oracle's C restatement with gcc into oracle/_build/ (checker only; never shipped)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libgrid_oracle.so")


def build(force=False):
    src = os.path.join(HERE, "grid_oracle.c")
    os.makedirs(OUT, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-fvisibility=hidden", "-ffp-contract=off", "-o", LIB,
                               src, "-lm"])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
