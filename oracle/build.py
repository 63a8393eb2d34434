"""Build the C restatement of the oracle (oracle/libdmp_oracle.so; TEST INFRASTRUCTURE ONLY).
The reference itself is pure Python, so there is nothing to compile into oracle/_ref/."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "dmp_oracle.c")
LIB = os.path.join(HERE, "libdmp_oracle.so")


def build_oracle(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-std=c11", "-o", LIB, SRC, "-lm"])
    return LIB


if __name__ == "__main__":
    print(build_oracle(force=True))
