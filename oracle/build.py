"""Builds the oracle's C restatement into oracle/_build/ (git-ignored; the .so
travels to the GPU box with the gpurun snapshot).  TEST INFRASTRUCTURE.

The reference itself is pure Python on this path: there is no
C/C++ reference source to compile into oracle/_ref/.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libpbo.so")


def build(force=False):
    src = os.path.join(HERE, "c", "hamming_knn.c")
    os.makedirs(OUT, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O3", "-fopenmp", "-mpopcnt", "-shared", "-fPIC",
                               "-o", LIB, src])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
