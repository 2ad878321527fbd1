"""Builds pybot_b200/libpybot_b200.so in-tree with nvcc for sm_100a.

    python -m pybot_b200.build [--force] [--verbose]

Cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU
box with the gpurun snapshot.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libpybot_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-diag-suppress", "128",
    "-cudart", "static",
]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libpybot_b200.so")
    return nvcc


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    hdrs.append(os.path.join(INCLUDE, "pybot_b200.h"))
    hdrs.append(os.path.abspath(__file__))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, obj, verbose):
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", INCLUDE, "-c", src, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return r.stderr if verbose else ""


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    dep = _deps_mtime()
    jobs = []
    objs = []
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        stale = (force or not os.path.exists(obj)
                 or os.path.getmtime(obj) < max(os.path.getmtime(src), dep))
        if stale:
            jobs.append((src, obj))
    logs = []
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
            for log in ex.map(lambda a: _compile(a[0], a[1], verbose), jobs):
                logs.append(log)
    if jobs or not os.path.exists(LIB):
        cmd = [_nvcc(), "-shared", "-cudart", "static",
               "-gencode", "arch=compute_100a,code=sm_100a",
               "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB, "\n".join(logs)


if __name__ == "__main__":
    lib, log = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    if log:
        print(log)
    print(lib)
