"""Builds A/B variants of libpybot_b200.so into scripts/probes/_bin/ (git-ignored, ships with gpurun):

    python scripts/build_variants.py NAME=-DFLAG=1,-DOTHER=2 [NAME2=...]

Select one at run time with PYBOT_B200_LIB=scripts/probes/_bin/libpbk_NAME.so.
"""
import concurrent.futures
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pybot_b200 import build as B  # noqa: E402

OUT = os.path.join(ROOT, "scripts", "probes", "_bin")


def one(spec):
    name, _, flags = spec.partition("=")
    flags = [f for f in flags.split(",") if f]
    d = os.path.join(OUT, "obj_" + name)
    os.makedirs(d, exist_ok=True)
    objs = []
    for src in B.sources():
        obj = os.path.join(d, os.path.basename(src)[:-3] + ".o")
        r = subprocess.run([B._nvcc()] + B.NVCC_FLAGS + flags + ["-I", B.INCLUDE, "-c", src, "-o", obj],
                           capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(r.stderr)
        objs.append(obj)
    lib = os.path.join(OUT, "libpbk_%s.so" % name)
    r = subprocess.run([B._nvcc(), "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
                        "-o", lib] + objs, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError(r.stderr)
    return lib


if __name__ == "__main__":
    with concurrent.futures.ThreadPoolExecutor(max_workers=6) as ex:
        for lib in ex.map(one, sys.argv[1:]):
            print(lib)
