"""Scan time against tiles per split (one GPU, 2000 queries, 37 splits x 4 query groups): how much
the exact-path-heavy start of every split costs.   python scripts/tc_tiles_probe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pybot_b200 import ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(1)
q = torch.randint(0, 256, (2000, 32), device="cuda", dtype=torch.uint8, generator=g)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts) * 1e3


print("tiles/split  rows      full us   no-selection us   difference us")
for tiles in (1, 2, 4, 8, 16, 32, 64, 128, 256, 528):
    nt = 37 * 128 * tiles
    t = torch.randint(0, 256, (nt, 32), device="cuda", dtype=torch.uint8, generator=g)
    signs = ops.hamming_tc_expand(t)
    full = timed(lambda: ops.hamming_knn2_keys_tc(q, signs, nt, 0))
    os.environ["PBK_TC_DEBUG"] = "1"
    nosel = timed(lambda: ops.hamming_knn2_keys_tc(q, signs, nt, 0))
    del os.environ["PBK_TC_DEBUG"]
    print("%6d  %9d  %9.1f  %14.1f  %14.1f" % (tiles, nt, full, nosel, full - nosel), flush=True)
