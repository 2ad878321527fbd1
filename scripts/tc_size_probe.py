"""Tensor-core Hamming scan against shard size (one GPU): fixed cost vs marginal rate.
    python scripts/tc_size_probe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pybot_b200 import ops  # noqa: E402

REPS = int(os.environ.get("REPS", 10))       # calls per event pair: the host runs ahead of the device
g = torch.Generator(device="cuda").manual_seed(1)
q = torch.randint(0, 256, (2000, 32), device="cuda", dtype=torch.uint8, generator=g)
prev = None
for nt in (160_000, 320_000, 625_000, 1_250_000, 2_500_000, 5_000_000, 10_000_000, 20_000_000):
    t = torch.randint(0, 256, (nt, 32), device="cuda", dtype=torch.uint8, generator=g)
    signs = ops.hamming_tc_expand(t)
    del t
    ops.hamming_knn2_keys_tc(q, signs, nt, 0)
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(REPS):
            ops.hamming_knn2_keys_tc(q, signs, nt, 0)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / REPS)
    ts.sort()
    ms = ts[len(ts) // 2]
    extra = "" if prev is None else "   marginal %.1f Gpairs/s" % (2000 * (nt - prev[0]) / (ms - prev[1]) / 1e6)
    print("%9d rows %8.3f ms  %8.1f Gpairs/s%s" % (nt, ms, 2000 * nt / ms / 1e6, extra))
    prev = (nt, ms)
    del signs
