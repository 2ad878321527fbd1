"""BA evaluate time against the number of observations (one GPU, L2 flushed, CUDA events):
separates the fixed cost (launch, prologue, ramp, tail, cost reduction) from the streaming rate."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pybot_b200 import ops, synthetic as syn  # noqa: E402

rv, tv, K, pts, oc, op, uv = syn.c4_problem()
fa = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
fb = torch.ones(256 << 20, dtype=torch.uint8, device="cuda")
for n in (1000, 62_500, 125_000, 250_000, 500_000, 1_000_000, 2_000_000, 4_000_000):
    p = ops.BAProblemDevice(rv, tv, K, pts, oc[:n], op[:n], uv[:n])
    for mode in ("one_kernel", "two_kernels"):
        ts = []
        for i in range(25):
            fa.fill_(1)
            fb.sum()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            p.evaluate(one_kernel=(mode == "one_kernel"))
            b.record()
            torch.cuda.synchronize()
            if i >= 5:
                ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        print("%9d obs %-11s median %7.2f us  min %7.2f us   (HBM time at 6552 GB/s: %6.2f us)" % (
            n, mode, ts[len(ts) // 2], ts[0], n * 96 / 6552e3))
