"""BA streaming-kernel probe: kernel-only time (CUDA events, L2 flushed) for output subsets.

    python scripts/ba_probe.py [reps]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pybot_b200 import ops, synthetic as syn  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
prob = syn.c4_problem()
fa = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
fb = torch.ones(256 << 20, dtype=torch.uint8, device="cuda")
for want in (("r", "Jc", "Jp"), (), ("Jc",), ("r", "Jp"), ("r",)):
    p = ops.BAProblemDevice(*prob, want=want)
    p.evaluate()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        fa.fill_(1)
        fb.sum()
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        p.evaluate(events=ev)
        torch.cuda.synchronize()
        ts.append(ev[0].elapsed_time(ev[1]))
    ts.sort()
    nbytes = p.n_obs * (16 + 8 * ("r" in want) + 48 * ("Jc" in want) + 24 * ("Jp" in want))
    print("want=%-18s kernel %.4f ms median  %.4f min   %.0f GB/s" % (
        ",".join(want) or "-", ts[len(ts) // 2], ts[0], nbytes / ts[len(ts) // 2] / 1e6))
    del p
