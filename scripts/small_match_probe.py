"""Per-frame matcher latency (C1: 2000 x 2000, CUDA-graph replay) and small-database knn2."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybot_b200 import ops, synthetic as syn  # noqa: E402

q, t = syn.orb_pair()
dq, dt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


pm = ops.PreparedMatcher(2000, 2000, 0.75)
print("prepared matcher 2000x2000 fwd+rev+ratio (graph replay): %.1f us per frame" % timed(lambda: pm.run(dq, dt), 200))
for n in (2000, 8000, 20000, 100000):
    db = torch.randint(0, 256, (n, 32), device="cuda", dtype=torch.uint8)
    pmn = ops.PreparedMatcher(2000, n, 0.75, mutual=False)
    print("%7d rows forward only (graph replay): %.1f us" % (n, timed(lambda: pmn.run(dq, db), 100)))
