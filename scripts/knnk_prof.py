"""hamming_knnk_kernel at the bench's size (2000 x 20 000, k = 8), without / with a 25 % mask;
standalone or under ncu (profiles/r02_knnk_ncu.md)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybot_b200 import ops, synthetic as syn  # noqa: E402

q, t = syn.orb_pair(seed=1007, nq=2000, nt=20_000)
mask = (np.random.default_rng(1007).random((2000, 20_000)) < 0.25).astype(np.uint8)
qd, td, md = (torch.from_numpy(x).cuda() for x in (q, t, mask))
for name, m in (("no mask", None), ("25 % mask", md)):
    for _ in range(2):
        ops.hamming_knnk(qd, td, 8, m)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        ops.hamming_knnk(qd, td, 8, m)
    b.record()
    torch.cuda.synchronize()
    print("%-10s %.1f us per call (L2-warm)" % (name, a.elapsed_time(b) / 5 * 1e3))
