"""Host-side cost of BFMatcher.knnMatch (cProfile) on a database small enough that the GPU is not
the bottleneck: python scripts/knnmatch_host_profile.py"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pybot_b200 import _ffi  # noqa: E402
from pybot_b200.vision.matcher import BFMatcher  # noqa: E402

per_kf, nkf = 2000, 50
g = torch.Generator(device="cuda").manual_seed(1)
flat = torch.randint(0, 256, (per_kf * nkf, 32), device="cuda", dtype=torch.uint8, generator=g)
q = _ffi.pinned_empty((2000, 32), np.uint8)
q[:] = np.random.default_rng(0).integers(0, 256, (2000, 32), dtype=np.uint8)
bf = BFMatcher.from_resident(flat, per_kf)
for _ in range(5):
    bf.knnMatch(q, k=2)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    bf.knnMatch(q, k=2)
print("knnMatch: %.1f us per call (100 k-row database)" % ((time.perf_counter() - t0) / 50 * 1e6))
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    bf.knnMatch(q, k=2)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
