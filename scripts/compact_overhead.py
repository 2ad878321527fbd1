"""Projection kernels (compacting, masks, plain): time vs size (fixed launch/memset/ramp overhead vs per-point cost).
    python scripts/compact_overhead.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pybot_b200 import ops, synthetic as syn  # noqa: E402
from pybot_b200.geometry import RigidTransform  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(1)
T = RigidTransform.from_rpyxyz(*syn.C2_RPYXYZ).to_matrix()
fa = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
fb = torch.ones(256 << 20, dtype=torch.uint8, device="cuda")
for n in (1024, 100_000, 1_000_000, 2_500_000, 5_000_000, 10_000_000, 20_000_000):
    X = torch.rand((n, 3), device="cuda", generator=g)
    X[:, 0] = X[:, 0] * 80 - 40
    X[:, 1] = X[:, 1] * 6 - 3
    X[:, 2] = X[:, 2] * 78 + 2
    for name, kw in (("compact", {}), ("masks", {"compact": False}), ("plain", None)):
        if kw is None:
            pr = ops.Projector(T[:3, :3], T[:3, 3], syn.kitti_K())
        else:
            pr = ops.Projector(T[:3, :3], T[:3, 3], syn.kitti_K(), shape=syn.KITTI_SHAPE, check_bounds=True,
                               check_depth=True, return_depth=True, return_valid=True, **kw)
        pr.run(X)
        torch.cuda.synchronize()
        ts = []
        for _ in range(15):
            fa.fill_(1)
            fb.sum()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            pr.run(X)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        print("n=%9d %-8s %.4f ms median %.4f min" % (n, name, ts[len(ts) // 2], ts[0]))
    del X
