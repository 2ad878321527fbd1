"""Launches each hot kernel at its BASELINE.json size with device-resident
inputs; meant to run under ncu (see profiles/README.md) or standalone.

    python scripts/prof_kernels.py [hamming|project|compact|transform|ba|reproject|depth|flow|all] [reps]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pybot_b200 import ops, synthetic as syn  # noqa: E402
from pybot_b200.geometry import RigidTransform  # noqa: E402
from pybot_b200.vision.camera_utils import StereoCamera  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = torch.Generator(device="cuda").manual_seed(1)


FLUSH = os.environ.get("PROF_FLUSH", "0") == "1"
_fa = _fb = None


def timed(name, fn):
    """PROF_FLUSH=1: L2 flushed between reps (as bench.py does), per-rep events, median + min."""
    global _fa, _fb
    fn()
    torch.cuda.synchronize()
    if FLUSH:
        if _fa is None:
            _fa = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
            _fb = torch.ones(256 << 20, dtype=torch.uint8, device="cuda")
        ts = []
        for _ in range(reps):
            _fa.fill_(1)
            _fb.sum()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        print("%-12s %.4f ms median  %.4f min (flushed, %d reps)" % (name, ts[len(ts) // 2], ts[0], reps))
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    print("%-12s %.3f ms/iter" % (name, a.elapsed_time(b) / reps))


if which in ("hamming", "all"):
    nt = int(os.environ.get("PROF_NT", 2_500_000))       # one 8-GPU shard of C3
    q = torch.randint(0, 256, (2000, 32), device="cuda", dtype=torch.uint8, generator=g)
    db = torch.randint(0, 256, (nt, 32), device="cuda", dtype=torch.uint8, generator=g)
    timed("hamming", lambda: ops.hamming_knn2_keys(q, db, 0))
    signs = ops.hamming_tc_expand(db)
    timed("hamming_tc (format %d)" % ops.hamming_tc_format(), lambda: ops.hamming_knn2_keys_tc(q, signs, nt, 0))
    other = 8 if ops.hamming_tc_format() == 4 else 4
    ops.hamming_tc_set_format(other)
    signs_o = ops.hamming_tc_expand(db)
    timed("hamming_tc (format %d)" % other, lambda: ops.hamming_knn2_keys_tc(q, signs_o, nt, 0))
    ops.hamming_tc_set_format(12 - other)
    del db, signs

if which in ("project", "compact", "transform", "all"):
    n = 10_000_000
    X = torch.rand((n, 3), device="cuda", generator=g)
    X[:, 0] = X[:, 0] * 80 - 40
    X[:, 1] = X[:, 1] * 6 - 3
    X[:, 2] = X[:, 2] * 78 + 2
    T = RigidTransform.from_rpyxyz(*syn.C2_RPYXYZ).to_matrix()
    K = syn.kitti_K()
    if which in ("project", "all"):
        pr = ops.Projector(T[:3, :3], T[:3, 3], K)
        timed("project", lambda: pr.run(X))
    if which in ("compact", "all"):
        pc = ops.Projector(T[:3, :3], T[:3, 3], K, shape=syn.KITTI_SHAPE, check_bounds=True,
                           check_depth=True, return_depth=True, return_valid=True)
        timed("compact", lambda: pc.run(X))
        pm = ops.Projector(T[:3, :3], T[:3, 3], K, shape=syn.KITTI_SHAPE, check_bounds=True,
                           check_depth=True, return_depth=True, return_valid=True, compact=False)
        timed("masks", lambda: pm.run(X))
    if which in ("transform", "all"):
        Y = torch.empty((n, 3), dtype=torch.float64, device="cuda")
        timed("transform", lambda: ops.transform_points(X, T[:3, :3], T[:3, 3], out=Y))
    del X

if which in ("ba", "all"):
    prob = syn.c4_problem()
    p = ops.BAProblemDevice(*prob)
    timed("ba", lambda: p.evaluate())
    del p

if which in ("reproject", "all"):
    B = int(os.environ.get("PROF_FRAMES", 16))
    H, W = syn.UHD_SHAPE
    cam = StereoCamera.from_calib_params(syn.UHD_FX, syn.UHD_FX, syn.UHD_CX, syn.UHD_CY,
                                         baseline=syn.UHD_BASELINE, shape=syn.UHD_SHAPE)
    disp = torch.rand((B, H, W), device="cuda", generator=g) * 250 + 1
    im = torch.randint(0, 256, (B, H, W, 3), device="cuda", dtype=torch.uint8, generator=g)
    timed("reproject", lambda: cam.reconstruct_with_texture(disp, im))

if which in ("depth", "flow", "all"):
    from pybot_b200.vision.camera_utils import CameraIntrinsic, DepthCamera
    from pybot_b200.vision.optflow_utils import SceneFlow
    H, W = syn.UHD_SHAPE
    K = np.array([[syn.UHD_FX, 0, syn.UHD_CX], [0, syn.UHD_FX, syn.UHD_CY], [0, 0, 1.0]])
    if which in ("depth", "all"):
        depth = torch.rand((16, H, W), device="cuda", generator=g) * 76 + 4
        dc = DepthCamera(K, shape=(H, W))
        timed("depth", lambda: dc.reconstruct(depth))
        del depth
    if which in ("flow", "all"):
        Z = torch.rand((H, W), device="cuda", generator=g) * 56 + 4
        xs = (torch.arange(W, device="cuda", dtype=torch.float32) - syn.UHD_CX) / syn.UHD_FX
        ys = (torch.arange(H, device="cuda", dtype=torch.float32) - syn.UHD_CY) / syn.UHD_FX
        dX = torch.stack([xs[None, :] * Z + 0.35, ys[:, None] * Z - 0.05, Z - 0.8], dim=2).contiguous()
        sf = SceneFlow(CameraIntrinsic(K, shape=(H, W)))
        timed("flow", lambda: sf.process(dX))
