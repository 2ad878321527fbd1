"""End-to-end (host arrays in, host arrays out) timing of the streaming path
against the single-chunk path.  python scripts/e2e_probe.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pybot_b200 import _ffi, _pipeline, synthetic as syn  # noqa: E402
from pybot_b200.geometry import RigidTransform  # noqa: E402
from pybot_b200.vision.camera_utils import Camera, CameraExtrinsic, CameraIntrinsic, StereoCamera  # noqa: E402


def timeit(fn, reps=5):
    fn()
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def both(name, fn, nbytes):
    saved = _pipeline.MIN_BYTES
    res = {}
    for label, mb in (("single", 1 << 60), ("streamed", saved)):
        _pipeline.MIN_BYTES = mb
        res[label] = timeit(fn)
    _pipeline.MIN_BYTES = saved
    print("%-34s single %8.3f ms  streamed %8.3f ms  (%.2fx, %.1f GB/s over PCIe)" % (
        name, res["single"], res["streamed"], res["single"] / res["streamed"], nbytes / res["streamed"] / 1e6))


def main():
    n = 10_000_000
    X = _ffi.pinned_empty((n, 3), np.float32)
    X[:] = syn.c2_points(n=n)
    pose = RigidTransform.from_rpyxyz(*syn.C2_RPYXYZ)
    cam = Camera.from_intrinsics_extrinsics(CameraIntrinsic(syn.kitti_K(), shape=syn.KITTI_SHAPE),
                                            CameraExtrinsic.from_rigid_transform(pose))
    both("C2 project 10M pinned", lambda: cam.project(X), n * 20)
    both("C2 transform 10M pinned (f64 out)", lambda: pose * X, n * 36)
    for cb in (12, 24, 32, 48, 64):
        _pipeline.COMPACT_CHUNK_BYTES = cb << 20
        both("C2 project x[valid]+depth+valid, %d MB chunks" % cb,
             lambda: cam.project(X, check_bounds=True, check_depth=True, return_depth=True, return_valid=True), n * 25)
    _pipeline.COMPACT_CHUNK_BYTES = 48 << 20
    Xu = np.array(X)                                     # pageable
    both("C2 project 10M pageable", lambda: cam.project(Xu), n * 20)
    H, W = syn.UHD_SHAPE
    sc = StereoCamera.from_calib_params(syn.UHD_FX, syn.UHD_FX, syn.UHD_CX, syn.UHD_CY,
                                        baseline=syn.UHD_BASELINE, shape=syn.UHD_SHAPE)
    B = 8
    d = _ffi.pinned_empty((B, H, W), np.float32)
    im = _ffi.pinned_empty((B, H, W, 3), np.uint8)
    rng = np.random.default_rng(5)
    d[:] = rng.uniform(1, 200, (B, H, W)).astype(np.float32)
    im[:] = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    both("C5 textured cloud, %d 4K frames" % B, lambda: sc.reconstruct_with_texture(d, im), B * H * W * 22)
    both("C5 plain cloud, %d 4K frames" % B, lambda: sc.reconstruct(d), B * H * W * 16)


if __name__ == "__main__":
    main()
