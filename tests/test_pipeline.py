"""Host-array streaming path (pybot_b200/_pipeline.py): chunk planning on CPU;
on the GPU the chunked upload/kernel/download pipeline must return exactly what
the single-launch device path returns (same kernels, same bits)."""
import numpy as np
import pytest

from pybot_b200 import _pipeline


def test_plan_covers_rows_once_and_respects_alignment():
    assert _pipeline.plan(0, 100) == []
    assert _pipeline.plan(1000, 20) == [(0, 1000)]                   # small: one chunk
    for n, bpr, align in ((10_000_000, 20, 1024), (1_000_003, 36, 1024), (16, 182_000_000, 1), (700_001, 36, 1024)):
        ch = _pipeline.plan(n, bpr, align)
        assert len(ch) > 1 and ch[0][0] == 0 and ch[-1][1] == n
        for (a, b), (c, d) in zip(ch[:-1], ch[1:]):
            assert b == c and a < b and a % align == 0
        assert len(ch) <= _pipeline.MAX_CHUNKS


def test_plan_keeps_every_plane_pointer_16_byte_aligned():
    """KITTI 376x1241 uint8 disparity is 466616 B per frame (mod 16 = 8) and its colour plane
    1399848 B (mod 16 = 8): chunks must start on frames where every plane's offset is a
    multiple of 16 (the kernels reject unaligned pointers)."""
    H, W = 376, 1241
    assert _pipeline.pointer_row_align([H * W]) == 2 and _pipeline.pointer_row_align([H * W * 4]) == 1
    assert _pipeline.pointer_row_align([1001 * 1001 * 4, 1001 * 1001 * 24]) == 4
    assert _pipeline.pointer_row_align([3, 12, 16]) == 16
    for B in (2, 5, 7, 33, 64):
        for planes in ([H * W, H * W * 12], [H * W, H * W * 12, H * W * 3, H * W * 3], [1001 * 1001 * 4, 1001 * 1001 * 24]):
            ch = _pipeline.plan(B, sum(planes) * 40, plane_row_bytes=planes)       # x40: force several chunks
            assert ch[0][0] == 0 and ch[-1][1] == B
            for lo, hi in ch:
                assert lo < hi and all((lo * p) % 16 == 0 for p in planes), (B, planes, ch)
            for (a, b), (c, d) in zip(ch[:-1], ch[1:]):
                assert b == c


@pytest.mark.gpu
@pytest.mark.parametrize("B", [2, 5])
def test_uint8_kitti_frames_with_colour_stream_through_unaligned_frame_sizes(B):
    """2- and 5-frame uint8 KITTI batches (frame stride 8 mod 16) with a colour plane."""
    from pybot_b200 import synthetic as syn
    from pybot_b200.vision.camera_utils import DepthCamera, StereoCamera
    H, W = syn.KITTI_SHAPE
    cam = StereoCamera.from_calib_params(syn.KITTI_FX, syn.KITTI_FX, syn.KITTI_CX, syn.KITTI_CY,
                                         baseline_px=syn.KITTI_BASELINE_PX, shape=(H, W))
    rng = np.random.default_rng(B)
    disp = rng.integers(0, 256, (B, H, W), dtype=np.uint8)
    im = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    old = _pipeline.MIN_BYTES
    _pipeline.MIN_BYTES = 1 << 20                         # make even 2 frames take the streamed path
    try:
        col, xyz = cam.reconstruct_with_texture(disp, im)
        plain = cam.reconstruct(disp)
        dc = DepthCamera(syn.kitti_K(), shape=(1001, 1001))
        depth = rng.uniform(0.3, 40, (B, 1001, 1001)).astype(np.float32)
        got = dc.reconstruct(depth)
    finally:
        _pipeline.MIN_BYTES = old
    for b in range(B):
        c1, x1 = cam.reconstruct_with_texture(disp[b], im[b])
        assert np.array_equal(col[b], c1) and np.array_equal(xyz[b], x1, equal_nan=True)
        assert np.array_equal(plain[b], cam.reconstruct(disp[b]), equal_nan=True)
        assert np.array_equal(got[b], dc.reconstruct(depth[b]))


@pytest.mark.gpu
@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_transform_and_project_host_stream_equals_device_path(pinned, dtype):
    import torch
    from pybot_b200 import _ffi, synthetic as syn
    from pybot_b200.geometry import RigidTransform
    from pybot_b200.vision.camera_utils import Camera, CameraExtrinsic, CameraIntrinsic
    n = 1_500_003
    X = syn.c2_points(n=n).astype(dtype)
    if pinned:
        Xp = _ffi.pinned_empty((n, 3), dtype)
        Xp[:] = X
        X = Xp
    assert len(_pipeline.plan(n, X.itemsize * 5, 1024)) > 1
    pose = RigidTransform.from_rpyxyz(*syn.C2_RPYXYZ)
    cam = Camera.from_intrinsics_extrinsics(CameraIntrinsic(syn.kitti_K(), shape=syn.KITTI_SHAPE),
                                            CameraExtrinsic.from_rigid_transform(pose))
    Xd = torch.from_numpy(np.ascontiguousarray(X)).cuda()
    Y = pose * X
    assert isinstance(Y, np.ndarray) and Y.dtype == np.float64
    assert np.array_equal(Y, (pose * Xd).cpu().numpy())
    uv = cam.project(X)
    assert uv.dtype == dtype and np.array_equal(uv, cam.project(Xd).cpu().numpy())
    uv2, depth = cam.project(X, return_depth=True)
    uvd, depthd = cam.project(Xd, return_depth=True)
    assert np.array_equal(uv2, uvd.cpu().numpy()) and np.array_equal(depth, depthd.cpu().numpy())
    # masked calls: every chunk is compacted on its own and lands at the running offset
    a = cam.project(X, check_bounds=True, check_depth=True)
    assert np.array_equal(a, cam.project(Xd, check_bounds=True, check_depth=True).cpu().numpy())
    x, dep, valid = cam.project(X, check_bounds=True, check_depth=True, return_depth=True, return_valid=True)
    xd, depd, validd = cam.project(Xd, check_bounds=True, check_depth=True, return_depth=True, return_valid=True)
    assert valid.dtype == np.bool_ and np.array_equal(valid, validd.cpu().numpy())
    assert np.array_equal(x, xd.cpu().numpy()) and np.array_equal(dep, depd.cpu().numpy())
    assert len(x) == int(valid.sum()) and 0.7 < valid.mean() < 0.78


@pytest.mark.gpu
def test_reconstruct_batch_host_stream_equals_per_frame():
    from pybot_b200 import synthetic as syn
    from pybot_b200.vision.camera_utils import DepthCamera, StereoCamera
    H, W = syn.KITTI_SHAPE
    cam = StereoCamera.from_calib_params(syn.KITTI_FX, syn.KITTI_FX, syn.KITTI_CX, syn.KITTI_CY,
                                         baseline_px=syn.KITTI_BASELINE_PX, shape=(H, W))
    rng = np.random.default_rng(77)
    B = 7
    disp = np.stack([syn.c1_frame() for _ in range(B)]) * rng.uniform(0.5, 1.5, (B, 1, 1)).astype(np.float32)
    im = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    assert len(_pipeline.plan(B, H * W * 22)) > 1
    col, xyz = cam.reconstruct_with_texture(disp, im)
    plain = cam.reconstruct(disp)
    col2, xyz2 = cam.reconstruct_with_texture(disp, im, sample=3)
    for b in range(B):
        c1, x1 = cam.reconstruct_with_texture(disp[b], im[b])
        assert np.array_equal(col[b], c1) and np.array_equal(xyz[b], x1, equal_nan=True)
        assert np.array_equal(plain[b], cam.reconstruct(disp[b]), equal_nan=True)
        c3, x3 = cam.reconstruct_with_texture(disp[b], im[b], sample=3)
        assert np.array_equal(col2[b], c3) and np.array_equal(xyz2[b], x3, equal_nan=True)
    dc = DepthCamera(syn.kitti_K(), shape=(H, W))
    depth = rng.uniform(0.3, 40, (B, H, W)).astype(np.float32)
    got = dc.reconstruct(depth)
    assert got.shape == (B, H, W, 3) and got.dtype == np.float64
    for b in range(B):
        assert np.array_equal(got[b], dc.reconstruct(depth[b]))
