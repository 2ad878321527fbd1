"""GPU parity: disparity -> XYZ through the C-ABI vs the oracle."""
import numpy as np
import pytest

from oracle import model, port
from pybot_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _same(a, b):
    """bit-exact on every non-NaN value, NaN where NaN."""
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb)
    assert np.array_equal(a[~na].view(np.uint32), b[~nb].view(np.uint32))


def _stereo():
    from pybot_b200.vision.camera_utils import StereoCamera
    return StereoCamera.from_calib_params(syn.KITTI_FX, syn.KITTI_FX, syn.KITTI_CX, syn.KITTI_CY,
                                          baseline_px=syn.KITTI_BASELINE_PX, shape=syn.KITTI_SHAPE)


def test_c1_kitti_frame_bit_exact():
    cam = _stereo()
    disp = syn.c1_frame()
    X = cam.reconstruct(disp)
    assert X.shape == (376, 1241, 3) and X.dtype == np.float32
    _same(X, model.reproject_image_to_3d(disp, cam.Q))
    _same(X, port.stereo_reconstruct(disp, cam.Q))          # cv2 itself when importable
    # specials: d=0 -> +-inf, NaN/inf -> NaN
    assert np.isnan(X[0, 1]).all() and np.isnan(X[0, 2]).all()
    assert np.isinf(X[disp == 0]).all()


def test_reciprocal_vs_divide_golden(golden_dir):
    g = np.load(golden_dir + "/reproject_divcases.npz")
    from pybot_b200 import ops
    _same(ops.reproject_disparity(g["disp"], g["Q"]), g["xyz"])


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.int32, np.float32])
def test_dtypes_and_general_Q(dtype):
    from pybot_b200 import ops
    rng = np.random.default_rng(11)
    H, W = 37, 203                                   # ragged: H*W % 128 != 0, W % 4 != 0
    if dtype == np.float32:
        disp = rng.uniform(-5, 200, (H, W)).astype(np.float32)
    else:
        info = np.iinfo(dtype)
        disp = rng.integers(max(info.min, -300), min(info.max, 3000), (H, W)).astype(dtype)
    Qs = [port.stereo_Q(700.0, 100.5, 20.25, 0.54), rng.normal(size=(4, 4)).astype(np.float32)]
    for Q in Qs:
        _same(ops.reproject_disparity(disp, Q), model.reproject_image_to_3d(disp, Q))


def test_float64_rejected_like_opencv():
    from pybot_b200 import ops
    with pytest.raises(NotImplementedError):
        ops.reproject_disparity(np.zeros((4, 4)), np.eye(4, dtype=np.float32))


@pytest.mark.parametrize("shape", [(1, 1), (3, 3), (2, 5), (1, 128), (1, 129), (7, 64)])
def test_small_and_tail_shapes(shape):
    from pybot_b200 import ops
    rng = np.random.default_rng(shape[0] * 100 + shape[1])
    disp = rng.uniform(0.5, 90, shape).astype(np.float32)
    Q = port.stereo_Q(500.0, 3.5, 2.5, 0.1)
    _same(ops.reproject_disparity(disp, Q), model.reproject_image_to_3d(disp, Q))


def test_empty():
    from pybot_b200 import ops
    out = ops.reproject_disparity(np.zeros((0, 16), np.float32), np.eye(4, dtype=np.float32))
    assert out.shape == (0, 16, 3)


@pytest.mark.parametrize("sample", [1, 2, 3])
def test_reconstruct_with_texture(sample):
    cam = _stereo()
    rng = np.random.default_rng(5)
    disp = syn.c1_frame(seed=77, shape=(94, 311))
    im = rng.integers(0, 256, (94, 311, 3), dtype=np.uint8)
    im_pub, X_pub = cam.reconstruct_with_texture(disp, im, sample=sample)
    rim, rX = port.stereo_reconstruct_with_texture(disp, im, cam.Q, sample=sample)
    assert np.array_equal(im_pub, rim)
    _same(X_pub, rX)


def test_batched_frames_match_per_frame():
    cam = _stereo()
    rng = np.random.default_rng(6)
    disp = np.stack([syn.disparity_image(rng, (45, 131), syn.KITTI_FX, syn.KITTI_BASELINE) for _ in range(5)])
    im = rng.integers(0, 256, (5, 45, 131, 3), dtype=np.uint8)
    Xb = cam.reconstruct(disp)
    for b in range(5):
        _same(Xb[b], model.reproject_image_to_3d(disp[b], cam.Q))
    cb, xb = cam.reconstruct_with_texture(disp, im)
    assert np.array_equal(cb, im.reshape(5, -1, 3))
    _same(xb, Xb.reshape(5, -1, 3))


def test_device_tensor_in_device_tensor_out():
    import torch
    cam = _stereo()
    disp = syn.c1_frame(seed=3, shape=(64, 256))
    out = cam.reconstruct(torch.from_numpy(disp).cuda())
    assert out.is_cuda
    _same(out.cpu().numpy(), model.reproject_image_to_3d(disp, cam.Q))


def test_reconstruct_sparse():
    cam = _stereo()
    rng = np.random.default_rng(8)
    xyd = np.stack([rng.uniform(0, 1241, 500), rng.uniform(0, 376, 500), rng.uniform(1, 90, 500)], 1)
    got = cam.reconstruct_sparse(xyd)
    ref = port.stereo_reconstruct_sparse(xyd, cam.Q)
    assert got.dtype == np.float64
    np.testing.assert_allclose(got, ref, rtol=1e-13, atol=0)


def test_full_size_property_4k_batch():
    """C5-sized property test (no oracle at this size): Z = fx*b/d, X/Z = (x-cx)/fx,
    and batch consistency, on 4 x 2160x3840 frames."""
    import torch
    from pybot_b200.vision.camera_utils import StereoCamera
    cam = StereoCamera.from_calib_params(syn.UHD_FX, syn.UHD_FX, syn.UHD_CX, syn.UHD_CY,
                                         baseline=syn.UHD_BASELINE, shape=syn.UHD_SHAPE)
    g = torch.Generator(device="cuda").manual_seed(5)
    disp = torch.rand((4, 2160, 3840), device="cuda", generator=g) * 250 + 1
    X = cam.reconstruct(disp)
    Q = cam.Q.astype(np.float64)
    Z = X[..., 2].double()
    ref_Z = (Q[2, 3] / (Q[3, 2] * disp.double()))
    assert torch.allclose(Z, ref_Z, rtol=2e-7, atol=0)
    xs = torch.arange(3840, device="cuda", dtype=torch.float64)
    ref_X = (Q[0, 3] - xs)[None, None, :] / (Q[3, 2] * disp.double())
    assert torch.allclose(X[..., 0].double(), ref_X, rtol=2e-7, atol=1e-9)
    one = cam.reconstruct(disp[2])
    assert torch.equal(one, X[2])
