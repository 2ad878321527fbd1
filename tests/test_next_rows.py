"""DepthCamera.reconstruct / reconstruct_sparse, SceneFlow.process,
get_object_bbox.  Golden outputs come from the real reference
(oracle/make_golden.py g6_next_rows).  CPU tests pin the oracle port; GPU tests
pin the CUDA path - bit-exact, all of it is either fp64 mul/div or integer."""
import numpy as np
import pytest

from oracle import model, port
from pybot_b200 import synthetic as syn


@pytest.fixture(scope="module")
def g6(golden_dir):
    return np.load(golden_dir + "/g6_next_rows.npz")


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (a.shape, b.shape, a.dtype, b.dtype)
    nan = np.isnan(b)
    assert np.array_equal(np.isnan(a), nan)
    assert np.array_equal(a[~nan], b[~nan])


# ------------------------------------------------------------------ oracle (CPU)
def test_port_depth_reconstruct_matches_reference(g6):
    K, shape = g6["K"], tuple(g6["shape"])
    for skip in (1, 2, 3):
        _same(port.depth_reconstruct(K, shape, skip, g6["d32"]), g6["depth_f32_s%d" % skip])
    _same(port.depth_reconstruct(K, shape, 1, g6["d16"]), g6["depth_u16_s1"])
    _same(port.depth_reconstruct_sparse(K, g6["sparse_pts"], g6["sparse_depth"]), g6["sparse_out"])


@pytest.mark.parametrize("use_cv2", [True, False])
def test_port_sceneflow_matches_reference(g6, use_cv2):
    K, shape = g6["K"], tuple(g6["shape"])
    _same(port.sceneflow_process(K, np.zeros(5), shape, g6["flow_X_f32"], use_cv2=use_cv2), g6["flow_f32"])
    _same(port.sceneflow_process(K, np.zeros(5), shape, g6["flow_X_f64"], use_cv2=use_cv2), g6["flow_f64"])
    _same(port.sceneflow_process(K, g6["flow_D"], shape, g6["flow_X_f32"], use_cv2=use_cv2), g6["flow_dist"])


# --------------------------------------------------------------------- CUDA (GPU)
@pytest.mark.gpu
def test_depth_camera_reconstruct_golden(g6):
    from pybot_b200.vision.camera_utils import DepthCamera
    K, shape = g6["K"], tuple(int(v) for v in g6["shape"])
    for skip in (1, 2, 3):
        _same(DepthCamera(K, shape=shape, skip=skip).reconstruct(g6["d32"]), g6["depth_f32_s%d" % skip])
    dc = DepthCamera(K, shape=shape)
    _same(dc.reconstruct(g6["d16"]), g6["depth_u16_s1"])
    _same(dc.reconstruct(g6["d32"].astype(np.float64)), g6["depth_f32_s1"])      # f32 values are exact in f64
    _same(dc.reconstruct_sparse(g6["sparse_pts"], g6["sparse_depth"]), g6["sparse_out"])
    with pytest.raises(AssertionError):
        dc.reconstruct(g6["d32"][:-1])


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 1), (3, 5), (2, 63), (1, 64), (7, 65), (480, 640)])
def test_depth_camera_shapes_and_batches(shape):
    import torch
    from pybot_b200 import synthetic as syn
    from pybot_b200.vision.camera_utils import DepthCamera
    rng = np.random.default_rng(sum(shape))
    K = syn.kitti_K()
    d = rng.uniform(0.3, 40, (3,) + shape).astype(np.float32)
    dc = DepthCamera(K, shape=shape)
    got = dc.reconstruct(d)                                           # batched
    for b in range(3):
        _same(got[b], port.depth_reconstruct(K, shape, 1, d[b]))
    dev = dc.reconstruct(torch.from_numpy(d[0]).cuda())               # device in -> device out
    assert dev.is_cuda
    _same(dev.cpu().numpy(), port.depth_reconstruct(K, shape, 1, d[0]))


@pytest.mark.gpu
def test_sceneflow_golden(g6):
    from pybot_b200.vision.camera_utils import CameraIntrinsic
    from pybot_b200.vision.optflow_utils import SceneFlow
    K, shape = g6["K"], tuple(int(v) for v in g6["shape"])
    sf = SceneFlow(CameraIntrinsic(K, shape=shape))
    _same(sf.process(g6["flow_X_f32"]), g6["flow_f32"])
    _same(sf.process(g6["flow_X_f64"]), g6["flow_f64"])
    sfd = SceneFlow(CameraIntrinsic(K, D=g6["flow_D"], shape=shape))
    _same(sfd.process(g6["flow_X_f32"]), g6["flow_dist"])
    with pytest.raises(ValueError):
        sf.process(g6["flow_X_f32"][:-1])


@pytest.mark.gpu
def test_sceneflow_kitti_frame_vs_port():
    """Full KITTI frame: a forward-moving camera over the C1 depth field."""
    from pybot_b200 import synthetic as syn
    from pybot_b200.vision.camera_utils import CameraIntrinsic, StereoCamera
    from pybot_b200.vision.optflow_utils import SceneFlow
    H, W = syn.KITTI_SHAPE
    K = syn.kitti_K()
    cam = StereoCamera.from_calib_params(syn.KITTI_FX, syn.KITTI_FX, syn.KITTI_CX, syn.KITTI_CY,
                                         baseline_px=syn.KITTI_BASELINE_PX, shape=(H, W))
    X = cam.reconstruct(syn.c1_frame())                               # (H,W,3) float32, with inf/NaN specials
    X = -X                                                            # Q negates X, Y, W: put the cloud in front
    X[..., 2] = np.abs(X[..., 2]) - 1.0                               # camera moved 1 m forward
    got = SceneFlow(CameraIntrinsic(K, shape=(H, W))).process(X)
    _same(got, port.sceneflow_process(K, np.zeros(5), (H, W), X))
    hit = ~np.isnan(got[..., 0])
    assert 0.2 < hit.mean() < 0.99


@pytest.mark.gpu
def test_get_object_bbox_golden(g6):
    from pybot_b200 import synthetic as syn
    from pybot_b200.geometry import RigidTransform
    from pybot_b200.vision.camera_utils import (Camera, CameraExtrinsic, CameraIntrinsic, get_object_bbox)
    pose = RigidTransform.from_matrix(g6["bbox_pose"])
    c = Camera.from_intrinsics_extrinsics(CameraIntrinsic(syn.kitti_K(), shape=(376, 1241)),
                                          CameraExtrinsic.from_rigid_transform(pose))
    p2, bbox, depth = get_object_bbox(c, g6["bbox_pts"], subsample=3, scale=1.2)
    assert np.array_equal(p2, g6["bbox_p2"]) and p2.dtype == np.int32
    assert np.array_equal(bbox, g6["bbox_box"]) and bbox.dtype == np.float32
    assert depth == float(g6["bbox_depth"])
    assert get_object_bbox(c, g6["bbox_pts"] - [0, 0, 200.0]) == [None] * 3     # behind the camera: nothing in view


# ------------------------------------------------- sampson error / wire colours
@pytest.fixture(scope="module")
def g7(golden_dir):
    return np.load(golden_dir + "/g7_epipolar_wire.npz")


def test_port_sampson_and_wire_match_reference(g7):
    assert np.array_equal(port.sampson_error(g7["F"], g7["p2"], g7["p1"]), g7["err"])
    _same(port.wire_colors(g7["wire_im"], False), g7["wire_rgb"])
    _same(port.wire_colors(g7["wire_im"], True), g7["wire_flip"])


def _cams(g7):
    from pybot_b200 import synthetic as syn
    from pybot_b200.geometry import RigidTransform
    from pybot_b200.vision.camera_utils import Camera, CameraExtrinsic, CameraIntrinsic
    mk = lambda rp: Camera.from_intrinsics_extrinsics(              # noqa: E731
        CameraIntrinsic(syn.kitti_K(), shape=(376, 1241)),
        CameraExtrinsic.from_rigid_transform(RigidTransform.from_rpyxyz(*rp)))
    return mk(g7["rp1"]), mk(g7["rp2"])


def test_fundamental_matrix_matches_reference(g7):
    """Camera.F is host-side (nine 4x4 determinants): same values as the reference."""
    c1, c2 = _cams(g7)
    np.testing.assert_allclose(c1.F(c2), g7["F"], rtol=1e-12, atol=1e-9 * np.abs(g7["F"]).max())


@pytest.mark.gpu
def test_sampson_error_golden(g7):
    from pybot_b200.vision.camera_utils import filter_sampson_error, sampson_error
    err = sampson_error(g7["F"], g7["p2"], g7["p1"])
    assert err.dtype == np.float64 and err.shape == g7["err"].shape
    np.testing.assert_allclose(err, g7["err"], rtol=1e-12)          # fp64; BLAS summation order is the only freedom
    c1, c2 = _cams(g7)
    f1, f2, ids = filter_sampson_error(c1, c2, g7["p1"], g7["p2"], np.arange(len(g7["p1"])), error=4)
    assert np.array_equal(ids, g7["inlier_ids"])
    assert np.array_equal(f1, g7["p1"][ids]) and np.array_equal(f2, g7["p2"][ids])
    big = np.random.default_rng(1).uniform(0, 1000, (1_000_003, 2))      # the reference would need an N x N matrix
    e = sampson_error(g7["F"], big, big[::-1].copy())
    assert np.isfinite(e).all() and (e >= 0).all()


@pytest.mark.gpu
def test_wire_colors_golden(g7):
    from pybot_b200 import ops
    _same(ops.wire_colors(g7["wire_im"], False), g7["wire_rgb"])
    _same(ops.wire_colors(g7["wire_im"], True), g7["wire_flip"])
    im = np.random.default_rng(2).integers(0, 256, (333, 517, 3), dtype=np.uint8)      # n_px not a multiple of 4
    _same(ops.wire_colors(im, True), port.wire_colors(im, True))


# ------------------------------------ undistort_points / ray / reconstruct / triangulate
@pytest.fixture(scope="module")
def g8(golden_dir):
    return np.load(golden_dir + "/g8_rays_triangulate.npz")


_DISTS = ("d0", "d5", "d4", "d8", "dneg")


def _pose_xyzw(g8):
    """Quaternion the reference ends up rotating with: pose -> Camera -> .extrinsics
    (two matrix round trips; host algebra only, pinned by g0)."""
    return _g8_cams(g8)[1].extrinsics.quat.q


@pytest.mark.parametrize("use_cv2", [True, False])
def test_port_rays_match_reference(g8, use_cv2):
    K, pts = g8["K"], g8["pts"]
    for name in _DISTS:
        D = g8["D_" + name]
        _same(port.undistort_points(K, D, pts, use_cv2), g8["und_" + name])
        _same(port.camera_ray(K, D, pts, use_cv2=use_cv2), g8["ray_" + name])
    D = g8["D_d5"]
    q = _pose_xyzw(g8)
    _same(port.camera_ray(K, D, pts, undistort=False), g8["ray_raw_f64"])
    _same(port.camera_ray(K, D, pts.astype(np.float32), undistort=False), g8["ray_raw_f32"])
    _same(port.camera_ray(K, D, pts, normalize=True, use_cv2=use_cv2), g8["ray_norm"])
    _same(port.camera_ray(K, D, pts, rotate_xyzw=q, use_cv2=use_cv2), g8["ray_rot"])
    _same(port.camera_ray(K, D, pts, undistort=False, rotate_xyzw=q, normalize=True), g8["ray_rot_norm"])
    _same(port.intrinsic_reconstruct(K, D, g8["xyZ"], use_cv2=use_cv2), g8["rec"])
    _same(port.intrinsic_reconstruct(K, D, g8["xyZ"], undistort=False), g8["rec_raw"])


def test_model_undistort_is_bit_exact_vs_cv2():
    cv2 = pytest.importorskip("cv2")
    from oracle import model
    from pybot_b200 import synthetic as syn
    K = syn.kitti_K()
    rng = np.random.default_rng(88)
    pts = rng.uniform([-100, -100], [1400, 500], (200_000, 2))
    for D in ([-0.3, 0.1, 0.001, -0.002, 0.01], [-0.37, 0.2, 0.0005, 0.0003],
              [0.1, -0.05, 0.001, 0.002, 0.003, 0.01, 0.02, -0.003], [-0.9, 0.1, 0, 0, 0]):
        D = np.float64(D)
        ref = cv2.undistortPoints(pts.reshape(-1, 1, 2).astype(np.float32), K, D, P=K).reshape(-1, 2)
        _same(model.undistort_points(pts, K, D, P=K), ref)


@pytest.mark.parametrize("use_cv2", [True, False])
def test_port_triangulate_matches_reference(g8, use_cv2):
    P1, P2 = g8["tri_P1"], g8["tri_P2"]
    for got, want, tol in (
            (port.triangulate_points(P1, P2, g8["tri_p1"], g8["tri_p2"], use_cv2), g8["tri_exact"], 1e-6),
            (port.triangulate_points(P1, P2, g8["tri_p1"], g8["tri_p2n"], use_cv2), g8["tri_noisy"], 1e-9)):
        assert got.dtype == want.dtype and got.shape == want.shape
        np.testing.assert_allclose(got, want, rtol=tol, atol=tol)
    np.testing.assert_allclose(g8["tri_exact"], g8["tri_X"], atol=1e-6)       # noise-free pairs recover the points
    got = port.triangulate_points(P1, P2, g8["tri_p1"].astype(np.float32), g8["tri_p2n"].astype(np.float32), use_cv2)
    assert got.dtype == np.float32
    np.testing.assert_allclose(got, g8["tri_f32"], rtol=2e-5, atol=1e-5)


def _g8_cams(g8):
    from pybot_b200.geometry import RigidTransform
    from pybot_b200.vision.camera_utils import Camera, CameraExtrinsic, CameraIntrinsic
    ci = CameraIntrinsic(g8["K"], D=g8["D_d5"], shape=(376, 1241))
    cam = Camera.from_intrinsics_extrinsics(
        ci, CameraExtrinsic.from_rigid_transform(RigidTransform.from_rpyxyz(*g8["pose_rpyxyz"])))
    return ci, cam


@pytest.mark.gpu
def test_undistort_and_rays_golden(g8):
    """Bit-exact against the reference's cv2.undistortPoints + numpy ray arithmetic."""
    from pybot_b200.vision.camera_utils import CameraIntrinsic
    K, pts = g8["K"], g8["pts"]
    for name in _DISTS:
        ci = CameraIntrinsic(K, D=g8["D_" + name], shape=(376, 1241))
        _same(ci.undistort_points(pts), g8["und_" + name])
        _same(ci.ray(pts), g8["ray_" + name])
    ci, cam = _g8_cams(g8)
    _same(ci.ray(pts, undistort=False), g8["ray_raw_f64"])
    _same(ci.ray(pts.astype(np.float32), undistort=False), g8["ray_raw_f32"])
    _same(ci.ray(pts, normalize=True), g8["ray_norm"])
    _same(cam.ray(pts, rotate=True), g8["ray_rot"])
    _same(cam.ray(pts, undistort=False, rotate=True, normalize=True), g8["ray_rot_norm"])
    _same(ci.reconstruct(g8["xyZ"]), g8["rec"])
    _same(ci.reconstruct(g8["xyZ"], undistort=False), g8["rec_raw"])
    with pytest.raises(AttributeError):
        ci.ray(pts, rotate=True)                                   # no extrinsics on a bare intrinsic, like the reference


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 255, 256, 257, 100_003])
def test_rays_ragged_sizes_vs_port(n, g8):
    import torch
    ci, cam = _g8_cams(g8)
    rng = np.random.default_rng(n)
    xyZ = np.hstack([rng.uniform([-50, -50], [1300, 430], (n, 2)), rng.uniform(0.5, 60, (n, 1))])
    K, D = g8["K"], g8["D_d5"]
    q = _pose_xyzw(g8)
    if n == 0:                       # cv2.undistortPoints asserts on an empty array; here: empty out
        assert ci.undistort_points(xyZ[:, :2]).shape == (0, 2)
        assert cam.ray(xyZ[:, :2], rotate=True).shape == (0, 3) and ci.reconstruct(xyZ).shape == (0, 3)
        return
    _same(ci.undistort_points(xyZ[:, :2]), port.undistort_points(K, D, xyZ[:, :2]).reshape(-1, 2))
    _same(cam.ray(xyZ[:, :2], rotate=True, normalize=True),
          port.camera_ray(K, D, xyZ[:, :2], rotate_xyzw=q, normalize=True).reshape(-1, 3))
    _same(ci.reconstruct(xyZ), port.intrinsic_reconstruct(K, D, xyZ).reshape(-1, 3))
    dev = ci.reconstruct(torch.from_numpy(xyZ).cuda())               # device in -> device out
    assert dev.is_cuda
    _same(dev.cpu().numpy(), port.intrinsic_reconstruct(K, D, xyZ).reshape(-1, 3))


@pytest.mark.gpu
def test_triangulate_points_golden(g8):
    """cv2.triangulatePoints is an SVD: parity is a tolerance (1e-9 relative in
    float64 on noisy pairs, where the null vector is well separated; the exact
    pairs are rank-deficient by construction, compared at 1e-6 like cv2 itself
    against the true points)."""
    from pybot_b200 import synthetic as syn
    from pybot_b200.geometry import RigidTransform
    from pybot_b200.vision.camera_utils import (Camera, CameraExtrinsic, CameraIntrinsic, triangulate_points)
    mk = lambda rp: Camera.from_intrinsics_extrinsics(              # noqa: E731
        CameraIntrinsic(syn.kitti_K(), shape=(376, 1241)),
        CameraExtrinsic.from_rigid_transform(RigidTransform.from_rpyxyz(*rp)))
    c1, c2 = mk(g8["tri_rp1"]), mk(g8["tri_rp2"])
    np.testing.assert_allclose(c1.P, g8["tri_P1"], rtol=1e-12, atol=1e-12)
    got = triangulate_points(c1, g8["tri_p1"], c2, g8["tri_p2n"])
    assert got.dtype == np.float64 and got.shape == g8["tri_noisy"].shape
    np.testing.assert_allclose(got, g8["tri_noisy"], rtol=1e-9, atol=1e-9)
    got = triangulate_points(c1, g8["tri_p1"], c2, g8["tri_p2"])
    np.testing.assert_allclose(got, g8["tri_exact"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(got, g8["tri_X"], atol=1e-6)
    got = triangulate_points(c1, g8["tri_p1"].astype(np.float32), c2, g8["tri_p2n"].astype(np.float32))
    assert got.dtype == np.float32
    np.testing.assert_allclose(got, g8["tri_f32"], rtol=2e-5, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 127, 128, 129, 200_001])
def test_triangulate_ragged_sizes_vs_port(n, g8):
    from pybot_b200 import ops
    rng = np.random.default_rng(n + 5)
    P1, P2 = g8["tri_P1"], g8["tri_P2"]
    X = rng.uniform([-10, -2, 5], [10, 2, 40], (n, 3))
    Xh = np.hstack([X, np.ones((n, 1))])
    p1 = (Xh @ P1.T)
    p2 = (Xh @ P2.T)
    p1 = p1[:, :2] / p1[:, 2:] + rng.normal(0, 0.7, (n, 2))
    p2 = p2[:, :2] / p2[:, 2:] + rng.normal(0, 0.7, (n, 2))
    got = ops.triangulate_points(P1, P2, p1, p2)
    assert got.shape == (n, 3) and got.dtype == np.float64
    if n:
        np.testing.assert_allclose(got, port.triangulate_points(P1, P2, p1, p2), rtol=1e-9, atol=1e-9)


# ------------------------------- epipolar lines and the small helpers around matching
@pytest.fixture(scope="module")
def g9(golden_dir):
    return np.load(golden_dir + "/g9_epipolar_helpers.npz")


def test_model_and_port_epipolar_lines_match_reference(g9):
    """The numpy restatement of cv2.computeCorrespondEpilines is bit-identical to what the
    reference's epipolar_line returned (float64 and float32 points, degenerate a = b = 0)."""
    x32 = g9["x64"].astype(np.float32)
    _same(model.epipolar_lines(g9["F"], g9["x64"]), g9["lines64"])
    _same(model.epipolar_lines(g9["F"], x32), g9["lines32"])
    _same(model.epipolar_lines(g9["Fz"], g9["x64"][:8]), g9["lines_z"])
    if port.cv2 is not None:
        _same(port.epipolar_line(g9["F"], g9["x64"]), g9["lines64"])
        _same(port.epipolar_line(g9["F"], x32), g9["lines32"])


def test_host_helpers_match_reference(g9):
    """colvec / rowvec / unproject / project / project_points / compute_essential / decompose_E and
    rigid_transform's normalize_vec / skew / tf_construct / tf_construct_3pt / tf_compose: same
    values as the reference.  camera_utils.skew, Camera.E and compute_epipole raise NameError
    there (missing imports), so those three are checked against their definitions."""
    from pybot_b200 import synthetic as syn
    from pybot_b200.geometry import rigid_transform as rt
    from pybot_b200.vision import camera_utils as cu
    X = g9["X"]
    _same(cu.project_points(X), g9["proj"])
    _same(cu.unproject(X[0]), g9["unproj"])
    _same(cu.project(X[0]), g9["deproj"])
    assert cu.colvec(X[0]).shape == (3, 1) and cu.rowvec(X[0]).shape == (1, 3)
    E = cu.compute_essential(g9["F"], syn.kitti_K())
    _same(E, g9["E"])
    R1, R2, t = cu.decompose_E(E)
    _same(R1, g9["R1"]), _same(R2, g9["R2"]), _same(t, g9["t"])
    v1, v2 = g9["v1"], g9["v2"]
    _same(rt.tf_construct(v1, v2), g9["tfc"])
    sk, dV = rt.skew(v1, return_dv=True)
    _same(sk, g9["skew_rt"]), _same(dV, g9["skew_dv"])
    assert sk.dtype == np.float32
    _same(rt.normalize_vec(v1 * 3.0), v1 * 3.0 / np.linalg.norm(v1 * 3.0))
    p3 = g9["p3"]
    _same(rt.tf_construct_3pt(p3[0], p3[1], p3[2]).to_matrix(), g9["T3"])
    _same(rt.tf_compose(g9["R1"], g9["t"]), g9["tcomp"])
    # the three the reference cannot run
    w = np.array([0.3, -1.2, 2.0])
    np.testing.assert_allclose(cu.skew(v1).dot(w), np.cross(v1, w), rtol=0, atol=1e-15)
    e = cu.compute_epipole(g9["F"])
    assert e[2] == 1.0 and np.abs(g9["F"].dot(e)).max() < 1e-9 * np.abs(g9["F"]).max() * np.abs(e).max()
    c1, c2 = _cams(g9)
    _same(c2.E(c1), cu.skew(c2.t).dot(c2.R))
    with pytest.raises(NotImplementedError):
        cu.recover_pose(E, None, None, None, None, None)


@pytest.mark.gpu
def test_epipolar_line_golden(g9):
    """epipolar_line (cv2.computeCorrespondEpilines) on the GPU: bit-identical lines."""
    from pybot_b200.vision.camera_utils import epipolar_line
    l64 = epipolar_line(g9["F"], g9["x64"])
    assert l64.dtype == np.float64 and l64.shape == (500, 3)
    _same(l64, g9["lines64"])
    l32 = epipolar_line(g9["F"], g9["x64"].astype(np.float32))
    assert l32.dtype == np.float32
    _same(l32, g9["lines32"])
    _same(epipolar_line(g9["Fz"], g9["x64"][:8]), g9["lines_z"])
    assert epipolar_line(g9["F"], np.zeros((0, 2))).shape == (0, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("n,dt", [(1, np.float64), (255, np.float32), (257, np.float64), (100_003, np.float32)])
def test_epipolar_line_ragged_sizes_vs_model(g9, n, dt):
    import torch
    from pybot_b200 import ops
    rng = np.random.default_rng(n)
    x = rng.uniform([-50, -50], [1300, 400], (n, 2)).astype(dt)
    want = model.epipolar_lines(g9["F"], x)
    _same(ops.epipolar_lines(g9["F"], x), want)
    got = ops.epipolar_lines(g9["F"], torch.from_numpy(x).cuda())         # device in -> device out
    assert got.is_cuda and np.array_equal(got.cpu().numpy(), want)
    if port.cv2 is not None:
        _same(port.epipolar_line(g9["F"], x), want)


@pytest.fixture(scope="module")
def g10(golden_dir):
    return np.load(golden_dir + "/g10_visibility.npz")


def _vis_camera(g10):
    from pybot_b200.geometry import RigidTransform
    from pybot_b200.vision.camera_utils import Camera, CameraExtrinsic, CameraIntrinsic
    pose = RigidTransform(g10["vis_pose_xyzw"], g10["vis_pose_tvec"])
    c = Camera.from_intrinsics_extrinsics(CameraIntrinsic(syn.kitti_K(), shape=(376, 1241)),
                                          CameraExtrinsic.from_rigid_transform(pose))
    assert np.array_equal(c.quat.q, g10["vis_cam_xyzw"]) and np.array_equal(c.tvec, g10["vis_cam_tvec"])
    return c


@pytest.mark.gpu
def test_check_visibility_golden(g10):
    """check_visibility (camera_utils.py:245-266) as one fused kernel against the mask the real
    reference returned: points behind the camera, beyond zmax, on the z limits, outside the fov."""
    import torch
    from pybot_b200.vision.camera_utils import check_visibility
    c = _vis_camera(g10)
    assert np.array_equal(c.fov, g10["vis_fov"]) and c.fov.dtype == np.float32
    pts = g10["vis_pts"]
    m = check_visibility(c, pts)
    assert m.dtype == np.bool_ and np.array_equal(m, g10["vis_mask"])
    assert np.array_equal(c.check_visibility(pts, zmin=5.0, zmax=30.0), g10["vis_mask_z"])
    assert np.array_equal(check_visibility(c, pts.astype(np.float32)), g10["vis_mask_f32"])
    md = check_visibility(c, torch.from_numpy(pts).cuda())            # device in -> device mask out
    assert md.is_cuda and md.dtype == torch.bool and np.array_equal(md.cpu().numpy(), g10["vis_mask"])
    assert 0.2 < m.mean() < 0.5 and check_visibility(c, pts[:0]).shape == (0,)


@pytest.mark.gpu
def test_median_depth_and_bounded_projection_golden(g10):
    from pybot_b200.vision.camera_utils import get_bounded_projection, get_median_depth
    c = _vis_camera(g10)
    pts = g10["vis_pts"]
    assert get_median_depth(c, pts, subsample=7) == float(g10["vis_median"])
    p2, valid = get_bounded_projection(c, pts, subsample=3)
    assert p2.dtype == np.float32 and np.array_equal(valid, g10["vis_bp_valid"])
    assert np.array_equal(p2, g10["vis_bp_x"])
