"""GPU parity: RigidTransform.__mul__, Sim3, Camera.project."""
import numpy as np
import pytest

from oracle import model, port
from pybot_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

POSES = ("c2", "ident", "big")


@pytest.fixture(params=["lookback", "two_kernel", "early"])
def compact_impl(request, monkeypatch):
    """Every compaction implementation behind Camera.project(check_*): the single-pass look-back
    kernel, the classify + dense-projection pair (PBK_COMPACT_IMPL=2) and the single pass with
    early aggregates (PBK_COMPACT_IMPL=3)."""
    monkeypatch.setenv("PBK_COMPACT_IMPL", {"lookback": "1", "two_kernel": "2", "early": "3"}[request.param])
    return request.param


def _same_bits(a, b):
    """torch.equal on the raw bits (NaN pixels of points that pass a depth-only check compare equal)."""
    import torch
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    it = {4: torch.int32, 8: torch.int64}[a.element_size()]
    return bool(torch.equal(a.contiguous().view(it), b.contiguous().view(it)))


def _bits_equal(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (a.shape, b.shape, a.dtype, b.dtype)
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb)
    return np.array_equal(a[~na], b[~nb])


@pytest.fixture(scope="module")
def g2(golden_dir):
    return np.load(golden_dir + "/g2_transform_project.npz")


def _camera(g2, name, D=None):
    from pybot_b200.vision.camera_utils import Camera, CameraIntrinsic, CameraExtrinsic
    from pybot_b200.geometry import RigidTransform
    p = RigidTransform(g2["pose_%s_xyzw" % name], g2["pose_%s_tvec" % name])
    intr = CameraIntrinsic(g2["K"], shape=syn.KITTI_SHAPE) if D is None else CameraIntrinsic(
        g2["K"], D=D, shape=syn.KITTI_SHAPE)
    cam = Camera.from_intrinsics_extrinsics(intr, CameraExtrinsic.from_rigid_transform(p))
    assert np.array_equal(cam.quat.q, g2["cam_%s_xyzw" % name])
    assert np.array_equal(cam.tvec, g2["cam_%s_tvec" % name])
    return cam


@pytest.mark.parametrize("name", POSES)
def test_rigid_transform_mul_matches_reference(g2, name):
    from pybot_b200.geometry import RigidTransform
    p = RigidTransform(g2["pose_%s_xyzw" % name], g2["pose_%s_tvec" % name])
    Y = p * g2["X"]
    ref = g2["mul_%s" % name]
    assert Y.dtype == np.float64 and Y.shape == ref.shape
    # tolerance of the north star: 1e-5 relative (+ abs for cancellation); the
    # FMA chain mirrors OpenBLAS so it is normally bit-exact
    np.testing.assert_allclose(Y, ref, rtol=1e-12, atol=1e-12)
    Y64 = p * g2["X"].astype(np.float64)
    np.testing.assert_allclose(Y64, g2["mul64_%s" % name], rtol=1e-12, atol=1e-12)
    Yf = p.transform_points(g2["X"], out_dtype=np.float32)
    assert Yf.dtype == np.float32
    np.testing.assert_allclose(Yf, ref, rtol=1e-5, atol=1e-5)


def test_transform_ragged_sizes_and_int_input():
    from pybot_b200.geometry import RigidTransform
    p = RigidTransform.from_rpyxyz(*syn.C2_RPYXYZ)
    T = port.rt_matrix(p.quat.q, p.tvec)
    rng = np.random.default_rng(0)
    for n in (0, 1, 3, 127, 128, 129, 1023, 1025, 5000):
        X = rng.normal(size=(n, 3)).astype(np.float32) * 20
        Y = p * X
        assert Y.shape == (n, 3)
        if n:
            np.testing.assert_allclose(Y, port.rt_mul_points(T, X), rtol=1e-12, atol=1e-12)
    Xi = rng.integers(-50, 50, (77, 3))
    np.testing.assert_allclose(p * Xi, port.rt_mul_points(T, Xi), rtol=1e-12, atol=1e-12)


def test_sim3_points():
    from pybot_b200.geometry import Sim3, RigidTransform
    base = RigidTransform.from_rpyxyz(0.3, -0.4, 1.0, 1.0, -2.0, 0.5)
    s = Sim3(xyzw=base.xyzw, tvec=base.tvec, scale=2.5)
    rng = np.random.default_rng(1)
    X = rng.normal(size=(999, 3)).astype(np.float32)
    np.testing.assert_allclose(s * X, port.sim3_mul_points(base.xyzw, base.tvec, 2.5, X),
                               rtol=1e-12, atol=1e-12)
    # consistent with oplus applied to a translation
    o = s.oplus(RigidTransform(tvec=X[0]))
    np.testing.assert_allclose((s * X[:1])[0], o.tvec, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("name", POSES)
def test_project_plain_matches_reference(g2, name):
    cam = _camera(g2, name)
    x = cam.project(g2["X"])
    ref = g2["proj_%s" % name]
    assert x.dtype == np.float32
    np.testing.assert_allclose(x, ref, rtol=1e-5, atol=1e-4)          # north-star bar
    assert _bits_equal(x, ref), "expected bit-exact float32 pixels"
    x64 = cam.project(g2["X"].astype(np.float64))
    assert x64.dtype == np.float64
    np.testing.assert_allclose(x64, g2["proj64_%s" % name], rtol=1e-9, atol=1e-9)   # z ~ 0 points amplify the last-ulp difference of the host Rodrigues round trip


@pytest.mark.parametrize("name", POSES)
def test_project_masks_and_compaction(g2, name, compact_impl):
    cam = _camera(g2, name)
    x, d, v = cam.project(g2["X"], check_bounds=True, check_depth=True, return_depth=True,
                          return_valid=True)
    assert v.dtype == np.bool_ and d.dtype == np.float64
    assert np.array_equal(v, g2["projv_%s_v" % name])                 # masks bit-exact
    assert _bits_equal(x, g2["projv_%s_x" % name])                    # order-preserving compaction
    np.testing.assert_allclose(d, g2["projv_%s_d" % name], rtol=1e-12, atol=1e-12)
    x2, v2 = cam.project(g2["X"], check_depth=True, return_valid=True, min_depth=2.0)
    assert np.array_equal(v2, g2["projd_%s_v" % name])
    assert _bits_equal(x2, g2["projd_%s_x" % name])


def test_project_edge_cases(g2, compact_impl):
    from pybot_b200.vision.camera_utils import Camera, CameraIntrinsic
    cam = Camera.from_intrinsics(CameraIntrinsic(g2["K"], shape=syn.KITTI_SHAPE))
    x = cam.project(g2["edge_X"])
    assert _bits_equal(x, g2["edge_proj"])
    x, d, v = cam.project(g2["edge_X"], check_bounds=True, check_depth=True, return_depth=True,
                          return_valid=True)
    assert np.array_equal(v, g2["edge_v"])
    assert _bits_equal(x, g2["edge_x"])
    np.testing.assert_allclose(d, g2["edge_d"], rtol=1e-12)


def test_project_distortion(g2):
    cam = _camera(g2, "c2", D=g2["dist_D"])
    x = cam.project(g2["X"])
    assert _bits_equal(x, g2["dist_proj"])


def test_check_bounds_without_shape_raises(g2):
    from pybot_b200.vision.camera_utils import Camera, CameraIntrinsic
    cam = Camera.from_intrinsics(CameraIntrinsic(g2["K"]))
    with pytest.raises(ValueError):
        cam.project(g2["X"], check_bounds=True)


def test_project_vs_oracle_seeded_200k(compact_impl):
    """C2 distribution at 200k points against the oracle port (cv2 when importable)."""
    from pybot_b200.vision.camera_utils import Camera, CameraIntrinsic, CameraExtrinsic
    from pybot_b200.geometry import RigidTransform
    X = syn.c2_points(n=200_000)
    p = RigidTransform.from_rpyxyz(*syn.C2_RPYXYZ)
    K = syn.kitti_K()
    cam = Camera.from_intrinsics_extrinsics(CameraIntrinsic(K, shape=syn.KITTI_SHAPE),
                                            CameraExtrinsic.from_rigid_transform(p))
    got = cam.project(X, check_bounds=True, check_depth=True, return_depth=True, return_valid=True)
    ref = port.camera_project(K, np.zeros(5), cam.quat.q, cam.tvec, syn.KITTI_SHAPE, X, True, True,
                              True, True)
    assert np.array_equal(got[2], ref[2])
    assert _bits_equal(got[0], ref[0])
    np.testing.assert_allclose(got[1], ref[1], rtol=1e-12, atol=1e-12)
    assert 0.70 < got[2].mean() < 0.78                      # 73.7 % in view
    ref2 = port.camera_project(K, np.zeros(5), cam.quat.q, cam.tvec, syn.KITTI_SHAPE, X, True, True,
                               True, True, use_cv2=False)    # pure-numpy model of OpenCV
    assert np.array_equal(got[2], ref2[2]) and _bits_equal(got[0], ref2[0])


def test_compaction_full_size_properties(compact_impl):
    """C2 full size (10M): compaction is order-preserving and complete; checked
    against the uncompacted output of the same kernel family."""
    import torch
    from pybot_b200 import ops
    from pybot_b200.geometry import RigidTransform
    g = torch.Generator(device="cuda").manual_seed(1002)
    n = 10_000_000
    X = torch.rand((n, 3), device="cuda", generator=g)
    X[:, 0] = X[:, 0] * 80 - 40
    X[:, 1] = X[:, 1] * 6 - 3
    X[:, 2] = X[:, 2] * 78 + 2
    T = RigidTransform.from_rpyxyz(*syn.C2_RPYXYZ).to_matrix()
    K = syn.kitti_K()
    kw = dict(shape=syn.KITTI_SHAPE, check_bounds=True, check_depth=True, return_depth=True,
              return_valid=True)
    uv_c, d_c, valid, m = ops.project_points(X, T[:3, :3], T[:3, 3], K, compact=True, **kw)
    uv_f, d_f, valid2, m2 = ops.project_points(X, T[:3, :3], T[:3, 3], K, compact=False, **kw)
    assert torch.equal(valid, valid2)
    assert m == int(valid.sum()) == m2
    assert uv_c.shape[0] == m
    assert torch.equal(uv_c, uv_f[valid]) and torch.equal(d_c, d_f[valid])
    inb = (uv_c[:, 0] >= 0) & (uv_c[:, 0] < 1241) & (uv_c[:, 1] >= 0) & (uv_c[:, 1] < 376)
    assert bool(inb.all()) and bool((d_c >= 0.1).all())


@pytest.mark.parametrize("n", [1, 2, 3, 1023, 1024, 1025, 2047, 3073, 444 * 1024 + 7, 2 * 444 * 1024 + 1021,
                               3 * 444 * 1024 + 2])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_compaction_ragged_sizes_equal_boolean_indexing(n, dtype, compact_impl):
    """The pipelined compaction kernel (1, 2, 3 ... generations of tiles per CTA,
    ragged last tile with a byte count that is not a multiple of 16) must equal
    boolean indexing of the uncompacted outputs, bit for bit."""
    import torch
    from pybot_b200 import ops
    from pybot_b200.geometry import RigidTransform
    g = torch.Generator(device="cuda").manual_seed(n)
    X = torch.rand((n, 3), device="cuda", generator=g, dtype=torch.float64)
    X[:, 0] = X[:, 0] * 80 - 40
    X[:, 1] = X[:, 1] * 6 - 3
    X[:, 2] = X[:, 2] * 78 + 2
    X = X.to(torch.float32 if dtype == np.float32 else torch.float64)
    T = RigidTransform.from_rpyxyz(*syn.C2_RPYXYZ).to_matrix()
    kw = dict(shape=syn.KITTI_SHAPE, check_bounds=True, check_depth=True, return_depth=True, return_valid=True)
    uv, dep, valid, count = ops.Projector(T[:3, :3], T[:3, 3], syn.kitti_K(), **kw).run(X)
    uv0, dep0, valid0, count0 = ops.Projector(T[:3, :3], T[:3, 3], syn.kitti_K(), compact=False, **kw).run(X)
    m = int(count.item())
    assert m == int(count0.item()) == int(valid0.sum().item())
    assert torch.equal(valid, valid0)
    keep = valid0.bool()
    assert torch.equal(uv[:m], uv0[keep]) and torch.equal(dep[:m], dep0[keep])


def test_compaction_all_and_none_valid(compact_impl):
    import torch
    from pybot_b200 import ops
    n = 50_000
    X = torch.zeros((n, 3), device="cuda")
    X[:, 2] = 10.0                                    # every point projects to (cx, cy): all valid
    K = syn.kitti_K()
    kw = dict(shape=syn.KITTI_SHAPE, check_bounds=True, check_depth=True, return_depth=True, return_valid=True)
    uv, dep, valid, count = ops.Projector(np.eye(3), np.zeros(3), K, **kw).run(X)
    assert int(count.item()) == n and bool(valid.all())
    assert torch.equal(uv[:, 0], torch.full((n,), np.float32(K[0, 2]), device="cuda"))
    X[:, 2] = -1.0                                    # behind the camera: none valid
    uv, dep, valid, count = ops.Projector(np.eye(3), np.zeros(3), K, **kw).run(X, fresh=True)
    assert int(count.item()) == 0 and not bool(valid.any())


def _near_threshold_points(n, seed, pose_T, K, shape, min_depth=0.1):
    """World points whose projections crowd the decision thresholds of Camera.project: pixels
    within +-0.02 px (and +-1e-6 px) of the four image borders, depths within +-1e-6 of
    min_depth, z ~ 0, z < 0, NaN and inf rows."""
    rng = np.random.default_rng(seed)
    H, W = shape
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    z = rng.uniform(0.5, 80, n)
    u = rng.uniform(0, W, n)
    v = rng.uniform(0, H, n)
    k = n // 8
    spread = np.where(rng.random(n) < 0.5, 0.02, 1e-6)
    edge = rng.uniform(-1, 1, n) * spread
    u[:k] = 0 + edge[:k]
    u[k:2 * k] = W + edge[k:2 * k]
    v[2 * k:3 * k] = 0 + edge[2 * k:3 * k]
    v[3 * k:4 * k] = H + edge[3 * k:4 * k]
    z[4 * k:5 * k] = min_depth + rng.uniform(-1, 1, k) * 1e-6
    z[5 * k:5 * k + 50] = rng.uniform(-1e-7, 1e-7, 50)
    z[5 * k + 50:5 * k + 100] = -rng.uniform(0.1, 50, 50)
    Xc = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1)
    Rm, t = pose_T[:3, :3], pose_T[:3, 3]
    Xw = (Xc - t) @ Rm                     # inverse of Xc = R Xw + t
    Xw[5 * k + 100] = np.nan
    Xw[5 * k + 101, 0] = np.inf
    Xw[5 * k + 102] = 0.0
    Xw[5 * k + 103] = [1e30, -1e30, 1e30]
    return Xw


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("checks", [(True, True), (True, False), (False, True)])
def test_compaction_classifier_at_the_thresholds(dtype, checks, monkeypatch, compact_impl):
    """The fp32 classifier of the compacting kernel must agree with the exact fp64 validity of
    the uncompacted kernel (itself pinned to the reference by golden g2) on points crowded
    around every threshold, with several chunks per launch."""
    import torch
    from pybot_b200 import ops
    from pybot_b200.geometry import RigidTransform
    monkeypatch.setenv("PBK_COMPACT_CHUNK_MB", "1")            # 7+ chunks: the barrier / offset path
    check_bounds, check_depth = checks
    K = syn.kitti_K()
    for pose in (RigidTransform.from_rpyxyz(*syn.C2_RPYXYZ), RigidTransform.identity(),
                 RigidTransform.from_rpyxyz(1.0, -0.7, 2.5, 30.0, -12.0, 4.0)):
        T = pose.to_matrix()
        Xw = _near_threshold_points(600_000, 7, T, K, syn.KITTI_SHAPE)
        X = torch.from_numpy(Xw.astype(dtype)).cuda()
        kw = dict(shape=syn.KITTI_SHAPE, check_bounds=check_bounds, check_depth=check_depth, return_depth=True,
                  return_valid=True)
        uv, dep, valid, count = ops.Projector(T[:3, :3], T[:3, 3], K, **kw).run(X)
        uv0, dep0, valid0, count0 = ops.Projector(T[:3, :3], T[:3, 3], K, compact=False, **kw).run(X)
        m = int(count.item())
        assert torch.equal(valid, valid0)
        assert m == int(count0.item()) == int(valid0.sum().item())
        keep = valid0.bool()
        assert _same_bits(uv[:m], uv0[keep]) and _same_bits(dep[:m], dep0[keep])
        assert 0.05 < m / len(Xw) < 0.95


def test_projector_runs_under_cuda_graph_capture(compact_impl):
    """Projector.run (plain and compacting) captured in a CUDA graph and replayed on new
    inputs: nothing in the call may read host memory asynchronously or allocate."""
    import torch
    from pybot_b200 import ops
    from pybot_b200.geometry import RigidTransform
    T = RigidTransform.from_rpyxyz(*syn.C2_RPYXYZ).to_matrix()
    K = syn.kitti_K()
    n = 300_000
    X = torch.from_numpy(syn.c2_points(n=n)).cuda()
    for kw in (dict(), dict(shape=syn.KITTI_SHAPE, check_bounds=True, check_depth=True, return_depth=True,
                            return_valid=True)):
        pr = ops.Projector(T[:3, :3], T[:3, 3], K, **kw)
        ref = [None if t is None else t.clone() for t in pr.run(X)]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            pr.run(X)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = pr.run(X)
        for t in out:
            if t is not None:
                t.zero_()
        g.replay()
        torch.cuda.synchronize()
        m = int(ref[3].item())
        assert int(out[3].item()) == m
        assert torch.equal(out[0][:m], ref[0][:m])
        if ref[2] is not None:
            assert torch.equal(out[2], ref[2]) and torch.equal(out[1][:m], ref[1][:m])
        # replay on different points through the same buffers
        X2 = X.flip(0).contiguous()
        want = [None if t is None else t.clone() for t in ops.Projector(T[:3, :3], T[:3, 3], K, **kw).run(X2)]
        X.copy_(X2)
        g.replay()
        torch.cuda.synchronize()
        m2 = int(want[3].item())
        assert int(out[3].item()) == m2 and torch.equal(out[0][:m2], want[0][:m2])
