"""Host facade names a one-line import swap relies on: DualQuaternion,
Frustum / HalfPlane / Camera.frustum, save / load, get_calib_params,
StereoCamera.from_intrinsics_extrinsics.  Golden g11 comes from the real reference wherever the
reference can execute the call (oracle/make_golden.py g11)."""
import numpy as np
import pytest

from pybot_b200 import synthetic as syn


@pytest.fixture(scope="module")
def g11(golden_dir):
    return np.load(golden_dir + "/g11_facade.npz")


def test_dual_quaternion_algebra_matches_rigid_transform(g11):
    from pybot_b200.geometry import DualQuaternion, RigidTransform
    assert "xyzw" in str(g11["dq_reference_error"])              # the reference class raises on construction
    a = DualQuaternion(g11["dq_a_xyzw"], g11["dq_a_tvec"])
    b = DualQuaternion(g11["dq_b_xyzw"], g11["dq_b_tvec"])
    np.testing.assert_allclose(a.to_matrix(), g11["dq_a_matrix"], atol=1e-12)
    np.testing.assert_allclose((a * b).to_matrix(), g11["dq_ab_matrix"], atol=1e-6)   # RigidTransform keeps float32 tvec
    np.testing.assert_allclose(a.inverse().to_matrix(), g11["dq_a_inv_matrix"], atol=1e-6)
    np.testing.assert_allclose((a * a.inverse()).to_matrix(), np.eye(4), atol=1e-12)
    R, t = a.to_Rt()
    np.testing.assert_allclose(DualQuaternion.from_Rt(R, t).to_matrix(), a.to_matrix(), atol=1e-12)
    np.testing.assert_allclose(DualQuaternion.from_matrix(g11["dq_a_matrix"]).translation, g11["dq_a_tvec"], atol=1e-12)
    assert np.array_equal(a.rotation.q, a.real.q) and np.array_equal(a.to_xyzw(), a.real.to_xyzw())
    assert a.dot(b) == a.real.dot(b.real) and abs(a.real.q.dot(a.dual)) < 1e-12
    blend = (a * 0.5 + b * 0.5).normalize()                       # dual-quaternion linear blending
    assert abs(np.linalg.norm(blend.real.q) - 1) < 1e-12 and abs(blend.real.q.dot(blend.dual)) < 1e-12
    np.testing.assert_allclose(DualQuaternion.identity().to_matrix(), np.eye(4))
    with pytest.raises(TypeError):
        a * 3
    with pytest.raises(NotImplementedError):
        2.0 * a
    assert isinstance(RigidTransform.identity(), RigidTransform)


@pytest.mark.gpu
def test_frustum_from_pose_and_planes(g11):
    from pybot_b200.geometry import RigidTransform
    from pybot_b200.vision.camera_utils import Frustum, HalfPlane
    pose = RigidTransform(g11["dq_a_xyzw"], g11["dq_a_tvec"])
    f = Frustum.from_pose(pose, zmin=0.5, zmax=4.0)
    np.testing.assert_allclose(f.vertices, g11["frustum_pose"], rtol=1e-12, atol=1e-12)
    pts, normals = f.points_and_normals
    assert pts.shape == (6, 3) and normals.shape == (6, 3) and f.planes.shape == (6, 4)
    np.testing.assert_allclose(np.linalg.norm(normals, axis=1), 1.0, atol=1e-12)
    # every plane contains its own anchor point, and the frustum's centroid lies on one side of all six
    np.testing.assert_allclose(np.sum(f.planes[:, :3] * pts, axis=1) + f.planes[:, 3], 0.0, atol=1e-9)
    side = f.planes[:, :3].dot(f.vertices.mean(axis=0)) + f.planes[:, 3]
    assert np.all(side > 0) or np.all(side < 0)
    with pytest.raises(ValueError):
        Frustum.from_pose(pose, zmin=0.001)
    with pytest.raises(ValueError):
        Frustum.from_pose(pose, fov=np.pi)
    hp = HalfPlane.from_points(np.float64([1, 0, 0]), np.float64([0, 1, 0]), np.float64([0, 0, 0]))
    np.testing.assert_allclose(hp.hp_, [0, 0, 1, 0])
    np.testing.assert_allclose(hp.signed_distance(np.float64([[0, 0, 2], [5, 5, -1]])), [2, -1])


@pytest.mark.gpu
def test_camera_frustum_golden(g11):
    from pybot_b200.geometry import RigidTransform
    from pybot_b200.vision.camera_utils import Camera, CameraExtrinsic, CameraIntrinsic, Frustum
    pose = RigidTransform(g11["dq_a_xyzw"], g11["dq_a_tvec"])
    c = Camera.from_intrinsics_extrinsics(CameraIntrinsic(syn.kitti_K(), D=g11["frustum_D"], shape=(376, 1241)),
                                          CameraExtrinsic.from_rigid_transform(pose))
    f = c.frustum(zmin=0.5, zmax=20.0)
    assert isinstance(f, Frustum)
    np.testing.assert_allclose(f.vertices, g11["frustum_cam"], rtol=1e-9, atol=1e-9)
    quad = np.float64([[100, 50], [100, 300], [900, 300], [900, 50]])
    np.testing.assert_allclose(c.frustum(zmin=1.0, zmax=2.0, pts=quad).vertices, g11["frustum_cam_pts"], rtol=1e-9, atol=1e-9)
    with pytest.raises(TypeError):
        Frustum.from_camera(pose)
    with pytest.raises(NotImplementedError):
        c.save("/tmp/x.yaml")


def test_intrinsic_yaml_matches_reference_file_and_round_trips(g11, tmp_path):
    from pybot_b200.vision.camera_utils import CameraExtrinsic, CameraIntrinsic, StereoCamera, get_calib_params
    ci = CameraIntrinsic(syn.kitti_K(), D=g11["frustum_D"], shape=(376, 1241))
    fn = str(tmp_path / "intrinsic.yaml")
    ci.save(fn)
    assert open(fn).read() == str(g11["intrinsic_yaml"])          # byte-identical to what the reference writes
    back = CameraIntrinsic.load(fn)
    assert np.array_equal(back.K, ci.K) and np.array_equal(back.D, ci.D)
    assert np.array_equal(back.shape, np.int32([1241, 376]))       # the reference's [width, height] order
    ce = CameraExtrinsic(*__import__("pybot_b200.geometry", fromlist=["x"]).RigidTransform.from_rpyxyz(
        0.1, 0.2, -0.3, 1.0, -2.0, 0.5).to_Rt())
    fe = str(tmp_path / "extrinsic.yaml")
    ce.save(fe)
    ce2 = CameraExtrinsic.load(fe)
    np.testing.assert_allclose(ce2.matrix, ce.matrix, atol=1e-7)
    with pytest.raises(RuntimeError, match="Deprecated"):
        get_calib_params(1, 1, 0, 0)
    with pytest.raises(NotImplementedError):
        StereoCamera.from_intrinsics_extrinsics(ci, ci, ce)


def test_epipole_and_essential_decomposition():
    from pybot_b200.vision.camera_utils import compute_epipole, decompose_E, skew
    rng = np.random.default_rng(4)
    t = rng.normal(size=3)
    A = rng.normal(size=(3, 3))
    Q, _ = np.linalg.qr(A)
    R = Q * np.sign(np.linalg.det(Q))
    E = skew(t).dot(R)
    R1, R2, tt = decompose_E(E)
    assert abs(np.linalg.det(R1) - 1) < 1e-9 and abs(np.linalg.det(R2) - 1) < 1e-9
    assert min(np.abs(R1 - R).max(), np.abs(R2 - R).max()) < 1e-6
    assert abs(abs(tt.dot(t / np.linalg.norm(t))) - 1) < 1e-6
    e = compute_epipole(E)
    assert abs(e[2] - 1) < 1e-12 and np.abs(E.dot(e)).max() < 1e-9
