"""Host-side pose algebra of the facade against known answers produced by the
real reference (tests/golden/g0_host_algebra.npz) and the reference's own
doctest vectors (pybot/geometry/transformations.py:1106-1233)."""
import numpy as np

from pybot_b200.geometry import RigidTransform, Sim3, Quaternion, Pose, rpyxyz
from pybot_b200.geometry import transformations as tf
from pybot_b200.geometry.rodrigues import rodrigues_round_trip, matrix_to_rvec, rvec_to_matrix


def test_reference_doctest_vectors():
    assert np.allclose(tf.quaternion_from_euler(1, 2, 3, "ryxz"), [0.310622, -0.718287, 0.444435, 0.435953])
    assert np.allclose(tf.quaternion_about_axis(0.123, (1, 0, 0)), [0.06146124, 0, 0, 0.99810947])
    assert np.allclose(tf.quaternion_multiply([1, -2, 3, 4], [-5, 6, 7, 8]), [-44, -14, 48, 28])
    assert np.allclose(tf.euler_from_quaternion([0.06146124, 0, 0, 0.99810947]), [0.123, 0, 0])
    R = tf.quaternion_matrix([0.06146124, 0, 0, 0.99810947])
    c, s = np.cos(0.123), np.sin(0.123)
    assert np.allclose(R[:3, :3], [[1, 0, 0], [0, c, -s], [0, s, c]])
    q = tf.quaternion_from_matrix(R)
    assert np.allclose(q, [0.06146124, 0, 0, 0.99810947])


def test_quaternion_self_check_like_reference_main():
    """quaternion.py:233-310: q.rotate(v) == q.R v, inverse == transpose."""
    rng = np.random.default_rng(0)
    for _ in range(20):
        q = Quaternion.from_rpy(*rng.uniform(-3, 3, 3))
        v = rng.normal(size=3)
        assert np.allclose(q.rotate(v), q.R @ v, atol=1e-12)
        assert np.allclose(q.inverse().R, q.R.T, atol=1e-12)
        th, ax = q.to_angle_axis()
        assert np.allclose(Quaternion.from_angle_axis(th, ax).R, q.R, atol=1e-9)


def test_pose_algebra_bit_identical_to_reference(golden_dir):
    g = np.load(golden_dir + "/g0_host_algebra.npz")
    prev = RigidTransform.identity()
    for i, a in enumerate(g["rpyxyz"]):
        p = rpyxyz(*a)
        assert np.array_equal(p.matrix, g["matrix"][i])
        assert np.array_equal(p.inverse().matrix, g["inverse"][i])
        assert np.array_equal((prev * p).matrix, g["compose"][i])
        assert np.array_equal(p.to_vector(), g["vector"][i])
        assert np.array_equal(RigidTransform.from_Rt(*p.to_Rt()).matrix, g["from_Rt"][i])
        assert p.tvec.dtype == np.float32 and p.quat.q.dtype == np.float64
        prev = p
    a = g["rpyxyz"][0]
    s = Sim3(xyzw=rpyxyz(*a).xyzw, tvec=a[3:], scale=2.5)
    assert np.array_equal(s.to_matrix(), g["sim3_matrix"])
    assert np.array_equal(s.oplus(rpyxyz(*g["rpyxyz"][1])).matrix, g["sim3_oplus"])
    for name, q in zip(g["axes_names"], g["axes_q"]):
        assert np.array_equal(tf.quaternion_from_euler(0.3, -0.7, 1.1, axes=str(name)), q), name


def test_misc_api_surface():
    p = RigidTransform.from_rpyxyz(0.1, 0.2, 0.3, 1, 2, 3)
    assert np.allclose(p.to_rpyxyz(), [0.1, 0.2, 0.3, 1, 2, 3], atol=1e-6)
    assert np.allclose(RigidTransform.from_vector(p.to_vector()).matrix, p.matrix)
    assert np.allclose(RigidTransform.from_matrix(p.matrix).matrix, p.matrix, atol=1e-7)
    assert np.allclose((p * p.inverse()).matrix, np.eye(4), atol=1e-6)
    assert np.allclose(p.interpolate(p, 0.5).matrix, p.matrix, atol=1e-6)
    assert isinstance(p.scaled(2.0), Sim3)
    assert Pose.from_rigid_transform(3, p).id == 3
    assert np.allclose(p.wrt(RigidTransform.identity()).matrix, p.matrix)
    assert np.allclose(p.rotate_vec(np.eye(3)), p.R.T, atol=1e-7)
    try:
        3 * p
        assert False
    except (NotImplementedError, TypeError):
        pass
    try:
        p.oplus(3)
        assert False
    except TypeError:
        pass


def test_rodrigues_round_trip_is_a_rotation_and_idempotent():
    rng = np.random.default_rng(1)
    for _ in range(20):
        R = RigidTransform.from_rpyxyz(*rng.uniform(-3, 3, 6)).R
        Rr = rodrigues_round_trip(R)
        assert np.abs(Rr - R).max() < 2e-13
        assert np.allclose(Rr @ Rr.T, np.eye(3), atol=1e-13)
    assert np.array_equal(rvec_to_matrix(np.zeros(3)), np.eye(3))
    assert np.allclose(matrix_to_rvec(np.eye(3)), 0)


def test_camera_constructors_match_reference_values():
    from pybot_b200.vision.camera_utils import StereoCamera, Camera, CameraIntrinsic, get_baseline
    from pybot_b200 import synthetic as syn
    cam = StereoCamera.from_calib_params(syn.KITTI_FX, syn.KITTI_FX, syn.KITTI_CX, syn.KITTI_CY,
                                         baseline_px=syn.KITTI_BASELINE_PX, shape=syn.KITTI_SHAPE)
    assert abs(cam.baseline - 0.5371657188644179) < 1e-15          # probed from the reference
    Q = cam.Q
    assert Q.dtype == np.float32 and Q[3, 2] == np.float32(-1.0 / cam.baseline)
    assert np.allclose(cam.right.tvec, [cam.baseline, 0, 0])
    assert get_baseline(2.0, baseline=0.5) == (0.5, 1.0)
    half = cam.scaled(0.5)
    assert tuple(half.shape) == (188, 620) and half.fx == syn.KITTI_FX * 0.5
    c = Camera.simulate()
    assert c.P.shape == (3, 4) and np.allclose(c.fov, 2 * np.arctan([0.64, 0.48]))
    try:
        StereoCamera(c, c, 0.1)
        assert False
    except ValueError:
        pass
