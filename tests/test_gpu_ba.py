"""GPU parity: BA reprojection residual + Jacobians."""
import numpy as np
import pytest

from oracle import model, port
from pybot_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _close(a, b, rtol=1e-5):
    """1e-5 relative to the largest magnitude of the block (fp32 storage)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape
    scale = np.abs(b).max(axis=-1, keepdims=True) + 1e-30
    assert (np.abs(a - b) / scale).max() < rtol, (np.abs(a - b) / scale).max()


def test_golden_jacobians_at_special_rvecs(golden_dir):
    """|rvec| in {~1, 1e-9, 0, pi-1e-3}: Jacobians cv2.projectPoints returns."""
    from pybot_b200.vision.ba import ba_residual_jacobian
    g = np.load(golden_dir + "/g5_ba.npz")
    X, K = g["X"], g["K"]
    n, C = len(X), len(g["rvecs"])
    obs_cam = np.repeat(np.arange(C, dtype=np.int32), n)
    obs_pt = np.tile(np.arange(n, dtype=np.int32), C)
    uv = np.zeros((C * n, 2), np.float32)
    r, Jc, Jp, cost = ba_residual_jacobian(g["rvecs"], g["tvecs"], K, X, obs_cam, obs_pt, uv)
    assert r.dtype == Jc.dtype == Jp.dtype == np.float32
    assert Jc.shape == (C * n, 2, 6) and Jp.shape == (C * n, 2, 3)
    x_ref = g["x"].reshape(-1, 2)
    np.testing.assert_allclose(r, x_ref, rtol=1e-5, atol=1e-4)           # 1e-4 px
    _close(Jc.reshape(-1, 12), g["Jcam"].reshape(-1, 12))
    np.testing.assert_allclose(cost, np.sum(x_ref.astype(np.float64) ** 2), rtol=1e-9)
    # point block = J_t * R
    for c in range(C):
        R = model.rodrigues_vec_to_mat(g["rvecs"][c].astype(np.float64))
        ref = np.einsum("nka,ab->nkb", g["Jcam"][c][:, :, 3:6], R)
        _close(Jp[c * n:(c + 1) * n].reshape(-1, 6), ref.reshape(-1, 6))


def test_c4_shaped_problem_vs_oracle():
    from pybot_b200.vision.ba import ba_residual_jacobian
    prob = syn.c4_problem(n_cams=40, n_points=5000)
    rv, tv, K, pts, oc, op, uv = prob
    r, Jc, Jp, cost = ba_residual_jacobian(*prob)
    rr, rJc, rJp, rcost = port.ba_residual_jacobian(rv, tv, K, pts, oc, op, uv)
    np.testing.assert_allclose(r, rr, rtol=1e-5, atol=1e-4)
    _close(Jc.reshape(-1, 12), rJc.reshape(-1, 12))
    _close(Jp.reshape(-1, 6), rJp.reshape(-1, 6))
    np.testing.assert_allclose(cost, rcost, rtol=1e-10)
    assert np.abs(rr).mean() < 2.0                                   # 0.5 px noise
    m = port.ba_residual_jacobian(rv, tv, K, pts, oc, op, uv, use_cv2=False)
    np.testing.assert_allclose(r, m[0], rtol=1e-5, atol=1e-4)
    _close(Jc.reshape(-1, 12), m[1].reshape(-1, 12))


def test_one_call_evaluate_equals_two_launch_form():
    """pbk_ba_evaluate (precompute + streaming kernel launched with programmatic stream
    serialization, the kernel's prologue overlapping the precompute) against the two separate
    calls: bit-identical, also when the cameras change between evaluations (the kernel must
    not read camera blocks of the previous evaluation)."""
    import torch
    from pybot_b200 import ops
    prob = syn.c4_problem(n_cams=300, n_points=40000)
    p = ops.BAProblemDevice(*prob)
    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    for it in range(4):
        p.rvec.mul_(1.0 + 0.01 * it)
        p.tvec.add_(0.001 * it)
        p.evaluate()
        got = [t.clone() for t in (p.r, p.Jc, p.Jp, p.cost)]
        p.cam_blocks.zero_()
        p.evaluate(events=ev)
        for a, b in zip(got, (p.r, p.Jc, p.Jp, p.cost)):
            assert torch.equal(a, b)
        assert float(p.cost.item()) > 0


@pytest.mark.parametrize("n_obs", [1, 31, 32, 33, 100, 257])
def test_ragged_observation_counts(n_obs):
    from pybot_b200.vision.ba import ba_residual_jacobian
    rv, tv, K, pts, oc, op, uv = syn.c4_problem(n_cams=6, n_points=300)
    sel = np.random.default_rng(n_obs).permutation(len(oc))[:n_obs]       # also unsorted cameras
    r, Jc, Jp, cost = ba_residual_jacobian(rv, tv, K, pts, oc[sel], op[sel], uv[sel])
    rr, rJc, rJp, rcost = port.ba_residual_jacobian(rv, tv, K, pts, oc[sel], op[sel], uv[sel])
    np.testing.assert_allclose(r, rr, rtol=1e-5, atol=1e-4)
    _close(Jc.reshape(-1, 12), rJc.reshape(-1, 12))
    _close(Jp.reshape(-1, 6), rJp.reshape(-1, 6))
    np.testing.assert_allclose(cost, rcost, rtol=1e-10)


def test_empty_and_bad_indices():
    from pybot_b200.vision.ba import ba_residual_jacobian
    rv, tv, K, pts, oc, op, uv = syn.c4_problem(n_cams=4, n_points=50)
    r, Jc, Jp, cost = ba_residual_jacobian(rv, tv, K, pts, oc[:0], op[:0], uv[:0])
    assert r.shape == (0, 2) and Jc.shape == (0, 2, 6) and cost == 0.0
    bad = oc.copy()
    bad[0] = 99
    with pytest.raises(ValueError):
        ba_residual_jacobian(rv, tv, K, pts, bad, op, uv)


def test_projectpoints_facade_matches_cv2_layout(golden_dir):
    from pybot_b200.vision.ba import project_points
    g = np.load(golden_dir + "/g5_ba.npz")
    x, J = project_points(g["X"], g["rvecs"][0], g["tvecs"][0], g["K"], np.zeros(5))
    n = len(g["X"])
    assert x.shape == (n, 1, 2) and J.shape == (2 * n, 6)
    np.testing.assert_allclose(x.reshape(n, 2), g["x"][0], rtol=1e-5, atol=1e-4)
    _close(J.reshape(n, 12), g["Jcam"][0].reshape(n, 12))
    with pytest.raises(NotImplementedError):
        project_points(g["X"], g["rvecs"][0], g["tvecs"][0], g["K"], np.array([0.1, 0, 0, 0, 0]))


def test_full_size_c4_properties():
    """C4 full size (4M obs): deterministic cost, finite outputs, and the
    Jacobian is consistent with a finite difference of the residual kernel."""
    import torch
    from pybot_b200.ops import BAProblemDevice
    prob = syn.c4_problem()
    rv, tv, K, pts, oc, op, uv = prob
    p = BAProblemDevice(*prob).evaluate()
    c1 = float(p.cost.item())
    r1 = p.r.clone()
    p.evaluate()
    assert float(p.cost.item()) == c1                      # fixed-order reduction
    assert torch.equal(r1, p.r)
    assert bool(torch.isfinite(p.Jc).all()) and bool(torch.isfinite(p.Jp).all())
    np.testing.assert_allclose(c1, float((r1.double() ** 2).sum()), rtol=1e-6)
    assert 0.3 < float(r1.abs().mean()) < 0.6              # N(0, 0.5 px) noise -> 0.4 mean abs
    # finite difference along tvec x of every camera
    eps = 1e-2
    tv2 = tv.copy()
    tv2[:, 0] += eps
    p2 = BAProblemDevice(rv, tv2, K, pts, oc, op, uv, want=("r",)).evaluate()
    fd = (p2.r - r1) / eps
    err = (fd - p.Jc[:, :, 3]).abs().max()
    assert float(err) < 0.5, float(err)                    # second-order term of a 1 cm step
