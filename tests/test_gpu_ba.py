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


def test_one_kernel_evaluate_equals_two_kernel_forms():
    """pbk_ba_evaluate: the default route (stand-alone precompute + streaming kernel with
    programmatic stream serialization), the one-kernel route (every warp derives the camera
    blocks of its own observation range in its prologue) and the two separate calls must be
    bit-identical - also when the cameras change between evaluations (no stale camera block
    may be read) and from a replayed CUDA graph."""
    import torch
    from pybot_b200 import ops
    prob = syn.c4_problem(n_cams=300, n_points=40000)
    p = ops.BAProblemDevice(*prob)
    assert p.sorted_by_camera
    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    for it in range(4):
        p.rvec.mul_(1.0 + 0.01 * it)
        p.tvec.add_(0.001 * it)
        p.cam_blocks.zero_()
        p.evaluate()
        got = [t.clone() for t in (p.r, p.Jc, p.Jp, p.cost)]
        for kw in (dict(one_kernel=True), dict(events=ev)):
            p.cam_blocks.zero_()
            p.evaluate(**kw)
            for a, b in zip(got, (p.r, p.Jc, p.Jp, p.cost)):
                assert torch.equal(a, b)
        p.cam_blocks.zero_()
        p.evaluate_graphed()
        torch.cuda.synchronize()
        for a, b in zip(got, (p.r, p.Jc, p.Jp, p.cost)):
            assert torch.equal(a, b)
        assert float(p.cost.item()) > 0


def test_few_observations_per_camera_and_camera_gaps():
    """One-kernel route when a CTA's observation range spans MANY cameras (2 observations per
    camera) and when some cameras have no observation at all."""
    import torch
    from pybot_b200 import ops
    rng = np.random.default_rng(3)
    rv, tv, K, pts, oc, op, uv = syn.c4_problem(n_cams=600, n_points=3000)
    keep = np.sort(rng.permutation(len(oc))[:1200])                     # ~2 per camera, sorted order kept
    keep = keep[(oc[keep] % 7) != 3]                                     # cameras 3, 10, 17 ... unobserved
    p = ops.BAProblemDevice(rv, tv, K, pts, oc[keep], op[keep], uv[keep])
    assert p.sorted_by_camera
    p.evaluate(one_kernel=True)
    rr, rJc, rJp, rcost = port.ba_residual_jacobian(rv, tv, K, pts, oc[keep], op[keep], uv[keep], use_cv2=False)
    np.testing.assert_allclose(p.r.cpu().numpy(), rr, rtol=1e-5, atol=1e-4)
    _close(p.Jc.cpu().numpy().reshape(-1, 12), rJc.reshape(-1, 12))
    _close(p.Jp.cpu().numpy().reshape(-1, 6), rJp.reshape(-1, 6))
    np.testing.assert_allclose(float(p.cost.item()), rcost, rtol=1e-10)


def _emulated_sharded_ba(prob, world):
    """`world` ranks of the fused-exchange BA on ONE GPU: one stream per rank, plain device
    buffers as the peers' exchange buffers (the protocol is the same over NVLink)."""
    import torch
    from pybot_b200 import ops
    from pybot_b200.distributed import PeerCostExchange, shard_range
    rv, tv, K, pts, oc, op, uv = prob
    xs = PeerCostExchange.emulated(world)
    probs = []
    for r in range(world):
        lo, hi = shard_range(len(oc), r, world)
        probs.append(ops.BAProblemDevice(rv, tv, K, pts, oc[lo:hi], op[lo:hi], uv[lo:hi]))
    streams = [torch.cuda.Stream() for _ in range(world)]
    return xs, probs, streams


def test_fused_cost_exchange_protocol_on_one_gpu():
    """The release/acquire exchange of the sharded BA, both ranks on one device: every rank ends
    with the same global cost, equal to the single-problem cost, call after call (slot parity,
    device-resident sequence), eagerly and from replayed CUDA graphs."""
    import torch
    from pybot_b200 import ops
    prob = syn.c4_problem(n_cams=50, n_points=20_000)
    whole = ops.BAProblemDevice(*prob).evaluate()
    ref = float(whole.cost.item())
    for world in (2, 3):
        xs, probs, streams = _emulated_sharded_ba(prob, world)
        torch.cuda.synchronize()
        for rep in range(5):
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    probs[r].evaluate(exchange=xs[r])
            torch.cuda.synchronize()
            costs = [float(p.cost.item()) for p in probs]
            assert all(c == costs[0] for c in costs), costs
            assert abs(costs[0] - ref) <= 1e-12 * ref
            for x in xs:
                x.check()
        lo = 0
        for p in probs:
            assert torch.equal(p.r, whole.r[lo:lo + p.n_obs]) and torch.equal(p.Jc, whole.Jc[lo:lo + p.n_obs])
            lo += p.n_obs


def test_fused_cost_exchange_timeout_raises():
    """A peer that never arrives: NaN cost, the status word is set and the wrapper raises."""
    import torch
    from pybot_b200 import _ffi
    prob = syn.c4_problem(n_cams=20, n_points=2000)
    xs, probs, streams = _emulated_sharded_ba(prob, 2)
    _ffi.check(_ffi.lib().pbk_set_exchange_timeout_ms(20.0))
    try:
        probs[0].evaluate(exchange=xs[0])                 # rank 1 is starved: it never calls
        torch.cuda.synchronize()
        assert np.isnan(float(probs[0].cost.item()))
        with pytest.raises(RuntimeError, match="did not arrive"):
            xs[0].check()
        xs[0].check()                                     # cleared
    finally:
        _ffi.check(_ffi.lib().pbk_set_exchange_timeout_ms(2000.0))


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
