"""Pins the CPU oracle: (a) the port against golden outputs of the REAL reference
(tests/golden, made by oracle/make_golden.py), (b) the pure-numpy models of the
OpenCV arithmetic against the installed cv2, bit for bit where stated."""
import numpy as np
import pytest

from oracle import model, port, c_oracle
from pybot_b200 import synthetic as syn

cv2 = pytest.importorskip("cv2")


def _bits(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype
    na = np.isnan(a) if a.dtype.kind == "f" else np.zeros(a.shape, bool)
    nb = np.isnan(b) if b.dtype.kind == "f" else np.zeros(b.shape, bool)
    assert np.array_equal(na, nb)
    assert np.array_equal(a[~na], b[~nb])


# ----------------------------------------------------------- reconstruct (a6-a8)
def test_reproject_model_bit_exact_vs_cv2_general_Q():
    rng = np.random.default_rng(0)
    Q = rng.normal(size=(4, 4)).astype(np.float32)
    disp = rng.uniform(0.5, 200, (200, 333)).astype(np.float32)
    _bits(model.reproject_image_to_3d(disp, Q), cv2.reprojectImageTo3D(disp, Q))


def test_reproject_golden_reference(golden_dir):
    g = np.load(golden_dir + "/g1_reconstruct.npz")
    Q = g["Q"]
    assert np.array_equal(Q, port.stereo_Q(syn.KITTI_FX, syn.KITTI_CX, syn.KITTI_CY, float(g["baseline"])))
    for name in ("f32", "u8", "s16", "s32"):
        for use_cv2 in (True, False):
            _bits(port.stereo_reconstruct(g["disp_" + name], Q, use_cv2=use_cv2), g["xyz_" + name])
    for s in (1, 2, 3):
        c, x = port.stereo_reconstruct_with_texture(g["disp_f32"], g["im"], Q, sample=s, use_cv2=False)
        assert np.array_equal(c, g["tex%d_im" % s])
        _bits(x, g["tex%d_xyz" % s])
    np.testing.assert_array_equal(port.stereo_reconstruct_sparse(g["sparse_xyd"], Q), g["sparse_xyz"])


def test_reproject_reciprocal_cases(golden_dir):
    g = np.load(golden_dir + "/reproject_divcases.npz")
    _bits(model.reproject_image_to_3d(g["disp"], g["Q"]), g["xyz"])
    _bits(cv2.reprojectImageTo3D(g["disp"], g["Q"]), g["xyz"])


def test_reproject_rejects_float64_like_cv2():
    with pytest.raises(TypeError):
        model.reproject_image_to_3d(np.zeros((2, 2)), np.eye(4))
    with pytest.raises(cv2.error):
        cv2.reprojectImageTo3D(np.zeros((2, 2)), np.eye(4))


# ------------------------------------------------------ transform / project (a1-a5)
def test_port_matches_reference_golden(golden_dir):
    g = np.load(golden_dir + "/g2_transform_project.npz")
    K, X = g["K"], g["X"]
    for name in ("c2", "ident", "big"):
        q, t = g["pose_%s_xyzw" % name], g["pose_%s_tvec" % name]
        np.testing.assert_array_equal(port.rt_mul_points(port.rt_matrix(q, t), X), g["mul_%s" % name])
        cq, ct = g["cam_%s_xyzw" % name], g["cam_%s_tvec" % name]
        for use_cv2 in (True, False):
            _bits(port.camera_project(K, np.zeros(5), cq, ct, syn.KITTI_SHAPE, X, use_cv2=use_cv2),
                  g["proj_%s" % name])
            x, d, v = port.camera_project(K, np.zeros(5), cq, ct, syn.KITTI_SHAPE, X, True, True,
                                          True, True, use_cv2=use_cv2)
            assert np.array_equal(v, g["projv_%s_v" % name])
            _bits(x, g["projv_%s_x" % name])
            np.testing.assert_array_equal(d, g["projv_%s_d" % name])
    ident = port.rt_from_Rt(np.eye(3), np.zeros(3))
    with np.errstate(all="ignore"):
        _bits(port.camera_project(K, np.zeros(5), ident[0], ident[1], syn.KITTI_SHAPE, g["edge_X"],
                                  use_cv2=False), g["edge_proj"])
    cq, ct = g["cam_c2_xyzw"], g["cam_c2_tvec"]
    _bits(port.camera_project(K, g["dist_D"], cq, ct, syn.KITTI_SHAPE, X, use_cv2=False), g["dist_proj"])


def test_project_model_bit_exact_fp64_vs_cv2():
    rng = np.random.default_rng(2)
    X = syn.c2_points(n=100_000).astype(np.float64)
    rvec = np.array([0.3, -0.2, 0.5])
    t = np.array([0.5, -0.1, 1.2])
    K = syn.kitti_K()
    R = model.rodrigues_vec_to_mat(rvec)
    for D in (np.zeros(5), np.array([-0.3, 0.1, 0.001, -0.002, 0.05]),
              np.array([-0.3, 0.1, 0.001, -0.002, 0.05, 0.01, -0.02, 0.003])):
        ref, _ = cv2.projectPoints(X, rvec, t, K, D)
        np.testing.assert_array_equal(model.project_points(X, R, t, K, D), ref.reshape(-1, 2))


def test_rodrigues_models_vs_cv2():
    rng = np.random.default_rng(3)
    for r in [np.array([0.3, -0.2, 0.5]), np.array([1e-9, 2e-9, -1e-9]), np.zeros(3),
              np.array([3.14, 0.01, -0.02]), rng.normal(size=3)]:
        R, J = model.rodrigues_vec_to_mat(r, True)
        Rc, Jc = cv2.Rodrigues(r)
        np.testing.assert_allclose(R, Rc, rtol=0, atol=1e-15)
        np.testing.assert_allclose(J, Jc, rtol=0, atol=1e-12)
        np.testing.assert_allclose(model.rodrigues_mat_to_vec(Rc), cv2.Rodrigues(Rc)[0].ravel(),
                                   rtol=0, atol=1e-12)


# ------------------------------------------------------------------ matching (a9)
def test_hamming_models_vs_cv2_golden(golden_dir):
    g = np.load(golden_dir + "/g3_matching.npz")
    idx, dist = model.hamming_knn(g["le_q"], g["le_t"], 2)
    assert np.array_equal(idx, g["le_idx"]) and np.array_equal(dist, g["le_dist"])
    cidx, cdist = c_oracle.hamming_knn2(g["le_q"], g["le_t"])
    assert np.array_equal(cidx, g["le_idx"]) and np.array_equal(cdist, g["le_dist"])
    _, _, _, m_ok, _ = model.ratio_mutual(g["le_q"], g["le_t"], 0.75)
    cross = np.stack([np.nonzero(m_ok)[0], idx[m_ok, 0]], 1)
    assert np.array_equal(cross, g["le_cross"])                # crossCheck=True semantics
    kfs = [g["kf_0"], g["kf_1"], g["kf_2"]]
    gidx, gdist = c_oracle.hamming_knn2(g["kf_q"], np.concatenate(kfs))
    offs = np.cumsum([0] + [len(k) for k in kfs])
    want = offs[g["kf_img"]] + g["kf_idx"]
    assert np.array_equal(gidx, want) and np.array_equal(gdist, g["kf_dist"])
    pidx, pimg, pdist = port.bf_knn_match(g["kf_q"], kfs, 2)
    assert np.array_equal(pidx, want) and np.array_equal(pimg, g["kf_img"])


def test_hamming_model_general_k_and_mask_vs_cv2_golden(golden_dir):
    """knnMatch beyond k = 2 (golden g12, generated by cv2.BFMatcher): any k, masks, and live cv2
    on a second seeded case."""
    g = np.load(golden_dir + "/g12_knnk.npz")
    q, t, mask = g["q"], g["t"], g["mask"]
    for k in (1, 3, 7, 40):
        idx, dist = model.hamming_knn(q, t, k)
        assert np.array_equal(idx, g["k%d_idx" % k]) and np.array_equal(dist, g["k%d_dist" % k])
    for k in (2, 5):
        idx, dist = model.hamming_knn(q, t, k, mask=mask)
        assert np.array_equal(idx, g["m%d_idx" % k]) and np.array_equal(dist, g["m%d_dist" % k])
        assert np.array_equal((idx >= 0).sum(1), g["m%d_len" % k])
    assert g["m5_len"][3] == 0 and g["m5_len"][5] == 1
    assert np.array_equal(np.nonzero(g["m5_len"])[0], g["m5_compact_query"])
    cuts = g["cuts"]
    for k, name, m in ((4, "c4", mask), (6, "c6", None)):
        idx, dist = model.hamming_knn(q, t, k, mask=m)
        want = np.where(g[name + "_idx"] >= 0, cuts[np.maximum(g[name + "_img"], 0)] + g[name + "_idx"], -1)
        assert np.array_equal(idx, want) and np.array_equal(dist, g[name + "_dist"])
    for k, m in ((1, None), (7, None), (40, None), (5, mask), (900, None), (900, mask)):   # the C restatement
        idx, dist = model.hamming_knn(q, t, k, mask=m)
        cidx, cdist, ccount = c_oracle.hamming_knnk(q, t, k, m)
        assert np.array_equal(cidx, idx) and np.array_equal(cdist, dist) and np.array_equal(ccount, (idx >= 0).sum(1))
    q2, t2 = syn.low_entropy_pair(seed=99, nq=40, nt=300)
    m2 = (np.random.default_rng(5).random((40, 300)) < 0.1).astype(np.uint8)
    res = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q2, t2, k=9, mask=m2)
    idx, dist = model.hamming_knn(q2, t2, 9, mask=m2)
    for i, r in enumerate(res):
        assert [m.trainIdx for m in r] == idx[i][idx[i] >= 0].tolist()
        assert [int(m.distance) for m in r] == dist[i][dist[i] >= 0].tolist()


def test_ratio_mutual_port_equals_model():
    q, t = syn.orb_pair(seed=5, nq=400, nt=700)
    a = model.ratio_mutual(q, t, 0.75)
    b = port.knn_ratio_mutual(q, [t[:300], t[300:]], 0.75)
    for x, y in zip(a, b):
        assert np.array_equal(np.asarray(x).astype(np.int64), np.asarray(y).astype(np.int64))
    assert 0.3 < a[4].mean() < 0.99


# ------------------------------------------------------------------------ BA (a10)
def test_ba_model_vs_cv2_golden(golden_dir):
    g = np.load(golden_dir + "/g5_ba.npz")
    X = g["X"].astype(np.float64)
    for c in range(len(g["rvecs"])):
        uv, Jc, Jp = model.project_points_jacobian(X, g["rvecs"][c], g["tvecs"][c], g["K"])
        np.testing.assert_allclose(uv, g["x"][c], rtol=1e-12, atol=1e-9)
        scale = np.abs(g["Jcam"][c]).max()
        assert np.abs(Jc - g["Jcam"][c]).max() < 1e-9 * scale


def test_ba_port_cv2_equals_model():
    prob = syn.c4_problem(n_cams=12, n_points=600)
    a = port.ba_residual_jacobian(*prob, use_cv2=True)
    b = port.ba_residual_jacobian(*prob, use_cv2=False)
    np.testing.assert_allclose(a[0], b[0], rtol=1e-10, atol=1e-9)
    for x, y in zip(a[1:3], b[1:3]):
        assert np.abs(x - y).max() < 1e-9 * np.abs(x).max()
    np.testing.assert_allclose(a[3], b[3], rtol=1e-10)


def test_synthetic_generators_are_seeded():
    assert np.array_equal(syn.c1_frame(shape=(20, 30)), syn.c1_frame(shape=(20, 30)), equal_nan=True)
    q1, d1 = syn.c3_database(nq=10, n_keyframes=3, per_kf=50)
    q2, d2 = syn.c3_database(nq=10, n_keyframes=3, per_kf=50)
    assert np.array_equal(q1, q2) and np.array_equal(d1, d2)
    rv, tv, K, pts, oc, op, uv = syn.c4_problem(n_cams=10, n_points=100)
    assert len(oc) == 400 and (np.diff(oc) >= 0).all()


def test_grouped_bfmatcher_equals_one_collection(monkeypatch):
    """OpenCV caps a collection at 8191 images (matchers.cpp:856); the port then runs BFMatcher
    over groups of images and merges by (distance, global row).  On tie-heavy descriptors the
    grouped result must equal the single-collection result exactly."""
    pytest.importorskip("cv2")
    from oracle import port
    rng = np.random.default_rng(5)
    q = rng.integers(0, 4, (300, 32), dtype=np.uint8)
    kfs = [rng.integers(0, 4, (int(n), 32), dtype=np.uint8) for n in rng.integers(1, 80, 29)]
    kfs[20][0] = kfs[3][1] = q[7]                         # duplicate across groups: lower image wins
    whole = port.bf_knn_match(q, kfs, 2)
    monkeypatch.setattr(port, "BF_MAX_IMAGES", 6)
    grouped = port.bf_knn_match(q, kfs, 2)
    for a, b in zip(whole, grouped):
        assert np.array_equal(a, b)
    assert (whole[2][:, 0] == whole[2][:, 1]).mean() > 0.2
    for a, b in zip(port.knn_ratio_mutual(q, kfs, 0.8), (lambda: (monkeypatch.undo(), port.knn_ratio_mutual(q, kfs, 0.8))[1])()):
        assert np.array_equal(a, b)
