"""Randomised pinning of the oracle's numpy / C restatements against the OpenCV calls the
reference makes (CPU, hypothesis): the restatements are what the GPU tests compare with at
sizes and in places where cv2 cannot be used (the GPU box has no /root/reference, and the
models say in which ORDER OpenCV rounds), so they are checked here beyond the fixed
golden vectors - bit for bit."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import c_oracle, model

cv2 = pytest.importorskip("cv2")

_SET = dict(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck))


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype
    nan = np.isnan(b)
    assert np.array_equal(np.isnan(a), nan) and np.array_equal(a[~nan], b[~nan])


@settings(**_SET)
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(1, 300), f32=st.booleans(), scale=st.sampled_from([1e-6, 1.0, 1e4]))
def test_epipolar_lines_model_equals_cv2(seed, n, f32, scale):
    rng = np.random.default_rng(seed)
    F = rng.normal(size=(3, 3)) * scale
    x = rng.uniform(-2000, 2000, (n, 2)).astype(np.float32 if f32 else np.float64)
    want = cv2.computeCorrespondEpilines(x.reshape(-1, 1, 2), 1, F).reshape(-1, 3)
    _same(model.epipolar_lines(F, x), want)


@settings(**_SET)
@given(seed=st.integers(0, 2**31 - 1), nq=st.integers(1, 60), nt=st.integers(2, 400), nbytes=st.integers(1, 32))
def test_hamming_knn_models_equal_bfmatcher(seed, nq, nt, nbytes):
    """Low-entropy descriptors (only `nbytes` random bytes) make ties the common case: the numpy
    model and the C restatement must break them like cv2.BFMatcher (lowest train index)."""
    rng = np.random.default_rng(seed)
    q = np.zeros((nq, 32), np.uint8)
    t = np.zeros((nt, 32), np.uint8)
    q[:, :nbytes] = rng.integers(0, 4, (nq, nbytes))
    t[:, :nbytes] = rng.integers(0, 4, (nt, nbytes))
    res = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
    widx = np.array([[m.trainIdx for m in r] for r in res], np.int64)
    wdist = np.array([[int(m.distance) for m in r] for r in res], np.int32)
    idx, dist = model.hamming_knn(q, t, 2)
    assert np.array_equal(np.asarray(idx, np.int64), widx) and np.array_equal(np.asarray(dist, np.int32), wdist)
    cidx, cdist = c_oracle.hamming_knn2(q, t)
    assert np.array_equal(cidx, widx) and np.array_equal(cdist, wdist)


@settings(**_SET)
@given(seed=st.integers(0, 2**31 - 1), nq=st.integers(1, 40), nt=st.integers(1, 300), nbytes=st.integers(1, 32),
       k=st.integers(1, 40), density=st.sampled_from([None, 0.02, 0.3, 1.0]))
def test_hamming_knn_model_general_k_and_mask_equals_bfmatcher(seed, nq, nt, nbytes, k, density):
    """any k, optional per-pair mask (knnMatch's full signature): rows shorter than k when fewer
    rows are allowed, empty when none is; order and ties as cv2.BFMatcher."""
    rng = np.random.default_rng(seed)
    q = np.zeros((nq, 32), np.uint8)
    t = np.zeros((nt, 32), np.uint8)
    q[:, :nbytes] = rng.integers(0, 4, (nq, nbytes))
    t[:, :nbytes] = rng.integers(0, 4, (nt, nbytes))
    mask = None if density is None else (rng.random((nq, nt)) < density).astype(np.uint8)
    res = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=k, mask=mask)
    idx, dist = model.hamming_knn(q, t, k, mask=mask)
    assert len(res) == nq
    for i, r in enumerate(res):
        assert [m.trainIdx for m in r] == idx[i][idx[i] >= 0].tolist()
        assert [int(m.distance) for m in r] == dist[i][dist[i] >= 0].tolist()
        assert (idx[i] >= 0).sum() == (dist[i] >= 0).sum() == min(k, nt if mask is None else int(mask[i].sum()))
    cidx, cdist, ccount = c_oracle.hamming_knnk(q, t, k, mask)                  # the C restatement too
    assert np.array_equal(cidx, idx) and np.array_equal(cdist, dist) and np.array_equal(ccount, (idx >= 0).sum(1))


@settings(**_SET)
@given(seed=st.integers(0, 2**31 - 1), h=st.integers(1, 24), w=st.integers(1, 40),
       dt=st.sampled_from([np.uint8, np.int16, np.int32, np.float32]))
def test_reproject_model_equals_cv2(seed, h, w, dt):
    rng = np.random.default_rng(seed)
    Q = np.float32([[1, 0, 0, -rng.uniform(1, 900)], [0, 1, 0, -rng.uniform(1, 500)],
                    [0, 0, 0, rng.uniform(100, 3000)], [0, 0, 1.0 / rng.uniform(0.05, 2.0), 0]])
    if dt == np.float32:
        d = rng.uniform(-2, 300, (h, w)).astype(np.float32)
        flat = d.reshape(-1)
        for v in (0.0, -1.0, np.nan, np.inf, 1e-30):
            flat[rng.integers(0, flat.size)] = v
    else:
        info = np.iinfo(dt)
        d = rng.integers(max(info.min, -5), min(info.max, 4000), (h, w)).astype(dt)
    _same(model.reproject_image_to_3d(d, Q), cv2.reprojectImageTo3D(d, Q))


@settings(**_SET)
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(1, 200), dist=st.booleans())
def test_project_points_model_equals_cv2_float64(seed, n, dist):
    rng = np.random.default_rng(seed)
    X = rng.uniform([-40, -3, 0.5], [40, 3, 80], (n, 3))
    rvec = rng.normal(size=3) * rng.choice([1e-9, 0.3, 1.5])
    t = rng.normal(size=3)
    K = np.array([[rng.uniform(300, 2500), 0, rng.uniform(100, 2000)], [0, rng.uniform(300, 2500), rng.uniform(100, 1200)],
                  [0, 0, 1.0]])
    D = rng.normal(size=5) * 0.05 if dist else np.zeros(5)
    R = cv2.Rodrigues(rvec)[0]
    want = cv2.projectPoints(X.reshape(-1, 1, 3), rvec, t, K, D)[0].reshape(-1, 2)
    _same(model.project_points(X, R, t, K, D), want)
