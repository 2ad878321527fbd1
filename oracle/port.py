"""Restatement ("port") of the reference's own Python on the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function cites the
/root/reference file:line it follows.  Where the reference calls OpenCV, this
port calls the same cv2 function when cv2 is importable (that IS the
reference's algorithm) and falls back to ``oracle.model`` (the pinned numpy
restatement of that OpenCV function) otherwise; ``use_cv2=False`` forces the
model.  tests/test_oracle_pinning.py checks this port against golden outputs
of the real reference (tests/golden/, produced by oracle/make_golden.py).
"""
import math

import numpy as np

from . import model

try:                                    # the reference imports cv2 at
    import cv2                          # camera_utils.py:4
except Exception:                       # pragma: no cover
    cv2 = None

_EPS = np.finfo(float).eps * 4.0        # transformations.py:_EPS


# ---------------------------------------------------------------- geometry
def quaternion_matrix(q):
    """transformations.py:1174-1193 (xyzw)."""
    q = np.array(q[:4], dtype=np.float64, copy=True)
    nq = np.dot(q, q)
    if nq < _EPS:
        return np.identity(4)
    q *= math.sqrt(2.0 / nq)
    q = np.outer(q, q)
    return np.array((
        (1.0 - q[1, 1] - q[2, 2], q[0, 1] - q[2, 3], q[0, 2] + q[1, 3], 0.0),
        (q[0, 1] + q[2, 3], 1.0 - q[0, 0] - q[2, 2], q[1, 2] - q[0, 3], 0.0),
        (q[0, 2] - q[1, 3], q[1, 2] + q[0, 3], 1.0 - q[0, 0] - q[1, 1], 0.0),
        (0.0, 0.0, 0.0, 1.0)), dtype=np.float64)


def quaternion_from_matrix(M):
    """transformations.py:1196-1225 (xyzw)."""
    q = np.empty((4,), dtype=np.float64)
    M = np.asarray(M, dtype=np.float64)[:4, :4]
    t = np.trace(M)
    if t > M[3, 3]:
        q[3] = t
        q[2] = M[1, 0] - M[0, 1]
        q[1] = M[0, 2] - M[2, 0]
        q[0] = M[2, 1] - M[1, 2]
    else:
        i, j, k = 0, 1, 2
        if M[1, 1] > M[0, 0]:
            i, j, k = 1, 2, 0
        if M[2, 2] > M[i, i]:
            i, j, k = 2, 0, 1
        t = M[i, i] - (M[j, j] + M[k, k]) + M[3, 3]
        q[i] = t
        q[j] = M[i, j] + M[j, i]
        q[k] = M[k, i] + M[i, k]
        q[3] = M[k, j] - M[j, k]
    q *= 0.5 / math.sqrt(t * M[3, 3])
    return q


def quaternion_from_rpy(roll, pitch, yaw):
    """transformations.py:1100-1154 for axes='rxyz' (the only sequence the
    path uses: rigid_transform.py:223-226 -> quaternion.py:175-177).
    'rxyz' -> firstaxis 2, parity 1, repetition 0, frame 1."""
    i, j, k = 2, 1, 0                     # _NEXT_AXIS walk for (2, 1, 0, 1)
    ai, aj, ak = yaw, -pitch, roll        # frame: swap ; parity: negate aj
    ai, aj, ak = ai / 2.0, aj / 2.0, ak / 2.0
    ci, si = math.cos(ai), math.sin(ai)
    cj, sj = math.cos(aj), math.sin(aj)
    ck, sk = math.cos(ak), math.sin(ak)
    cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
    q = np.empty(4)
    q[i] = cj * sc - sj * cs
    q[j] = cj * ss + sj * cc
    q[k] = cj * cs - sj * sc
    q[3] = cj * cc + sj * ss
    q[j] *= -1
    return q


def normalize_quaternion(q):
    """quaternion.py:18-27,49-53: float64, renormalise iff |norm-1| > 1e-6."""
    q = np.array(q, np.float64)
    n = np.linalg.norm(q)
    if abs(n - 1) > 1e-6:
        q /= n
    return q


def rt_matrix(xyzw, tvec, scale=None):
    """RigidTransform.to_matrix rigid_transform.py:202-206 with the float32
    translation of __init__ (:116-119); Sim3.to_matrix :318-323 if scale."""
    T = quaternion_matrix(normalize_quaternion(xyzw))
    T[:3, 3] = np.float32(tvec)
    if scale is not None:
        T[3, 3] = 1.0 / scale
        T[:3, :3] /= scale
    return T


def rt_from_Rt(R, t):
    """RigidTransform.from_Rt rigid_transform.py:228-232 -> (xyzw, tvec f32)."""
    T = np.eye(4)
    T[:3, :3] = np.asarray(R, np.float64)
    return normalize_quaternion(quaternion_from_matrix(T)), np.float32(t)


def rt_mul_points(T, X):
    """RigidTransform.__mul__ ndarray branch, rigid_transform.py:139-141."""
    X = np.hstack([X, np.ones((len(X), 1))]).T
    return (np.dot(T, X).T)[:, :3]


def sim3_mul_points(xyzw, tvec, scale, X):
    """Sim3 point application.  NOT in the reference (rigid_transform.py:339-342
    raises NotImplementedError); the build adopts Y = s*R*X + t, consistent
    with Sim3.oplus (:344-348).  PARITY UNPINNED."""
    T = rt_matrix(xyzw, tvec)
    R = T[:3, :3]
    return scale * (np.asarray(X, np.float64) @ R.T) + T[:3, 3]


# ------------------------------------------------------------ Camera.project
def camera_project(K, D, xyzw, tvec, shape, X, check_bounds=False,
                   check_depth=False, return_depth=False, return_valid=False,
                   min_depth=0.1, use_cv2=True):
    """Camera.project camera_utils.py:687-734 for a Camera whose extrinsic
    state is (xyzw, tvec) (Camera.__init__ :647-649 routes (R, t) through
    RigidTransform.from_Rt, so the state is always a quaternion + f32 tvec).
    """
    T = rt_matrix(xyzw, tvec)
    R, t = T[:3, :3].copy(), T[:3, 3].copy()                   # to_Rt :214-217
    X = np.asarray(X)
    if use_cv2 and cv2 is not None:
        rvec, _ = cv2.Rodrigues(R)                             # :694
        proj, _ = cv2.projectPoints(X, rvec, t, np.asarray(K, np.float64),
                                    np.asarray(D, np.float64))  # :697
        x = proj.reshape(-1, 2)
    else:
        Rr = model.rodrigues_vec_to_mat(model.rodrigues_mat_to_vec(R))
        x = model.project_points(X, Rr, t, K, D)
    if return_depth or check_depth:
        depths = camera_depth(xyzw, tvec, X)                   # :705-706
    if check_depth:
        valid = depths >= min_depth                            # :708-710
    else:
        valid = np.ones(len(x), dtype=bool)                    # :712
    if check_bounds:
        if shape is None:
            raise ValueError(
                'check_bounds cannot proceed. Camera.shape is not set')
        valid = np.bitwise_and(valid, np.bitwise_and(          # :720-725
            np.bitwise_and(x[:, 0] >= 0, x[:, 0] < shape[1]),
            np.bitwise_and(x[:, 1] >= 0, x[:, 1] < shape[0])))
    output = [x[valid]]                                        # :727
    if return_depth:
        output.append(depths[valid])
    if return_valid:
        output.append(valid)
    return output if len(output) > 1 else output[0]


def camera_depth(xyzw, tvec, X):
    """Camera.depth_from_projection camera_utils.py:736-743: the extrinsics
    property (:683-685) rebuilds CameraExtrinsic(self.R, self.t), i.e. one more
    matrix -> quaternion -> matrix round trip, then __mul__ and column 2."""
    T = rt_matrix(xyzw, tvec)
    q2, t2 = rt_from_Rt(T[:3, :3], np.float32(tvec))
    return rt_mul_points(rt_matrix(q2, t2), X)[:, 2]


# -------------------------------------------------------------- StereoCamera
def stereo_Q(fx, cx, cy, baseline):
    """StereoCamera.Q camera_utils.py:884-896 (float32)."""
    q43 = -1.0 / baseline
    return np.float32([[-1, 0, 0, cx], [0, -1, 0, cy], [0, 0, 0, -fx],
                       [0, 0, q43, 0]])


def get_baseline(fx, baseline=None, baseline_px=None):
    """camera_utils.py:93-105."""
    assert baseline is not None or baseline_px is not None
    if baseline is not None:
        baseline_px = baseline * fx
    else:
        baseline = baseline_px / fx
    return baseline, baseline_px


def stereo_reconstruct(disp, Q, use_cv2=True):
    """StereoCamera.reconstruct camera_utils.py:964-969."""
    if use_cv2 and cv2 is not None:
        return cv2.reprojectImageTo3D(disp, Q)
    return model.reproject_image_to_3d(disp, Q)


def stereo_reconstruct_with_texture(disp, im, Q, sample=1, use_cv2=True):
    """StereoCamera.reconstruct_with_texture camera_utils.py:982-990."""
    assert im.ndim == 3
    X = stereo_reconstruct(disp, Q, use_cv2)
    im_pub = np.copy(im[::sample, ::sample]).reshape(-1, 3)
    X_pub = np.copy(X[::sample, ::sample]).reshape(-1, 3)
    return im_pub, X_pub


def stereo_reconstruct_sparse(xyd, Q):
    """StereoCamera.reconstruct_sparse camera_utils.py:971-980."""
    N = xyd.shape[0]
    xyd1 = np.hstack([xyd, np.ones(shape=(N, 1))])
    XYZW = np.dot(Q, xyd1.T).T
    W = (XYZW[:, 3]).reshape(-1, 1)
    return (XYZW / W)[:, :3]


# ------------------------------------------------- matching (a9, via OpenCV)
# ------------------------------------------------- next rows
def depth_mesh(K, shape, skip=1):
    """DepthCamera._build_mesh camera_utils.py:1020-1030 -> (xs1d, ys1d)."""
    K = np.asarray(K, np.float64)
    H, W = shape
    xs, ys = np.arange(0, W), np.arange(0, H)
    return ((xs - K[0, 2]) * 1.0 / K[0, 0])[::skip], ((ys - K[1, 2]) * 1.0 / K[1, 1])[::skip]


def depth_reconstruct(K, shape, skip, depth):
    """DepthCamera.reconstruct camera_utils.py:1032-1036."""
    xs1d, ys1d = depth_mesh(K, shape, skip)
    xs, ys = np.meshgrid(xs1d, ys1d)
    depth_sampled = np.asarray(depth)[::skip, ::skip]
    assert depth_sampled.shape == xs.shape
    return np.dstack([xs * depth_sampled, ys * depth_sampled, depth_sampled])


def depth_reconstruct_sparse(K, pts, depth):
    """DepthCamera.reconstruct_sparse camera_utils.py:1038-1041 (fx on both axes)."""
    K = np.asarray(K, np.float64)
    fx, cx, cy = K[0, 0], K[0, 2], K[1, 2]
    return np.vstack([(pts[:, 0] - cx) * 1.0 / fx * depth,
                      (pts[:, 1] - cy) * 1.0 / fx * depth,
                      depth]).T


def sceneflow_process(K, D, shape, dX, use_cv2=True):
    """SceneFlow.process optflow_utils.py:129-150 for a camera with identity
    extrinsics (:122)."""
    H, W = int(shape[0]), int(shape[1])
    xs_, ys_ = np.meshgrid(np.arange(0, W), np.arange(0, H))
    grid_ = np.dstack([xs_, ys_]).astype(np.float32)
    x = camera_project(K, D, [0., 0., 0., 1.], [0., 0., 0.], (H, W), np.asarray(dX).reshape(-1, 3),
                       use_cv2=use_cv2)
    with np.errstate(invalid="ignore"):
        x2 = x.astype(np.int32)                                # :137
    valid = np.bitwise_and(np.bitwise_and(x2[:, 0] >= 0, x2[:, 0] < W),
                           np.bitwise_and(x2[:, 1] >= 0, x2[:, 1] < H))
    xs, ys = x2[:, 0], x2[:, 1]
    gxs, gys = xs_.ravel(), ys_.ravel()
    new_grid = np.copy(grid_) * np.nan
    new_grid[ys[valid], xs[valid]] = grid_[gys[valid], gxs[valid]]
    return new_grid - grid_


def sampson_error(F, pts1, pts2):
    """camera_utils.py:52-68, verbatim arithmetic (N x N product and its diagonal)."""
    x1 = np.hstack([pts1, np.ones((len(pts1), 1))]).T          # unproject_points :37-39
    x2 = np.hstack([pts2, np.ones((len(pts2), 1))]).T
    Fx1 = np.dot(F, x1)
    Fx2 = np.dot(F, x2)
    denom = Fx1[0]**2 + Fx1[1]**2 + Fx2[0]**2 + Fx2[1]**2
    return (np.diag(x1.T.dot(Fx2)))**2 / denom


def epipolar_line(F_10, x_0):
    """camera_utils.py:352-357: the same cv2 call the reference makes."""
    return cv2.computeCorrespondEpilines(x_0.reshape(-1, 1, 2), 1, F_10).reshape(-1, 3)


def triangulate_points(P1, P2, pts1, pts2, use_cv2=True):
    """camera_utils.py:108-116."""
    if use_cv2 and cv2 is not None:
        X = cv2.triangulatePoints(P1, P2, pts1.reshape(-1, 1, 2), pts2.reshape(-1, 1, 2)).T
    else:
        X = model.triangulate_points(P1, P2, pts1, pts2).T
    return X[:, :3] / X[:, 3].reshape(-1, 1)


def undistort_points(K, D, pts, use_cv2=True):
    """CameraIntrinsic.undistort_points camera_utils.py:529-538."""
    if use_cv2 and cv2 is not None:
        out = cv2.undistortPoints(pts.reshape(-1, 1, 2).astype(np.float32), K, D, P=K)
        return out.reshape(-1, 2)
    return model.undistort_points(pts, K, D, P=K)


def quaternion_rotate(q, v):
    """Quaternion.rotate quaternion.py:69-85."""
    qx, qy, qz, qw = q
    ab, ac, ad = qw * qx, qw * qy, qw * qz
    nbb, bc, bd = -qx * qx, qx * qy, qx * qz
    ncc, cd, ndd = -qy * qy, qy * qz, -qz * qz
    return np.array((2 * ((ncc + ndd) * v[0] + (bc - ad) * v[1] + (ac + bd) * v[2]) + v[0],
                     2 * ((ad + bc) * v[0] + (nbb + ndd) * v[1] + (cd - ab) * v[2]) + v[1],
                     2 * ((bd - ac) * v[0] + (ab + cd) * v[1] + (nbb + ncc) * v[2]) + v[2]))


def camera_ray(K, D, pts, undistort=True, rotate_xyzw=None, normalize=False, use_cv2=True):
    """CameraIntrinsic.ray camera_utils.py:499-517 (rotate: extrinsics.rotate_vec,
    rigid_transform.py:181-186)."""
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    upts = undistort_points(K, D, pts, use_cv2) if undistort else pts
    ret = np.hstack([np.hstack([(upts[:, :1] - cx) / fx, (upts[:, 1:2] - cy) / fy]),
                     np.ones((len(upts), 1))])                 # unproject_points :37-39
    if rotate_xyzw is not None:
        ret = np.vstack([quaternion_rotate(rotate_xyzw, v_) for v_ in ret])
    if normalize:
        ret = ret / np.linalg.norm(ret, axis=1)[:, np.newaxis]
    return ret


def intrinsic_reconstruct(K, D, xyZ, undistort=True, use_cv2=True):
    """CameraIntrinsic.reconstruct camera_utils.py:519-524."""
    return camera_ray(K, D, xyZ[:, :2], undistort=undistort, use_cv2=use_cv2) * xyZ[:, 2:3]


def wire_colors(carr, flip_rb=False):
    """get_color_arr externals/draw_helpers.py:35-53 for an array argument."""
    carr = carr.reshape(-1, 3) if carr.ndim == 3 else carr      # reshape_arr :29-33
    carr = carr.copy()
    if flip_rb:
        b, r = carr[:, 0], carr[:, 2]
        carr[:, 0], carr[:, 2] = r.copy(), b.copy()
    return carr.astype(np.float32) / 255.0 if carr.dtype == np.uint8 else carr.astype(np.float32)


# OpenCV asserts imgCount * 2^18 < INT_MAX in knnMatchImpl (matchers.cpp:856): a collection may
# hold at most 8191 images.  The C3 database has 10 000 keyframes.
BF_MAX_IMAGES = 8000


def _bf_knn_match_grouped(q, train_list, k):
    """A collection of more than 8191 images: BFMatcher over consecutive groups of images, and
    the groups' top-k merged by (distance, global row) - the order BFMatcher itself produces
    inside a collection (ties to the lowest (imgIdx, trainIdx))."""
    offs = np.concatenate([[0], np.cumsum([len(t) for t in train_list])])
    parts = []
    for g0 in range(0, len(train_list), BF_MAX_IMAGES):
        idx, img, dist = bf_knn_match(q, train_list[g0:g0 + BF_MAX_IMAGES], k)
        idx = np.where(idx >= 0, idx + offs[g0], -1)
        img = np.where(img >= 0, img + g0, -1)
        parts.append((idx, img, dist))
    idx = np.concatenate([p[0] for p in parts], 1)
    img = np.concatenate([p[1] for p in parts], 1)
    dist = np.concatenate([p[2] for p in parts], 1)
    big = np.iinfo(np.int64).max
    key = np.where(idx >= 0, dist.astype(np.int64) * (1 << 40) + idx, big)
    order = np.argsort(key, axis=1, kind="stable")[:, :k]
    take = lambda a: np.take_along_axis(a, order, 1)      # noqa: E731
    return take(idx), take(img), take(dist)


def bf_knn_match(q, train_list, k=2, threads=None):
    """cv2.BFMatcher(NORM_HAMMING).knnMatch over a keyframe collection
    (no call site in the reference).  A single matrix of
    >= 262144 rows is rejected by OpenCV, hence the list form.  Returns
    global idx (Nq,k) int64 (= keyframe row offset + trainIdx), imgIdx,
    dist (Nq,k) int32; -1 padded."""
    if cv2 is None:
        t = np.concatenate(train_list, 0)
        idx, dist = model.hamming_knn(q, t, k)
        return idx.astype(np.int64), np.zeros_like(idx), dist
    if threads is not None:
        cv2.setNumThreads(threads)
    if len(train_list) > BF_MAX_IMAGES:
        return _bf_knn_match_grouped(q, train_list, k)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    bf.add([np.ascontiguousarray(t) for t in train_list])
    res = bf.knnMatch(np.ascontiguousarray(q), k=k)
    offs = np.concatenate([[0], np.cumsum([len(t) for t in train_list])])
    Nq = len(q)
    idx = np.full((Nq, k), -1, np.int64)
    img = np.full((Nq, k), -1, np.int32)
    dist = np.full((Nq, k), -1, np.int32)
    for i, row in enumerate(res):
        for j, m in enumerate(row):
            idx[i, j] = offs[m.imgIdx] + m.trainIdx
            img[i, j] = m.imgIdx
            dist[i, j] = int(m.distance)
    return idx, img, dist


def knn_ratio_mutual(q, train_list, ratio=0.75, threads=None):
    """Forward k=2 + Lowe ratio + mutual check with cv2 as the engine: the
    user-side pipeline the north star describes.  Same outputs as
    model.ratio_mutual but with global indices over the collection."""
    idx, img, dist = bf_knn_match(q, train_list, 2, threads)
    Nq = len(q)
    has2 = idx[:, 1] >= 0
    ratio_ok = np.zeros(Nq, bool)
    ratio_ok[has2] = (dist[has2, 0].astype(np.float64)
                      < ratio * dist[has2, 1].astype(np.float64))
    mutual_ok = np.zeros(Nq, bool)
    has1 = idx[:, 0] >= 0
    if has1.any():
        offs = np.concatenate([[0], np.cumsum([len(t) for t in train_list])])
        g = idx[has1, 0]
        im = np.searchsorted(offs, g, side="right") - 1
        rows = np.stack([train_list[a][b] for a, b in zip(im, g - offs[im])])
        ridx, _, _ = bf_knn_match(rows, [np.asarray(q)], 1, threads)
        mutual_ok[np.nonzero(has1)[0]] = ridx[:, 0] == np.nonzero(has1)[0]
    return idx, dist, ratio_ok, mutual_ok, ratio_ok & mutual_ok


# --------------------------------------------------- BA (a10, via OpenCV)
def ba_residual_jacobian(rvecs, tvecs, K, points, obs_cam, obs_pt, obs_uv,
                         use_cv2=True):
    """Reprojection residual r = project(X) - z and its Jacobians, using the
    Jacobian cv2.projectPoints computes and Camera.project discards
    (camera_utils.py:697).  Looped per camera like a pybot user would.
    Returns r (O,2), J_cam (O,2,6), J_pt (O,2,3) float64 and cost = sum r^2.
    """
    O = len(obs_cam)
    r = np.empty((O, 2))
    Jc = np.empty((O, 2, 6))
    Jp = np.empty((O, 2, 3))
    K = np.asarray(K, np.float64)
    order = np.argsort(obs_cam, kind="stable")
    cams, starts = np.unique(obs_cam[order], return_index=True)
    ends = list(starts[1:]) + [O]
    for c, s, e in zip(cams, starts, ends):
        sel = order[s:e]
        X = np.asarray(points, np.float64)[obs_pt[sel]]
        rv = np.asarray(rvecs[c], np.float64)
        tv = np.asarray(tvecs[c], np.float64)
        if use_cv2 and cv2 is not None:
            x, J = cv2.projectPoints(X.reshape(-1, 1, 3), rv, tv, K, np.zeros(5))
            x = x.reshape(-1, 2)
            J = J.reshape(-1, 2, 15)
            Jc[sel] = J[:, :, :6]
            R, _ = cv2.Rodrigues(rv)
            Jp[sel] = np.einsum("nka,ab->nkb", J[:, :, 3:6], R)
        else:
            x, jc, jp = model.project_points_jacobian(X, rv, tv, K)
            Jc[sel], Jp[sel] = jc, jp
        r[sel] = x - np.asarray(obs_uv, np.float64)[sel]
    return r, Jc, Jp, float(np.sum(r * r))
