"""Seeded synthetic inputs for the five BASELINE.json configs.

KITTI calibration constants are the ones the reference ships in
pybot/utils/dataset/kitti.py:110-112.  Every generator takes scale-down
arguments so the parity tests can run the same distributions at sizes the CPU
oracle finishes in seconds.
"""
import numpy as np

KITTI_FX = 718.856
KITTI_CX = 607.1928
KITTI_CY = 185.2157
KITTI_BASELINE_PX = 386.1448
KITTI_SHAPE = (376, 1241)
KITTI_BASELINE = KITTI_BASELINE_PX / KITTI_FX

UHD_FX = 2224.35
UHD_CX = 1919.5
UHD_CY = 1079.5
UHD_BASELINE = 0.5372
UHD_SHAPE = (2160, 3840)


def kitti_K():
    K = np.eye(3)
    K[0, 0] = K[1, 1] = KITTI_FX
    K[0, 2], K[1, 2] = KITTI_CX, KITTI_CY
    return K


def disparity_image(rng, shape, fx, baseline, specials=True):
    """fx*b/Z with Z~U(4,80) m, 5% zeros, 1% -1 (SGBM invalid) and, if asked,
    four planted specials (NaN, +inf, 1e-30, -0.0) in the first row."""
    H, W = shape
    Z = rng.uniform(4.0, 80.0, size=(H, W))
    disp = (fx * baseline / Z).astype(np.float32)
    u = rng.random((H, W))
    disp[u < 0.05] = 0.0
    disp[(u >= 0.05) & (u < 0.06)] = -1.0
    if specials and W >= 8:
        disp[0, 1] = np.nan
        disp[0, 2] = np.inf
        disp[0, 3] = 1e-30
        disp[0, 4] = -0.0
    return disp


def c1_frame(seed=1001, shape=KITTI_SHAPE):
    rng = np.random.default_rng(seed)
    return disparity_image(rng, shape, KITTI_FX, KITTI_BASELINE)


def orb_pair(seed=1001, nq=2000, nt=2000, match_frac=0.6, flip_p=0.08,
             n_dup=16):
    """q uniform; t: match_frac rows are distinct query rows with
    Binomial(256, flip_p) bit flips, the rest uniform, n_dup exact duplicate
    rows, shuffled."""
    rng = np.random.default_rng(seed + 7)
    q = rng.integers(0, 256, size=(nq, 32), dtype=np.uint8)
    t = rng.integers(0, 256, size=(nt, 32), dtype=np.uint8)
    nm = min(int(match_frac * nt), nq)
    src = rng.permutation(nq)[:nm]
    flips = rng.random((nm, 256)) < flip_p
    t[:nm] = q[src] ^ np.packbits(flips, axis=1)
    n_dup = min(n_dup, nt // 2)
    if n_dup:
        a = rng.permutation(nt)[:2 * n_dup]
        t[a[n_dup:]] = t[a[:n_dup]]
    t = t[rng.permutation(nt)]
    return q, np.ascontiguousarray(t)


def low_entropy_pair(seed=7, nq=500, nt=700, nbytes=4):
    """Descriptors with only `nbytes` random bytes: ~40% of queries have a
    tied top-2, exercising lowest-train-index tie-breaking."""
    rng = np.random.default_rng(seed)
    q = np.zeros((nq, 32), np.uint8)
    t = np.zeros((nt, 32), np.uint8)
    q[:, :nbytes] = rng.integers(0, 256, (nq, nbytes))
    t[:, :nbytes] = rng.integers(0, 256, (nt, nbytes))
    return q, t


def c2_points(seed=1002, n=10_000_000):
    rng = np.random.default_rng(seed)
    X = np.empty((n, 3), np.float32)
    X[:, 0] = rng.uniform(-40, 40, n)
    X[:, 1] = rng.uniform(-3, 3, n)
    X[:, 2] = rng.uniform(2, 80, n)
    return X


C2_RPYXYZ = (0.01, -0.02, 0.03, 0.5, -0.1, 1.2)


def c3_database(seed=1003, nq=2000, n_keyframes=10_000, per_kf=2000,
                n_revisit=5, revisit_frac=0.7, flip_p=0.08):
    """Queries + a keyframe database of n_keyframes x per_kf descriptors as one
    contiguous (n_keyframes*per_kf, 32) uint8 array.  n_revisit keyframes hold
    noisy copies of revisit_frac of the queries."""
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 256, size=(nq, 32), dtype=np.uint8)
    nt = n_keyframes * per_kf
    db = np.empty((nt, 32), np.uint8)
    step = 1 << 20
    for s in range(0, nt, step):
        e = min(nt, s + step)
        db[s:e] = rng.integers(0, 256, size=(e - s, 32), dtype=np.uint8)
    n_revisit = min(n_revisit, n_keyframes)
    kfs = rng.permutation(n_keyframes)[:n_revisit]
    nm = min(int(revisit_frac * nq), per_kf)
    for kf in kfs:
        src = rng.permutation(nq)[:nm]
        rows = kf * per_kf + rng.permutation(per_kf)[:nm]
        flips = rng.random((nm, 256)) < flip_p
        db[rows] = q[src] ^ np.packbits(flips, axis=1)
    return q, db


def c4_problem(seed=1004, n_cams=1000, n_points=1_000_000, obs_per_point=4,
               noise_px=0.5):
    """Forward trajectory of n_cams cameras (1 m steps, yaw random walk
    sigma=0.01), points 5-60 m ahead of a random camera, each observed by
    obs_per_point consecutive cameras; obs sorted by camera.
    Returns rvecs (C,3) f32, tvecs (C,3) f32, K, points (P,3) f32,
    obs_cam (O,) i32, obs_pt (O,) i32, obs_uv (O,2) f32."""
    rng = np.random.default_rng(seed)
    yaw = np.cumsum(rng.normal(0, 0.01, n_cams))
    pos = np.zeros((n_cams, 3))
    pos[1:, 0] = np.cumsum(np.sin(yaw[:-1]))
    pos[1:, 2] = np.cumsum(np.cos(yaw[:-1]))
    # world->camera: R_cw = Ry(yaw)^T ; rvec = (0, -yaw, 0) ; t = -R_cw * pos
    c, s = np.cos(yaw), np.sin(yaw)
    tvecs = np.stack([-(c * pos[:, 0] - s * pos[:, 2]),
                      np.zeros(n_cams),
                      -(s * pos[:, 0] + c * pos[:, 2])], 1)
    rvecs = np.stack([rng.normal(0, 1e-3, n_cams), -yaw,
                      rng.normal(0, 1e-3, n_cams)], 1)
    rvecs = rvecs.astype(np.float32)
    tvecs = tvecs.astype(np.float32)
    opp = min(obs_per_point, n_cams)
    cam0 = rng.integers(0, n_cams - opp + 1, n_points)
    depth = rng.uniform(5, 60, n_points)
    lat = rng.uniform(-0.4, 0.4, n_points) * depth
    ver = rng.uniform(-0.12, 0.12, n_points) * depth
    pts_c = np.stack([lat, ver, depth], 1)
    # camera -> world with the exact yaw-only pose of cam0
    c0, s0 = c[cam0], s[cam0]
    pts_w = np.stack([c0 * pts_c[:, 0] + s0 * pts_c[:, 2],
                      pts_c[:, 1],
                      -s0 * pts_c[:, 0] + c0 * pts_c[:, 2]], 1) + pos[cam0]
    points = pts_w.astype(np.float32)
    obs_cam = (cam0[:, None] + np.arange(opp)[None, :]).ravel().astype(np.int32)
    obs_pt = np.repeat(np.arange(n_points, dtype=np.int32), opp)
    order = np.argsort(obs_cam, kind="stable")
    obs_cam, obs_pt = obs_cam[order], obs_pt[order]
    K = kitti_K()
    # noisy observations from an fp64 pinhole (generator-side only)
    th = np.linalg.norm(rvecs.astype(np.float64), axis=1)
    uv = np.empty((len(obs_cam), 2), np.float32)
    chunk = 1 << 20
    for a in range(0, len(obs_cam), chunk):
        b = min(len(obs_cam), a + chunk)
        cc = obs_cam[a:b]
        r = rvecs[cc].astype(np.float64)
        X = points[obs_pt[a:b]].astype(np.float64)
        t_ = th[cc][:, None]
        k = r / np.maximum(t_, 1e-300)
        ct, st = np.cos(t_), np.sin(t_)
        Xr = X * ct + np.cross(k, X) * st + k * np.sum(k * X, 1, keepdims=True) * (1 - ct)
        Xc = Xr + tvecs[cc].astype(np.float64)
        uv[a:b, 0] = K[0, 0] * Xc[:, 0] / Xc[:, 2] + K[0, 2]
        uv[a:b, 1] = K[1, 1] * Xc[:, 1] / Xc[:, 2] + K[1, 2]
    uv += rng.normal(0, noise_px, uv.shape).astype(np.float32)
    return rvecs, tvecs, K, points, obs_cam, obs_pt, uv


def c5_batch(seed=1005, batch=64, shape=UHD_SHAPE):
    """batch x (H,W) f32 disparity + batch x (H,W,3) u8 BGR."""
    rng = np.random.default_rng(seed)
    H, W = shape
    disp = np.empty((batch, H, W), np.float32)
    im = np.empty((batch, H, W, 3), np.uint8)
    for b in range(batch):
        disp[b] = disparity_image(rng, shape, UHD_FX, UHD_BASELINE,
                                  specials=(b == 0))
        im[b] = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    return disp, im
