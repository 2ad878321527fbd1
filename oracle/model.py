"""Pure-numpy models of the third-party arithmetic on the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference (spillai/pybot) delegates the arithmetic of this path to OpenCV
(pinned ``opencv3=3.1.0`` in /root/reference/environment.yml:53; installed and
probed here: opencv-python-headless 4.13.0.92):

    cv2.reprojectImageTo3D   pybot/vision/camera_utils.py:968, :987
    cv2.Rodrigues            pybot/vision/camera_utils.py:694
    cv2.projectPoints        pybot/vision/camera_utils.py:697
    cv2.BFMatcher            (no call site in the reference)
    cv2.undistortPoints      pybot/vision/camera_utils.py:536-537
    cv2.triangulatePoints    pybot/vision/camera_utils.py:114-115

OpenCV's sources are not under /root/reference, so its published algorithms are
restated here and pinned bit-for-bit against the installed cv2 by
tests/test_oracle_pinning.py (every model below states what "pinned" means
for it).  The CUDA kernels are compared against these models and against cv2
itself where cv2 is importable.
"""
import numpy as np


# --------------------------------------------------------------------------
# cv2.reprojectImageTo3D(disp, Q, handleMissingValues=False)
# --------------------------------------------------------------------------
def reproject_image_to_3d(disp, Q):
    """Model of cv2.reprojectImageTo3D (call sites camera_utils.py:968,987).

    Pinned bit-exact (all float32 bit patterns) on random dense 4x4 Q and on
    the reference's sparse Q, including the planted cases that separate
    "divide" from "multiply by reciprocal":

        Qd   = double(Q)
        h_i  = ((Qd[i,0]*x + Qd[i,1]*y) + Qd[i,2]*double(d)) + Qd[i,3]   (no FMA)
        out  = float32( double(float32(h_0..2)) * (1.0 / h_3) )

    i.e. the numerator is rounded to float32 FIRST, then scaled by the fp64
    reciprocal of W and rounded again.  Accepts uint8/int16/int32/float32
    disparity like OpenCV (integer types are converted to float32 first);
    float64 raises like OpenCV does.
    """
    disp = np.asarray(disp)
    if disp.ndim != 2:
        raise ValueError("disparity must be HxW")
    if disp.dtype not in (np.uint8, np.int16, np.int32, np.float32):
        raise TypeError("unsupported disparity dtype %s" % disp.dtype)
    Qd = np.asarray(Q).astype(np.float64)
    if Qd.shape != (4, 4):
        raise ValueError("Q must be 4x4")
    H, W = disp.shape
    d = disp.astype(np.float32).astype(np.float64)
    x = np.arange(W, dtype=np.float64)[None, :]
    y = np.arange(H, dtype=np.float64)[:, None]
    with np.errstate(all="ignore"):
        h = [((Qd[i, 0] * x + Qd[i, 1] * y) + Qd[i, 2] * d) + Qd[i, 3]
             for i in range(4)]
        iw = 1.0 / h[3]
        out = np.empty((H, W, 3), np.float32)
        for i in range(3):
            out[..., i] = (h[i].astype(np.float32).astype(np.float64) * iw
                           ).astype(np.float32)
    return out


# --------------------------------------------------------------------------
# cv2.Rodrigues
# --------------------------------------------------------------------------
def rodrigues_vec_to_mat(rvec, jacobian=False):
    """Rotation vector -> matrix (and dR/dr, 3x9, OpenCV layout).

    Published algorithm of cv::Rodrigues: theta = |r|; theta < DBL_EPSILON ->
    identity; else R = c*I + (1-c)*rr^T + s*[r]_x with r normalised.
    Pinned against cv2.Rodrigues to <= 4 ulp of fp64 on R and 1e-12 on dR/dr
    (transcendental libm differences prevent a bit-exact pin).
    jacobian[i, 3*a+b] = d R[a,b] / d r_i.
    """
    r = np.asarray(rvec, np.float64).reshape(3)
    theta = float(np.sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]))
    I = np.eye(3)
    if theta < np.finfo(np.float64).eps:
        R = I.copy()
        if not jacobian:
            return R
        J = np.zeros((3, 9))
        J[0, 5], J[0, 7] = -1, 1
        J[1, 2], J[1, 6] = 1, -1
        J[2, 1], J[2, 3] = -1, 1
        return R, J
    c, s = np.cos(theta), np.sin(theta)
    c1 = 1.0 - c
    itheta = 1.0 / theta
    u = r * itheta
    rrt = np.outer(u, u)
    rx = np.array([[0, -u[2], u[1]], [u[2], 0, -u[0]], [-u[1], u[0], 0]])
    R = c * I + c1 * rrt + s * rx
    if not jacobian:
        return R
    ux, uy, uz = u
    drrt = np.array([
        [ux + ux, uy, uz, uy, 0, 0, uz, 0, 0],
        [0, ux, 0, ux, uy + uy, uz, 0, uz, 0],
        [0, 0, ux, 0, 0, uy, ux, uy, uz + uz]], np.float64)
    d_r_x = np.array([
        [0, 0, 0, 0, 0, -1, 0, 1, 0],
        [0, 0, 1, 0, 0, 0, -1, 0, 0],
        [0, -1, 0, 1, 0, 0, 0, 0, 0]], np.float64)
    J = np.empty((3, 9))
    for i in range(3):
        ri = u[i]
        a0 = -s * ri
        a1 = (s - 2 * c1 * itheta) * ri
        a2 = c1 * itheta
        a3 = (c - s * itheta) * ri
        a4 = s * itheta
        J[i] = (a0 * I.ravel() + a1 * rrt.ravel() + a2 * drrt[i]
                + a3 * rx.ravel() + a4 * d_r_x[i])
    return R, J


def rodrigues_mat_to_vec(R):
    """Rotation matrix -> vector.  cv::Rodrigues first projects R onto SO(3)
    by SVD (R <- U V^T), then r = axis * angle from the skew part, with the
    theta ~ pi branch taken from the symmetric part.
    Pinned against cv2.Rodrigues to 1e-12 absolute.
    """
    R = np.asarray(R, np.float64).reshape(3, 3)
    U, _, Vt = np.linalg.svd(R)
    R = U @ Vt
    r = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = np.sqrt((r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) * 0.25)
    c = (R[0, 0] + R[1, 1] + R[2, 2] - 1) * 0.5
    c = min(1.0, max(-1.0, c))
    theta = np.arccos(c)
    if s < 1e-5:
        if c > 0:
            return np.zeros(3)
        t = (R[0, 0] + 1) * 0.5
        rx = np.sqrt(max(t, 0.0))
        t = (R[1, 1] + 1) * 0.5
        ry = np.sqrt(max(t, 0.0)) * (-1.0 if R[0, 1] < 0 else 1.0)
        t = (R[2, 2] + 1) * 0.5
        rz = np.sqrt(max(t, 0.0)) * (-1.0 if R[0, 2] < 0 else 1.0)
        if (abs(rx) < abs(ry) and abs(rx) < abs(rz)
                and (R[1, 2] > 0) != (ry * rz > 0)):
            rz = -rz
        v = np.array([rx, ry, rz])
        v *= theta / np.sqrt(v @ v)
        return v
    vth = 1.0 / (2.0 * s)
    vth *= theta
    return r * vth


# --------------------------------------------------------------------------
# cv2.projectPoints(X, rvec, tvec, K, D) -> (x, J)
# --------------------------------------------------------------------------
def project_points(X, R, t, K, D=None, out_dtype=None):
    """Pinhole (+ Brown/rational distortion) projection given the 3x3 R that
    cv::projectPoints derives from rvec.

    Pinned BIT-EXACT in fp64 against cv2.projectPoints (1e6 points, with and
    without distortion) - OpenCV's per-point loop, no FMA, in this order:

        x = ((R00*X + R01*Y) + R02*Z) + t0      (same for y, z)
        z = z ? 1/z : 1 ; x *= z ; y *= z
        r2 = x*x + y*y ; r4 = r2*r2 ; r6 = r4*r2
        a1 = 2*x*y ; a2 = r2 + 2*x*x ; a3 = r2 + 2*y*y
        cdist   = 1 + k1*r2 + k2*r4 + k3*r6
        icdist2 = 1/(1 + k4*r2 + k5*r4 + k6*r6)
        xd = x*cdist*icdist2 + p1*a1 + p2*a2 ; yd = y*cdist*icdist2 + p1*a3 + p2*a1
        u = xd*fx + cx ; v = yd*fy + cy         (skew K[0,1] ignored)

    D = [k1,k2,p1,p2,k3,(k4,k5,k6)] as in construct_D (camera_utils.py:129-133).
    Output dtype follows X (float32 in -> one rounding to float32), like cv2.
    """
    X = np.asarray(X)
    odt = out_dtype or (np.float32 if X.dtype == np.float32 else np.float64)
    Xd = X.astype(np.float64).reshape(-1, 3)
    R = np.asarray(R, np.float64).reshape(3, 3)
    t = np.asarray(t, np.float64).reshape(3)
    K = np.asarray(K, np.float64)
    k = np.zeros(8)
    if D is not None:
        Dv = np.asarray(D, np.float64).ravel()
        if Dv.size not in (0, 4, 5, 8):
            raise ValueError("distortion must have 4, 5 or 8 coefficients")
        k[:Dv.size] = Dv
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    with np.errstate(all="ignore"):
        x = ((R[0, 0] * Xd[:, 0] + R[0, 1] * Xd[:, 1]) + R[0, 2] * Xd[:, 2]) + t[0]
        y = ((R[1, 0] * Xd[:, 0] + R[1, 1] * Xd[:, 1]) + R[1, 2] * Xd[:, 2]) + t[1]
        z = ((R[2, 0] * Xd[:, 0] + R[2, 1] * Xd[:, 1]) + R[2, 2] * Xd[:, 2]) + t[2]
        z = np.where(z != 0, 1.0 / z, 1.0)
        x = x * z
        y = y * z
        r2 = x * x + y * y
        r4 = r2 * r2
        r6 = r4 * r2
        a1 = 2 * x * y
        a2 = r2 + 2 * x * x
        a3 = r2 + 2 * y * y
        cdist = 1 + k[0] * r2 + k[1] * r4 + k[4] * r6
        icdist2 = 1.0 / (1 + k[5] * r2 + k[6] * r4 + k[7] * r6)
        xd = x * cdist * icdist2 + k[2] * a1 + k[3] * a2
        yd = y * cdist * icdist2 + k[2] * a3 + k[3] * a1
        u = xd * fx + cx
        v = yd * fy + cy
    return np.stack([u, v], 1).astype(odt)


def project_points_jacobian(X, rvec, t, K):
    """x and the rvec/tvec/point Jacobians of the distortion-free pinhole
    projection: the (2N x 15) Jacobian that cv2.projectPoints
    computes and Camera.project discards (camera_utils.py:697), columns 0:6,
    plus the point block d x / d X = J_t . R.

    Returns x (N,2) f64, J_cam (N,2,6) f64 [rvec | tvec], J_pt (N,2,3) f64.
    Pinned against cv2.projectPoints's jacobian to 1e-9 relative (the rvec
    block goes through dR/dr, which carries libm differences).
    """
    Xd = np.asarray(X, np.float64).reshape(-1, 3)
    R, dRdr = rodrigues_vec_to_mat(rvec, jacobian=True)
    t = np.asarray(t, np.float64).reshape(3)
    K = np.asarray(K, np.float64)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    Xc = Xd @ R.T + t
    z = np.where(Xc[:, 2] != 0, 1.0 / Xc[:, 2], 1.0)
    x = Xc[:, 0] * z
    y = Xc[:, 1] * z
    uv = np.stack([x * fx + cx, y * fy + cy], 1)
    N = len(Xd)
    Jt = np.zeros((N, 2, 3))
    Jt[:, 0, 0] = fx * z
    Jt[:, 0, 2] = -fx * x * z
    Jt[:, 1, 1] = fy * z
    Jt[:, 1, 2] = -fy * y * z
    # d(R X)/d r_i = (dR/dr_i) X ; dRdr[i] is dR/dr_i flattened row-major
    dRX = np.einsum("iab,nb->nai", dRdr.reshape(3, 3, 3), Xd)   # (N, 3, 3): [a, i]
    Jr = np.einsum("nka,nai->nki", Jt, dRX)
    Jcam = np.concatenate([Jr, Jt], axis=2)
    Jpt = np.einsum("nka,ab->nkb", Jt, R)
    return uv, Jcam, Jpt


# --------------------------------------------------------------------------
# cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch
# --------------------------------------------------------------------------
_POPC8 = np.array([bin(i).count("1") for i in range(256)], np.uint8)


def hamming_matrix(q, t):
    """(Nq, Nt) int32 Hamming distances between uint8 descriptor rows."""
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    out = np.zeros((len(q), len(t)), np.int32)
    for b in range(q.shape[1]):
        out += _POPC8[q[:, b, None] ^ t[None, :, b]]
    return out


def hamming_knn(q, t, k=2, mask=None):
    """Model of BFMatcher(NORM_HAMMING).knnMatch(q, t, k, mask): the k smallest
    distances per query, ties broken by LOWEST train index.
    mask (Nq, Nt), non-zero = allowed pair (OpenCV's convention); excluded
    pairs never match.  Rows are padded with idx=-1, dist=-1 when fewer than
    k rows are allowed (Nt < k, or masked out).
    Pinned bit-exact (indices and distances) against cv2 on tie-heavy
    low-entropy descriptors, for k up to 40 and with masks (golden g12).
    """
    Nq = len(q)
    idx = np.full((Nq, k), -1, np.int32)
    dist = np.full((Nq, k), -1, np.int32)
    if len(t) == 0 or Nq == 0:
        return idx, dist
    excluded = 1 << 20                     # above any distance
    for s in range(0, Nq, 256):            # bound the distance matrix
        d = hamming_matrix(q[s:s + 256], t)
        if mask is not None:
            d[np.asarray(mask[s:s + 256]) == 0] = excluded
        order = np.argsort(d, axis=1, kind="stable")[:, :k]
        kk = order.shape[1]
        dd = np.take_along_axis(d, order, 1)
        idx[s:s + 256, :kk] = np.where(dd < excluded, order, -1)
        dist[s:s + 256, :kk] = np.where(dd < excluded, dd, -1)
    return idx, dist


def ratio_mutual(q, t, ratio=0.75):
    """Forward kNN (k=2) + Lowe ratio + mutual (cross) check, the pipeline the
    north star names.  Returns idx (Nq,2), dist (Nq,2), ratio_ok (Nq,),
    mutual_ok (Nq,), valid = ratio_ok & mutual_ok.

    ratio:  float(d1) < ratio * float(d2) in IEEE fp64; fewer than 2
            neighbours fails.
    mutual: rev[idx1] == i where rev[j] is the lowest-index best QUERY for
            train row j  ==  cv2.BFMatcher(crossCheck=True).match semantics.
    """
    idx, dist = hamming_knn(q, t, 2)
    Nq = len(q)
    ratio_ok = np.zeros(Nq, bool)
    mutual_ok = np.zeros(Nq, bool)
    has2 = idx[:, 1] >= 0
    ratio_ok[has2] = (dist[has2, 0].astype(np.float64)
                      < ratio * dist[has2, 1].astype(np.float64))
    has1 = idx[:, 0] >= 0
    if has1.any():
        rows = np.unique(idx[has1, 0])
        ridx, _ = hamming_knn(np.asarray(t)[rows], q, 1)
        rev = dict(zip(rows.tolist(), ridx[:, 0].tolist()))
        for i in np.nonzero(has1)[0]:
            mutual_ok[i] = rev[int(idx[i, 0])] == i
    return idx, dist, ratio_ok, mutual_ok, ratio_ok & mutual_ok


# --------------------------------------------------------------------------
# cv2.undistortPoints(pts_f32, K, D, P=P)   (default criteria: 5 iterations)
# --------------------------------------------------------------------------
def undistort_points(pts, K, D, P=None, iters=5):
    """Published algorithm (imgproc undistort.dispatch / cvUndistortPointsInternal):
    x0 = (u-cx)/fx (as a product with 1/fx); five fixed-point iterations
    x <- (x0 - tangential(x)) * icdist(x) with icdist the ratio of the two radial
    polynomials (Horner form); a negative icdist aborts to x0; then the 3x3 new
    camera matrix with a perspective divide; one rounding to float32.  Pinned:
    bit-identical to cv2 4.13 on 200 k points for 4/5/8-coefficient models
    (tests/test_next_rows.py)."""
    K = np.asarray(K, np.float64)
    P = K if P is None else np.asarray(P, np.float64)
    k = np.zeros(8)
    k[:len(np.ravel(D))] = np.ravel(D)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    ifx, ify = 1.0 / fx, 1.0 / fy
    p = np.asarray(pts).reshape(-1, 2).astype(np.float32).astype(np.float64)
    x, y = (p[:, 0] - cx) * ifx, (p[:, 1] - cy) * ify
    x0, y0 = x.copy(), y.copy()
    done = np.zeros(len(x), bool)
    with np.errstate(all="ignore"):
        for _ in range(iters):
            r2 = x * x + y * y
            icd = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2)
            neg = (icd < 0) & ~done
            dx = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x)
            dy = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y
            xn = np.where(neg, x0, (x0 - dx) * icd)
            yn = np.where(neg, y0, (y0 - dy) * icd)
            x, y = np.where(done, x, xn), np.where(done, y, yn)
            done |= neg
        xx = P[0, 0] * x + P[0, 1] * y + P[0, 2]
        yy = P[1, 0] * x + P[1, 1] * y + P[1, 2]
        ww = 1.0 / (P[2, 0] * x + P[2, 1] * y + P[2, 2])
        return np.stack([xx * ww, yy * ww], 1).astype(np.float32)


# --------------------------------------------------------------------------
# cv2.triangulatePoints(P1, P2, x1, x2)
# --------------------------------------------------------------------------
def triangulate_points(P1, P2, x1, x2):
    """Published algorithm (calib3d triangulate.cpp): per point the 4x4 DLT matrix
    [x*P[2]-P[0]; y*P[2]-P[1]] for both views, SVD, last right singular vector;
    output in float32 only if both point sets are float32.  An SVD has no
    bit-exact contract: pinned to 1e-9 relative on well-conditioned pairs."""
    P1, P2 = np.asarray(P1, np.float64), np.asarray(P2, np.float64)
    a, b = np.asarray(x1).reshape(-1, 2), np.asarray(x2).reshape(-1, 2)
    f32 = a.dtype == np.float32 and b.dtype == np.float32
    a, b = a.astype(np.float64), b.astype(np.float64)
    A = np.stack([a[:, :1] * P1[2] - P1[0], a[:, 1:] * P1[2] - P1[1],
                  b[:, :1] * P2[2] - P2[0], b[:, 1:] * P2[2] - P2[1]], axis=1)     # (n,4,4)
    X = np.linalg.svd(A)[2][:, 3, :]
    return (X.astype(np.float32) if f32 else X).T


def epipolar_lines(F, pts):
    """cv2.computeCorrespondEpilines(pts, 1, F) restated (calib3d fundam.cpp): per point
    l = F [x y 1]^T in double, products and sums rounded separately left to right, scaled by
    1 / sqrt(a^2 + b^2) (by 1 when that is zero), one rounding to the point dtype."""
    pts = np.asarray(pts)
    f = np.asarray(F, np.float64).reshape(9)
    x, y = pts[:, 0].astype(np.float64), pts[:, 1].astype(np.float64)
    l = [(f[3 * r] * x + f[3 * r + 1] * y) + f[3 * r + 2] for r in range(3)]
    nu = l[0] * l[0] + l[1] * l[1]
    with np.errstate(divide="ignore"):
        nu = np.where(nu != 0, 1.0 / np.sqrt(nu), 1.0)
    return np.stack([l[0] * nu, l[1] * nu, l[2] * nu], 1).astype(pts.dtype)
