"""Pinhole / stereo camera models with GPU projection and reconstruction.

Same class, method and argument names as the hot-path part of
/root/reference/pybot/vision/camera_utils.py (:93-133, :400-1007), so
``from pybot_b200.vision.camera_utils import Camera, StereoCamera`` is a
one-line swap.  Intrinsics / extrinsics bookkeeping is host-side scalar work;
the array work runs on the B200:

    Camera.project                      -> pbk_project_points   (was cv2.projectPoints
                                           + np.dot + masks + fancy indexing, :687-734)
    Camera.depth_from_projection,
    CameraExtrinsic.c2w / w2c           -> pbk_transform_points (was np.dot, :585-595,:736-743)
    StereoCamera.reconstruct*           -> pbk_reproject_disp   (was cv2.reprojectImageTo3D
                                           + np.copy passes, :964-990)

Unlike the reference, failures raise (the reference prints and continues at
:696-701).  Everything needs a CUDA device; there is no CPU fallback.
"""
import numpy as np

from ..geometry.rigid_transform import RigidTransform
from ..geometry.rodrigues import rodrigues_round_trip
from .. import ops


def get_baseline(fx, baseline=None, baseline_px=None):
    """baseline [m] <-> baseline [px] through the focal length (reference :93-105)."""
    assert baseline is not None or baseline_px is not None
    if baseline is not None:
        baseline_px = baseline * fx
    else:
        baseline = baseline_px / fx
    return baseline, baseline_px


def get_calib_params(fx, fy, cx, cy, baseline=None, baseline_px=None):
    """Deprecated in the reference too (:148-149 raises before doing anything)."""
    raise RuntimeError("Deprecated, see camera_utils.StereoCamera")


def _save_yaml(filename, d):
    import os
    import yaml
    with open(os.path.expanduser(filename), "w") as fp:
        fp.write(yaml.dump(d, default_flow_style=False))          # the layout AttrDict.save_yaml writes


def _load_yaml(filename):
    import os
    import yaml
    with open(os.path.expanduser(filename), "r") as fp:
        return yaml.safe_load(fp)


def construct_K(fx=500.0, fy=500.0, cx=319.5, cy=239.5):
    K = np.eye(3)
    K[0, 0], K[1, 1] = fx, fy
    K[0, 2], K[1, 2] = cx, cy
    return K


def construct_D(k1=0, k2=0, k3=0, p1=0, p2=0):
    """OpenCV order [k1, k2, p1, p2, k3] (reference :129-133)."""
    return np.float64([k1, k2, p1, p2, k3])


def check_visibility(camera, pts_w, zmin=0, zmax=100):
    """Field-of-view test (reference :245-266) as ONE fused kernel: transform into the
    camera's frame, direction angles and the four comparisons on the device; only the
    boolean mask comes back."""
    pts = pts_w.reshape(-1, 3)
    M = camera.matrix                                       # c2w(X) = camera * X (reference :585-589)
    fov = camera.fov                                        # float32 pair, like the reference's
    return ops.fov_mask(pts, M[:3, :3], M[:3, 3], float(fov[0] * 0.5), float(fov[1] * 0.5), zmin, zmax)


def get_median_depth(camera, pts, subsample=10):
    """Median z of camera * pts[::subsample] (reference :269-276): only the z column of the
    sampled points is computed and brought back."""
    M = camera.matrix
    return np.median(np.asarray(ops.transform_depth(pts, M[2, :3], M[2, 3], subsample)))


def get_bounded_projection(camera, pts, subsample=10, return_depth=False):
    """Reference :279-288, with the bounds test and compaction fused into the
    projection kernel instead of a second numpy pass."""
    return camera.project(np.asarray(pts[::subsample], np.float32), check_bounds=True,
                          return_valid=True)


def triangulate_points(cam1, pts1, cam2, pts2):
    """Triangulate 2-D correspondences from two views (reference :108-116); the
    per-point DLT + 4x4 SVD of cv2.triangulatePoints runs on the GPU."""
    assert isinstance(cam1, Camera) and isinstance(cam2, Camera)
    assert isinstance(pts1, np.ndarray) and isinstance(pts2, np.ndarray)
    return ops.triangulate_points(np.asarray(cam1.P), np.asarray(cam2.P), pts1, pts2)


def colvec(vec):
    """Reference :17-19."""
    return vec.reshape(-1, 1)


def rowvec(vec):
    """Reference :22-24."""
    return vec.reshape(1, -1)


def unproject(a):
    """Homogenise a vector (reference :27-29)."""
    return np.append(a, 1)


def project(a):
    """De-homogenise a vector (reference :32-34)."""
    return a[:-1] / float(a[-1])


def project_points(pts):
    """Rows [x, y, z] -> [x / z, y / z] (reference :42-44)."""
    return pts[:, :2] / colvec(pts[:, 2])


def skew(a):
    """[a]_x with a x v = [a]_x v (reference :47-49)."""
    return np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])


def compute_essential(F, K):
    """E = K^T F K (reference :229-231)."""
    return (K.T).dot(F).dot(K)


def compute_epipole(F):
    """Right epipole e with F e = 0, scaled so that e[2] = 1 (what reference :234-242 documents;
    its own body needs an un-imported `linalg`).  The null vector is the last right singular
    vector of F."""
    null_vec = np.linalg.svd(np.asarray(F, np.float64))[2][-1]
    return null_vec / null_vec[2]


def decompose_E(E):
    """E = [t]x R  ->  (R1, R2, t): the two rotation candidates U W V^T / U W^T V^T (flipped to
    det +1) and the unit baseline direction U[:, 2]; as in reference :177-203 no cheirality test
    chooses between the four (R, +-t) combinations."""
    U, _, Vt = np.linalg.svd(E)
    quarter_turn = np.float32([[0, -1, 0], [1, 0, 0], [0, 0, 1]])
    rotations = []
    for W in (quarter_turn, quarter_turn.T):
        R = U.dot(W).dot(Vt)
        rotations.append(-R if np.linalg.det(R) < 0 else R)
    t = U[:, 2] / np.linalg.norm(U[:, 2])
    return rotations[0], rotations[1], t


def recover_pose(E, pts1, pts2, K, distance_threshold, mask):
    """Reference :206-207 raises too."""
    raise NotImplementedError()


def epipolar_line(F_10, x_0):
    """l_1 = F_10 x_0, scaled to a^2 + b^2 = 1 (reference :352-357,
    cv2.computeCorrespondEpilines(x, 1, F)): one GPU pass, bit-identical lines."""
    return ops.epipolar_lines(F_10, x_0)


def unproject_points(pts):
    """Reference :37-39."""
    pts = np.asarray(pts)
    assert pts.ndim == 2 and pts.shape[1] == 2
    return np.hstack([pts, np.ones((len(pts), 1))])


def sampson_error(F, pts1, pts2):
    """Squared Sampson error of reference :52-68 (the same expression, including
    its use of F*x2); only the diagonal of x1^T (F x2) is formed, on the GPU."""
    return ops.sampson_error(F, pts1, pts2)


def filter_sampson_error(cam1, cam2, pts1, pts2, matched_ids, error=4):
    """Reference :71-90: keep the matches whose Sampson error is below `error`."""
    F = cam1.F(cam2)
    err = np.asarray(sampson_error(F, pts2, pts1))
    inliers, = np.where(np.fabs(err) < error)
    pts1, pts2, matched_ids = pts1[inliers], pts2[inliers], matched_ids[inliers]
    assert len(pts1) == len(pts2) == len(matched_ids)
    return pts1, pts2, matched_ids


def get_object_bbox(camera, pts, subsample=10, scale=1.0, min_height=10, min_width=10):
    """Reference :309-349: projected points, their [l, t, r, b] box and median
    depth, or [None]*3.  The projection + bounds compaction runs fused on the
    GPU (get_bounded_projection); the box logic is scalar host work."""
    pts2d, valid = get_bounded_projection(camera, pts, subsample=subsample)
    if not len(pts2d):
        return [None] * 3
    x0, x1 = int(max(0, np.min(pts2d[:, 0]))), int(min(camera.shape[1] - 1, np.max(pts2d[:, 0])))
    y0, y1 = int(max(0, np.min(pts2d[:, 1]))), int(min(camera.shape[0] - 1, np.max(pts2d[:, 1])))
    xmed, ymed = np.median(pts2d[:, 0]), np.median(pts2d[:, 1])
    if (xmed >= 0 and ymed >= 0 and xmed <= camera.shape[1] and ymed < camera.shape[0]) and \
       (y1 - y0) >= min_height and (x1 - x0) >= min_width:
        depth = get_median_depth(camera, pts, subsample=subsample)
        if depth < 0:
            return [None] * 3
        if scale != 1.0:
            w2, h2 = (scale - 1.0) * (x1 - x0) / 2, (scale - 1.0) * (y1 - y0) / 2
            x0, x1 = int(max(0, x0 - w2)), int(min(x1 + w2, camera.shape[1] - 1))
            y0, y1 = int(max(0, y0 - h2)), int(min(y1 + h2, camera.shape[0] - 1))
        return pts2d.astype(np.int32), np.float32([x0, y0, x1, y1]), depth
    return [None] * 3


class CameraIntrinsic(object):
    def __init__(self, K, D=np.zeros(5, dtype=np.float64), shape=None):
        """K: 3x3 calibration; D: distortion; shape: (H, W[, C])."""
        self.K = K
        self.D = D
        self.shape = np.int32(shape) if shape is not None else None

    def __repr__(self):
        return ("CameraIntrinsic: fx {:3.2f} fy {:3.2f} cx {:3.2f} cy {:3.2f} shape {} D {}"
                .format(self.fx, self.fy, self.cx, self.cy, self.shape,
                        np.array_str(np.asarray(self.D), precision=2, suppress_small=True)))

    @classmethod
    def simulate(cls):
        return cls.from_calib_params(500., 500., 320., 240., shape=(480, 640))

    @classmethod
    def from_calib_params(cls, fx, fy, cx, cy, k1=0, k2=0, k3=0, p1=0, p2=0, shape=None):
        return cls(construct_K(fx, fy, cx, cy), D=construct_D(k1, k2, k3, p1, p2), shape=shape)

    @classmethod
    def from_calib_params_fov(cls, fov, cx, cy, D=np.zeros(5, dtype=np.float64), shape=None):
        return cls(construct_K(cx / np.tan(fov), cy / np.tan(fov), cx, cy), D=D, shape=shape)

    fx = property(lambda self: self.K[0, 0])
    fy = property(lambda self: self.K[1, 1])
    cx = property(lambda self: self.K[0, 2])
    cy = property(lambda self: self.K[1, 2])
    skew = property(lambda self: self.K[0, 1])
    k1 = property(lambda self: self.D[0])
    k2 = property(lambda self: self.D[1])
    p1 = property(lambda self: self.D[2])
    p2 = property(lambda self: self.D[3])
    k3 = property(lambda self: self.D[4])

    @property
    def fov(self):
        return np.float32([np.arctan(self.shape[1] * 0.5 / self.fx),
                           np.arctan(self.shape[0] * 0.5 / self.fy)]) * 2.0

    def scaled(self, scale):
        shape = np.int32(self.shape * scale) if self.shape is not None else None
        return CameraIntrinsic.from_calib_params(
            self.fx * scale, self.fy * scale, self.cx * scale, self.cy * scale,
            k1=self.k1, k2=self.k2, k3=self.k3, p1=self.p1, p2=self.p2, shape=shape)

    def in_view(self, x):
        x = np.asarray(x)
        return np.where((x[:, 0] >= 0) & (x[:, 0] < self.shape[1])
                        & (x[:, 1] >= 0) & (x[:, 1] < self.shape[0]))[0]

    # ------------------------------------------- rays (reference :499-538), on the GPU
    def undistort_points(self, pts):
        """cv2.undistortPoints(pts.astype(float32), K, D, P=K) -> (N,2) float32 (:529-538)."""
        return ops.undistort_points(pts, self.K, self.D, P=self.K)

    def _rotate_rows(self):
        """Coefficient rows of Quaternion.rotate (quaternion.py:69-85) for the
        extrinsics' quaternion, formed in the reference's own order."""
        qx, qy, qz, qw = self.extrinsics.quat.q
        ab, ac, ad = qw * qx, qw * qy, qw * qz
        nbb, bc, bd = -qx * qx, qx * qy, qx * qz
        ncc, cd, ndd = -qy * qy, qy * qz, -qz * qz
        return np.float64([[ncc + ndd, bc - ad, ac + bd],
                           [ad + bc, nbb + ndd, cd - ab],
                           [bd - ac, ab + cd, nbb + ncc]])

    def save(self, filename):
        """YAML with the reference's keys (:547-557): fx fy cx cy k1 k2 k3 p1 p2 width height."""
        try:
            height, width = self.shape[:2]
        except Exception:
            height, width = "", ""
        _save_yaml(filename, dict(
            fx=float(self.fx), fy=float(self.fy), cx=float(self.cx), cy=float(self.cy),
            k1=float(self.D[0]), k2=float(self.D[1]), k3=float(self.D[4]), p1=float(self.D[2]), p2=float(self.D[3]),
            width=int(width), height=int(height)))

    @classmethod
    def load(cls, filename):
        """Reads what save() wrote.  Like the reference (:559-565) the shape comes back as
        int32([width, height]) - the reference's order, kept for file and caller compatibility."""
        db = _load_yaml(filename)
        shape = np.int32([db["width"], db["height"]]) if "width" in db and "height" in db else None
        return cls.from_calib_params(db["fx"], db["fy"], db["cx"], db["cy"], k1=db["k1"], k2=db["k2"],
                                     k3=db["k3"], p1=db["p1"], p2=db["p2"], shape=shape)

    def ray(self, pts, undistort=True, rotate=False, normalize=False):
        """Rays [(u-cx)/fx, (v-cy)/fy, 1] through the (optionally undistorted)
        pixels, optionally rotated into the camera's viewpoint and normalised
        (:499-517); one kernel, (N,3) float64."""
        return ops.camera_rays(pts, self.K, self.D, undistort=undistort,
                               rot=self._rotate_rows() if rotate else None, normalize=normalize)

    def reconstruct(self, xyZ, undistort=True):
        """(N,3) [x, y, Z] -> ray(x, y) * Z (:519-524)."""
        return ops.camera_rays(xyZ, self.K, self.D, undistort=undistort, scale_by_z=True)


class CameraExtrinsic(RigidTransform):
    def __init__(self, R=np.eye(3), t=np.zeros(3)):
        """Pose p_cw (world with respect to the camera)."""
        p = RigidTransform.from_Rt(R, t)
        RigidTransform.__init__(self, xyzw=p.quat.to_xyzw(), tvec=p.tvec)
        self.__cached_inverse = None

    def inverse(self):
        if self.__cached_inverse is None:
            self.__cached_inverse = super(CameraExtrinsic, self).inverse()
        return self.__cached_inverse

    def __repr__(self):
        return "CameraExtrinsic: pose = {:}".format(RigidTransform.__repr__(self))

    def c2w(self, X):
        return self * X

    def w2c(self, X):
        return self.inverse() * X

    @classmethod
    def from_rigid_transform(cls, p):
        R, t = p.to_Rt()
        return cls(R, t)

    R = property(lambda self: self.quat.R)
    Rt = property(lambda self: self.matrix[:3])
    t = property(lambda self: self.tvec)
    o2w = property(lambda self: self)
    w2o = property(lambda self: self.inverse())

    @classmethod
    def identity(cls):
        return cls()

    @classmethod
    def simulate(cls):
        return cls.identity()

    def save(self, filename):
        """YAML {R: 3x3 nested list, t: list}: the file reference :636-638 means to write (its
        body reads an undefined name) and :640-643 reads back."""
        R, t = self.to_Rt()
        _save_yaml(filename, dict(R=np.asarray(R, np.float64).tolist(), t=np.asarray(t, np.float64).tolist()))

    @classmethod
    def load(cls, filename):
        db = _load_yaml(filename)
        return cls(R=np.float64(db["R"]), t=np.float64(db["t"]))


class Camera(CameraIntrinsic, CameraExtrinsic):
    def __init__(self, K, R, t, D=np.zeros(5, dtype=np.float64), shape=None):
        CameraIntrinsic.__init__(self, K, D, shape=shape)
        CameraExtrinsic.__init__(self, R, t)

    def __repr__(self):
        return "{:}\n{:}".format(CameraIntrinsic.__repr__(self), CameraExtrinsic.__repr__(self))

    @property
    def P(self):
        return self.K.dot(self.Rt)

    def F(self, other):
        """Fundamental matrix with respect to `other` (reference :777-808, H&Z
        eq. 17.3): F[i, j] = det of the rows of P left after removing row j of
        self.P stacked on those left after removing row i of other.P.  Host-side
        scalar work (nine 4x4 determinants)."""
        rows = ([1, 2], [2, 0], [0, 1])
        P1, P2 = np.asarray(self.P), np.asarray(other.P)
        return np.float64([[np.linalg.det(np.vstack([P1[rows[j], :], P2[rows[i], :]])) for j in range(3)]
                           for i in range(3)])

    @classmethod
    def simulate(cls):
        return cls.from_intrinsics_extrinsics(CameraIntrinsic.simulate(), CameraExtrinsic.simulate())

    @classmethod
    def from_intrinsics_extrinsics(cls, intrinsic, extrinsic):
        return cls(intrinsic.K, extrinsic.R, extrinsic.t, D=intrinsic.D, shape=intrinsic.shape)

    @classmethod
    def from_intrinsics(cls, intrinsic):
        return cls.from_intrinsics_extrinsics(intrinsic, CameraExtrinsic.identity())

    def frustum(self, zmin=0.01, zmax=0.1, pts=None):
        """Viewing frustum between zmin and zmax (reference :820-822)."""
        assert self.shape is not None
        return Frustum.from_camera(self, zmin=zmin, zmax=zmax, pts=pts)

    def save(self, filename):
        raise NotImplementedError()                       # reference :832-833

    @classmethod
    def load(cls, filename):
        raise NotImplementedError()                       # reference :835-837

    @classmethod
    def from_KD(cls, K, D, shape=None):
        return cls.from_intrinsics(CameraIntrinsic(K, D=D, shape=shape))

    @property
    def intrinsics(self):
        return CameraIntrinsic(self.K, self.D, shape=self.shape)

    @property
    def extrinsics(self):
        return CameraExtrinsic(self.R, self.t)

    # ---------------------------------------------------------------- hot path
    def _projection_params(self):
        """Rotation for the pixels (Rodrigues round trip, reference :693-697)
        and depth row (extrinsics round trip, :683-685, :743)."""
        R, t = self.to_Rt()
        Rp = rodrigues_round_trip(R)
        Tz = self.extrinsics.to_matrix()
        return Rp, t, Tz[2, :3].copy(), float(Tz[2, 3])

    def project(self, X, check_bounds=False, check_depth=False, return_depth=False,
                return_valid=False, min_depth=0.1):
        """Project (N,3) points to (M,2) pixels; M <= N when a check is on.
        Output list/array structure and dtypes follow the reference (:727-734):
        x has X's float dtype, depths float64, valid bool."""
        if check_bounds and self.shape is None:
            raise ValueError("check_bounds cannot proceed. Camera.shape is not set")
        Rp, t, Rz, tz = self._projection_params()
        x, depths, valid, _ = ops.project_points(
            X, Rp, t, self.K, self.D, Rz=Rz, tz=tz,
            shape=None if self.shape is None else (int(self.shape[0]), int(self.shape[1])),
            check_bounds=check_bounds, check_depth=check_depth, return_depth=return_depth,
            return_valid=return_valid, min_depth=min_depth, compact=True)
        output = [x]
        if return_depth:
            output.append(depths)
        if return_valid:
            output.append(valid)
        return output if len(output) > 1 else output[0]

    def depth_from_projection(self, X):
        """z of p_cw * X (reference :736-743)."""
        return (self.extrinsics * X)[:, 2]

    def E(self, other):
        """[t]_x R of this camera (reference :809-815; `other` is unused there too)."""
        return skew(self.t).dot(self.R)

    def triangulate(self, pts, other_cam, other_pts):
        """Reference :817-818."""
        return triangulate_points(self, pts, other_cam, other_pts)

    def center(self):
        if self.cx is None:
            raise AssertionError("cx, cy is not set")
        return np.matrix([self.cx, self.cy])

    def set_pose(self, pose):
        self.quat = pose.quat
        self.tvec = pose.tvec

    def scaled(self, scale):
        return Camera.from_intrinsics_extrinsics(self.intrinsics.scaled(scale), self.extrinsics)

    def check_visibility(self, pts, zmin=0, zmax=100):
        return check_visibility(self, pts, zmin=zmin, zmax=zmax)


class StereoCamera(Camera):
    def __init__(self, lcamera, rcamera, baseline):
        if not isinstance(lcamera, Camera) or not isinstance(rcamera, Camera):
            raise TypeError("lcamera, rcamera are not Camera type")
        Camera.__init__(self, lcamera.K, lcamera.R, lcamera.t, D=lcamera.D, shape=lcamera.shape)
        if np.linalg.norm(lcamera.t - rcamera.t) < 1e-2:
            raise ValueError("Right camera does not have an offset from the left camera")
        self.right = rcamera
        self.baseline = baseline

    @classmethod
    def simulate(cls):
        return cls.from_left_with_baseline(Camera.simulate(), baseline=0.1)

    @property
    def left(self):
        return Camera.from_intrinsics_extrinsics(self.intrinsics, self.extrinsics)

    def __repr__(self):
        return "StereoCamera: baseline {:} m\nLEFT {:}\nRIGHT {:}".format(
            self.baseline, Camera.__repr__(self), self.right)

    @property
    def baseline_px(self):
        return self.baseline * self.left.fx

    @property
    def Q(self):
        """float32 reprojection matrix of the reference (:884-896):
        X = (x-cx)*b/d, Y = (y-cy)*b/d, Z = fx*b/d."""
        q43 = -1.0 / self.baseline
        return np.float32([[-1, 0, 0, self.left.cx], [0, -1, 0, self.left.cy],
                           [0, 0, 0, -self.left.fx], [0, 0, q43, 0]])

    @classmethod
    def from_calib_params(cls, fx, fy, cx, cy, k1=0, k2=0, k3=0, p1=0, p2=0, baseline=None,
                          baseline_px=None, shape=None):
        baseline, baseline_px = get_baseline(fx, baseline=baseline, baseline_px=baseline_px)
        lcamera = Camera.from_intrinsics(CameraIntrinsic.from_calib_params(
            fx, fy, cx, cy, k1=k1, k2=k2, k3=k3, p1=p1, p2=p2, shape=shape))
        return cls.from_left_with_baseline(lcamera, baseline=baseline)

    @classmethod
    def from_intrinsics_extrinsics(cls, lintrinsics, rintrinsics, rextrinsics):
        """Left extrinsics assumed identity; not implemented in the reference either (:925-928)."""
        raise NotImplementedError()

    @classmethod
    def from_left_with_baseline(cls, lcamera, baseline):
        """Left extrinsics is assumed to be identity."""
        rcamera = Camera.from_intrinsics_extrinsics(
            lcamera.intrinsics, CameraExtrinsic.from_rigid_transform(
                lcamera.oplus(RigidTransform.from_rpyxyz(0, 0, 0, baseline, 0, 0))))
        return cls(lcamera, rcamera, baseline=baseline)

    def scaled(self, scale):
        left = super(StereoCamera, self).scaled(scale)
        return StereoCamera(left, self.right.scaled(scale), baseline=self.baseline)

    def disparity_from_plane(self, rows, height):
        Z = height * self.left.fy / (np.arange(rows) - self.left.cy + 1e-9)
        return self.left.fx * self.baseline / Z

    def depth_from_disparity(self, disp):
        return self.left.fx * self.baseline / disp

    def disparity_from_depth(self, depth):
        return self.left.fx * self.baseline / depth

    # ---------------------------------------------------------------- hot path
    def reconstruct(self, disp):
        """(H,W) disparity -> (H,W,3) float32 XYZ.  Also accepts a (B,H,W)
        batch of frames from the same camera (one launch)."""
        return ops.reproject_disparity(disp, self.Q)

    def reconstruct_sparse(self, xyd):
        """(N,3) [x, y, d] rows -> (N,3) float64 XYZ."""
        return ops.reproject_sparse(xyd, self.Q)

    def reconstruct_with_texture(self, disp, im, sample=1):
        """Returns (im_pub (N,3) uint8, X_pub (N,3) float32), N = ceil(H/s)*ceil(W/s);
        batched inputs give (B,N,3)."""
        if im.ndim != disp.ndim + 1:
            raise AssertionError("image must be (H, W, 3)")
        xyz, col = ops.reproject_disparity(disp, self.Q, sample=sample, bgr=im)
        lead = tuple(xyz.shape[:-3])
        return col.reshape(lead + (-1, 3)), xyz.reshape(lead + (-1, 3))

    def set_pose(self, left_pose):
        super(StereoCamera, self).set_pose(left_pose)
        self.right.set_pose(self.left.oplus(RigidTransform.from_rpyxyz(0, 0, 0, self.baseline, 0, 0)))


class DepthCamera(CameraIntrinsic):
    """Reference :1010-1048.  reconstruct() turns a depth image into an organised
    float64 cloud dstack([xs*d, ys*d, d]) on the GPU (pbk_depth_reconstruct) with
    the same precomputed meshes as the reference's _build_mesh."""

    def __init__(self, K, shape=(480, 640), skip=1, D=np.zeros(5, dtype=np.float64)):
        CameraIntrinsic.__init__(self, K, D)
        self.shape = shape
        self.skip = skip
        self._build_mesh(shape=shape)

    def _build_mesh(self, shape):
        H, W = shape
        xs, ys = np.arange(0, W), np.arange(0, H)
        self.xs1d = ((xs - self.cx) * 1.0 / self.fx)[::self.skip]       # reference :1024-1027
        self.ys1d = ((ys - self.cy) * 1.0 / self.fy)[::self.skip]
        self.xs, self.ys = np.meshgrid(self.xs1d, self.ys1d)
        self._mesh_dev = None

    def reconstruct(self, depth):
        """(H,W) [or (B,H,W)] depth -> (Hs,Ws,3) float64."""
        if getattr(self, "_mesh_dev", None) is None:
            from .. import _ffi
            self._mesh_dev = (_ffi.to_device(np.ascontiguousarray(self.xs1d, np.float64))[0],
                              _ffi.to_device(np.ascontiguousarray(self.ys1d, np.float64))[0])
        return ops.depth_reconstruct(depth, self._mesh_dev[0], self._mesh_dev[1], self.skip)

    def reconstruct_sparse(self, pts, depth):
        return ops.depth_reconstruct_sparse(pts, depth, self.fx, self.cx, self.cy)

    def save(self, filename):
        raise NotImplementedError()

    @classmethod
    def load(cls, filename):
        raise NotImplementedError()


class HalfPlane(object):
    """Plane through p, q, r (counter-clockwise) as [a, b, c, d] with a x + b y + c z + d = 0 and
    the normal (p - r) x (q - r) (reference :1069-1099; its from_points calls a method numpy
    arrays do not have, so the definition in its docstring is what is implemented)."""

    def __init__(self, hp):
        hp = np.asarray(hp)
        assert hp.ndim == 1 and len(hp) == 4
        self.hp_ = hp

    @classmethod
    def from_points(cls, p, q, r):
        p, q, r = (np.asarray(v, np.float64) for v in (p, q, r))
        abc = np.cross(p - r, q - r)
        return cls(np.r_[abc, -abc.dot(r)])

    def signed_distance(self, pts):
        """(N,) a x + b y + c z + d for (N,3) points (positive on the normal's side)."""
        return np.asarray(pts, np.float64).dot(self.hp_[:3]) + self.hp_[3]

    def intersect(self):
        pass                                              # a stub in the reference as well (:1090-1099)


class Frustum(object):
    """Eight corner points of a viewing frustum: four on the near plane, then four on the far
    plane, each quad in the same winding (reference :1106-1214)."""

    def __init__(self, vertices):
        self.vertices_ = vertices

    @property
    def _front_back_vertices(self):
        half = len(self.vertices_) // 2                   # the reference divides with `/`: a float index under Python 3
        return self.vertices_[:half], self.vertices_[half:]

    @classmethod
    def from_pose(cls, pose, zmin=0.01, zmax=0.1, fov=np.deg2rad(60)):
        if fov >= np.pi:
            raise ValueError("Frustum fov cannot be {} radians".format(fov))
        if zmin < 0.01:
            raise ValueError("zmin needs to be finite > 0.01")
        # the reference ignores `fov` and uses the half-extents of a 640x480, f = 500 camera
        rx, ry = 0.638, 0.478
        corners = np.array([[-rx, -ry, 1.0], [-rx, ry, 1.0], [rx, ry, 1.0], [rx, -ry, 1.0]])
        return cls(pose * np.vstack([corners * zmin, corners * zmax]))

    @classmethod
    def from_camera(cls, c, zmin=0.01, zmax=0.1, pts=None):
        if not isinstance(c, Camera):
            raise TypeError("c is not of Camera type, cannot instantiate frustum")
        if zmin < 0.01:
            raise ValueError("zmin needs to be finite > 0.01")
        if pts is None:                                   # the full field of view: the image corners
            H, W = c.shape
            pts = np.vstack([[0, 0], [0, H - 1], [W - 1, H - 1], [W - 1, 0]])
        rays = c.ray(pts, undistort=True, rotate=False)
        return cls(c.w2c(np.vstack([rays * zmin, rays * zmax])))

    @property
    def vertices(self):
        return self.vertices_

    @property
    def points_and_normals(self):
        """One (point, unit normal) per bounding plane: the four side planes, then near and far."""
        near, far = self._front_back_vertices
        near_next, far_next = np.roll(near, -1, axis=0), np.roll(far, -1, axis=0)
        e1 = np.vstack([far - near, near_next[0] - near[0], far[2] - far[1]])
        e2 = np.vstack([far_next - far, near_next[1] - near_next[0], far[1] - far[0]])
        pts = np.vstack([far, near[0], far[1]])
        e1 = e1 / np.linalg.norm(e1, axis=1).reshape(-1, 1)
        e2 = e2 / np.linalg.norm(e2, axis=1).reshape(-1, 1)
        normals = np.cross(e1, e2)
        normals /= np.linalg.norm(normals, axis=1).reshape(-1, 1)
        return pts, normals

    @property
    def planes(self):
        """(6,4) rows [n, -n . p]."""
        pts, normals = self.points_and_normals
        return np.hstack([normals, -np.sum(pts * normals, axis=1).reshape(-1, 1)])
