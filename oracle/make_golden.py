"""Regenerates tests/golden/*.npz by running the REAL reference (and the cv2
calls a pybot user makes for matching / BA) on small seeded inputs.

TEST INFRASTRUCTURE.  Needs /root/reference (authoring container only):

    python -m oracle.make_golden

The fixtures pin (a) the oracle port/model against the reference
(tests/test_oracle_pinning.py, CPU) and (b) the CUDA path against the
reference (tests/test_gpu_*.py) on the GPU box, where /root/reference does
not exist.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")

from oracle import ref_loader  # noqa: E402
from pybot_b200 import synthetic as syn  # noqa: E402


def g1_reconstruct(ns):
    """G1: KITTI-calibrated frame with specials, all accepted dtypes, texture."""
    cam = ns.StereoCamera.from_calib_params(syn.KITTI_FX, syn.KITTI_FX, syn.KITTI_CX, syn.KITTI_CY,
                                            baseline_px=syn.KITTI_BASELINE_PX, shape=(96, 317))
    rng = np.random.default_rng(101)
    disp = syn.disparity_image(rng, (96, 317), syn.KITTI_FX, syn.KITTI_BASELINE)
    im = rng.integers(0, 256, (96, 317, 3), dtype=np.uint8)
    out = dict(Q=cam.Q, baseline=cam.baseline, disp_f32=disp, im=im,
               xyz_f32=cam.reconstruct(disp))
    for name, dt in (("u8", np.uint8), ("s16", np.int16), ("s32", np.int32)):
        d = np.clip(np.nan_to_num(disp, nan=0, posinf=0), -1, 250).astype(dt)
        out["disp_" + name] = d
        out["xyz_" + name] = cam.reconstruct(d)
    for s in (1, 2, 3):
        c, x = cam.reconstruct_with_texture(disp, im, sample=s)
        out["tex%d_im" % s], out["tex%d_xyz" % s] = c, x
    xyd = np.stack([rng.uniform(0, 317, 64), rng.uniform(0, 96, 64), rng.uniform(1, 90, 64)], 1)
    out["sparse_xyd"] = xyd
    out["sparse_xyz"] = cam.reconstruct_sparse(xyd)
    np.savez_compressed(os.path.join(OUT, "g1_reconstruct.npz"), **out)


def g1b_divcases():
    """Pixels where float(n / W) != float(n * (1/W)): cv2 multiplies by the
    reciprocal.  Found by random search; checked here."""
    import cv2
    q43 = np.float32(-1.0 / 0.5371657)
    cases = [(2610, 15.590840339660645), (3549, 224.04904174804688),
             (3470, 172.59080505371094), (3201, 189.9710235595703)]
    Q = np.float32([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, q43, 0]])
    disp = np.ones((2, 4096), np.float32)
    for x, d in cases:
        disp[0, x] = np.float32(d)
    xyz = cv2.reprojectImageTo3D(disp, Q)
    for x, d in cases:
        W = float(q43) * float(np.float32(d))
        assert np.float32(x / W) != np.float32(x * (1.0 / W))
        assert xyz[0, x, 0] == np.float32(x * (1.0 / W))
    np.savez_compressed(os.path.join(OUT, "reproject_divcases.npz"), Q=Q, disp=disp, xyz=xyz)


def g2_transform_project(ns):
    """G2: transform + project with z in {<0, 0, tiny}, border pixels, min_depth edge."""
    rng = np.random.default_rng(102)
    n = 4000
    X = np.stack([rng.uniform(-40, 40, n), rng.uniform(-3, 3, n), rng.uniform(-5, 80, n)], 1)
    X = X.astype(np.float32)
    K = syn.kitti_K()
    out = dict(K=K, X=X)
    poses = {"c2": syn.C2_RPYXYZ, "ident": (0, 0, 0, 0, 0, 0), "big": (2.5, -1.2, 0.7, -3.0, 4.0, 9.0)}
    for name, rp in poses.items():
        p = ns.RigidTransform.from_rpyxyz(*rp)
        out["pose_%s_xyzw" % name] = p.quat.q.copy()
        out["pose_%s_tvec" % name] = p.tvec.copy()
        out["mul_%s" % name] = p * X
        out["mul64_%s" % name] = p * X.astype(np.float64)
        inv = p.inverse()
        out["inv_%s_xyzw" % name], out["inv_%s_tvec" % name] = inv.quat.q.copy(), inv.tvec.copy()
        cam = ns.Camera.from_intrinsics_extrinsics(
            ns.CameraIntrinsic(K, shape=syn.KITTI_SHAPE), ns.CameraExtrinsic.from_rigid_transform(p))
        out["cam_%s_xyzw" % name], out["cam_%s_tvec" % name] = cam.quat.q.copy(), cam.tvec.copy()
        out["proj_%s" % name] = cam.project(X)
        x, d, v = cam.project(X, check_bounds=True, check_depth=True, return_depth=True,
                              return_valid=True)
        out["projv_%s_x" % name], out["projv_%s_d" % name], out["projv_%s_v" % name] = x, d, v
        x, v = cam.project(X, check_depth=True, return_valid=True, min_depth=2.0)
        out["projd_%s_x" % name], out["projd_%s_v" % name] = x, v
        out["proj64_%s" % name] = cam.project(X.astype(np.float64))
    # identity camera, planted edge cases
    cam = ns.Camera.from_intrinsics(ns.CameraIntrinsic(K, shape=syn.KITTI_SHAPE))
    W, H = 1241.0, 376.0
    fx, cx, cy = syn.KITTI_FX, syn.KITTI_CX, syn.KITTI_CY
    edge = np.float32([
        [1, 1, 0], [0, 0, 0], [1, 2, 1e-12], [1, 1, -4], [0, 0, 0.1], [0, 0, 0.099999994],
        [(W - cx) / fx * 10, 0, 10], [(W - 1e-3 - cx) / fx * 10, 0, 10], [(0 - cx) / fx * 10, 0, 10],
        [0, (H - cy) / fx * 10, 10], [0, (0 - cy) / fx * 10, 10], [np.nan, 0, 5], [np.inf, 1, 5],
        [3e38, 3e38, 1e-30]])
    out["edge_X"] = edge
    with np.errstate(all="ignore"):
        out["edge_proj"] = cam.project(edge)
        x, d, v = cam.project(edge, check_bounds=True, check_depth=True, return_depth=True,
                              return_valid=True)
    out["edge_x"], out["edge_d"], out["edge_v"] = x, d, v
    # distortion
    Dcam = ns.Camera.from_intrinsics_extrinsics(
        ns.CameraIntrinsic.from_calib_params(fx, fx, cx, cy, k1=-0.3, k2=0.1, k3=0.05, p1=0.001,
                                             p2=-0.002, shape=syn.KITTI_SHAPE),
        ns.CameraExtrinsic.from_rigid_transform(ns.RigidTransform.from_rpyxyz(*syn.C2_RPYXYZ)))
    out["dist_D"] = Dcam.D
    out["dist_proj"] = Dcam.project(X)
    np.savez_compressed(os.path.join(OUT, "g2_transform_project.npz"), **out)


def g3_matching():
    """G3/G4: tie-heavy descriptors, duplicates, Nt<k, multi-keyframe collection."""
    import cv2
    out = {}
    q, t = syn.low_entropy_pair(seed=7, nq=300, nt=450)
    t[17] = t[3]
    t[200] = q[5]
    t[9] = q[5]
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    res = bf.knnMatch(q, t, k=2)
    out["le_q"], out["le_t"] = q, t
    out["le_idx"] = np.array([[m.trainIdx for m in r] for r in res], np.int32)
    out["le_dist"] = np.array([[int(m.distance) for m in r] for r in res], np.int32)
    cc = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(q, t)
    out["le_cross"] = np.array([[m.queryIdx, m.trainIdx] for m in cc], np.int32)
    # Nt = 1 (< k)
    r1 = bf.knnMatch(q[:4], t[:1], k=2)
    out["nt1_len"] = np.array([len(r) for r in r1], np.int32)
    # collection with a duplicate across keyframes
    q2, t2 = syn.orb_pair(seed=33, nq=200, nt=600)
    kfs = [t2[:150].copy(), t2[150:400].copy(), t2[400:].copy()]
    kfs[1][0] = kfs[0][5]
    kfs[0][5] = q2[10]
    kfs[1][0] = q2[10]
    bf2 = cv2.BFMatcher(cv2.NORM_HAMMING)
    bf2.add(kfs)
    res = bf2.knnMatch(q2, k=2)
    out["kf_q"] = q2
    out["kf_0"], out["kf_1"], out["kf_2"] = kfs
    out["kf_img"] = np.array([[m.imgIdx for m in r] for r in res], np.int32)
    out["kf_idx"] = np.array([[m.trainIdx for m in r] for r in res], np.int32)
    out["kf_dist"] = np.array([[int(m.distance) for m in r] for r in res], np.int32)
    np.savez_compressed(os.path.join(OUT, "g3_matching.npz"), **out)


def g5_ba():
    """G5: cv2.projectPoints Jacobians at |rvec| in {~1, 1e-9, 0, pi-1e-3}."""
    import cv2
    rng = np.random.default_rng(105)
    K = syn.kitti_K()
    n = 300
    X = np.stack([rng.uniform(-10, 10, n), rng.uniform(-3, 3, n), rng.uniform(4, 60, n)], 1)
    X = X.astype(np.float32)
    ax = np.array([0.3, -0.5, 0.8])
    ax /= np.linalg.norm(ax)
    rvecs = np.float32([[0.3, -0.2, 0.5], [1e-9, 2e-9, -1e-9], [0, 0, 0], ax * (np.pi - 1e-3),
                        [0.01, -1.2, 0.02]])
    tvecs = np.float32([[0.5, -0.1, 1.2], [0, 0, 0], [1, 2, 3], [0.2, 0.1, 70.0], [-3, 0.2, 5]])
    xs, Js = [], []
    for rv, tv in zip(rvecs, tvecs):
        x, J = cv2.projectPoints(X.astype(np.float64), rv.astype(np.float64), tv.astype(np.float64),
                                 K, np.zeros(5))
        xs.append(x.reshape(-1, 2))
        Js.append(J.reshape(n, 2, 15)[:, :, :6])
    np.savez_compressed(os.path.join(OUT, "g5_ba.npz"), K=K, X=X, rvecs=rvecs, tvecs=tvecs,
                        x=np.stack(xs), Jcam=np.stack(Js))


def g0_host_algebra(ns):
    """Pose algebra known answers from the reference itself (host-side glue)."""
    rng = np.random.default_rng(100)
    rp = rng.uniform(-3, 3, (12, 6))
    mats, invs, comps, vecs, rts = [], [], [], [], []
    prev = ns.RigidTransform.identity()
    for a in rp:
        p = ns.RigidTransform.from_rpyxyz(*a)
        mats.append(p.matrix)
        invs.append(p.inverse().matrix)
        comps.append((prev * p).matrix)
        vecs.append(p.to_vector())
        rts.append(ns.RigidTransform.from_Rt(*p.to_Rt()).matrix)
        prev = p
    s = ns.Sim3(xyzw=ns.RigidTransform.from_rpyxyz(*rp[0]).xyzw, tvec=rp[0, 3:], scale=2.5)
    np.savez_compressed(
        os.path.join(OUT, "g0_host_algebra.npz"), rpyxyz=rp, matrix=np.stack(mats),
        inverse=np.stack(invs), compose=np.stack(comps), vector=np.stack(vecs),
        from_Rt=np.stack(rts), sim3_matrix=s.to_matrix(),
        sim3_oplus=s.oplus(ns.RigidTransform.from_rpyxyz(*rp[1])).matrix,
        axes_q=np.stack([ns.tf.quaternion_from_euler(0.3, -0.7, 1.1, axes=a)
                         for a in sorted(ns.tf._AXES2TUPLE)]),
        axes_names=np.array(sorted(ns.tf._AXES2TUPLE)))


def g6_next_rows(ns):
    """G6: DepthCamera.reconstruct / reconstruct_sparse, SceneFlow.process and
    get_object_bbox of the real reference."""
    import importlib
    of = importlib.import_module("pybot.vision.optflow_utils")
    cu = importlib.import_module("pybot.vision.camera_utils")
    rng = np.random.default_rng(606)
    K = syn.kitti_K()
    H, W = 47, 155
    K = K.copy()
    K[:2] *= 0.5                                    # half-resolution KITTI camera
    out = {"K": K, "shape": np.int32([H, W])}
    d32 = rng.uniform(0.4, 30.0, (H, W)).astype(np.float32)
    d32[3, 5] = 0.0
    d32[4, 6] = np.nan
    d16 = (rng.uniform(0.4, 30.0, (H, W)) * 1000).astype(np.uint16)
    out["d32"], out["d16"] = d32, d16
    for skip in (1, 2, 3):
        dc = ns.DepthCamera(K, shape=(H, W), skip=skip)
        out["depth_f32_s%d" % skip] = dc.reconstruct(d32)
    out["depth_u16_s1"] = ns.DepthCamera(K, shape=(H, W)).reconstruct(d16)
    dc = ns.DepthCamera(K, shape=(H, W))
    pts = rng.uniform(0, 150, (257, 2))
    dep = rng.uniform(0.5, 20, 257)
    out["sparse_pts"], out["sparse_depth"] = pts, dep
    out["sparse_out"] = dc.reconstruct_sparse(pts, dep)
    # scene flow: a camera translating forward/sideways over a random depth field,
    # with collisions (several points landing on one pixel), off-image points and specials
    cam = ns.CameraIntrinsic(K, shape=(H, W))
    sf = of.SceneFlow(cam)
    Z = rng.uniform(4, 60, (H, W))
    xs, ys = np.meshgrid(np.arange(W), np.arange(H))
    for name, dt in (("f32", np.float32), ("f64", np.float64)):
        X = np.dstack([(xs - K[0, 2]) / K[0, 0] * Z + 0.35, (ys - K[1, 2]) / K[1, 1] * Z - 0.05, Z - 0.8]).astype(dt)
        X[0, 0] = [np.nan, 1, 1]
        X[0, 1] = [1e30, 1, 1e-3]
        X[0, 2] = [0, 0, 0]
        X[0, 3] = [1, 1, -2]
        X[5, :40] = X[6, :40]                       # forced collisions: later source wins
        out["flow_X_" + name] = X
        out["flow_" + name] = sf.process(X)
    camd = ns.CameraIntrinsic(K, D=np.float64([-0.3, 0.1, 0.001, -0.002, 0.01]), shape=(H, W))
    out["flow_D"] = camd.D
    out["flow_dist"] = of.SceneFlow(camd).process(out["flow_X_f32"])
    # get_object_bbox on a cluster of points in front of a posed camera
    pose = ns.RigidTransform.from_rpyxyz(0.02, -0.05, 0.01, 0.3, -0.2, 0.5)
    c = ns.Camera.from_intrinsics_extrinsics(ns.CameraIntrinsic(syn.kitti_K(), shape=(376, 1241)),
                                             ns.CameraExtrinsic.from_rigid_transform(pose))
    obj = rng.normal(0, 1.0, (1500, 3)) * [1.5, 0.6, 1.0] + [2.0, 0.0, 14.0]
    p2, bbox, depth = cu.get_object_bbox(c, obj, subsample=3, scale=1.2)
    out.update(bbox_pts=obj, bbox_pose=pose.matrix, bbox_p2=p2, bbox_box=bbox, bbox_depth=np.float64(depth))
    np.savez_compressed(os.path.join(OUT, "g6_next_rows.npz"), **out)


def _reference_functions(relpath, names, extra_globals):
    """Executes selected top-level function definitions of a reference module
    whose import fails here (pybot.externals needs protobuf 3.4 generated code):
    the function bodies that run are the reference's own source."""
    import ast
    src = open(os.path.join(ref_loader.REFERENCE_ROOT, relpath)).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    mod = ast.Module(body=keep, type_ignores=[])
    g = dict(extra_globals)
    exec(compile(mod, relpath, "exec"), g)
    return g


def g7_epipolar_wire(ns):
    """G7: sampson_error / filter_sampson_error / Camera.F and the viewer colour
    wire format (get_color_arr) of the real reference."""
    import importlib
    cu = importlib.import_module("pybot.vision.camera_utils")
    rng = np.random.default_rng(707)
    K = syn.kitti_K()
    mk = lambda *rp: ns.Camera.from_intrinsics_extrinsics(          # noqa: E731
        ns.CameraIntrinsic(K, shape=(376, 1241)),
        ns.CameraExtrinsic.from_rigid_transform(ns.RigidTransform.from_rpyxyz(*rp)))
    rp1, rp2 = (0.01, -0.02, 0.03, 0.1, 0.0, 0.2), (-0.02, 0.05, 0.0, -0.4, 0.05, 0.9)
    c1, c2 = mk(*rp1), mk(*rp2)
    F = c1.F(c2)
    X = rng.uniform([-10, -2, 5], [10, 2, 40], (300, 3))
    p1 = c1.project(X)
    p2 = c2.project(X) + rng.normal(0, 1.5, (300, 2))
    err = cu.sampson_error(F, p2, p1)
    f1, f2, ids = cu.filter_sampson_error(c1, c2, p1, p2, np.arange(300), error=4)
    g = _reference_functions("pybot/externals/draw_helpers.py", ("reshape_arr", "get_color_arr"), {"np": np})
    im = rng.integers(0, 256, (7, 9, 3), dtype=np.uint8)
    np.savez_compressed(
        os.path.join(OUT, "g7_epipolar_wire.npz"), rp1=np.float64(rp1), rp2=np.float64(rp2), F=F, p1=p1, p2=p2,
        err=err, inlier_ids=ids, wire_im=im, wire_rgb=g["get_color_arr"](im.copy(), im.size // 3, flip_rb=False),
        wire_flip=g["get_color_arr"](im.copy(), im.size // 3, flip_rb=True))


def g8_rays_triangulate(ns):
    """G8: CameraIntrinsic.undistort_points / ray / reconstruct and
    triangulate_points of the real reference."""
    import importlib
    cu = importlib.import_module("pybot.vision.camera_utils")
    rng = np.random.default_rng(808)
    K = syn.kitti_K()
    out = {"K": K}
    pts = rng.uniform([-60, -40], [1300, 420], (600, 2))
    pts[:4] = [[K[0, 2], K[1, 2]], [0, 0], [1240.5, 375.5], [-1e4, 3e4]]
    out["pts"] = pts
    rp = (0.3, -0.2, 0.7, 0.5, -0.1, 1.2)
    pose = ns.RigidTransform.from_rpyxyz(*rp)
    out["pose_rpyxyz"] = np.float64(rp)
    dists = {"d0": np.zeros(5), "d5": np.float64([-0.3, 0.1, 0.001, -0.002, 0.01]),
             "d4": np.float64([-0.37, 0.2, 0.0005, 0.0003]),
             "d8": np.float64([0.1, -0.05, 0.001, 0.002, 0.003, 0.01, 0.02, -0.003]),
             "dneg": np.float64([-0.9, 0.1, 0, 0, 0])}            # icdist < 0 branch far from the centre
    for name, D in dists.items():
        out["D_" + name] = D
        ci = ns.CameraIntrinsic(K, D=D, shape=(376, 1241))
        out["und_" + name] = ci.undistort_points(pts)
        out["ray_" + name] = ci.ray(pts)
    ci = ns.CameraIntrinsic(K, D=dists["d5"], shape=(376, 1241))
    cam = ns.Camera.from_intrinsics_extrinsics(ci, ns.CameraExtrinsic.from_rigid_transform(pose))
    out["ray_raw_f64"] = ci.ray(pts, undistort=False)
    out["ray_raw_f32"] = ci.ray(pts.astype(np.float32), undistort=False)
    out["ray_norm"] = ci.ray(pts, normalize=True)
    out["ray_rot"] = cam.ray(pts, rotate=True)
    out["ray_rot_norm"] = cam.ray(pts, undistort=False, rotate=True, normalize=True)
    xyZ = np.hstack([pts, rng.uniform(0.5, 40, (len(pts), 1))])
    out["xyZ"] = xyZ
    out["rec"] = ci.reconstruct(xyZ)
    out["rec_raw"] = ci.reconstruct(xyZ, undistort=False)
    # triangulation: two posed KITTI cameras, exact and noisy correspondences, f64 and f32
    mk = lambda *rp: ns.Camera.from_intrinsics_extrinsics(          # noqa: E731
        ns.CameraIntrinsic(K, shape=(376, 1241)),
        ns.CameraExtrinsic.from_rigid_transform(ns.RigidTransform.from_rpyxyz(*rp)))
    rp1, rp2 = (0.01, -0.02, 0.03, 0.1, 0.0, 0.2), (-0.02, 0.05, 0.0, -0.4, 0.05, 0.9)
    c1, c2 = mk(*rp1), mk(*rp2)
    X = rng.uniform([-10, -2, 5], [10, 2, 40], (500, 3))
    p1, p2 = c1.project(X), c2.project(X)
    p2n = p2 + rng.normal(0, 1.0, p2.shape)
    out.update(tri_rp1=np.float64(rp1), tri_rp2=np.float64(rp2), tri_P1=c1.P, tri_P2=c2.P, tri_X=X,
               tri_p1=p1, tri_p2=p2, tri_p2n=p2n,
               tri_exact=cu.triangulate_points(c1, p1, c2, p2),
               tri_noisy=cu.triangulate_points(c1, p1, c2, p2n),
               tri_f32=cu.triangulate_points(c1, p1.astype(np.float32), c2, p2n.astype(np.float32)))
    np.savez_compressed(os.path.join(OUT, "g8_rays_triangulate.npz"), **out)


def g9_epipolar_helpers(ns):
    """G9: the small callers either side of matching in the real reference - epipolar_line
    (cv2.computeCorrespondEpilines), Camera.E, compute_essential / compute_epipole /
    decompose_E, project_points / unproject / skew, and the frame helpers of
    rigid_transform.py (tf_construct, tf_construct_3pt, tf_compose, normalize_vec, skew).
    camera_utils.skew (:47-49) and with it Camera.E (:809-815) raise NameError in the reference
    (`array` is never imported) and so does compute_epipole (:234-242, `linalg`): nothing to
    pin for those three."""
    import importlib
    cu = importlib.import_module("pybot.vision.camera_utils")
    rt = importlib.import_module("pybot.geometry.rigid_transform")
    rng = np.random.default_rng(909)
    K = syn.kitti_K()
    mk = lambda *rp: ns.Camera.from_intrinsics_extrinsics(          # noqa: E731
        ns.CameraIntrinsic(K, shape=(376, 1241)),
        ns.CameraExtrinsic.from_rigid_transform(ns.RigidTransform.from_rpyxyz(*rp)))
    rp1, rp2 = (0.01, -0.02, 0.03, 0.1, 0.0, 0.2), (-0.02, 0.05, 0.0, -0.4, 0.05, 0.9)
    c1, c2 = mk(*rp1), mk(*rp2)
    F = c1.F(c2)
    x64 = rng.uniform([0, 0], [1241, 376], (500, 2))
    x64[:3] = [[0, 0], [1240.5, 375.5], [607.1928, 185.2157]]
    x32 = x64.astype(np.float32)
    Fz = np.zeros((3, 3))
    Fz[2] = [1.0, 2.0, 3.0]                                   # a = b = 0: lines left unscaled
    E = cu.compute_essential(F, K)
    R1, R2, t = cu.decompose_E(E)
    X = rng.uniform([-10, -2, 5], [10, 2, 40], (50, 3))
    v1, v2 = rt.normalize_vec(rng.normal(size=3)), rt.normalize_vec(rng.normal(size=3))
    p3 = rng.normal(size=(3, 3))
    sk, dV = rt.skew(v1, return_dv=True)
    T3 = rt.tf_construct_3pt(p3[0], p3[1], p3[2])
    np.savez_compressed(
        os.path.join(OUT, "g9_epipolar_helpers.npz"), rp1=np.float64(rp1), rp2=np.float64(rp2), F=F, x64=x64,
        lines64=cu.epipolar_line(F, x64), lines32=cu.epipolar_line(F, x32), Fz=Fz,
        lines_z=cu.epipolar_line(Fz, x64[:8]), E=E, R1=R1, R2=R2,
        t=t, X=X, proj=cu.project_points(X), unproj=cu.unproject(X[0]), deproj=cu.project(X[0]),
        v1=v1, v2=v2, tfc=rt.tf_construct(v1, v2), skew_rt=sk, skew_dv=dV, p3=p3,
        T3=T3.to_matrix(), tcomp=rt.tf_compose(R1, t))


def g10_visibility(ns):
    """G10: check_visibility / get_median_depth / get_bounded_projection of the real reference
    (camera_utils.py:245-288) around a posed KITTI camera: points in
    front, behind, beyond zmax, exactly on zmin / zmax, far outside the field of view."""
    import importlib
    cu = importlib.import_module("pybot.vision.camera_utils")
    rng = np.random.default_rng(1010)
    pose = ns.RigidTransform.from_rpyxyz(0.02, -0.05, 0.01, 0.3, -0.2, 0.5)
    c = ns.Camera.from_intrinsics_extrinsics(ns.CameraIntrinsic(syn.kitti_K(), shape=(376, 1241)),
                                             ns.CameraExtrinsic.from_rigid_transform(pose))
    n = 20000
    pts = rng.normal(0, 1.0, (n, 3)) * [30.0, 8.0, 40.0] + [0.0, 0.0, 20.0]
    pts[:50] = rng.normal(0, 1.0, (50, 3)) * [3, 3, 3] + [0, 0, 150.0]        # beyond zmax = 100
    cw = pose.matrix                                                         # c2w(X) = pose * X
    inv = np.linalg.inv(cw)
    edge = np.array([[0.0, 0.0, 100.0], [0.0, 0.0, 0.0], [1.0, 0.5, 5.0], [0.0, 0.0, 99.999999]])
    pts[50:54] = (edge - cw[:3, 3]) @ inv[:3, :3].T                          # ~ on the z limits after c2w
    out = {"vis_pose": pose.matrix, "vis_pose_xyzw": pose.quat.q, "vis_pose_tvec": pose.tvec,
           "vis_cam_xyzw": c.quat.q, "vis_cam_tvec": c.tvec, "vis_pts": pts}
    out["vis_mask"] = cu.check_visibility(c, pts)
    out["vis_mask_z"] = cu.check_visibility(c, pts, zmin=5.0, zmax=30.0)
    out["vis_mask_f32"] = cu.check_visibility(c, pts.astype(np.float32))
    out["vis_median"] = np.float64(cu.get_median_depth(c, pts, subsample=7))
    p2, valid = cu.get_bounded_projection(c, pts, subsample=3)
    out["vis_bp_x"], out["vis_bp_valid"] = p2, valid
    out["vis_fov"] = c.fov
    np.savez_compressed(os.path.join(OUT, "g10_visibility.npz"), **out)


def g11_facade(ns):
    """G11: the parts of the host facade the reference can actually execute - Frustum.from_pose /
    Camera.frustum vertices, the YAML CameraIntrinsic.save writes (camera_utils.py:547-557,
    :820-822, :1106-1176) - and, for DualQuaternion (rigid_transform.py:371-480, whose
    constructor raises), the RigidTransform matrices the same poses give."""
    import contextlib
    import importlib
    import io
    import tempfile
    rt = importlib.import_module("pybot.geometry.rigid_transform")
    cu = importlib.import_module("pybot.vision.camera_utils")
    pa = ns.RigidTransform.from_rpyxyz(0.3, -0.2, 1.0, 1.0, 2.0, 3.0)
    pb = ns.RigidTransform.from_rpyxyz(-0.5, 0.1, 0.4, -2.0, 0.5, 1.0)
    try:                                             # the reference class cannot be instantiated:
        rt.DualQuaternion(pa.quat.q, pa.tvec)        # its __init__ calls Quaternion(xyzw=...) (:380),
        dq_error = ""                                # a keyword Quaternion.__init__ does not have
    except TypeError as e:
        dq_error = str(e)
    out = dict(dq_a_xyzw=pa.quat.q, dq_a_tvec=pa.tvec, dq_b_xyzw=pb.quat.q, dq_b_tvec=pb.tvec,
               dq_a_matrix=pa.matrix, dq_ab_matrix=(pa * pb).matrix, dq_a_inv_matrix=pa.inverse().matrix,
               dq_reference_error=np.array(dq_error))
    out["frustum_pose"] = cu.Frustum.from_pose(pa, zmin=0.5, zmax=4.0).vertices
    c = ns.Camera.from_intrinsics_extrinsics(
        ns.CameraIntrinsic(syn.kitti_K(), D=np.float64([-0.3, 0.1, 0.001, -0.002, 0.01]), shape=(376, 1241)),
        ns.CameraExtrinsic.from_rigid_transform(pa))
    out["frustum_cam"] = c.frustum(zmin=0.5, zmax=20.0).vertices
    out["frustum_cam_pts"] = c.frustum(zmin=1.0, zmax=2.0, pts=np.float64([[100, 50], [100, 300], [900, 300], [900, 50]])).vertices
    out["frustum_D"] = c.D
    with tempfile.TemporaryDirectory() as d:
        fn = os.path.join(d, "intrinsic.yaml")
        with contextlib.redirect_stdout(io.StringIO()):          # AttrDict.save_yaml prints the dict
            ns.CameraIntrinsic(syn.kitti_K(), D=c.D, shape=(376, 1241)).save(fn)
        out["intrinsic_yaml"] = np.array(open(fn).read())
    np.savez_compressed(os.path.join(OUT, "g11_facade.npz"), **out)


def g12_knnk():
    """cv2.BFMatcher.knnMatch beyond the hot path's k = 2: k in {1, 3, 7, 40}, a sparse mask (with a
    fully masked query), compactResult, and the collection overload with per-image masks."""
    import cv2
    out = {}
    rng = np.random.default_rng(12)
    q, t = syn.low_entropy_pair(seed=12, nq=120, nt=700)
    t[40] = t[2]
    t[333] = q[9]
    t[11] = q[9]
    out["q"], out["t"] = q, t
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)

    def rows(res, k, field):
        a = np.full((len(res), k), -1, np.int32)
        for i, r in enumerate(res):
            a[i, :len(r)] = [int(getattr(m, field)) for m in r]
        return a

    for k in (1, 3, 7, 40):
        res = bf.knnMatch(q, t, k=k)
        out["k%d_idx" % k], out["k%d_dist" % k] = rows(res, k, "trainIdx"), rows(res, k, "distance")
    mask = (rng.random((len(q), len(t))) < 0.01).astype(np.uint8)
    mask[3] = 0                                   # a query with no allowed row
    mask[5, :] = 0
    mask[5, 600] = 1                              # exactly one allowed row
    out["mask"] = mask
    for k in (2, 5):
        res = bf.knnMatch(q, t, k=k, mask=mask)
        out["m%d_idx" % k], out["m%d_dist" % k] = rows(res, k, "trainIdx"), rows(res, k, "distance")
        out["m%d_len" % k] = np.array([len(r) for r in res], np.int32)
    res = bf.knnMatch(q, t, k=5, mask=mask, compactResult=True)
    out["m5_compact_query"] = np.array([r[0].queryIdx for r in res], np.int32)
    cuts = [0, 250, 400, 700]      # (a 1-row image next to masks makes cv2 4.13 return uninitialised imgIdx/trainIdx)
    kfs = [t[a:b].copy() for a, b in zip(cuts[:-1], cuts[1:])]
    bf2 = cv2.BFMatcher(cv2.NORM_HAMMING)
    bf2.add(kfs)
    res = bf2.knnMatch(q, k=4, masks=[mask[:, a:b].copy() for a, b in zip(cuts[:-1], cuts[1:])])
    out["cuts"] = np.array(cuts, np.int64)
    out["c4_img"], out["c4_idx"], out["c4_dist"] = rows(res, 4, "imgIdx"), rows(res, 4, "trainIdx"), rows(res, 4, "distance")
    res = bf2.knnMatch(q, k=6)
    out["c6_img"], out["c6_idx"], out["c6_dist"] = rows(res, 6, "imgIdx"), rows(res, 6, "trainIdx"), rows(res, 6, "distance")
    np.savez_compressed(os.path.join(OUT, "g12_knnk.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = ref_loader.load()
    if len(sys.argv) > 1 and sys.argv[1] in ("g10", "g11", "g12"):
        {"g10": g10_visibility, "g11": g11_facade, "g12": lambda ns: g12_knnk()}[sys.argv[1]](ns)
        return
    g0_host_algebra(ns)
    g1_reconstruct(ns)
    g1b_divcases()
    g2_transform_project(ns)
    g3_matching()
    g5_ba()
    g6_next_rows(ns)
    g7_epipolar_wire(ns)
    g8_rays_triangulate(ns)
    g9_epipolar_helpers(ns)
    g10_visibility(ns)
    g11_facade(ns)
    g12_knnk()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
