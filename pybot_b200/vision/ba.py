"""Bundle-adjustment reprojection residual and Jacobians on the GPU.

The reference has no BA: the nearest arithmetic is the
(2N x 15) Jacobian ``cv2.projectPoints`` returns and ``Camera.project``
discards (/root/reference/pybot/vision/camera_utils.py:697).  ``project_points``
keeps that OpenCV call shape for the distortion-free pinhole; the array API
evaluates a whole BA problem (cameras x points x observations) in one launch.
"""
import numpy as np

from .. import _ffi, ops

BAProblemDevice = ops.BAProblemDevice


def ba_residual_jacobian(rvecs, tvecs, K, points, obs_cam, obs_pt, obs_uv):
    """r (O,2) = project(X_p; rvec_c, tvec_c, K) - uv, J_cam (O,2,6) = [d/drvec | d/dtvec],
    J_pt (O,2,3) = d/dX, all float32, and cost = sum ||r||^2 (python float, fp64 sum)."""
    return ops.ba_residual_jacobian(rvecs, tvecs, K, points, obs_cam, obs_pt, obs_uv)


def project_points(objectPoints, rvec, tvec, cameraMatrix, distCoeffs=None):
    """cv2.projectPoints(X, rvec, tvec, K, D) -> (imagePoints (N,1,2), jacobian (2N,6)).

    Only the [rvec | tvec] Jacobian block (OpenCV columns 0:6, rows 2i = u,
    2i+1 = v) is produced, in float64 like OpenCV; the intrinsic / distortion
    columns are not.  Non-zero distortion raises."""
    if distCoeffs is not None and np.any(np.asarray(distCoeffs, np.float64) != 0):
        raise NotImplementedError("Jacobians with lens distortion are not implemented")
    X = np.asarray(objectPoints, np.float64).reshape(-1, 3)
    n = len(X)
    zeros = np.zeros((n, 2), np.float32)
    r, Jc, _, _ = ops.ba_residual_jacobian(
        np.asarray(rvec, np.float32).reshape(1, 3), np.asarray(tvec, np.float32).reshape(1, 3),
        cameraMatrix, X.astype(np.float32), np.zeros(n, np.int32), np.arange(n, dtype=np.int32), zeros)
    return r.astype(np.float64).reshape(n, 1, 2), Jc.astype(np.float64).reshape(2 * n, 6)
