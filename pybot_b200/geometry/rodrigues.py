"""Rotation vector <-> matrix (host side, float64).

``Camera.project`` in the reference converts its rotation matrix to a rotation
vector with cv2.Rodrigues and cv2.projectPoints converts it back
(/root/reference/pybot/vision/camera_utils.py:694, :697).  The kernel needs the
matrix AFTER that round trip, so the same two conversions are done here, from
the published axis-angle formulas, for the nine numbers that parameterise a
launch.
"""
import math

import numpy as np


def matrix_to_rvec(R):
    R = np.asarray(R, np.float64).reshape(3, 3)
    U, _, Vt = np.linalg.svd(R)               # nearest rotation, as OpenCV does
    R = U @ Vt
    ax = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = math.sqrt((ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]) * 0.25)
    c = min(1.0, max(-1.0, (R[0, 0] + R[1, 1] + R[2, 2] - 1) * 0.5))
    theta = math.acos(c)
    if s < 1e-5:
        if c > 0:
            return np.zeros(3)
        # theta ~ pi: recover the axis from the symmetric part
        v = np.sqrt(np.maximum((np.diag(R) + 1) * 0.5, 0.0))
        if R[0, 1] < 0:
            v[1] = -v[1]
        if R[0, 2] < 0:
            v[2] = -v[2]
        if abs(v[0]) < abs(v[1]) and abs(v[0]) < abs(v[2]) and (R[1, 2] > 0) != (v[1] * v[2] > 0):
            v[2] = -v[2]
        return v * (theta / math.sqrt(float(v @ v)))
    return ax * (theta / (2.0 * s))


def rvec_to_matrix(rvec):
    r = np.asarray(rvec, np.float64).reshape(3)
    theta = math.sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2])
    if theta < np.finfo(np.float64).eps:
        return np.eye(3)
    c, s = math.cos(theta), math.sin(theta)
    u = r * (1.0 / theta)
    K = np.array([[0, -u[2], u[1]], [u[2], 0, -u[0]], [-u[1], u[0], 0]])
    return c * np.eye(3) + (1.0 - c) * np.outer(u, u) + s * K


def rodrigues_round_trip(R):
    """The rotation cv2.projectPoints actually applies for Camera.project."""
    return rvec_to_matrix(matrix_to_rvec(R))
