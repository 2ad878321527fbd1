"""Quaternion / Euler helpers used by the pose classes (host side, float64).

Only the functions the hot path reaches are provided, with the reference's
names and its ROS-tf **xyzw** quaternion order
(/root/reference/pybot/geometry/transformations.py:1031-1260).  Written from
the underlying formulas; the floating-point evaluation order of
``quaternion_matrix`` / ``quaternion_from_matrix`` / ``quaternion_multiply`` is
kept identical to the reference so a pose's 4x4 matrix - the parameter block
every kernel launch receives - is bit-identical to the reference's.
"""
import math

import numpy as np

_EPS = np.finfo(float).eps * 4.0

# axis sequence -> (first axis, parity, repetition, frame); 's' static, 'r' rotating
_SEQUENCES = {}
for _first, _name in enumerate("xyz"):
    _n1, _n2 = "xyz"[(_first + 1) % 3], "xyz"[(_first + 2) % 3]
    _SEQUENCES["s" + _name + _n1 + _n2] = (_first, 0, 0, 0)
    _SEQUENCES["s" + _name + _n1 + _name] = (_first, 0, 1, 0)
    _SEQUENCES["s" + _name + _n2 + _n1] = (_first, 1, 0, 0)
    _SEQUENCES["s" + _name + _n2 + _name] = (_first, 1, 1, 0)
for _k, (_f, _p, _r, _) in list(_SEQUENCES.items()):
    # rotating-frame sequence = static sequence read backwards
    _SEQUENCES["r" + _k[:0:-1]] = (_f, _p, _r, 1)
_AXES2TUPLE = dict(_SEQUENCES)
_NEXT = (1, 2, 0, 1)


def _sequence(axes):
    if isinstance(axes, str):
        return _SEQUENCES[axes.lower()]
    f, p, r, fr = axes
    return int(f), int(p), int(r), int(fr)


def quaternion_matrix(quaternion):
    """4x4 homogeneous rotation from an xyzw quaternion (reference :1174-1193)."""
    q = np.array(quaternion, dtype=np.float64).ravel()[:4].copy()
    nq = float(np.dot(q, q))
    if nq < _EPS:
        return np.identity(4)
    q *= math.sqrt(2.0 / nq)
    x, y, z, w = (float(v) for v in q)
    xx, yy, zz = x * x, y * y, z * z
    xy, xz, xw = x * y, x * z, x * w
    yz, yw, zw = y * z, y * w, z * w
    M = np.identity(4)
    M[0, 0] = 1.0 - yy - zz
    M[0, 1] = xy - zw
    M[0, 2] = xz + yw
    M[1, 0] = xy + zw
    M[1, 1] = 1.0 - xx - zz
    M[1, 2] = yz - xw
    M[2, 0] = xz - yw
    M[2, 1] = yz + xw
    M[2, 2] = 1.0 - xx - yy
    return M


def quaternion_from_matrix(matrix):
    """xyzw quaternion from a 4x4 homogeneous rotation (reference :1196-1225)."""
    M = np.asarray(matrix, dtype=np.float64)[:4, :4]
    q = np.empty(4)
    tr = M[0, 0] + M[1, 1] + M[2, 2] + M[3, 3]
    if tr > M[3, 3]:
        t = tr
        q[3] = t
        q[2] = M[1, 0] - M[0, 1]
        q[1] = M[0, 2] - M[2, 0]
        q[0] = M[2, 1] - M[1, 2]
    else:
        i = 0
        if M[1, 1] > M[0, 0]:
            i = 1
        if M[2, 2] > M[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        t = M[i, i] - (M[j, j] + M[k, k]) + M[3, 3]
        q[i] = t
        q[j] = M[i, j] + M[j, i]
        q[k] = M[k, i] + M[i, k]
        q[3] = M[k, j] - M[j, k]
    q *= 0.5 / math.sqrt(t * M[3, 3])
    return q


def quaternion_multiply(quaternion1, quaternion0):
    """Hamilton product q1*q0, xyzw (reference :1228-1242)."""
    x0, y0, z0, w0 = quaternion0
    x1, y1, z1, w1 = quaternion1
    return np.array([x1 * w0 + y1 * z0 - z1 * y0 + w1 * x0,
                     -x1 * z0 + y1 * w0 + z1 * x0 + w1 * y0,
                     x1 * y0 - y1 * x0 + z1 * w0 + w1 * z0,
                     -x1 * x0 - y1 * y0 - z1 * z0 + w1 * w0], dtype=np.float64)


def quaternion_conjugate(quaternion):
    q = np.array(quaternion, dtype=np.float64)
    return np.array([-q[0], -q[1], -q[2], q[3]])


def quaternion_inverse(quaternion):
    return quaternion_conjugate(quaternion) / np.dot(quaternion, quaternion)


def quaternion_about_axis(angle, axis):
    q = np.zeros(4)
    q[:3] = np.asarray(axis, np.float64)[:3]
    n = math.sqrt(float(np.dot(q, q)))
    if n > _EPS:
        q *= math.sin(angle / 2.0) / n
    q[3] = math.cos(angle / 2.0)
    return q


def quaternion_from_euler(ai, aj, ak, axes="sxyz"):
    """xyzw quaternion from Euler angles for any of the 24 axis sequences
    (reference :1100-1154)."""
    first, parity, repetition, frame = _sequence(axes)
    i = first
    j = _NEXT[i + parity]
    k = _NEXT[i - parity + 1]
    if frame:
        ai, ak = ak, ai
    if parity:
        aj = -aj
    ai, aj, ak = ai / 2.0, aj / 2.0, ak / 2.0
    ci, si = math.cos(ai), math.sin(ai)
    cj, sj = math.cos(aj), math.sin(aj)
    ck, sk = math.cos(ak), math.sin(ak)
    cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
    q = np.empty(4)
    if repetition:
        q[i] = cj * (cs + sc)
        q[j] = sj * (cc + ss)
        q[k] = sj * (cs - sc)
        q[3] = cj * (cc - ss)
    else:
        q[i] = cj * sc - sj * cs
        q[j] = cj * ss + sj * cc
        q[k] = cj * cs - sj * sc
        q[3] = cj * cc + sj * ss
    if parity:
        q[j] *= -1
    return q


def euler_from_matrix(matrix, axes="sxyz"):
    """Euler angles of a rotation matrix (reference :1031-1086)."""
    first, parity, repetition, frame = _sequence(axes)
    i = first
    j = _NEXT[i + parity]
    k = _NEXT[i - parity + 1]
    M = np.asarray(matrix, dtype=np.float64)[:3, :3]
    if repetition:
        sy = math.sqrt(M[i, j] * M[i, j] + M[i, k] * M[i, k])
        if sy > _EPS:
            ax = math.atan2(M[i, j], M[i, k])
            ay = math.atan2(sy, M[i, i])
            az = math.atan2(M[j, i], -M[k, i])
        else:
            ax = math.atan2(-M[j, k], M[j, j])
            ay = math.atan2(sy, M[i, i])
            az = 0.0
    else:
        cy = math.sqrt(M[i, i] * M[i, i] + M[j, i] * M[j, i])
        if cy > _EPS:
            ax = math.atan2(M[k, j], M[k, k])
            ay = math.atan2(-M[k, i], cy)
            az = math.atan2(M[j, i], M[i, i])
        else:
            ax = math.atan2(-M[j, k], M[j, j])
            ay = math.atan2(-M[k, i], cy)
            az = 0.0
    if parity:
        ax, ay, az = -ax, -ay, -az
    if frame:
        ax, az = az, ax
    return ax, ay, az


def euler_from_quaternion(quaternion, axes="sxyz"):
    return euler_from_matrix(quaternion_matrix(quaternion), axes)
