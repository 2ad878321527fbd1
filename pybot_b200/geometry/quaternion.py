"""Unit quaternion (x, y, z, w) - host-side scalar algebra.

API of /root/reference/pybot/geometry/quaternion.py:12-228 (same method and
property names, same float64 storage and the same "renormalise only when the
norm is off by more than 1e-6" rule, :49-53), written fresh on numpy.
"""
import math

import numpy as np

from . import transformations as tf


class Quaternion(object):
    def __init__(self, q=(0, 0, 0, 1)):
        if isinstance(q, Quaternion):
            self.q = q.q.copy()
        else:
            try:
                self.q = np.array(q, np.float64)
            except Exception:
                raise TypeError("Quaternion can not be initialized from {:}".format(type(q)))
        self.normalize()

    def __repr__(self):
        return "%s" % self.q

    def __getitem__(self, i):
        return self.q[i]

    def copy(self):
        return Quaternion(self.q.copy())

    # ------------------------------------------------------------ algebra
    def __mul__(self, other):
        if isinstance(other, float):
            return Quaternion(self.q * other)
        if isinstance(other, Quaternion):
            return Quaternion(tf.quaternion_multiply(self.q, other.q))
        raise TypeError("Quaternion multiply error")

    def normalize(self):
        n = self.norm()
        if abs(n - 1) > 1e-6:
            self.q /= n

    def norm(self):
        return np.linalg.norm(self.q)

    def dot(self, other):
        return self.q.dot(other.q)

    def inverse(self):
        return Quaternion(tf.quaternion_conjugate(self.q))

    conjugate = inverse

    def rotate(self, v):
        """Rotates a 3-vector; the expanded sandwich product of the reference
        (:69-85), same term order so pose compounding is bit-identical."""
        qx, qy, qz, qw = self.q
        ab, ac, ad = qw * qx, qw * qy, qw * qz
        nbb, bc, bd = -qx * qx, qx * qy, qx * qz
        ncc, cd, ndd = -qy * qy, qy * qz, -qz * qz
        return np.array((
            2 * ((ncc + ndd) * v[0] + (bc - ad) * v[1] + (ac + bd) * v[2]) + v[0],
            2 * ((ad + bc) * v[0] + (nbb + ndd) * v[1] + (cd - ab) * v[2]) + v[1],
            2 * ((bd - ac) * v[0] + (ab + cd) * v[1] + (nbb + ncc) * v[2]) + v[2]))

    def interpolate(self, other, this_weight):
        """Spherical interpolation in wxyz space (reference :100-125)."""
        q0, q1 = np.roll(self.q, 1), np.roll(other.q, 1)
        u = 1 - this_weight
        assert 0 <= u <= 1
        cos_omega = float(np.dot(q0, q1))
        if cos_omega < 0:
            start, cos_omega = -q0, -cos_omega
        else:
            start = q0.copy()
        cos_omega = min(cos_omega, 1)
        omega = math.acos(cos_omega)
        sin_omega = math.sin(omega)
        if abs(sin_omega) < 1e-6:
            mixed = start * this_weight + q1 * u
            mixed /= math.sqrt(np.dot(mixed, mixed))
        else:
            a = math.sin((1 - u) * omega) / sin_omega
            b = math.sin(u * omega) / sin_omega
            mixed = start * a + q1 * b
        return Quaternion(np.roll(mixed, -1))

    # -------------------------------------------------------- conversions
    def to_wxyz(self):
        return np.roll(self.q, 1)

    def to_xyzw(self):
        return self.q

    def to_rpy(self, axes="rxyz"):
        return np.array(tf.euler_from_quaternion(self.q, axes=axes))

    def to_angle_axis(self):
        w = np.roll(self.q, 1)
        half = math.acos(w[0])
        if abs(half) < 1e-12:
            return 0, np.array((0, 0, 1))
        return half * 2, np.array(w[1:4]) / math.sin(half)

    def to_matrix(self):
        return tf.quaternion_matrix(self.q)

    @classmethod
    def from_wxyz(cls, q):
        return cls(np.roll(q, -1))

    @classmethod
    def from_xyzw(cls, q):
        return cls(q)

    @classmethod
    def from_matrix(cls, matrix):
        # the reference returns the bare ndarray here (:166-169); callers wrap it
        return tf.quaternion_from_matrix(matrix)

    @classmethod
    def from_rpy(cls, roll, pitch, yaw, axes="rxyz"):
        return cls(tf.quaternion_from_euler(roll, pitch, yaw, axes=axes))

    @classmethod
    def from_angle_axis(cls, theta, axis):
        x, y, z = axis
        n = math.sqrt(x * x + y * y + z * z)
        if n == 0:
            return cls([0, 0, 0, 1])
        t = math.sin(theta / 2) / n
        return cls([x * t, y * t, z * t, math.cos(theta / 2)])

    @classmethod
    def identity(cls):
        return cls()

    # ---------------------------------------------------------- properties
    matrix = property(lambda self: self.to_matrix())
    R = property(lambda self: self.to_matrix()[:3, :3])
    x = property(lambda self: self.q[0])
    y = property(lambda self: self.q[1])
    z = property(lambda self: self.q[2])
    w = property(lambda self: self.q[3])
    wxyz = property(lambda self: self.to_wxyz())
    xyzw = property(lambda self: self.to_xyzw())
    rpy = property(lambda self: self.to_rpy())
