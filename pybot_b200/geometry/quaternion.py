"""Unit quaternion (x, y, z, w) - host-side scalar algebra.

API of /root/reference/pybot/geometry/quaternion.py:12-228: the method and
property names, the float64 (x, y, z, w) storage and the "renormalise only when
the norm is off by more than 1e-6" rule (:49-53) are the reference's, because
callers and the bit-exact golden g0 (tests/test_host_geometry.py) depend on them.
Where the golden pins bits (rotate, interpolate) the floating-point operations
are applied in the reference's order; the code itself is vectorised numpy.
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
        """R(q) v for a 3-vector as  v + 2 M v  with M = R(q) / 2 - I / 2 written out in products
        of the components (reference :69-85).  Each row is summed left to right and doubled
        before v is added - the order the golden pose compositions are pinned to."""
        x, y, z, w = self.q
        xx, yy, zz = x * x, y * y, z * z
        wx, wy, wz = w * x, w * y, w * z
        xy, xz, yz = x * y, x * z, y * z
        half = np.array(((-yy - zz, xy - wz, wy + xz),
                         (wz + xy, -xx - zz, yz - wx),
                         (xz - wy, wx + yz, -xx - yy)))
        v = np.asarray(v)
        return 2 * ((half[:, 0] * v[0] + half[:, 1] * v[1]) + half[:, 2] * v[2]) + v[:3]

    def interpolate(self, other, this_weight):
        """Slerp between self (weight this_weight) and other, taking the short way round
        (reference :100-125: evaluated on the w-first vectors)."""
        far_weight = 1 - this_weight
        assert 0 <= far_weight <= 1
        mine, theirs = np.roll(self.q, 1), np.roll(other.q, 1)
        c = float(np.dot(mine, theirs))
        if c < 0:                                   # antipodal representation of the same rotation
            mine, c = -mine, -c
        angle = math.acos(min(c, 1))
        s = math.sin(angle)
        if abs(s) < 1e-6:                           # (nearly) the same rotation: blend and renormalise
            out = mine * this_weight + theirs * far_weight
            out = out / math.sqrt(np.dot(out, out))
        else:
            out = mine * (math.sin((1 - far_weight) * angle) / s) + theirs * (math.sin(far_weight * angle) / s)
        return Quaternion(np.roll(out, -1))

    # -------------------------------------------------------- conversions
    def to_wxyz(self):
        return np.roll(self.q, 1)

    def to_xyzw(self):
        return self.q

    def to_rpy(self, axes="rxyz"):
        return np.array(tf.euler_from_quaternion(self.q, axes=axes))

    def to_angle_axis(self):
        """(angle, unit axis); the identity reports angle 0 about +z."""
        half_angle = math.acos(self.q[3])
        if abs(half_angle) < 1e-12:
            return 0, np.array((0, 0, 1))
        return half_angle * 2, self.q[:3] / math.sin(half_angle)

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
        """Rotation by theta about `axis` (any length; the zero axis gives the identity)."""
        ax, ay, az = axis
        length = math.sqrt(ax * ax + ay * ay + az * az)
        if length == 0:
            return cls.identity()
        scale = math.sin(theta / 2) / length
        return cls([ax * scale, ay * scale, az * scale, math.cos(theta / 2)])

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
