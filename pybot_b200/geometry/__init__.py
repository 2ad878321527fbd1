"""pybot.geometry drop-in: same names as /root/reference/pybot/geometry/__init__.py:1."""
from .rigid_transform import DualQuaternion, Pose, RigidTransform, Quaternion, Sim3, rpyxyz  # noqa: F401
