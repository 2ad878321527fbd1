"""Import the unmodified reference (spillai/pybot) from /root/reference.

Test infrastructure (see oracle/__init__.py).  Only works in the authoring
container; on the GPU box /root/reference does not exist and ``available()``
returns False.  Recipe from SURVEY.md Appendix B: the reference's
``pybot.utils.db_utils`` imports PyTables and ``pybot.vision.color_utils``
imports matplotlib, neither of which is installed, so inert stubs are injected.
"""
import os
import sys
import types
import unittest.mock

REFERENCE_ROOT = os.environ.get("PYBOT_REFERENCE_ROOT", "/root/reference")

_loaded = None


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pybot"))


def load():
    """Returns a namespace with the reference classes on the hot path."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    for name in ("tables", "matplotlib", "matplotlib.pyplot", "matplotlib.cm",
                 "matplotlib.colors"):
        if name not in sys.modules:
            sys.modules[name] = unittest.mock.MagicMock()
    sys.modules["tables"].NaturalNameWarning = type(
        "NaturalNameWarning", (Warning,), {})
    from pybot.geometry.rigid_transform import RigidTransform, Sim3, Pose
    from pybot.geometry.quaternion import Quaternion
    from pybot.geometry import transformations as tf
    from pybot.vision.camera_utils import (
        Camera, StereoCamera, CameraIntrinsic, CameraExtrinsic, DepthCamera)
    ns = types.SimpleNamespace(
        RigidTransform=RigidTransform, Sim3=Sim3, Pose=Pose,
        Quaternion=Quaternion, tf=tf, Camera=Camera, StereoCamera=StereoCamera,
        CameraIntrinsic=CameraIntrinsic, CameraExtrinsic=CameraExtrinsic,
        DepthCamera=DepthCamera)
    _loaded = ns
    return ns
