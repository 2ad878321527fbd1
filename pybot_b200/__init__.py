"""pybot_b200 - B200-native (sm_100a) implementation of pybot's data-parallel
geometry-and-matching hot path, behind pybot's own Python API.

    from pybot_b200.geometry import RigidTransform, Sim3
    from pybot_b200.vision.camera_utils import Camera, StereoCamera
    from pybot_b200.vision.matcher import BFMatcher
    from pybot_b200.vision.ba import ba_residual_jacobian

Compute goes through libpybot_b200.so (hand-written CUDA, C-ABI in
include/pybot_b200.h).  There is no CPU fallback.
"""
__version__ = "0.1.0"
