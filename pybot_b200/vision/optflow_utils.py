"""Scene flow from relative pose and scene depth - the hot-path caller
``SceneFlow`` of /root/reference/pybot/vision/optflow_utils.py:114-150, same
class and method names.

The reference projects H*W points with cv2.projectPoints and then makes five
numpy passes (int cast, bounds mask, two gathers, scatter, subtract).  Here the
projection, cast, mask and scatter are one kernel and the subtraction another
(pbk_sceneflow); only the (H,W,2) flow comes back.
"""
import numpy as np

from .. import _ffi, ops
from .camera_utils import Camera, CameraExtrinsic


class SceneFlow(object):
    """Computes the scene flow vectors given relative pose and scene depth
    (static scenes)."""

    def __init__(self, cam):
        assert cam.shape is not None
        H, W = cam.shape[:2]
        self.cam_ = Camera.from_intrinsics_extrinsics(cam, CameraExtrinsic.identity())
        self.xs_, self.ys_ = np.meshgrid(np.arange(0, W), np.arange(0, H))
        self.grid_ = np.dstack([self.xs_, self.ys_]).astype(np.float32)
        Rp, t, Rz, tz = self.cam_._projection_params()
        self._proj = ops.Projector(Rp, t, self.cam_.K, self.cam_.D, Rz=Rz, tz=tz)

    def process(self, dX):
        """dX: (H, W, 3) scene points in the camera frame -> (H, W, 2) float32
        displacement of every pixel that some point lands on, NaN elsewhere."""
        shp = tuple(int(v) for v in dX.shape[:2])
        if shp != (int(self.cam_.shape[0]), int(self.cam_.shape[1])):
            raise ValueError("dX shape does not agree with cam.shape, dX.shape needs to be [H x W x 3] "
                             "dX: {}, cam.shape: {}".format(np.int32(shp), self.cam_.shape[:2]))
        return ops.sceneflow(dX, self._proj.params)
