"""ctypes binding of libpybot_b200.so (include/pybot_b200.h).

This is the only place the shared library is loaded.  There is no CPU
fallback: if the library is missing or no CUDA device is usable, every compute
call raises.  PyTorch is used purely as a carrier (device memory, pinned host
memory, streams).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PYBOT_B200_LIB", os.path.join(_HERE, "libpybot_b200.so"))

PBK_OK, PBK_EINVAL, PBK_ECUDA, PBK_EALIGN, PBK_EUNSUPPORTED = 0, -1, -2, -3, -4
PBK_U8, PBK_S16, PBK_S32, PBK_F32, PBK_F64, PBK_U16 = 0, 1, 2, 3, 4, 5
PBK_KEY_NONE = 0xFFFFFFFFFFFFFFFF
PBK_STATUS_PEER_TIMEOUT = 1
PBK_BA_CAM_DOUBLES = 40

_NP2PBK = {np.dtype(np.uint8): PBK_U8, np.dtype(np.int16): PBK_S16,
           np.dtype(np.int32): PBK_S32, np.dtype(np.float32): PBK_F32,
           np.dtype(np.float64): PBK_F64, np.dtype(np.uint16): PBK_U16}


class ProjectParams(ctypes.Structure):
    """struct pbk_project_params."""
    _fields_ = [
        ("R", ctypes.c_double * 9), ("t", ctypes.c_double * 3),
        ("Rz", ctypes.c_double * 3), ("tz", ctypes.c_double),
        ("fx", ctypes.c_double), ("fy", ctypes.c_double),
        ("cx", ctypes.c_double), ("cy", ctypes.c_double),
        ("k", ctypes.c_double * 8), ("min_depth", ctypes.c_double),
        ("width", ctypes.c_int32), ("height", ctypes.c_int32),
        ("check_depth", ctypes.c_int32), ("check_bounds", ctypes.c_int32),
        ("has_distortion", ctypes.c_int32), ("compact", ctypes.c_int32),
    ]


class RayParams(ctypes.Structure):
    """struct pbk_ray_params."""
    _fields_ = [
        ("fx", ctypes.c_double), ("fy", ctypes.c_double), ("cx", ctypes.c_double), ("cy", ctypes.c_double),
        ("k", ctypes.c_double * 8), ("P", ctypes.c_double * 9), ("rot", ctypes.c_double * 9),
        ("undistort", ctypes.c_int32), ("rotate", ctypes.c_int32),
        ("normalize", ctypes.c_int32), ("scale_by_z", ctypes.c_int32),
    ]


_vp, _i, _i64, _u64, _d, _sz = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int64,
                                ctypes.c_uint64, ctypes.c_double, ctypes.c_size_t)

# name -> (restype, argtypes); must list every symbol include/pybot_b200.h declares
SIGNATURES = {
    "pbk_version": (_i, []),
    "pbk_last_error_string": (ctypes.c_char_p, []),
    "pbk_device_props": (_i, [_i, ctypes.POINTER(_i64)]),
    "pbk_hbm_probe": (_i, [_i, _vp, _vp, _i64, _vp, _vp]),
    "pbk_reproject_disp": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "pbk_reproject_sparse": (_i, [_vp, _i64, _vp, _vp, _vp]),
    "pbk_transform_points": (_i, [_vp, _i, _i64, _vp, _d, _vp, _i, _vp]),
    "pbk_set_exchange_timeout_ms": (_i, [_d]),
    "pbk_project_scratch_bytes": (_sz, [_i64, _i]),
    "pbk_project_points": (_i, [_vp, _i, _i64, ctypes.POINTER(ProjectParams),
                                _vp, _vp, _vp, _vp, _vp, _vp]),
    "pbk_hamming_plan": (_i, [_i64, _i64, ctypes.POINTER(_i), ctypes.POINTER(_sz)]),
    "pbk_hamming_knn2": (_i, [_vp, _i64, _vp, _i64, _u64, _vp, _vp, _vp]),
    "pbk_hamming_merge_top2": (_i, [_vp, _i, _i64, _vp, _vp]),
    "pbk_hamming_gather_best": (_i, [_vp, _i64, _u64, _vp, _i64, _vp, _vp]),
    "pbk_hamming_ratio_mutual": (_i, [_vp, _vp, _i64, _d, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pbk_hamming_knnk": (_i, [_vp, _i64, _vp, _i64, _i, _vp, _u64, _vp, _vp, _vp, _vp]),
    "pbk_hamming_exchange_words": (_sz, [_i64, _i]),
    "pbk_hamming_knn2_dist": (_i, [_vp, _i64, _vp, _i64, _u64, _vp, _vp, _vp, _i, _i, _u64, _vp, _vp, _vp]),
    "pbk_hamming_gather_best_dist": (_i, [_vp, _i64, _u64, _vp, _i64, _vp, _vp, _i, _i, _u64, _vp, _vp, _vp]),
    "pbk_hamming_tc_format": (_i, []),
    "pbk_hamming_tc_set_format": (_i, [_i]),
    "pbk_hamming_tc_signs_bytes": (_sz, [_i64]),
    "pbk_hamming_tc_expand": (_i, [_vp, _i64, _i64, _vp, _vp]),
    "pbk_hamming_tc_plan": (_i, [_i64, _i64, ctypes.POINTER(_i), ctypes.POINTER(_sz)]),
    "pbk_hamming_tc_knn2": (_i, [_vp, _i64, _vp, _i64, _u64, _vp, _vp, _vp]),
    "pbk_hamming_tc_knn2_dist": (_i, [_vp, _i64, _vp, _i64, _u64, _vp, _vp, _vp, _i, _i, _u64, _vp, _vp, _vp]),
    "pbk_hamming_tc_peak_probe": (_i, [_i, _vp, ctypes.POINTER(_d), _vp]),
    "pbk_popc_peak_probe": (_i, [_i, _i, _i, _vp, ctypes.POINTER(_d), _vp]),
    "pbk_depth_reconstruct": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "pbk_depth_reconstruct_sparse": (_i, [_vp, _vp, _i64, _d, _d, _d, _vp, _vp]),
    "pbk_sceneflow": (_i, [_vp, _i, _i, _i, ctypes.POINTER(ProjectParams), _vp, _vp, _vp]),
    "pbk_fov_mask": (_i, [_vp, _i, _i64, _vp, _d, _d, _d, _d, _vp, _vp]),
    "pbk_transform_depth": (_i, [_vp, _i, _i64, _i64, _vp, _vp, _vp]),
    "pbk_sampson_error": (_i, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "pbk_epipolar_lines": (_i, [_vp, _vp, _i, _i64, _vp, _vp]),
    "pbk_triangulate_points": (_i, [_vp, _vp, _vp, _vp, _i, _i64, _vp, _vp]),
    "pbk_undistort_points": (_i, [_vp, _i, _i, _i64, ctypes.POINTER(RayParams), _vp, _vp]),
    "pbk_camera_rays": (_i, [_vp, _i, _i, _i64, ctypes.POINTER(RayParams), _vp, _vp]),
    "pbk_wire_colors": (_i, [_vp, _i64, _i, _vp, _vp]),
    "pbk_ba_camera_precompute": (_i, [_vp, _vp, _i, _vp, _vp]),
    "pbk_ba_cost_slots": (_i, [_i64]),
    "pbk_ba_residual_jacobian": (_i, [_vp, _i64, _vp, _i64, _vp, _i, _d, _d, _d, _d,
                                      _vp, _vp, _vp, _vp, _vp, _vp]),
    "pbk_ba_residual_jacobian_dist": (_i, [_vp, _i64, _vp, _i64, _vp, _i, _d, _d, _d, _d,
                                           _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _u64, _vp, _vp]),
    "pbk_ba_evaluate": (_i, [_vp, _i64, _i, _vp, _i64, _vp, _vp, _i, _vp, _d, _d, _d, _d,
                             _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _u64, _vp, _vp]),
}

_lib = None


def lib():
    """Loads (once) and returns the ctypes handle; raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libpybot_b200.so not found at %s - run `python -m pybot_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        h = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)        # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


class PbkError(RuntimeError):
    pass


def check(rc):
    if rc == PBK_OK:
        return
    msg = lib().pbk_last_error_string().decode("utf-8", "replace")
    if rc == PBK_EINVAL or rc == PBK_EALIGN:
        raise ValueError("libpybot_b200: " + msg)
    if rc == PBK_EUNSUPPORTED:
        raise NotImplementedError("libpybot_b200: " + msg)
    raise PbkError("libpybot_b200: " + msg)


# ------------------------------------------------------------------ carriers
_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


def require_cuda():
    t = torch()
    if not t.cuda.is_available():
        raise RuntimeError("pybot_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    lib()
    return t


def stream_ptr():
    return ctypes.c_void_p(torch().cuda.current_stream().cuda_stream)


def ptr(tensor):
    return ctypes.c_void_p(tensor.data_ptr()) if tensor is not None else ctypes.c_void_p(0)


def pbk_dtype(np_dtype):
    try:
        return _NP2PBK[np.dtype(np_dtype)]
    except KeyError:
        raise TypeError("unsupported dtype %s" % np_dtype)


def is_tensor(x):
    t = _torch
    return t is not None and isinstance(x, t.Tensor)


def to_device(x, dtype=None):
    """numpy array / torch tensor -> contiguous CUDA tensor (+ whether the
    caller passed a device tensor, in which case results stay on device)."""
    t = require_cuda()
    if isinstance(x, t.Tensor):
        was_dev = x.is_cuda
        if dtype is not None:
            x = x.to(getattr(t, np.dtype(dtype).name))
        return x.cuda().contiguous(), was_dev
    a = np.ascontiguousarray(x if dtype is None else np.asarray(x, dtype))
    h = t.from_numpy(a)
    return h.to("cuda", non_blocking=h.is_pinned()), False


def to_host(tensor):
    """CUDA tensor -> numpy through pinned memory, synchronising the stream."""
    t = torch()
    out = t.empty(tensor.shape, dtype=tensor.dtype, pin_memory=True)
    out.copy_(tensor, non_blocking=True)
    t.cuda.current_stream().synchronize()
    return out.numpy()


class StatusWord(object):
    """uint32 status word the fused-exchange kernels write on failure: pinned host memory
    mapped into the device (UVA), so the host polls it without a copy or a synchronisation."""

    def __init__(self):
        t = require_cuda()
        self.word = t.zeros(4, dtype=t.int32, pin_memory=True)
        self.ptr = ctypes.c_void_p(self.word.data_ptr())

    def check(self, what):
        """Raises if a kernel reported a failure (call after the stream has been synchronised,
        or at the next call: the word is sticky until cleared here)."""
        code = int(self.word[0])
        if code:
            self.word[0] = 0
            if code == PBK_STATUS_PEER_TIMEOUT:
                raise PbkError("libpybot_b200: %s: a peer rank did not arrive within the exchange timeout "
                               "(outputs of that call are poisoned: no-match keys / zero rows / NaN cost)" % what)
            raise PbkError("libpybot_b200: %s failed with device status %d" % (what, code))


def pinned_empty(shape, dtype):
    """numpy array backed by pinned host memory (fast H2D/D2H)."""
    t = require_cuda()
    return t.empty(tuple(np.atleast_1d(shape)), dtype=getattr(t, np.dtype(dtype).name),
                   pin_memory=True).numpy()


def host_f32(a, n):
    arr = np.ascontiguousarray(np.asarray(a, np.float32).reshape(-1))
    assert arr.size == n
    return arr, arr.ctypes.data_as(ctypes.c_void_p)


def host_f64(a, n):
    arr = np.ascontiguousarray(np.asarray(a, np.float64).reshape(-1))
    assert arr.size == n
    return arr, arr.ctypes.data_as(ctypes.c_void_p)
