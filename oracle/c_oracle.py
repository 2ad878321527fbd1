"""ctypes wrapper of oracle/c/hamming_knn.c.  TEST INFRASTRUCTURE ONLY."""
import ctypes

import numpy as np

from . import build as _build

_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
        _lib.pbo_hamming_knn2.restype = None
        _lib.pbo_hamming_knn2.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                          ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
        _lib.pbo_hamming_knnk.restype = None
        _lib.pbo_hamming_knnk.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                          ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p]
    return _lib


def hamming_knn2(q, t):
    """(idx (nq,2) int64, dist (nq,2) int32) with BFMatcher tie semantics."""
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    assert q.ndim == 2 and q.shape[1] == 32 and t.ndim == 2 and t.shape[1] == 32
    idx = np.empty((len(q), 2), np.int64)
    dist = np.empty((len(q), 2), np.int32)
    _load().pbo_hamming_knn2(q.ctypes.data, len(q), t.ctypes.data, len(t), idx.ctypes.data,
                             dist.ctypes.data)
    return idx, dist


def hamming_knnk(q, t, k, mask=None):
    """knnMatch(q, t, k, mask): (idx (nq,k) int64, dist (nq,k) int32, count (nq,) int32), -1 past
    the neighbours found; mask (nq, nt), non-zero = allowed."""
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    assert q.ndim == 2 and q.shape[1] == 32 and t.ndim == 2 and t.shape[1] == 32 and k >= 1
    m = None
    if mask is not None:
        m = np.ascontiguousarray(np.asarray(mask) != 0).view(np.uint8)
        assert m.shape == (len(q), len(t))
    idx = np.empty((len(q), k), np.int64)
    dist = np.empty((len(q), k), np.int32)
    count = np.empty((len(q),), np.int32)
    _load().pbo_hamming_knnk(q.ctypes.data, len(q), t.ctypes.data, len(t), int(k),
                             m.ctypes.data if m is not None else None, idx.ctypes.data, dist.ctypes.data,
                             count.ctypes.data)
    return idx, dist, count
