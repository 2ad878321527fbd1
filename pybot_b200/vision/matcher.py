"""Brute-force Hamming matcher with cv2.BFMatcher semantics, on the GPU.

The reference never matches descriptors itself; a pybot user
writes ``cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)`` + Lowe ratio +
mutual check next to it.  This module keeps that OpenCV surface
(``BFMatcher.add / clear / getTrainDescriptors / knnMatch / match``, ``DMatch``)
and adds an array API for the fused pipeline.  Semantics pinned against
OpenCV 4.13: ascending distance, ties to the lowest
(imgIdx, trainIdx); rows shorter than k when the train set is smaller.
"""
import numpy as np

from .. import _ffi, ops

NORM_HAMMING = 6          # cv2.NORM_HAMMING


class DMatch(object):
    """Same four fields as cv2.DMatch."""
    __slots__ = ("queryIdx", "trainIdx", "imgIdx", "distance")

    def __init__(self, queryIdx=-1, trainIdx=-1, imgIdx=-1, distance=float("inf")):
        self.queryIdx, self.trainIdx, self.imgIdx, self.distance = (
            int(queryIdx), int(trainIdx), int(imgIdx), float(distance))

    def __repr__(self):
        return "DMatch(queryIdx=%d, trainIdx=%d, imgIdx=%d, distance=%.1f)" % (
            self.queryIdx, self.trainIdx, self.imgIdx, self.distance)


class KeyframeDatabase(object):
    """A collection of per-keyframe descriptor matrices resident in HBM as one
    contiguous (N,32) uint8 tensor + row offsets (global row = offset[imgIdx]
    + trainIdx, the ordering BFMatcher.add([...]) uses).

    Loop closure adds one keyframe (2000 rows, 64 KB) and then matches against
    everything so far, thousands of times: rows are appended in place into a
    buffer that grows geometrically, so an add costs its own bytes and `flat` is a
    view - never a concatenation of the whole database."""

    MIN_CAPACITY = 1 << 16          # rows (2 MB)

    def __init__(self):
        self._host = []
        self._buf = None            # (capacity, 32) uint8 on the device
        self._n = 0
        self.offsets = np.zeros(1, np.int64)

    def reserve(self, rows):
        """Make room for `rows` descriptors in total (e.g. the expected session length)."""
        T = _ffi.require_cuda()
        cap = 0 if self._buf is None else self._buf.shape[0]
        if rows <= cap:
            return
        new_cap = max(int(rows), self.MIN_CAPACITY, cap + cap // 2)
        buf = T.empty((new_cap, 32), dtype=T.uint8, device="cuda")
        if self._n:
            buf[:self._n].copy_(self._buf[:self._n])
        self._buf = buf

    def add(self, descriptors):
        _ffi.require_cuda()
        for d in descriptors:
            dd = ops._desc_in(d)
            k = int(dd.shape[0])
            if k:
                self.reserve(self._n + k)
                self._buf[self._n:self._n + k].copy_(dd)
                self._n += k
            self._host.append(d)
            self.offsets = np.append(self.offsets, self.offsets[-1] + k)

    def clear(self):
        self.__init__()

    def descriptors(self):
        return list(self._host)

    def __len__(self):
        return int(self.offsets[-1])

    @property
    def flat(self):
        T = _ffi.require_cuda()
        if self._buf is None:
            return T.empty((0, 32), dtype=T.uint8, device="cuda")
        return self._buf[:self._n]

    def split_global(self, g):
        """global rows -> (imgIdx, trainIdx)."""
        g = np.asarray(g, np.int64)
        img = np.searchsorted(self.offsets, g, side="right") - 1
        img = np.clip(img, 0, max(len(self.offsets) - 2, 0))
        return img.astype(np.int32), (g - self.offsets[img]).astype(np.int32)


class BFMatcher(object):
    """cv2.BFMatcher(normType=cv2.NORM_HAMMING, crossCheck=False) replacement."""

    def __init__(self, normType=NORM_HAMMING, crossCheck=False):
        if normType != NORM_HAMMING:
            raise NotImplementedError("only NORM_HAMMING (256-bit ORB descriptors) is implemented")
        self.crossCheck = bool(crossCheck)
        self.db = KeyframeDatabase()

    # -- collection management (cv2.DescriptorMatcher)
    def add(self, descriptors):
        self.db.add(descriptors)

    def clear(self):
        self.db.clear()

    def empty(self):
        return len(self.db) == 0

    def getTrainDescriptors(self):
        return self.db.descriptors()

    # -- array API ---------------------------------------------------------
    def knn_arrays(self, queryDescriptors, trainDescriptors=None, ratio=None, mutual=False):
        """Top-2 as arrays: dict(idx (nq,2) global rows, imgIdx, trainIdx,
        dist (nq,2) int32, ratio_ok, mutual_ok, valid)."""
        if trainDescriptors is not None:
            db = KeyframeDatabase()
            db.add([trainDescriptors])
        else:
            db = self.db
        idx, dist, r_ok, m_ok, valid = ops.knn_ratio_mutual(
            queryDescriptors, db.flat, ratio=1.0 if ratio is None else ratio, mutual=mutual)
        idx, dist = np.asarray(_np(idx)), np.asarray(_np(dist))
        img, tr = db.split_global(np.maximum(idx, 0))
        img[idx < 0] = -1
        tr[idx < 0] = -1
        return dict(idx=idx, imgIdx=img, trainIdx=tr, dist=dist, ratio_ok=_np(r_ok),
                    mutual_ok=_np(m_ok), valid=_np(valid))

    # -- OpenCV API --------------------------------------------------------
    def knnMatch(self, queryDescriptors, trainDescriptors=None, k=2, mask=None, compactResult=False):
        if mask is not None:
            raise NotImplementedError("masks are not supported")
        if k not in (1, 2):
            raise NotImplementedError("k must be 1 or 2")
        if self.crossCheck and k != 1:
            raise ValueError("crossCheck=True needs k == 1 (OpenCV asserts the same)")
        r = self.knn_arrays(queryDescriptors, trainDescriptors, mutual=self.crossCheck)
        out = []
        for i in range(len(r["idx"])):
            row = []
            for j in range(k):
                if r["idx"][i, j] < 0:
                    break
                if self.crossCheck and not r["mutual_ok"][i]:
                    break
                row.append(DMatch(i, r["trainIdx"][i, j], r["imgIdx"][i, j], r["dist"][i, j]))
            if row or not compactResult:
                out.append(row)
        return out

    def match(self, queryDescriptors, trainDescriptors=None, mask=None):
        return [row[0] for row in self.knnMatch(queryDescriptors, trainDescriptors, k=1, mask=mask)
                if row]


def _np(x):
    if _ffi.is_tensor(x):
        return _ffi.to_host(x) if x.is_cuda else x.numpy()
    return x


def knn_ratio_mutual(q, t, ratio=0.75):
    """Array pipeline of the north star: forward k=2, Lowe ratio, mutual check.
    Returns idx (nq,2) int64, dist (nq,2) int32, ratio_ok, mutual_ok, valid (bool)."""
    return ops.knn_ratio_mutual(q, t, ratio=ratio, mutual=True)
