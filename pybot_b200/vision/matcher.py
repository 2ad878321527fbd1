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


class KnnMatches(object):
    """What knnMatch returns: a read-only sequence of rows of DMatch (row i = the <= k
    neighbours of query i, ascending distance), exactly the shape cv2 returns, backed by the
    result ARRAYS.  DMatch objects are built when a row is indexed or iterated, not in the
    call: 4000 Python objects per frame cost more host time than the whole 8-GPU database
    scan.  `.arrays` exposes the columns for vectorised use (Lowe ratio in one numpy line).
    Compares equal to the equivalent list of lists."""
    __slots__ = ("_q", "_tr", "_img", "_dist", "_n", "arrays")

    def __init__(self, query_rows, trainIdx, imgIdx, dist, counts, arrays=None):
        self._q, self._tr, self._img, self._dist, self._n = query_rows, trainIdx, imgIdx, dist, counts
        self.arrays = arrays

    def __len__(self):
        return len(self._q)

    def _row(self, r):
        i = int(self._q[r])
        return [DMatch(i, self._tr[i, j], self._img[i, j], self._dist[i, j]) for j in range(int(self._n[r]))]

    def __getitem__(self, r):
        if isinstance(r, slice):
            return [self._row(x) for x in range(*r.indices(len(self)))]
        if r < 0:
            r += len(self)
        if not 0 <= r < len(self):
            raise IndexError("match row out of range")
        return self._row(r)

    def __iter__(self):
        for r in range(len(self)):
            yield self._row(r)

    def __eq__(self, other):
        try:
            return len(other) == len(self) and all(
                len(a) == len(b) and all(_same_match(x, y) for x, y in zip(a, b)) for a, b in zip(self, other))
        except TypeError:
            return NotImplemented

    def tolist(self):
        return list(self)

    def __repr__(self):
        return "KnnMatches(%d query rows)" % len(self)


def _same_match(a, b):
    return (a.queryIdx, a.trainIdx, a.imgIdx, a.distance) == (b.queryIdx, b.trainIdx, b.imgIdx, b.distance)


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
        self._signs = None          # expanded copy for the tensor-core scan (ops.hamming_tc_expand)
        self._signs_n = 0           # rows it is valid for
        self._signs_fmt = 0         # ops.hamming_tc_format() it was expanded in

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
            if self._host is not None:
                self._host.append(d)
            self.offsets = np.append(self.offsets, self.offsets[-1] + k)

    @classmethod
    def from_flat(cls, flat, rows_per_image):
        """Adopt a device-resident (N,32) uint8 tensor that holds N / rows_per_image keyframes
        of equal size back to back (no copy; getTrainDescriptors() is not available)."""
        db = cls()
        n = int(flat.shape[0])
        if n % int(rows_per_image):
            raise ValueError("the flat database is not a whole number of keyframes")
        db._buf, db._n = flat, n
        db.offsets = np.arange(0, n + 1, int(rows_per_image), dtype=np.int64)
        db._host = None
        return db

    def clear(self):
        self.__init__()

    def descriptors(self):
        if self._host is None:
            raise RuntimeError("this database adopted a device tensor; the host descriptors were never held")
        return list(self._host)

    def __len__(self):
        return int(self.offsets[-1])

    @property
    def flat(self):
        T = _ffi.require_cuda()
        if self._buf is None:
            return T.empty((0, 32), dtype=T.uint8, device="cuda")
        return self._buf[:self._n]

    def signs(self):
        """The expanded copy (128 B per row, e2m1) the tensor-core scan reads, or None when this
        database takes the integer-pipe route (small, or PBK_HAMMING_ENGINE=popc).  Kept
        resident and extended in place: an append only expands the rows from the last whole
        block of 512 on."""
        if self._n == 0 or ops.hamming_engine(self._n) != "tc":
            return None
        fmt = ops.hamming_tc_format()
        if self._signs_fmt != fmt:                   # expanded in the other encoding: start over
            self._signs, self._signs_n, self._signs_fmt = None, 0, fmt
        if self._signs_n == self._n:
            return self._signs
        T = _ffi.require_cuda()
        need = int(_ffi.lib().pbk_hamming_tc_signs_bytes(self._n))
        first = 0
        if self._signs is None or self._signs.numel() < need:
            cap = int(_ffi.lib().pbk_hamming_tc_signs_bytes(max(self._buf.shape[0], self._n)))
            grown = T.empty((cap,), dtype=T.int8, device="cuda")
            if self._signs is not None and self._signs_n:
                keep = (self._signs_n // 512) * 512 * ops.hamming_tc_row_bytes()
                grown[:keep].copy_(self._signs[:keep])
                first = (self._signs_n // 512) * 512
            self._signs = grown
        else:
            first = (self._signs_n // 512) * 512
        ops.hamming_tc_expand(self.flat, self._signs, first)
        self._signs_n = self._n
        return self._signs

    def split_global(self, g):
        """global rows -> (imgIdx, trainIdx)."""
        return _split_global(self.offsets, g)


class BFMatcher(object):
    """cv2.BFMatcher(normType=cv2.NORM_HAMMING, crossCheck=False) replacement.

    group: a torch.distributed process group (or True for the default group) shards the
    collection across the ranks by contiguous keyframe ranges: every rank makes the SAME
    add() / knnMatch() calls with the same arguments (the whole list of keyframes, the
    replicated queries) and receives the same, global result - imgIdx / trainIdx index the
    whole collection exactly as on one GPU.  The partition is (re)built by train() or by the
    first match after an add()."""

    def __init__(self, normType=NORM_HAMMING, crossCheck=False, group=None):
        if normType != NORM_HAMMING:
            raise NotImplementedError("only NORM_HAMMING (256-bit ORB descriptors) is implemented")
        self.crossCheck = bool(crossCheck)
        self.db = KeyframeDatabase()
        self.group = group
        self._pending = []          # sharded mode: host keyframes not yet partitioned
        self._sharded = None        # (ShardedKeyframeMatcher, global offsets)

    # -- collection management (cv2.DescriptorMatcher)
    def add(self, descriptors):
        if self.group is None:
            self.db.add(descriptors)
            return
        for d in descriptors:
            d = np.ascontiguousarray(d)
            if d.dtype != np.uint8 or d.ndim != 2 or d.shape[1] != 32:
                raise ValueError("descriptors must be (N, 32) uint8 (256-bit ORB)")
            self._pending.append(d)
        self._sharded = None

    def clear(self):
        self.db.clear()
        self._pending, self._sharded = [], None

    def empty(self):
        return (len(self.db) == 0) if self.group is None else (len(self._pending) == 0 and self._sharded is None)

    def getTrainDescriptors(self):
        return self.db.descriptors() if self.group is None else list(self._pending)

    def train(self):
        """cv2.DescriptorMatcher.train(): sharded mode uploads this rank's contiguous keyframe
        range of the collection; a no-op on one GPU."""
        if self.group is None or self._sharded is not None:
            return
        import torch.distributed as dist
        from ..distributed import ShardedKeyframeMatcher, shard_range
        T = _ffi.require_cuda()
        group = None if self.group is True else self.group
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        offsets = np.concatenate([[0], np.cumsum([len(d) for d in self._pending])]).astype(np.int64)
        lo, hi = shard_range(len(self._pending), rank, world)
        mine = self._pending[lo:hi]
        rows = np.concatenate(mine) if mine else np.zeros((0, 32), np.uint8)
        shard = T.from_numpy(np.ascontiguousarray(rows)).cuda()
        self._sharded = (ShardedKeyframeMatcher(shard, int(offsets[lo]), group=group), offsets)

    @classmethod
    def from_resident(cls, flat, rows_per_image, shard_base=0, n_images_total=None, group=None, crossCheck=False):
        """A matcher over descriptors ALREADY resident in HBM: `flat` (N,32) uint8 CUDA tensor of
        equal-sized keyframes.  With `group`, `flat` is this rank's contiguous range starting at
        global row `shard_base` of a collection of n_images_total keyframes."""
        m = cls(NORM_HAMMING, crossCheck, group)
        if group is None:
            m.db = KeyframeDatabase.from_flat(flat, rows_per_image)
            return m
        from ..distributed import ShardedKeyframeMatcher
        offsets = np.arange(0, (int(n_images_total) + 1) * int(rows_per_image), int(rows_per_image), dtype=np.int64)
        m._sharded = (ShardedKeyframeMatcher(flat, int(shard_base), group=None if group is True else group), offsets)
        return m

    # -- array API ---------------------------------------------------------
    def knn_arrays(self, queryDescriptors, trainDescriptors=None, ratio=None, mutual=False):
        """Top-2 as arrays: dict(idx (nq,2) global rows, imgIdx, trainIdx,
        dist (nq,2) int32, ratio_ok, mutual_ok, valid)."""
        r = 1.0 if ratio is None else ratio
        if trainDescriptors is None and self.group is not None:
            self.train()
            sm, offsets = self._sharded
            T = _ffi.require_cuda()
            qd = ops._desc_in(queryDescriptors)
            idx, dist, r_ok, m_ok, valid = sm.match_host(qd, r, mutual)
        else:
            if trainDescriptors is not None:
                db = KeyframeDatabase()
                db.add([trainDescriptors])
            else:
                db = self.db
            offsets = db.offsets
            res = ops.knn_ratio_mutual(queryDescriptors, db.flat, ratio=r, mutual=mutual, signs=db.signs())
            idx, dist, r_ok, m_ok, valid = res.to_host() if hasattr(res, "to_host") else (_np(x) for x in res)
        idx, dist = np.asarray(idx), np.asarray(dist)
        img, tr = _split_global(offsets, np.maximum(idx, 0))
        img[idx < 0] = -1
        tr[idx < 0] = -1
        return dict(idx=idx, imgIdx=img, trainIdx=tr, dist=dist, ratio_ok=r_ok, mutual_ok=m_ok, valid=valid)

    # -- OpenCV API --------------------------------------------------------
    def knnMatch(self, queryDescriptors, trainDescriptors=None, k=2, mask=None, compactResult=False, masks=None):
        """cv2's two overloads in one: knnMatch(query, train, k[, mask[, compactResult]]) against one
        train matrix and knnMatch(query, k[, masks[, compactResult]]) against the added collection
        (an integer in second position is k, as in OpenCV).  k = 1 / 2 without a mask is the hot
        path; any other k (<= ops.KNNK_MAX_K) or a mask goes through pbk_hamming_knnk (sharded
        collections: per-shard k best + one all-gather + merge), same order
        (ascending distance, ties to the lowest (imgIdx, trainIdx)); masked-out pairs never match
        and a query may come back with fewer than k rows - dropped when compactResult is set.
        mask: (nq, nt) uint8 / bool, non-zero = allowed; for the collection either `masks`, a list
        with one (nq, n_i) mask (or None = all allowed) per added image, or one (nq, total) matrix."""
        if isinstance(trainDescriptors, (int, np.integer)) and not isinstance(trainDescriptors, bool):
            trainDescriptors, k = None, int(trainDescriptors)        # the collection overload: (query, k, ...)
        if masks is not None:
            if mask is not None:
                raise ValueError("give mask or masks, not both")
            mask = masks
        k = int(k)
        if k < 1:
            raise ValueError("k must be >= 1 (OpenCV asserts knn > 0)")
        if self.crossCheck and k != 1:
            raise ValueError("crossCheck=True needs k == 1 (OpenCV asserts the same)")
        if mask is None and k <= 2:
            r = self.knn_arrays(queryDescriptors, trainDescriptors, mutual=self.crossCheck)
            counts = (r["idx"][:, :k] >= 0).sum(1)              # fewer than k rows when the train set is smaller
            if self.crossCheck:
                counts = np.where(r["mutual_ok"], counts, 0)
        else:
            r, counts = self._knn_general(queryDescriptors, trainDescriptors, k, mask)
        rows = np.arange(len(counts))
        if compactResult:
            rows = rows[counts > 0]
            counts = counts[counts > 0]
        return KnnMatches(rows, r["trainIdx"], r["imgIdx"], r["dist"], counts, arrays=r)

    def _knn_general(self, queryDescriptors, trainDescriptors, k, mask):
        """any k, optional mask: one launch of pbk_hamming_knnk over the flat collection."""
        if self.crossCheck:
            raise NotImplementedError("crossCheck together with a mask is not implemented")
        sharded = trainDescriptors is None and self.group is not None
        if sharded:
            self.train()
            offsets = self._sharded[1]
        elif trainDescriptors is not None:
            db = KeyframeDatabase()
            db.add([trainDescriptors])
            offsets = db.offsets
        else:
            db = self.db
            offsets = db.offsets
        nq, total = int(np.shape(queryDescriptors)[0]), int(offsets[-1])
        if isinstance(mask, (list, tuple)):
            if len(mask) != len(offsets) - 1:
                raise ValueError("masks must hold one entry per added image (%d), got %d" % (len(offsets) - 1, len(mask)))
            full = np.ones((nq, total), np.uint8)
            for i, m in enumerate(mask):
                if m is None:
                    continue
                m = np.asarray(_np(m))
                lo, hi = int(offsets[i]), int(offsets[i + 1])
                if m.shape != (nq, hi - lo):
                    raise ValueError("masks[%d] must be (%d, %d), got %s" % (i, nq, hi - lo, m.shape))
                full[:, lo:hi] = m != 0
            mask = full
        if sharded:       # every rank: its shard's k best, one all-gather, the k smallest by (distance, row)
            res = self._sharded[0].knnk(ops._desc_in(queryDescriptors), k, mask)
        else:
            res = ops.hamming_knnk(queryDescriptors, db.flat, k, mask)
        idx, dist, counts = (np.asarray(_np(x)) for x in res)
        img, tr = _split_global(offsets, np.maximum(idx, 0))
        img[idx < 0] = -1
        tr[idx < 0] = -1
        return dict(idx=idx, imgIdx=img, trainIdx=tr, dist=dist), counts

    def match(self, queryDescriptors, trainDescriptors=None, mask=None):
        return [row[0] for row in self.knnMatch(queryDescriptors, trainDescriptors, k=1, mask=mask)
                if row]


_UNIFORM = {}     # id(offsets) -> (offsets, rows per image or 0): equal-sized keyframes split by one division


def _split_global(offsets, g):
    """global rows -> (imgIdx, trainIdx) for a collection with row offsets `offsets`."""
    g = np.asarray(g, np.int64)
    hit = _UNIFORM.get(id(offsets))
    if hit is None or hit[0] is not offsets:
        step = int(offsets[1] - offsets[0]) if len(offsets) > 1 else 0
        if step <= 0 or int(offsets[0]) != 0 or not np.array_equal(offsets, np.arange(len(offsets), dtype=np.int64) * step):
            step = 0
        if len(_UNIFORM) > 64:
            _UNIFORM.clear()
        hit = _UNIFORM[id(offsets)] = (offsets, step)
    if hit[1]:                                   # (a binary search over 10 001 offsets costs 120 us per 4000 rows)
        img = np.minimum(g // hit[1], len(offsets) - 2)
        return img.astype(np.int32), (g - img * hit[1]).astype(np.int32)
    img = np.searchsorted(offsets, g, side="right") - 1
    img = np.clip(img, 0, max(len(offsets) - 2, 0))
    return img.astype(np.int32), (g - offsets[img]).astype(np.int32)


def _np(x):
    if _ffi.is_tensor(x):
        return _ffi.to_host(x) if x.is_cuda else x.numpy()
    return x


def knn_ratio_mutual(q, t, ratio=0.75):
    """Array pipeline of the north star: forward k=2, Lowe ratio, mutual check.
    Returns idx (nq,2) int64, dist (nq,2) int32, ratio_ok, mutual_ok, valid (bool)."""
    return ops.knn_ratio_mutual(q, t, ratio=ratio, mutual=True)
