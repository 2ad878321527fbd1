"""Array-level wrappers over the C-ABI (one function per kernel family).

Inputs may be numpy arrays (copied host->device, results copied back through
pinned memory) or torch CUDA tensors (zero-copy; results stay on the device).
torch is only the carrier: every arithmetic result comes from
libpybot_b200.so.  Nothing here falls back to the CPU.
"""
import ctypes

import numpy as np

from . import _ffi, _pipeline

_NP = {"float32": np.float32, "float64": np.float64}


def _tdtype(t, npdt):
    return getattr(t, np.dtype(npdt).name)


def _points_in(X):
    """(n,3) float32/float64 device tensor from any float array-like."""
    t = _ffi.require_cuda()
    if isinstance(X, t.Tensor):
        if X.dtype not in (t.float32, t.float64):
            X = X.to(t.float64)
    else:
        X = np.asarray(X)
        if X.dtype not in (np.float32, np.float64):
            X = X.astype(np.float64)         # like np.hstack with float64 ones
    if X.ndim != 2 or X.shape[1] != 3:
        raise ValueError("expected an (N, 3) point array, got %s" % (tuple(X.shape),))
    return _ffi.to_device(X)


def _finish(tensor, on_device):
    return tensor if on_device else _ffi.to_host(tensor)


# ---------------------------------------------------------------- transform
def _host_points(X):
    """Host (n,3) float32/float64 array for the pipelined path, or None when X is a tensor."""
    T = _ffi.torch()
    if isinstance(X, T.Tensor):
        return None
    X = np.asarray(X)
    if X.dtype not in (np.float32, np.float64):
        X = X.astype(np.float64)             # like np.hstack with float64 ones
    if X.ndim != 2 or X.shape[1] != 3:
        raise ValueError("expected an (N, 3) point array, got %s" % (tuple(X.shape),))
    return _pipeline.as_host_tensor(X)


def transform_points(X, R, t, scale=1.0, out_dtype=np.float64, out=None):
    """Y = scale * R @ X + t for (N,3) points.  Reference:
    RigidTransform.__mul__ (rigid_transform.py:139-141) returns float64.
    `out`: optional preallocated (N,3) CUDA tensor of out_dtype.  Host arrays
    are streamed in row chunks (upload, kernel and download overlapped)."""
    T = _ffi.require_cuda()
    Rt = np.concatenate([np.asarray(R, np.float64).reshape(9), np.asarray(t, np.float64).reshape(3)])
    Rt, Rt_p = _ffi.host_f64(Rt, 12)
    L = _ffi.lib()

    def launch(Xd, Y):
        _ffi.check(L.pbk_transform_points(
            _ffi.ptr(Xd), _ffi.pbk_dtype(_NP[str(Xd.dtype).split(".")[-1]]), Xd.shape[0], Rt_p,
            float(scale), _ffi.ptr(Y), _ffi.pbk_dtype(out_dtype), _ffi.stream_ptr()))

    Xh = _host_points(X) if out is None else None
    if Xh is not None:
        n = Xh.shape[0]
        chunks = _pipeline.plan(n, Xh.element_size() * 3 + np.dtype(out_dtype).itemsize * 3, 1024)
        if len(chunks) > 1:
            Xd = T.empty((n, 3), dtype=Xh.dtype, device="cuda")
            Y = T.empty((n, 3), dtype=_tdtype(T, out_dtype), device="cuda")
            res = _pipeline.stream_rows(chunks, [Xh], [Xd], lambda lo, hi: launch(Xd[lo:hi], Y[lo:hi]), [Y],
                                        [_pipeline.host_out((n, 3), Y.dtype)])
            return res[0]
    Xd, on_dev = _points_in(X)
    Y = out if out is not None else T.empty((Xd.shape[0], 3), dtype=_tdtype(T, out_dtype), device=Xd.device)
    launch(Xd, Y)
    return _finish(Y, on_dev)


# ------------------------------------------------------------------ project
class Projector(object):
    """Prepared Camera.project launch: the parameter block is built once and
    output buffers are reused, so a call costs one ctypes call (a few us of
    host time) - per-frame callers and the benchmark use this directly."""

    def __init__(self, R, t, K, D=None, Rz=None, tz=None, shape=None, check_bounds=False,
                 check_depth=False, return_depth=False, return_valid=False, min_depth=0.1,
                 compact=True):
        P = _ffi.ProjectParams()
        R = np.asarray(R, np.float64).reshape(9)
        t = np.asarray(t, np.float64).reshape(3)
        P.R[:] = R.tolist()
        P.t[:] = t.tolist()
        rz = R[6:9] if Rz is None else np.asarray(Rz, np.float64).reshape(3)
        P.Rz[:] = rz.tolist()
        P.tz = float(t[2] if tz is None else tz)
        K = np.asarray(K, np.float64)
        P.fx, P.fy, P.cx, P.cy = float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])
        k = np.zeros(8)
        if D is not None:
            Dv = np.asarray(D, np.float64).ravel()
            if Dv.size > 8:
                raise NotImplementedError("thin-prism / tilt distortion (>8 coefficients) is not supported")
            if Dv.size not in (0, 4, 5, 8):
                raise ValueError("distortion must have 4, 5 or 8 coefficients")
            k[:Dv.size] = Dv
        P.k[:] = k.tolist()
        P.has_distortion = int(np.any(k != 0))
        P.min_depth = float(min_depth)
        if check_bounds and shape is None:
            raise ValueError("check_bounds cannot proceed. Camera.shape is not set")
        P.height, P.width = (int(shape[0]), int(shape[1])) if shape is not None else (0, 0)
        P.check_depth = int(bool(check_depth))
        P.check_bounds = int(bool(check_bounds))
        self.masked = bool(check_depth or check_bounds)
        self.compact = bool(compact and self.masked)
        P.compact = int(self.compact)
        self.params = P
        self.return_depth, self.return_valid = bool(return_depth), bool(return_valid)
        self._bufs = None

    def _buffers(self, Xd):
        T = _ffi.torch()
        n = Xd.shape[0]
        key = (n, Xd.dtype, Xd.device)
        if self._bufs is None or self._bufs[0] != key:
            uv = T.empty((n, 2), dtype=Xd.dtype, device=Xd.device)
            depth = T.empty((n,), dtype=T.float64, device=Xd.device) if self.return_depth else None
            valid = T.empty((n,), dtype=T.uint8, device=Xd.device) if self.return_valid else None
            count = T.empty((1,), dtype=T.int64, device=Xd.device)
            scratch = None
            if self.compact:
                nbytes = _ffi.lib().pbk_project_scratch_bytes(
                    n, _ffi.PBK_F32 if Xd.dtype == T.float32 else _ffi.PBK_F64)
                scratch = T.empty((max(nbytes, 16),), dtype=T.uint8, device=Xd.device)
            self._bufs = (key, uv, depth, valid, count, scratch)
        return self._bufs[1:]

    def run(self, Xd, fresh=False):
        """Xd: (n,3) float32/float64 CUDA tensor.  Returns device tensors
        (uv, depth, valid_u8, count); asynchronous.  Buffers are reused across
        calls unless fresh=True."""
        T = _ffi.torch()
        if fresh:
            self._bufs = None
        uv, depth, valid, count, scratch = self._buffers(Xd)
        in_dt = _ffi.PBK_F32 if Xd.dtype == T.float32 else _ffi.PBK_F64
        _ffi.check(_ffi.lib().pbk_project_points(
            _ffi.ptr(Xd), in_dt, Xd.shape[0], ctypes.byref(self.params), _ffi.ptr(uv), _ffi.ptr(depth),
            _ffi.ptr(valid), _ffi.ptr(count), _ffi.ptr(scratch), _ffi.stream_ptr()))
        return uv, depth, valid, count


def project_points(X, R, t, K, D=None, Rz=None, tz=None, shape=None, check_bounds=False,
                   check_depth=False, return_depth=False, return_valid=False, min_depth=0.1,
                   compact=True):
    """Fused cv2.projectPoints + depth + masks + x[valid] compaction.

    R, t      rotation/translation used for the pixel coordinates
    Rz, tz    third row / z-translation used for the depth (defaults to R[2], t[2])
    Returns (uv, depth or None, valid or None, count).  With compact=True uv
    and depth hold only the valid points, in input order (camera_utils.py:727-732).
    """
    pr = Projector(R, t, K, D, Rz, tz, shape, check_bounds, check_depth, return_depth, return_valid,
                   min_depth, compact)
    Xh = _host_points(X) if not (pr.masked or return_valid) else None
    if Xh is not None:
        # no masks -> rows are independent: stream the host array through in chunks
        n = Xh.shape[0]
        chunks = _pipeline.plan(n, Xh.element_size() * 5 + (8 if return_depth else 0), 1024)
        if len(chunks) > 1:
            T = _ffi.torch()
            Xd = T.empty((n, 3), dtype=Xh.dtype, device="cuda")
            uv = T.empty((n, 2), dtype=Xh.dtype, device="cuda")
            dev_out, host_res = [uv], [_pipeline.host_out((n, 2), uv.dtype)]
            dp = None
            if return_depth:
                dp = T.empty((n,), dtype=T.float64, device="cuda")
                dev_out.append(dp)
                host_res.append(_pipeline.host_out((n,), T.float64))
            in_dt = _ffi.PBK_F32 if Xd.dtype == T.float32 else _ffi.PBK_F64
            L = _ffi.lib()

            def launch(lo, hi):
                _ffi.check(L.pbk_project_points(
                    _ffi.ptr(Xd[lo:hi]), in_dt, hi - lo, ctypes.byref(pr.params), _ffi.ptr(uv[lo:hi]),
                    _ffi.ptr(dp[lo:hi]) if dp is not None else None, None, None, None, _ffi.stream_ptr()))
            res = _pipeline.stream_rows(chunks, [Xh], [Xd], launch, dev_out, host_res)
            return res[0], (res[1] if return_depth else None), None, n
    Xh = _host_points(X) if (pr.masked and pr.compact) else None
    if Xh is not None:
        res = _project_compact_streamed(pr, Xh, return_depth, return_valid)
        if res is not None:
            return res
    Xd, on_dev = _points_in(X)
    n = Xd.shape[0]
    uv, depth, valid, count = pr.run(Xd)
    m = n
    if pr.masked:
        m = int(count.item())            # one 8-byte D2H; also orders the stream
    if pr.compact:
        uv = uv[:m]
        if depth is not None:
            depth = depth[:m]
    if valid is not None:
        valid = valid.view(_ffi.torch().bool) if on_dev else valid
    outs = [_finish(uv, on_dev),
            _finish(depth, on_dev) if depth is not None else None,
            (_finish(valid, on_dev) if valid is not None else None)]
    if not on_dev and outs[2] is not None:
        outs[2] = outs[2].view(np.bool_)
    return outs[0], outs[1], outs[2], m


def _project_compact_streamed(pr, Xh, return_depth, return_valid):
    """Camera.project with masks from a HOST array: x[valid] is a global prefix, so
    every row chunk is compacted on its own (one launch, its own count) and the
    chunk's valid rows are downloaded to the running offset of the host result as
    soon as its 8-byte count has arrived.  All uploads and launches are enqueued
    first; the host then only waits for counts, so the copies of chunk i-1 (down)
    and chunk i+1 (up) overlap chunk i's kernel.  Returns None for small inputs."""
    T = _ffi.torch()
    n = Xh.shape[0]
    es = Xh.element_size()
    chunks = _pipeline.plan(n, es * 5 + (8 if return_depth else 0) + 1, 1024, _pipeline.COMPACT_CHUNK_BYTES)
    if len(chunks) < 2:
        return None
    L = _ffi.lib()
    cur = T.cuda.current_stream()
    s_in, s_out = _pipeline._side_streams(T.device("cuda", T.cuda.current_device()))
    s_in.wait_stream(cur)
    s_out.wait_stream(cur)
    Xd = T.empty((n, 3), dtype=Xh.dtype, device="cuda")
    uv = T.empty((n, 2), dtype=Xh.dtype, device="cuda")
    dp = T.empty((n,), dtype=T.float64, device="cuda") if return_depth else None
    vd = T.empty((n,), dtype=T.uint8, device="cuda") if return_valid else None
    counts = T.zeros((len(chunks),), dtype=T.int64, device="cuda")
    counts_h = _pipeline.host_out((len(chunks),), T.int64)
    rows = max(hi - lo for lo, hi in chunks)
    scratch = T.empty((max(L.pbk_project_scratch_bytes(
        rows, _ffi.PBK_F32 if Xh.dtype == T.float32 else _ffi.PBK_F64), 16),), dtype=T.uint8, device="cuda")
    uv_h = _pipeline.host_out((n, 2), uv.dtype)
    dp_h = _pipeline.host_out((n,), T.float64) if return_depth else None
    vd_h = _pipeline.host_out((n,), T.uint8) if return_valid else None
    Xs = Xh if Xh.is_pinned() else _pipeline.host_out(Xh.shape, Xh.dtype)
    in_dt = _ffi.PBK_F32 if Xd.dtype == T.float32 else _ffi.PBK_F64
    done = []
    for c, (lo, hi) in enumerate(chunks):
        if Xs is not Xh:
            _pipeline._stage_rows(Xs, Xh, lo, hi)
        with T.cuda.stream(s_in):
            Xd[lo:hi].copy_(Xs[lo:hi], non_blocking=True)
            ev_in = s_in.record_event()
        cur.wait_event(ev_in)
        _ffi.check(L.pbk_project_points(
            _ffi.ptr(Xd[lo:hi]), in_dt, hi - lo, ctypes.byref(pr.params), _ffi.ptr(uv[lo:hi]),
            _ffi.ptr(dp[lo:hi]) if dp is not None else None, _ffi.ptr(vd[lo:hi]) if vd is not None else None,
            _ffi.ptr(counts[c:c + 1]), _ffi.ptr(scratch), _ffi.stream_ptr()))
        counts_h[c:c + 1].copy_(counts[c:c + 1], non_blocking=True)
        done.append(cur.record_event())
    m = 0
    for c, (lo, hi) in enumerate(chunks):
        done[c].synchronize()
        mc = int(counts_h[c])
        s_out.wait_event(done[c])
        with T.cuda.stream(s_out):
            if mc:
                uv_h[m:m + mc].copy_(uv[lo:lo + mc], non_blocking=True)
                if dp is not None:
                    dp_h[m:m + mc].copy_(dp[lo:lo + mc], non_blocking=True)
            if vd is not None:
                vd_h[lo:hi].copy_(vd[lo:hi], non_blocking=True)
        m += mc
    s_out.synchronize()
    return (uv_h.numpy()[:m], dp_h.numpy()[:m] if dp_h is not None else None,
            vd_h.numpy().view(np.bool_) if vd_h is not None else None, m)


# ---------------------------------------------------------------- reproject
_DISP_DTYPES = (np.uint8, np.int16, np.int32, np.float32)


def reproject_disparity(disp, Q, sample=1, bgr=None):
    """cv2.reprojectImageTo3D(disp, Q) for (H,W) or (B,H,W) disparity; with
    `bgr` also returns the sub-sampled colour like reconstruct_with_texture.
    Returns xyz (..., Hs, Ws, 3) float32 [, bgr_out (..., Hs, Ws, 3) uint8]."""
    T = _ffi.require_cuda()
    is_t = isinstance(disp, T.Tensor)
    dt = np.dtype(str(disp.dtype).split(".")[-1]) if is_t else np.asarray(disp).dtype
    if dt == np.float64:
        raise NotImplementedError(
            "float64 disparity is not supported (cv2.reprojectImageTo3D rejects it as well)")
    if dt.type not in _DISP_DTYPES:
        raise TypeError("disparity dtype must be uint8, int16, int32 or float32; got %s" % dt)
    s = int(sample)
    if s < 1:
        raise ValueError("sample must be >= 1")
    Qh, Qp = _ffi.host_f32(Q, 16)
    L = _ffi.lib()

    def launch(d, c, xyz, c_out):
        _ffi.check(L.pbk_reproject_disp(
            _ffi.ptr(d), _ffi.pbk_dtype(dt), d.shape[0], int(d.shape[1]), int(d.shape[2]), Qp, s, _ffi.ptr(xyz),
            _ffi.ptr(c), _ffi.ptr(c_out), _ffi.stream_ptr()))

    if not is_t and np.ndim(disp) == 3 and (bgr is None or not isinstance(bgr, T.Tensor)):
        # a batch of host frames: frames are independent -> stream them through in groups
        dh = _pipeline.as_host_tensor(disp)
        B, H, W = (int(v) for v in dh.shape)
        Hs, Ws = -(-H // s), -(-W // s)
        ch = None
        if bgr is not None:
            ch = _pipeline.as_host_tensor(bgr)
            if ch.dtype != T.uint8 or tuple(ch.shape) != (B, H, W, 3):
                raise ValueError("colour image must be uint8 (H, W, 3) matching the disparity")
        per_frame = H * W * (dh.element_size() + (3 if ch is not None else 0)) + Hs * Ws * (12 + (3 if ch is not None else 0))
        # chunk starts must keep every plane's pointer 16-byte aligned (odd-sized uint8 frames)
        planes = [H * W * dh.element_size(), Hs * Ws * 12] + ([H * W * 3, Hs * Ws * 3] if ch is not None else [])
        chunks = _pipeline.plan(B, per_frame, plane_row_bytes=planes)
        if len(chunks) > 1:
            d = T.empty((B, H, W), dtype=dh.dtype, device="cuda")
            xyz = T.empty((B, Hs, Ws, 3), dtype=T.float32, device="cuda")
            host_in, dev_in, dev_out = [dh], [d], [xyz]
            host_res = [_pipeline.host_out(xyz.shape, T.float32)]
            c = c_out = None
            if ch is not None:
                c = T.empty((B, H, W, 3), dtype=T.uint8, device="cuda")
                c_out = T.empty((B, Hs, Ws, 3), dtype=T.uint8, device="cuda")
                host_in.append(ch)
                dev_in.append(c)
                dev_out.append(c_out)
                host_res.append(_pipeline.host_out(c_out.shape, T.uint8))
            res = _pipeline.stream_rows(
                chunks, host_in, dev_in,
                lambda lo, hi: launch(d[lo:hi], c[lo:hi] if c is not None else None, xyz[lo:hi],
                                      c_out[lo:hi] if c_out is not None else None),
                dev_out, host_res)
            return res[0] if ch is None else (res[0], res[1])
    d, on_dev = _ffi.to_device(disp)
    if d.ndim not in (2, 3):
        raise ValueError("disparity must be (H,W) or (B,H,W)")
    batched = d.ndim == 3
    B = d.shape[0] if batched else 1
    H, W = int(d.shape[-2]), int(d.shape[-1])
    Hs, Ws = -(-H // s), -(-W // s)
    xyz = T.empty((B, Hs, Ws, 3), dtype=T.float32, device=d.device)
    c = c_out = None
    if bgr is not None:
        c, _ = _ffi.to_device(bgr)
        if c.dtype != T.uint8 or tuple(c.shape[-3:]) != (H, W, 3):
            raise ValueError("colour image must be uint8 (H, W, 3) matching the disparity")
        c_out = T.empty((B, Hs, Ws, 3), dtype=T.uint8, device=d.device)
    launch(d.reshape(B, H, W), c, xyz, c_out)
    if not batched:
        xyz = xyz[0]
        c_out = c_out[0] if c_out is not None else None
    if c_out is None:
        return _finish(xyz, on_dev)
    return _finish(xyz, on_dev), _finish(c_out, on_dev)


def reproject_sparse(xyd, Q):
    """StereoCamera.reconstruct_sparse (camera_utils.py:971-980): float64."""
    T = _ffi.require_cuda()
    x, on_dev = _ffi.to_device(xyd, np.float64)
    if x.ndim != 2 or x.shape[1] != 3:
        raise ValueError("expected (N,3) [x, y, d] rows")
    Qh, Qp = _ffi.host_f32(Q, 16)
    out = T.empty_like(x)
    _ffi.check(_ffi.lib().pbk_reproject_sparse(_ffi.ptr(x), x.shape[0], Qp, _ffi.ptr(out),
                                               _ffi.stream_ptr()))
    return _finish(out, on_dev)


# ------------------------------------------------------------------ hamming
def _desc_in(a):
    T = _ffi.require_cuda()
    if isinstance(a, T.Tensor):
        if a.dtype != T.uint8:
            raise TypeError("descriptors must be uint8")
    else:
        a = np.asarray(a)
        if a.dtype != np.uint8:
            raise TypeError("descriptors must be uint8")
    if a.ndim != 2 or a.shape[1] != 32:
        raise ValueError("descriptors must be (N, 32) uint8 (256-bit ORB)")
    return _ffi.to_device(a)[0]


def hamming_knn2_keys(q_dev, t_dev, train_base=0):
    """Top-2 packed keys (nq,2) uint64-as-int64 device tensor for device inputs."""
    T = _ffi.torch()
    nq, nt = q_dev.shape[0], t_dev.shape[0]
    splits = ctypes.c_int(0)
    nbytes = ctypes.c_size_t(0)
    L = _ffi.lib()
    _ffi.check(L.pbk_hamming_plan(nq, nt, ctypes.byref(splits), ctypes.byref(nbytes)))
    scratch = T.empty((max(int(nbytes.value), 16),), dtype=T.uint8, device=q_dev.device)
    keys = T.empty((max(nq, 1), 2), dtype=T.int64, device=q_dev.device)[:nq]
    _ffi.check(L.pbk_hamming_knn2(_ffi.ptr(q_dev), nq, _ffi.ptr(t_dev) if nt else None, nt,
                                  int(train_base), _ffi.ptr(keys), _ffi.ptr(scratch),
                                  _ffi.stream_ptr()))
    return keys


# Below these sizes the integer-pipe kernel's latency wins (per-frame 2000 x 2000 matching); measured with
# 2000 queries by scripts/tc_size_probe.py: 16 k rows 50 us (popc) / 36 us (tc, resident copy) / 55 us (tc,
# copy expanded inside the call); 32 k rows 89 / 37 / 52 us.
TC_MIN_ROWS = 1 << 14          # the database keeps a resident expanded copy
TC_MIN_ROWS_EXPAND = 1 << 15   # the copy has to be made inside the call


def hamming_engine(nt, resident=True):
    """'tc' (tcgen05 scan over the expanded database) or 'popc' (integer-pipe kernel).
    PBK_HAMMING_ENGINE=popc|tc forces one; by default databases of >= 16384 rows (32768 when the
    expanded copy is not resident) take the tensor-core route.  Both return bit-identical keys."""
    import os
    e = os.environ.get("PBK_HAMMING_ENGINE", "").lower()
    if e in ("popc", "tc"):
        return e
    return "tc" if nt >= (TC_MIN_ROWS if resident else TC_MIN_ROWS_EXPAND) else "popc"


def hamming_tc_format():
    """Encoding of the expanded descriptors the tensor-core scan reads: 4 = e2m1 nibbles for
    tcgen05.mma kind::mxf4 (128 B per row, the default), 8 = signed bytes for kind::i8 (256 B per
    row).  PBK_TC_FORMAT=4|8 in the environment or hamming_tc_set_format() selects it."""
    return int(_ffi.lib().pbk_hamming_tc_format())


def hamming_tc_set_format(fmt):
    """Selects the encoding for buffers expanded FROM NOW ON; a resident expanded copy must be
    scanned with the format it was expanded in (KeyframeDatabase re-expands when it differs)."""
    _ffi.check(_ffi.lib().pbk_hamming_tc_set_format(int(fmt)))


def hamming_tc_row_bytes():
    """Bytes of one expanded row in the selected format (128 or 256)."""
    return int(_ffi.lib().pbk_hamming_tc_signs_bytes(512)) // 512


def hamming_tc_signs_rows(n):
    """Rows the expanded buffer of an n-row database holds (padded to 512)."""
    return -(-int(n) // 512) * 512


def hamming_tc_expand(t_dev, signs=None, first_row=0):
    """(n,32) uint8 descriptors -> the expanded copy the tensor-core scan reads (int8 tensor of
    hamming_tc_row_bytes() B per row, tile layout of csrc/hamming_tc4.cu / hamming_tc.cu).  Done once per database; with
    `signs` (a buffer of at least pbk_hamming_tc_signs_bytes(n) bytes) and first_row (a multiple
    of 512) only the rows from first_row on are (re)written - appending to a database."""
    T = _ffi.torch()
    n = int(t_dev.shape[0])
    L = _ffi.lib()
    need = max(int(L.pbk_hamming_tc_signs_bytes(n)), 16)
    if signs is None:
        signs = T.empty((need,), dtype=T.int8, device=t_dev.device)
    elif signs.numel() < need:
        raise ValueError("signs buffer too small")
    _ffi.check(L.pbk_hamming_tc_expand(_ffi.ptr(t_dev) if n else None, n, int(first_row), _ffi.ptr(signs),
                                       _ffi.stream_ptr()))
    return signs


def hamming_knn2_keys_tc(q_dev, t_signs, nt, train_base=0):
    """Top-2 packed keys (nq,2) through the tcgen05 scan; t_signs from hamming_tc_expand (same
    format) for exactly these nt rows."""
    T = _ffi.torch()
    nq = int(q_dev.shape[0])
    parts, nbytes = ctypes.c_int(0), ctypes.c_size_t(0)
    L = _ffi.lib()
    _ffi.check(L.pbk_hamming_tc_plan(nq, int(nt), ctypes.byref(parts), ctypes.byref(nbytes)))
    scratch = T.empty((max(int(nbytes.value), 16),), dtype=T.uint8, device=q_dev.device)
    keys = T.empty((max(nq, 1), 2), dtype=T.int64, device=q_dev.device)[:nq]
    _ffi.check(L.pbk_hamming_tc_knn2(_ffi.ptr(q_dev), nq, _ffi.ptr(t_signs) if nt else None, int(nt), int(train_base),
                                     _ffi.ptr(keys), _ffi.ptr(scratch), _ffi.stream_ptr()))
    return keys


def hamming_scan_keys(q_dev, t_dev, train_base=0, signs=None):
    """Top-2 keys of q against the database t: the tensor-core scan when `signs` (the expanded
    copy) is given or the engine rule picks it (expanding on the fly: 128 B per row written once
    is still cheaper than the integer-pipe scan of a large database), else the integer-pipe kernel."""
    nt = int(t_dev.shape[0])
    if nt > 0 and int(q_dev.shape[0]) > 0 and hamming_engine(nt, resident=signs is not None) == "tc":
        if signs is None:
            signs = hamming_tc_expand(t_dev)
        return hamming_knn2_keys_tc(q_dev, signs, nt, train_base)
    return hamming_knn2_keys(q_dev, t_dev, train_base)


def hamming_merge(partial_dev):
    """(parts, nq, 2) keys -> (nq, 2) global top-2."""
    T = _ffi.torch()
    parts, nq = int(partial_dev.shape[0]), int(partial_dev.shape[1])
    keys = T.empty((max(nq, 1), 2), dtype=T.int64, device=partial_dev.device)[:nq]
    _ffi.check(_ffi.lib().pbk_hamming_merge_top2(_ffi.ptr(partial_dev.contiguous()), parts, nq,
                                                 _ffi.ptr(keys), _ffi.stream_ptr()))
    return keys


def hamming_gather_best(t_dev, keys, train_base=0):
    T = _ffi.torch()
    nq = int(keys.shape[0])
    rows = T.empty((max(nq, 1), 32), dtype=T.uint8, device=keys.device)[:nq]
    nt = int(t_dev.shape[0])
    _ffi.check(_ffi.lib().pbk_hamming_gather_best(_ffi.ptr(t_dev) if nt else None, nt, int(train_base),
                                                  _ffi.ptr(keys), nq, _ffi.ptr(rows), _ffi.stream_ptr()))
    return rows


KNNK_MAX_K = 1024      # PBK_HAMMING_KNNK_MAX_K


def hamming_knnk(q, t, k, mask=None, train_base=0):
    """General-k / masked kNN in cv2.BFMatcher's order (ascending distance, ties to the lowest
    train row): the rest of knnMatch(query, train, k, mask)'s signature; k = 2 without a mask is
    the hot path above.  q (nq,32), t (nt,32) uint8; mask None or (nq, nt), non-zero = allowed.
    Returns (idx (nq,k) int64, dist (nq,k) int32, count (nq,) int32): numpy, or device tensors
    if q is a CUDA tensor; entries past count are -1."""
    T = _ffi.require_cuda()
    on_device = isinstance(q, T.Tensor) and q.is_cuda
    qd, td = _desc_in(q), _desc_in(t)
    nq, nt, k = int(qd.shape[0]), int(td.shape[0]), int(k)
    if k < 1:
        raise ValueError("k must be >= 1")
    md = None
    if mask is not None:
        if not isinstance(mask, T.Tensor):
            mask = np.ascontiguousarray(np.asarray(mask))
            if mask.dtype == np.bool_:
                mask = mask.view(np.uint8)
        elif mask.dtype == T.bool:
            mask = mask.view(T.uint8)
        if tuple(mask.shape) != (nq, nt):
            raise ValueError("mask must be (%d, %d), got %s" % (nq, nt, tuple(mask.shape)))
        if (mask.dtype != T.uint8) if isinstance(mask, T.Tensor) else (mask.dtype != np.uint8):
            raise TypeError("mask must be uint8 or bool (OpenCV: CV_8UC1)")
        md = _ffi.to_device(mask)[0].contiguous()
    m = max(nq, 1)
    idx = T.empty((m, k), dtype=T.int64, device=qd.device)[:nq]
    dist = T.empty((m, k), dtype=T.int32, device=qd.device)[:nq]
    count = T.empty((m,), dtype=T.int32, device=qd.device)[:nq]
    _ffi.check(_ffi.lib().pbk_hamming_knnk(
        _ffi.ptr(qd), nq, _ffi.ptr(td) if nt else None, nt, k, _ffi.ptr(md) if md is not None and nq * nt else None,
        int(train_base), _ffi.ptr(idx), _ffi.ptr(dist), _ffi.ptr(count), _ffi.stream_ptr()))
    if on_device:
        return idx, dist, count
    return _ffi.to_host(idx), _ffi.to_host(dist), _ffi.to_host(count)


class MatchResult(tuple):
    """(idx (nq,2) int64, dist (nq,2) int32, ratio_ok, mutual_ok, valid) as device tensors that are
    views of ONE buffer, so the whole result crosses PCIe in one copy (`to_host`) instead of five
    copies and five stream synchronisations."""

    def __new__(cls, items, packed, nq):
        self = super().__new__(cls, items)
        self.packed, self.nq = packed, nq
        return self

    @staticmethod
    def _views(base, m, nq, T):
        idx = base[:16 * m].view(T.int64).view(m, 2)[:nq]
        dist = base[16 * m:24 * m].view(T.int32).view(m, 2)[:nq]
        fb = base[24 * m:27 * m].view(3, m).view(T.bool)
        return idx, dist, fb[0, :nq], fb[1, :nq], fb[2, :nq]

    def to_host(self):
        """numpy (idx, dist, ratio_ok, mutual_ok, valid): one pinned copy, one synchronisation."""
        T = _ffi.torch()
        host = T.empty(self.packed.shape, dtype=T.uint8, pin_memory=True)
        host.copy_(self.packed, non_blocking=True)
        T.cuda.current_stream().synchronize()
        return tuple(v.numpy() for v in self._views(host, max(self.nq, 1), self.nq, T))


def hamming_ratio_mutual(keys, rev_keys, ratio):
    T = _ffi.torch()
    nq = int(keys.shape[0])
    m = max(nq, 1)
    base = T.empty((27 * m,), dtype=T.uint8, device=keys.device)   # the kernel writes 0/1 bytes as flags
    out = MatchResult._views(base, m, nq, T)
    idx, dist = out[0], out[1]
    flags = base[24 * m:27 * m].view(3, m)
    _ffi.check(_ffi.lib().pbk_hamming_ratio_mutual(
        _ffi.ptr(keys), _ffi.ptr(rev_keys), nq, float(ratio), _ffi.ptr(base), _ffi.ptr(base[16 * m:]),
        _ffi.ptr(flags[0]), _ffi.ptr(flags[1]), _ffi.ptr(flags[2]), _ffi.stream_ptr()))
    return MatchResult(out, base, nq)


def knn_ratio_mutual(q, t, ratio=0.75, train_base=0, mutual=True, signs=None):
    """Forward k=2 + Lowe ratio + mutual check on one GPU.
    Returns numpy (idx (nq,2) int64, dist (nq,2) int32, ratio_ok, mutual_ok, valid)
    or device tensors if q is a CUDA tensor.  signs: the database's expanded copy
    (hamming_tc_expand) if the caller keeps one resident."""
    T = _ffi.require_cuda()
    on_dev = isinstance(q, T.Tensor) and q.is_cuda
    qd, td = _desc_in(q), _desc_in(t)
    keys = hamming_scan_keys(qd, td, train_base, signs)
    rev = None
    if mutual:
        rows = hamming_gather_best(td, keys, train_base)
        rev = hamming_knn2_keys(rows, qd, 0)
    out = hamming_ratio_mutual(keys, rev, ratio)
    if on_dev:
        return out
    return out.to_host()


class PreparedMatcher(object):
    """Fixed-shape forward k=2 + ratio + mutual for per-frame use (C1: 2000 x 2000).

    At this size the work is ~20 us of GPU time but six kernel launches, five
    allocations and four ctypes calls: launch-latency-bound.  The buffers are
    allocated once and the launch sequence is captured in a CUDA graph, so a frame
    costs two small copies and one graph replay.
    """

    def __init__(self, nq, nt, ratio=0.75, mutual=True, use_graph=True):
        T = _ffi.require_cuda()
        self.T, self.nq, self.nt, self.ratio, self.mutual = T, int(nq), int(nt), float(ratio), bool(mutual)
        dev = "cuda"
        L = _ffi.lib()

        def scratch(a, b):
            splits, nbytes = ctypes.c_int(0), ctypes.c_size_t(0)
            _ffi.check(L.pbk_hamming_plan(a, b, ctypes.byref(splits), ctypes.byref(nbytes)))
            return T.empty((max(int(nbytes.value), 16),), dtype=T.uint8, device=dev)
        m = max(self.nq, 1)
        self.q = T.zeros((m, 32), dtype=T.uint8, device=dev)
        self.t = T.zeros((max(self.nt, 1), 32), dtype=T.uint8, device=dev)
        self.scratch_f, self.scratch_r = scratch(self.nq, self.nt), scratch(self.nq, self.nq)
        self.keys = T.empty((m, 2), dtype=T.int64, device=dev)
        self.rev = T.empty((m, 2), dtype=T.int64, device=dev)
        self.rows = T.empty((m, 32), dtype=T.uint8, device=dev)
        self.idx = T.empty((m, 2), dtype=T.int64, device=dev)
        self.dist = T.empty((m, 2), dtype=T.int32, device=dev)
        self.flags = T.empty((3, m), dtype=T.uint8, device=dev)
        self.graph = None
        if use_graph and self.nq > 0:
            side = T.cuda.Stream()
            side.wait_stream(T.cuda.current_stream())
            with T.cuda.stream(side):
                self._launch()                      # warm-up outside capture (attributes, lazy init)
            T.cuda.current_stream().wait_stream(side)
            T.cuda.synchronize()
            g = T.cuda.CUDAGraph()
            with T.cuda.graph(g):
                self._launch()
            self.graph = g

    def _launch(self):
        L, s = _ffi.lib(), _ffi.stream_ptr()
        nq, nt = self.nq, self.nt
        _ffi.check(L.pbk_hamming_knn2(_ffi.ptr(self.q), nq, _ffi.ptr(self.t) if nt else None, nt, 0,
                                      _ffi.ptr(self.keys), _ffi.ptr(self.scratch_f), s))
        rev = None
        if self.mutual:
            _ffi.check(L.pbk_hamming_gather_best(_ffi.ptr(self.t) if nt else None, nt, 0, _ffi.ptr(self.keys), nq,
                                                 _ffi.ptr(self.rows), s))
            _ffi.check(L.pbk_hamming_knn2(_ffi.ptr(self.rows), nq, _ffi.ptr(self.q), nq, 0, _ffi.ptr(self.rev),
                                          _ffi.ptr(self.scratch_r), s))
            rev = self.rev
        _ffi.check(L.pbk_hamming_ratio_mutual(
            _ffi.ptr(self.keys), _ffi.ptr(rev), nq, self.ratio, _ffi.ptr(self.idx), _ffi.ptr(self.dist),
            _ffi.ptr(self.flags[0]), _ffi.ptr(self.flags[1]), _ffi.ptr(self.flags[2]), s))

    def run(self, q, t):
        """q (nq,32), t (nt,32) uint8, numpy (pinned or not) or CUDA tensors.  Returns device
        tensors (idx, dist, ratio_ok, mutual_ok, valid) that are OVERWRITTEN by the next run."""
        T = self.T
        qs = q if isinstance(q, T.Tensor) else T.from_numpy(np.ascontiguousarray(q))
        ts = t if isinstance(t, T.Tensor) else T.from_numpy(np.ascontiguousarray(t))
        if tuple(qs.shape) != (self.nq, 32) or tuple(ts.shape) != (self.nt, 32):
            raise ValueError("PreparedMatcher was built for %d x %d descriptors" % (self.nq, self.nt))
        self.q[:self.nq].copy_(qs, non_blocking=True)
        self.t[:self.nt].copy_(ts, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        elif self.nq > 0:
            self._launch()
        n = self.nq
        fb = self.flags.view(T.bool)
        return self.idx[:n], self.dist[:n], fb[0, :n], fb[1, :n], fb[2, :n]


def tc_peak_probe(iters=4000):
    """Launches the int8 tensor-pipe probe; returns the multiply-adds of the launch."""
    T = _ffi.require_cuda()
    sink = T.zeros((1024,), dtype=T.int32, device="cuda")
    n = ctypes.c_double(0)
    _ffi.check(_ffi.lib().pbk_hamming_tc_peak_probe(int(iters), _ffi.ptr(sink), ctypes.byref(n), _ffi.stream_ptr()))
    return n.value


def popc_probe(mode=0, ctas=None, iters=2000):
    """Launches the popc-pipe probe; returns the POPC count of the launch."""
    T = _ffi.require_cuda()
    if ctas is None:
        ctas = T.cuda.get_device_properties(0).multi_processor_count * 8
    sink = T.zeros((8 + ctas * 256,), dtype=T.int32, device="cuda")
    n = ctypes.c_double(0)
    _ffi.check(_ffi.lib().pbk_popc_peak_probe(int(mode), int(ctas), int(iters), _ffi.ptr(sink),
                                              ctypes.byref(n), _ffi.stream_ptr()))
    return n.value


# ------------------------------------------------------- next rows (8f)
def depth_reconstruct(depth, xs, ys, skip=1):
    """DepthCamera.reconstruct (camera_utils.py:1032-1036).  depth: (H,W) or
    (B,H,W) uint16/float32/float64 (other dtypes are promoted to float64 like
    numpy does); xs (Ws,), ys (Hs,) float64 meshes.  Returns (..., Hs, Ws, 3) float64."""
    T = _ffi.require_cuda()
    is_t = isinstance(depth, T.Tensor)
    dt = np.dtype(str(depth.dtype).split(".")[-1]) if is_t else np.asarray(depth).dtype
    if dt.type not in (np.uint16, np.float32, np.float64):
        depth = depth.to(T.float64) if is_t else np.asarray(depth, np.float64)
        dt = np.dtype(np.float64)
    s = int(skip)
    L = _ffi.lib()

    def meshes(H, W):
        Hs, Ws = -(-H // s), -(-W // s)
        # meshes may be passed as float64 CUDA tensors (DepthCamera caches them per device)
        xs_d = xs if isinstance(xs, T.Tensor) else _ffi.to_device(np.ascontiguousarray(xs, np.float64))[0]
        ys_d = ys if isinstance(ys, T.Tensor) else _ffi.to_device(np.ascontiguousarray(ys, np.float64))[0]
        if xs_d.numel() != Ws or ys_d.numel() != Hs:
            raise AssertionError("depth shape %s does not match the camera mesh (%d, %d)" % ((H, W), Hs, Ws))
        return Hs, Ws, xs_d, ys_d

    def launch(d, out, xs_d, ys_d):
        _ffi.check(L.pbk_depth_reconstruct(_ffi.ptr(d), _ffi.pbk_dtype(dt), d.shape[0], int(d.shape[1]),
                                           int(d.shape[2]), s, _ffi.ptr(xs_d), _ffi.ptr(ys_d), _ffi.ptr(out),
                                           _ffi.stream_ptr()))

    if not is_t and np.ndim(depth) == 3:
        # a batch of host frames: stream frame groups (upload / kernel / download overlapped)
        dh = _pipeline.as_host_tensor(depth)
        B, H, W = (int(v) for v in dh.shape)
        Hs, Ws, xs_d, ys_d = meshes(H, W)
        chunks = _pipeline.plan(B, H * W * dh.element_size() + Hs * Ws * 24,
                                plane_row_bytes=[H * W * dh.element_size(), Hs * Ws * 24])
        if len(chunks) > 1:
            d = T.empty((B, H, W), dtype=dh.dtype, device="cuda")
            out = T.empty((B, Hs, Ws, 3), dtype=T.float64, device="cuda")
            return _pipeline.stream_rows(chunks, [dh], [d], lambda lo, hi: launch(d[lo:hi], out[lo:hi], xs_d, ys_d),
                                         [out], [_pipeline.host_out(out.shape, T.float64)])[0]
    d, on_dev = _ffi.to_device(depth)
    if d.ndim not in (2, 3):
        raise ValueError("depth must be (H,W) or (B,H,W)")
    batched = d.ndim == 3
    B = d.shape[0] if batched else 1
    H, W = int(d.shape[-2]), int(d.shape[-1])
    Hs, Ws, xs_d, ys_d = meshes(H, W)
    out = T.empty((B, Hs, Ws, 3), dtype=T.float64, device=d.device)
    launch(d.reshape(B, H, W), out, xs_d, ys_d)
    if not batched:
        out = out[0]
    return _finish(out, on_dev)


def depth_reconstruct_sparse(pts, depth, fx, cx, cy):
    """DepthCamera.reconstruct_sparse (camera_utils.py:1038-1041)."""
    T = _ffi.require_cuda()
    p, on_dev = _ffi.to_device(pts, np.float64)
    d, _ = _ffi.to_device(depth, np.float64)
    if p.ndim != 2 or p.shape[1] < 2 or d.shape[0] != p.shape[0]:
        raise ValueError("expected pts (N,2) and depth (N,)")
    p = p[:, :2].contiguous()
    out = T.empty((p.shape[0], 3), dtype=T.float64, device=p.device)
    _ffi.check(_ffi.lib().pbk_depth_reconstruct_sparse(_ffi.ptr(p), _ffi.ptr(d.contiguous()), p.shape[0], float(fx),
                                                       float(cx), float(cy), _ffi.ptr(out), _ffi.stream_ptr()))
    return _finish(out, on_dev)


def sceneflow(dX, params):
    """SceneFlow.process (optflow_utils.py:129-150) for (H,W,3) float32/float64
    scene points and a prepared ProjectParams block.  Returns (H,W,2) float32."""
    T = _ffi.require_cuda()
    is_t = isinstance(dX, T.Tensor)
    if not is_t:
        dX = np.asarray(dX)
        if dX.dtype not in (np.float32, np.float64):
            dX = dX.astype(np.float64)
    elif dX.dtype not in (T.float32, T.float64):
        dX = dX.to(T.float64)
    d, on_dev = _ffi.to_device(dX)
    if d.ndim != 3 or d.shape[2] != 3:
        raise ValueError("dX must be (H, W, 3)")
    H, W = int(d.shape[0]), int(d.shape[1])
    winner = T.empty((max(H * W, 1),), dtype=T.int32, device=d.device)
    flow = T.empty((H, W, 2), dtype=T.float32, device=d.device)
    dt = _ffi.PBK_F32 if d.dtype == T.float32 else _ffi.PBK_F64
    _ffi.check(_ffi.lib().pbk_sceneflow(_ffi.ptr(d), dt, H, W, ctypes.byref(params), _ffi.ptr(winner),
                                        _ffi.ptr(flow), _ffi.stream_ptr()))
    return _finish(flow, on_dev)


def fov_mask(X, R, t, half_fov_h, half_fov_v, zmin, zmax):
    """check_visibility (camera_utils.py:245-266) fused: bool mask (n,) of the points whose
    camera-frame direction lies inside the field of view and zmin <= z <= zmax."""
    T = _ffi.require_cuda()
    Xd, on_dev = _points_in(X)
    Rt = np.concatenate([np.asarray(R, np.float64).reshape(9), np.asarray(t, np.float64).reshape(3)])
    Rt, Rt_p = _ffi.host_f64(Rt, 12)
    mask = T.empty((max(Xd.shape[0], 1),), dtype=T.uint8, device=Xd.device)[:Xd.shape[0]]
    in_dt = _ffi.PBK_F32 if Xd.dtype == T.float32 else _ffi.PBK_F64
    _ffi.check(_ffi.lib().pbk_fov_mask(_ffi.ptr(Xd), in_dt, Xd.shape[0], Rt_p, float(half_fov_h), float(half_fov_v),
                                       float(zmin), float(zmax), _ffi.ptr(mask), _ffi.stream_ptr()))
    return mask.view(T.bool) if on_dev else _ffi.to_host(mask).view(np.bool_)


def transform_depth(X, Rz, tz, stride=1):
    """z column of T * X[::stride] (get_median_depth, camera_utils.py:269-276): (ceil(n/stride),) float64."""
    T = _ffi.require_cuda()
    Xd, on_dev = _points_in(X)
    n = Xd.shape[0]
    n_out = -(-n // int(stride))
    v = np.concatenate([np.asarray(Rz, np.float64).reshape(3), [float(tz)]])
    v, v_p = _ffi.host_f64(v, 4)
    z = T.empty((max(n_out, 1),), dtype=T.float64, device=Xd.device)[:n_out]
    in_dt = _ffi.PBK_F32 if Xd.dtype == T.float32 else _ffi.PBK_F64
    _ffi.check(_ffi.lib().pbk_transform_depth(_ffi.ptr(Xd), in_dt, n, int(stride), v_p, _ffi.ptr(z), _ffi.stream_ptr()))
    return _finish(z, on_dev)


def sampson_error(F, pts1, pts2):
    """camera_utils.py:52-68 for (N,2) point arrays; returns (N,) float64."""
    T = _ffi.require_cuda()
    a, on_dev = _ffi.to_device(pts1, np.float64)
    b, _ = _ffi.to_device(pts2, np.float64)
    if a.ndim != 2 or a.shape[1] != 2 or tuple(b.shape) != tuple(a.shape):
        raise AssertionError("pts1 and pts2 must both be (N, 2)")
    Fh, Fp = _ffi.host_f64(F, 9)
    err = T.empty((a.shape[0],), dtype=T.float64, device=a.device)
    _ffi.check(_ffi.lib().pbk_sampson_error(Fp, _ffi.ptr(a), _ffi.ptr(b), a.shape[0], _ffi.ptr(err),
                                            _ffi.stream_ptr()))
    return _finish(err, on_dev)


def epipolar_lines(F, pts):
    """camera_utils.py:352-357 (cv2.computeCorrespondEpilines(x, 1, F)): (N,2) float32 /
    float64 points -> (N,3) normalised lines of the same dtype."""
    T = _ffi.require_cuda()
    is32 = (pts.dtype == T.float32) if isinstance(pts, T.Tensor) else (np.asarray(pts).dtype == np.float32)
    dt = np.float32 if is32 else np.float64
    a, on_dev = _ffi.to_device(pts, dt)
    a = a.reshape(-1, 2).contiguous()
    Fh, Fp = _ffi.host_f64(F, 9)
    out = T.empty((a.shape[0], 3), dtype=a.dtype, device=a.device)
    _ffi.check(_ffi.lib().pbk_epipolar_lines(Fp, _ffi.ptr(a), _ffi.PBK_F32 if is32 else _ffi.PBK_F64,
                                             a.shape[0], _ffi.ptr(out), _ffi.stream_ptr()))
    return _finish(out, on_dev)


def triangulate_points(P1, P2, pts1, pts2):
    """camera_utils.py:108-116: (N,2) correspondences + two 3x4 projection matrices
    -> (N,3) points, float32 or float64 like the inputs (cv2.triangulatePoints
    keeps float32 only when both point sets are float32)."""
    T = _ffi.require_cuda()

    def _is32(x):
        return (x.dtype == T.float32) if isinstance(x, T.Tensor) else (np.asarray(x).dtype == np.float32)
    dt = np.float32 if (_is32(pts1) and _is32(pts2)) else np.float64
    a, on_dev = _ffi.to_device(pts1, dt)
    b, _ = _ffi.to_device(pts2, dt)
    a, b = a.reshape(-1, 2).contiguous(), b.reshape(-1, 2).contiguous()
    if tuple(a.shape) != tuple(b.shape):
        raise AssertionError("pts1 and pts2 must have the same shape")
    P1h, p1 = _ffi.host_f64(P1, 12)                  # host arrays stay alive through the call
    P2h, p2 = _ffi.host_f64(P2, 12)
    out = T.empty((a.shape[0], 3), dtype=a.dtype, device=a.device)
    _ffi.check(_ffi.lib().pbk_triangulate_points(p1, p2, _ffi.ptr(a), _ffi.ptr(b), _ffi.pbk_dtype(dt), a.shape[0],
                                                 _ffi.ptr(out), _ffi.stream_ptr()))
    del P1h, P2h
    return _finish(out, on_dev)


def _ray_params(K, D, P=None, rot=None, undistort=True, normalize=False, scale_by_z=False):
    prm = _ffi.RayParams()
    K = np.asarray(K, np.float64)
    prm.fx, prm.fy, prm.cx, prm.cy = float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])
    k = np.zeros(8)
    if D is not None:
        Dv = np.asarray(D, np.float64).ravel()
        if Dv.size > 8:
            raise NotImplementedError("thin-prism / tilt distortion (>8 coefficients) is not supported")
        if Dv.size not in (0, 4, 5, 8):
            raise ValueError("distortion must have 4, 5 or 8 coefficients")
        k[:Dv.size] = Dv
    prm.k[:] = k.tolist()
    prm.P[:] = np.asarray(K if P is None else P, np.float64)[:3, :3].reshape(9).tolist()
    prm.rot[:] = (np.zeros(9) if rot is None else np.asarray(rot, np.float64).reshape(9)).tolist()
    prm.undistort, prm.rotate = int(bool(undistort)), int(rot is not None)
    prm.normalize, prm.scale_by_z = int(bool(normalize)), int(bool(scale_by_z))
    return prm


def _rows_in(pts, min_cols):
    """(n, c>=min_cols) float32/float64 device rows; other dtypes go through float64."""
    T = _ffi.require_cuda()
    if isinstance(pts, T.Tensor):
        x = pts if pts.dtype in (T.float32, T.float64) else pts.to(T.float64)
    else:
        x = np.asarray(pts)
        if x.dtype not in (np.float32, np.float64):
            x = x.astype(np.float64)
    if x.ndim != 2 or x.shape[1] < min_cols:
        raise ValueError("expected an (N, >=%d) array, got %s" % (min_cols, tuple(x.shape)))
    return _ffi.to_device(x)


def undistort_points(pts, K, D, P=None):
    """cv2.undistortPoints(pts.astype(float32), K, D, P=K) (camera_utils.py:529-538)
    -> (N,2) float32, bit-exact."""
    T = _ffi.require_cuda()
    if not isinstance(pts, T.Tensor):
        pts = np.asarray(pts).reshape(-1, 2)
    x, on_dev = _rows_in(pts, 2)
    prm = _ray_params(K, D, P)
    out = T.empty((x.shape[0], 2), dtype=T.float32, device=x.device)
    dt = np.float32 if x.dtype == T.float32 else np.float64
    _ffi.check(_ffi.lib().pbk_undistort_points(_ffi.ptr(x), _ffi.pbk_dtype(dt), x.shape[1], x.shape[0],
                                               ctypes.byref(prm), _ffi.ptr(out), _ffi.stream_ptr()))
    return _finish(out, on_dev)


def camera_rays(pts, K, D, undistort=True, rot=None, normalize=False, scale_by_z=False):
    """CameraIntrinsic.ray / reconstruct (camera_utils.py:499-524) -> (N,3) float64.
    rot: the 3x3 coefficient rows of Quaternion.rotate or None."""
    T = _ffi.require_cuda()
    x, on_dev = _rows_in(pts, 3 if scale_by_z else 2)
    prm = _ray_params(K, D, None, rot, undistort, normalize, scale_by_z)
    out = T.empty((x.shape[0], 3), dtype=T.float64, device=x.device)
    dt = np.float32 if x.dtype == T.float32 else np.float64
    _ffi.check(_ffi.lib().pbk_camera_rays(_ffi.ptr(x), _ffi.pbk_dtype(dt), x.shape[1], x.shape[0],
                                          ctypes.byref(prm), _ffi.ptr(out), _ffi.stream_ptr()))
    return _finish(out, on_dev)


def wire_colors(im, flip_rb=False):
    """uint8 (..., 3) colours -> float32 (N,3) in [0,1] (get_color_arr, draw_helpers.py:35-53)."""
    T = _ffi.require_cuda()
    c, on_dev = _ffi.to_device(im)
    if c.dtype != T.uint8 or c.shape[-1] != 3:
        raise ValueError("colours must be uint8 (..., 3)")
    n = c.numel() // 3
    out = T.empty((n, 3), dtype=T.float32, device=c.device)
    _ffi.check(_ffi.lib().pbk_wire_colors(_ffi.ptr(c), n, int(bool(flip_rb)), _ffi.ptr(out), _ffi.stream_ptr()))
    return _finish(out, on_dev)


def hbm_probe(mode, src, dst, sink):
    """Launches the HBM probe (0 read / 1 write / 2 copy) over the given uint8 CUDA tensors."""
    ref = dst if mode == 1 else src
    _ffi.check(_ffi.lib().pbk_hbm_probe(int(mode), _ffi.ptr(src), _ffi.ptr(dst), int(ref.numel()),
                                        _ffi.ptr(sink), _ffi.stream_ptr()))


# ----------------------------------------------------------------------- BA
class BAProblemDevice(object):
    """Device-resident BA inputs + outputs (allocated once, reused per eval)."""

    def __init__(self, rvecs, tvecs, K, points, obs_cam, obs_pt, obs_uv, want=("r", "Jc", "Jp")):
        T = _ffi.require_cuda()
        self.T = T
        self.rvec, _ = _ffi.to_device(rvecs, np.float32)
        self.tvec, _ = _ffi.to_device(tvecs, np.float32)
        self.points, _ = _ffi.to_device(points, np.float32)
        self.n_cams, self.n_points = int(self.rvec.shape[0]), int(self.points.shape[0])
        K = np.asarray(K, np.float64)
        self.fx, self.fy, self.cx, self.cy = (float(K[0, 0]), float(K[1, 1]), float(K[0, 2]),
                                              float(K[1, 2]))
        if isinstance(obs_cam, T.Tensor):
            oc, op, ouv = obs_cam.cuda().int(), obs_pt.cuda().int(), obs_uv.cuda().float()
            if oc.numel() and (int(oc.min()) < 0 or int(oc.max()) >= self.n_cams
                               or int(op.min()) < 0 or int(op.max()) >= self.n_points):
                raise ValueError("observation index out of range")
            obs = T.empty((oc.shape[0], 4), dtype=T.int32, device="cuda")
            obs[:, 0], obs[:, 1] = oc, op
            obs[:, 2:] = ouv.contiguous().view(T.int32)
            self.obs = obs
            self.sorted_by_camera = bool(oc.numel() < 2 or bool((oc[1:] >= oc[:-1]).all()))
        else:
            oc = np.asarray(obs_cam, np.int32)
            op = np.asarray(obs_pt, np.int32)
            if oc.size and (oc.min() < 0 or oc.max() >= self.n_cams or op.min() < 0
                            or op.max() >= self.n_points):
                raise ValueError("observation index out of range")
            packed = np.empty((len(oc), 4), np.int32)
            packed[:, 0], packed[:, 1] = oc, op
            packed[:, 2:] = np.ascontiguousarray(obs_uv, np.float32).view(np.int32)
            self.obs, _ = _ffi.to_device(packed)
            # camera-sorted observations (the usual BA layout) take the one-kernel route
            self.sorted_by_camera = bool(oc.size < 2 or np.all(oc[1:] >= oc[:-1]))
        self.n_obs = int(self.obs.shape[0])
        O = max(self.n_obs, 1)
        dev = "cuda"
        self.cam_blocks = T.empty((max(self.n_cams, 1), _ffi.PBK_BA_CAM_DOUBLES), dtype=T.float64, device=dev)
        self.r = T.empty((O, 2), dtype=T.float32, device=dev)[:self.n_obs] if "r" in want else None
        self.Jc = T.empty((O, 2, 6), dtype=T.float32, device=dev)[:self.n_obs] if "Jc" in want else None
        self.Jp = T.empty((O, 2, 3), dtype=T.float32, device=dev)[:self.n_obs] if "Jp" in want else None
        slots = _ffi.lib().pbk_ba_cost_slots(self.n_obs)
        if slots < 0:
            _ffi.check(_ffi.PBK_ECUDA)
        self.partials = T.zeros((max(slots, 2),), dtype=T.float64, device=dev)
        self.cost = T.zeros((1,), dtype=T.float64, device=dev)
        self._graph = None

    def evaluate(self, events=None, exchange=None, one_kernel=False):
        """One residual + Jacobian evaluation (asynchronous on the current stream).
        Without `events` it is ONE C-ABI call (pbk_ba_evaluate): the per-camera precompute, then
        the streaming kernel launched with programmatic stream serialization so that its
        prologue (barriers, first record tiles, point gathers) overlaps the precompute.
        one_kernel=True (camera-sorted observations only): no precompute launch - every warp of
        the streaming kernel derives the camera blocks of its own observation range in its
        prologue.  Bit-identical; measured 1 us SLOWER at every size from 1 k to 4 M
        observations (scripts/ba_size_probe.py: the warp idles through obs -> rvec -> sincos
        while the two-launch form hides the precompute behind the launch of the second kernel),
        so it is not the default.
        `events`: optional (start, end) CUDA events recorded around the streaming
        kernel alone (two separate calls, the precompute stays outside), for roofline timing.
        `exchange`: a distributed.PeerCostExchange - the kernel then also performs the
        sum-all-reduce of the cost over NVLink peer memory and self.cost is GLOBAL."""
        L = _ffi.lib()
        s = _ffi.stream_ptr()
        out = (_ffi.ptr(self.r), _ffi.ptr(self.Jc), _ffi.ptr(self.Jp), _ffi.ptr(self.partials),
               _ffi.ptr(self.cost))
        K = (self.fx, self.fy, self.cx, self.cy)
        if events is None:
            # seq = 0: the call sequence of the fused exchange lives on the device, so the
            # same call can be captured in a CUDA graph and replayed
            ex = ((exchange.ptrs, exchange.rank, exchange.world, 0, exchange.status.ptr)
                  if exchange is not None else (None, 0, 1, 0, None))
            _ffi.check(L.pbk_ba_evaluate(
                _ffi.ptr(self.obs), self.n_obs, int(bool(one_kernel) and self.sorted_by_camera),
                _ffi.ptr(self.points), self.n_points,
                _ffi.ptr(self.rvec), _ffi.ptr(self.tvec), self.n_cams, _ffi.ptr(self.cam_blocks),
                *(K + out + ex + (s,))))
            return self
        _ffi.check(L.pbk_ba_camera_precompute(_ffi.ptr(self.rvec), _ffi.ptr(self.tvec), self.n_cams,
                                              _ffi.ptr(self.cam_blocks), s))
        events[0].record()
        head = (_ffi.ptr(self.obs), self.n_obs, _ffi.ptr(self.points), self.n_points,
                _ffi.ptr(self.cam_blocks), self.n_cams) + K + out
        if exchange is not None:
            _ffi.check(L.pbk_ba_residual_jacobian_dist(
                *(head + (exchange.ptrs, exchange.rank, exchange.world, 0, exchange.status.ptr, s))))
        else:
            _ffi.check(L.pbk_ba_residual_jacobian(*(head + (s,))))
        events[1].record()
        return self

    def evaluate_graphed(self, exchange=None):
        """evaluate() replayed from a CUDA graph captured on first use (fixed buffers; the
        inputs rvec / tvec / points are updated in place between calls).  One graph launch
        instead of a ctypes call + kernel launch per LM iteration."""
        T = self.T
        if self._graph is None or self._graph[1] is not exchange:
            side = T.cuda.Stream()
            side.wait_stream(T.cuda.current_stream())
            with T.cuda.stream(side):
                self.evaluate(exchange=exchange)           # warm-up outside capture (attributes, lazy init)
            T.cuda.current_stream().wait_stream(side)
            T.cuda.synchronize()
            g = T.cuda.CUDAGraph()
            with T.cuda.graph(g):
                self.evaluate(exchange=exchange)
            self._graph = (g, exchange)
            return self                                     # the warm-up call WAS this evaluation
        self._graph[0].replay()
        return self


def ba_residual_jacobian(rvecs, tvecs, K, points, obs_cam, obs_pt, obs_uv):
    """r (O,2), J_cam (O,2,6), J_pt (O,2,3) float32 and cost = sum ||r||^2 (float)."""
    p = BAProblemDevice(rvecs, tvecs, K, points, obs_cam, obs_pt, obs_uv).evaluate()
    cost = float(p.cost.item())
    return _ffi.to_host(p.r), _ffi.to_host(p.Jc), _ffi.to_host(p.Jp), cost
