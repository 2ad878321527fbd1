"""Host-array callers: chunked H2D -> kernel -> D2H over three streams.

The reference's callers hold numpy arrays (rigid_transform.py:139-141,
camera_utils.py:687-734, :964-990), so a drop-in call pays a host->device copy
of the inputs and a device->host copy of the results.  Done back to back those
two copies serialise on PCIe (C2: 120 MB in + 80 MB out = 3.7 ms for a 45 us
kernel).  PCIe is full duplex: rows are independent on this path, so the array
is cut into row chunks and chunk i+1's upload, chunk i's kernel and chunk i-1's
download run concurrently on an upload stream, the caller's stream and a
download stream.  The end-to-end time drops from H2D + D2H to about
max(H2D, D2H).

torch is only the carrier here (streams, events, pinned memory); the kernels are
launched by the `launch(lo, hi)` callback through the C-ABI.
"""
import concurrent.futures
import os

import numpy as np

from . import _ffi

MIN_BYTES = 8 << 20          # below this a single chunk (latency, not bandwidth, dominates)
CHUNK_BYTES = 12 << 20       # per-chunk bytes moved over PCIe (in + out)
MAX_CHUNKS = 32
COMPACT_CHUNK_BYTES = 48 << 20   # masked project: the host waits for a count per chunk, so fewer, larger chunks

STAGE_THREADS = max(1, min(8, (os.cpu_count() or 2) // 2))

_streams = {}
_pool = None


def _stage_rows(dst, src, lo, hi):
    """Pageable rows -> pinned staging rows with a few threads (numpy releases the
    GIL while copying); a single-threaded copy or the driver's own pageable path
    tops out near 10-15 GB/s, well under PCIe."""
    global _pool
    if _pool is None:
        _pool = concurrent.futures.ThreadPoolExecutor(max_workers=STAGE_THREADS, thread_name_prefix="pbk-stage")
    d, s = dst.numpy(), src.numpy()
    step = max(1, -(-(hi - lo) // STAGE_THREADS))
    futs = [_pool.submit(np.copyto, d[a:min(a + step, hi)], s[a:min(a + step, hi)]) for a in range(lo, hi, step)]
    for f in futs:
        f.result()


def _side_streams(device):
    T = _ffi.torch()
    key = (device.type, device.index)
    if key not in _streams:
        _streams[key] = (T.cuda.Stream(device=device), T.cuda.Stream(device=device))
    return _streams[key]


def pointer_row_align(plane_row_bytes):
    """Smallest row count r such that r * b is a multiple of 16 for every per-row byte stride b:
    the kernels require 16-byte aligned pointers, and a chunk starts at row lo of every plane."""
    import math
    r = 1
    for b in plane_row_bytes:
        b = int(b)
        if b > 0:
            need = 16 // math.gcd(16, b)
            r = r * need // math.gcd(r, need)
    return r


def plan(n_rows, bytes_per_row, row_align=1, chunk_bytes=None, plane_row_bytes=()):
    """Chunk boundaries [(lo, hi), ...] for n_rows rows; one chunk when small.
    plane_row_bytes: the per-row byte stride of EVERY input and output plane the chunks are
    sliced from (a KITTI uint8 frame is 466616 B: chunks must then start on even frames)."""
    import math
    total = n_rows * bytes_per_row
    if n_rows <= 0:
        return []
    if total < MIN_BYTES:
        return [(0, n_rows)]
    pa = pointer_row_align(plane_row_bytes)
    row_align = row_align * pa // math.gcd(row_align, pa)
    k = int(min(MAX_CHUNKS, max(2, total // (chunk_bytes or CHUNK_BYTES))))
    rows = -(-n_rows // k)
    rows = -(-rows // row_align) * row_align
    return [(lo, min(lo + rows, n_rows)) for lo in range(0, n_rows, rows)]


def host_out(shape, torch_dtype):
    """Pinned host tensor for results (torch caches pinned blocks, so repeated
    calls do not pay cudaHostAlloc again)."""
    return _ffi.torch().empty(tuple(shape), dtype=torch_dtype, pin_memory=True)


def stream_rows(chunks, host_in, dev_in, launch, dev_out, host_res):
    """Runs the pipeline.

    chunks    [(lo, hi)] from plan()
    host_in   list of host torch tensors (views of the caller's numpy arrays), first dim = rows
    dev_in    list of device tensors of the same shapes
    launch    callable(lo, hi): enqueue the kernel(s) for those rows on the CURRENT stream
    dev_out   list of device tensors, first dim = rows
    host_res  list of pinned host tensors of the same shapes (filled on return)
    """
    T = _ffi.torch()
    cur = T.cuda.current_stream()
    s_in, s_out = _side_streams(dev_in[0].device if dev_in else dev_out[0].device)
    s_in.wait_stream(cur)            # the device buffers may be recycled blocks with work pending on `cur`
    s_out.wait_stream(cur)
    # pageable inputs go through pinned staging so the upload is a real async DMA
    staged = [h if h.is_pinned() else host_out(h.shape, h.dtype) for h in host_in]
    for lo, hi in chunks:
        for h, p in zip(host_in, staged):
            if p is not h:
                _stage_rows(p, h, lo, hi)
        with T.cuda.stream(s_in):
            for h, d in zip(staged, dev_in):
                d[lo:hi].copy_(h[lo:hi], non_blocking=True)
            ev_in = s_in.record_event()
        cur.wait_event(ev_in)
        launch(lo, hi)
        ev_k = cur.record_event()
        s_out.wait_event(ev_k)
        with T.cuda.stream(s_out):
            for d, h in zip(dev_out, host_res):
                h[lo:hi].copy_(d[lo:hi], non_blocking=True)
    s_out.synchronize()              # results are in host memory; every stream is done with the buffers
    return [h.numpy() for h in host_res]


def as_host_tensor(a, dtype=None):
    """numpy array -> host torch tensor sharing its memory (pinned stays pinned)."""
    a = np.ascontiguousarray(a if dtype is None else np.asarray(a, dtype))
    return _ffi.torch().from_numpy(a)
