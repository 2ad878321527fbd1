"""Where the end-to-end C3 step spends its time beyond the device-resident step (one GPU):

    python scripts/c3_e2e_timeline.py [keyframes]

Per phase of ShardedKeyframeMatcher.match_host (upload, scan, gather + reverse + ratio, packed
download, synchronise) the CUDA-event time and the host time, then whole-call variants timed the
way bench.py times `e2e` (events around the call, L2 flushed before it)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pybot_b200 import _ffi, ops, synthetic as syn  # noqa: E402
from pybot_b200.distributed import ShardedKeyframeMatcher  # noqa: E402
import bench  # noqa: E402


def main():
    nkf = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    nq, per_kf = 2000, 2000
    q_np, db_np = syn.c3_database(nq=nq, n_keyframes=nkf, per_kf=per_kf)
    shard = torch.from_numpy(db_np).cuda()
    q_pin = _ffi.pinned_empty((nq, 32), np.uint8)
    q_pin[:] = q_np
    q_dev = torch.from_numpy(q_np).cuda()
    m = ShardedKeyframeMatcher(shard, 0)
    flush = bench.make_flush(torch)
    E = lambda: torch.cuda.Event(enable_timing=True)   # noqa: E731

    def whole(fn, reps=20):
        for _ in range(3):
            flush()
            fn()
        ms, host = [], []
        for _ in range(reps):
            flush()
            a, b = E(), E()
            t0 = time.perf_counter()
            a.record()
            fn()
            b.record()
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
            host.append((t1 - t0) * 1e3)
        return float(np.median(ms)), float(np.median(host))

    def resident():
        keys = m.local_keys(q_dev)
        rows = ops.hamming_gather_best(shard, keys, m.shard_base)
        rev = ops.hamming_knn2_keys(rows, q_dev, 0)
        return ops.hamming_ratio_mutual(keys, rev, 0.75)

    def e2e():
        return m.match_host(torch.from_numpy(q_pin).to("cuda", non_blocking=True))

    host_out = torch.empty((27 * nq,), dtype=torch.uint8, pin_memory=True)
    q_buf = torch.empty((nq, 32), dtype=torch.uint8, device="cuda")
    q_pin_t = torch.from_numpy(q_pin)

    def e2e_prealloc():
        q_buf.copy_(q_pin_t, non_blocking=True)
        r = m.match(q_buf, 0.75)
        host_out.copy_(r.packed, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host_out

    def e2e_no_sync_inside():
        q_buf.copy_(q_pin_t, non_blocking=True)
        r = m.match(q_buf, 0.75)
        host_out.copy_(r.packed, non_blocking=True)
        return host_out

    for name, fn in (("device-resident step", resident), ("match_host (bench e2e)", e2e),
                     ("preallocated pinned/device buffers", e2e_prealloc),
                     ("same, synchronise after the end event", e2e_no_sync_inside)):
        ms, host = whole(fn)
        print("%-42s events %8.3f ms   host %8.3f ms" % (name, ms, host))

    # phase timeline of match_host
    rows_ev, rows_host = [], []
    for it in range(13):
        flush()
        ev = [E() for _ in range(6)]
        t = [time.perf_counter()]
        ev[0].record()
        qd = torch.from_numpy(q_pin).to("cuda", non_blocking=True)
        ev[1].record(); t.append(time.perf_counter())
        keys = m.local_keys(qd)
        ev[2].record(); t.append(time.perf_counter())
        rws = ops.hamming_gather_best(shard, keys, m.shard_base)
        rev = ops.hamming_knn2_keys(rws, qd, 0)
        r = ops.hamming_ratio_mutual(keys, rev, 0.75)
        ev[3].record(); t.append(time.perf_counter())
        host = torch.empty(r.packed.shape, dtype=torch.uint8, pin_memory=True)
        host.copy_(r.packed, non_blocking=True)
        ev[4].record(); t.append(time.perf_counter())
        torch.cuda.current_stream().synchronize()
        t.append(time.perf_counter())
        ev[5].record()
        torch.cuda.synchronize()
        if it >= 3:
            rows_ev.append([ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(5)])
            rows_host.append([(t[i + 1] - t[i]) * 1e6 for i in range(5)])
    names = ["upload", "scan (expand+scan+merge)", "gather+reverse+ratio", "packed download", "synchronise"]
    med_e, med_h = np.median(np.array(rows_ev), 0), np.median(np.array(rows_host), 0)
    print("\nphase                         device us (events)   host us")
    for n, e, h in zip(names, med_e, med_h):
        print("%-30s %12.1f %14.1f" % (n, e, h))
    print("%-30s %12.1f %14.1f" % ("sum", med_e.sum(), med_h.sum()))


    # the device-resident step replayed from a CUDA graph: what the launch gaps between its six kernels cost
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                resident()
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = resident()
        ms, host = whole(g.replay)
        print("%-42s events %8.3f ms   host %8.3f ms" % ("device-resident step, CUDA graph replay", ms, host))
        ref = resident()
        g.replay()
        torch.cuda.synchronize()
        print("graph result equals eager result:", bool(torch.equal(out.packed, ref.packed)))
    except Exception as e:                                    # noqa: BLE001
        print("graph capture of the resident step failed:", repr(e)[:300])


if __name__ == "__main__":
    main()
