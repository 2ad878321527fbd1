"""Does the 20 M-row scan run slower when a host<->device copy sits next to it?  (Follow-up of
scripts/c3_e2e_timeline.py, which saw the scan phase of the end-to-end step at 2.88 ms against
2.58 ms in the device-resident step.)   python scripts/c3_scan_after_copy.py [keyframes]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pybot_b200 import _ffi, ops, synthetic as syn  # noqa: E402
from pybot_b200.distributed import ShardedKeyframeMatcher  # noqa: E402
import bench  # noqa: E402


def smi():
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        return "sm %d MHz, mem %d MHz, %.0f W, %d C" % (
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM),
            pynvml.nvmlDeviceGetPowerUsage(h) / 1e3, pynvml.nvmlDeviceGetTemperature(h, 0))
    except Exception as e:                                    # noqa: BLE001
        return repr(e)[:80]


def main():
    nkf = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    nq, per_kf = 2000, 2000
    q_np, db_np = syn.c3_database(nq=nq, n_keyframes=nkf, per_kf=per_kf)
    shard = torch.from_numpy(db_np).cuda()
    q_pin = _ffi.pinned_empty((nq, 32), np.uint8)
    q_pin[:] = q_np
    q_pin_t = torch.from_numpy(q_pin)
    q_dev = torch.from_numpy(q_np).cuda()
    q_buf = torch.from_numpy(q_np).cuda()
    host_out = torch.empty((27 * nq,), dtype=torch.uint8, pin_memory=True)
    m = ShardedKeyframeMatcher(shard, 0)
    flush = bench.make_flush(torch)
    E = lambda: torch.cuda.Event(enable_timing=True)   # noqa: E731

    def step(q, upload, download, sync_inside, fresh=False):
        evs = [E(), E()]
        if upload:
            if fresh:
                q = q_pin_t.to("cuda", non_blocking=True)
            else:
                q_buf.copy_(q_pin_t, non_blocking=True)
        evs[0].record()
        keys = m.local_keys(q)
        evs[1].record()
        rows = ops.hamming_gather_best(shard, keys, m.shard_base)
        rev = ops.hamming_knn2_keys(rows, q, 0)
        r = ops.hamming_ratio_mutual(keys, rev, 0.75)
        if download:
            host_out.copy_(r.packed, non_blocking=True)
        if sync_inside:
            torch.cuda.current_stream().synchronize()
        return evs

    def run(name, reps=16, use_flush=True, **kw):
        for _ in range(3):
            if use_flush:
                flush()
            step(**kw)
        whole, scan = [], []
        for _ in range(reps):
            if use_flush:
                flush()
            a, b = E(), E()
            a.record()
            evs = step(**kw)
            b.record()
            torch.cuda.synchronize()
            whole.append(a.elapsed_time(b))
            scan.append(evs[0].elapsed_time(evs[1]))
        print("%-64s step %7.3f ms (min %7.3f max %7.3f)  scan %7.3f ms   [%s]" % (
            name, np.median(whole), min(whole), max(whole), np.median(scan), smi()), flush=True)

    run("A  resident q_dev", q=q_dev, upload=False, download=False, sync_inside=False)
    run("B  upload into q_buf, scan q_buf", q=q_buf, upload=True, download=False, sync_inside=False)
    run("C  resident q_dev + packed download", q=q_dev, upload=False, download=True, sync_inside=False)
    run("D  resident q_dev again", q=q_dev, upload=False, download=False, sync_inside=False)
    run("E  q_buf without an upload", q=q_buf, upload=False, download=False, sync_inside=False)
    run("F  upload into q_buf, scan q_dev", q=q_dev, upload=True, download=False, sync_inside=False)
    run("G  upload to a fresh tensor + download + synchronise inside", q=None, upload=True, download=True, sync_inside=True, fresh=True)
    run("H  resident q_dev, synchronise inside", q=q_dev, upload=False, download=False, sync_inside=True)
    run("I  resident q_dev, no L2 flush", use_flush=False, q=q_dev, upload=False, download=False, sync_inside=False)
    run("J  upload + download + synchronise, no L2 flush", use_flush=False, q=q_buf, upload=True, download=True, sync_inside=True)
    run("K  resident q_dev once more", q=q_dev, upload=False, download=False, sync_inside=False)


if __name__ == "__main__":
    main()
