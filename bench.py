#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c4|c5]
                    [--impl reference] [--no-also]

Default workload = C3, loop-closure matching: 2000 ORB query descriptors vs a
10k-keyframe database (20M descriptors, 640 MB), k=2 + Lowe ratio + mutual
check, database sharded by contiguous keyframe ranges over N GPUs (strong
scaling: the database size is fixed).  One JSON line on stdout (rank 0).

A "step" is one pass of the path over one batch: for C3 one query frame
matched against the whole database.  `value` is measured with the inputs
resident in HBM; `e2e` goes through the public API with pinned HOST buffers
(H2D of the step's inputs and D2H of its results inside the timed region).
With --no-also only the headline workload runs; otherwise the single-GPU
workloads C2 (transform+project 10M), C4 (BA 4M obs) and C5 (4K x 64 coloured
cloud) are measured too and reported under "also" (N=1 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pybot_b200 import synthetic as syn  # noqa: E402

L2_FLUSH_BYTES = 256 << 20


# ------------------------------------------------------------------ helpers
def ncu_traffic(kernel, scale=1.0):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` from the
    committed `ncu --set full` capture (profiles/ncu_traffic.json, written by
    scripts/ncu_summary.py), scaled to this launch's size where the capture ran a
    smaller problem; None if there is no capture."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)
        return float(_traffic_entry(d, kernel)["dram_bytes"]) * scale
    except Exception:
        return None


def _traffic_entry(d, kernel):
    """Exact kernel name, else the first captured kernel whose name starts with it
    (template arguments differ between kernel variants)."""
    ks = d["kernels"]
    if kernel in ks:
        return ks[kernel]
    return next(v for k, v in ks.items() if k.startswith(kernel))


def digest(*arrays):
    """sha256 over the raw bytes of the result arrays (numpy or CUDA tensors), in order."""
    import hashlib
    h = hashlib.sha256()
    for a in arrays:
        if not isinstance(a, np.ndarray):
            a = a.detach().cpu().numpy()
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode())
        h.update(a.view(np.uint8).reshape(-1).data)
    return h.hexdigest()


def source_sha(fname):
    """sha256 (first 16 hex) of a kernel source file: ncu captures record it, so a capture
    made from a different version of the kernel is reported as stale."""
    import hashlib
    try:
        with open(os.path.join(ROOT, "pybot_b200", "csrc", fname), "rb") as f:
            return hashlib.sha256(f.read()).hexdigest()[:16]
    except OSError:
        return None


def ncu_traffic_stale(source_file):
    """True when profiles/ncu_traffic.json was captured from another version of `source_file`
    (or does not say), None without a capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            d = json.load(f)
    except Exception:
        return None
    return d.get("source_sha", {}).get(source_file) != source_sha(source_file)


def ncu_traffic_note(kernel, scale=1.0):
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)
        e = _traffic_entry(d, kernel)
        return ("dram read+write bytes per launch from profiles/%s_ncu_full_summary.md (captured on: %s)%s"
                % (d["tag"], e.get("problem", ""), "" if scale == 1.0 else ", scaled x%g to this launch" % scale))
    except Exception:
        return "no ncu capture"


def measured_bf16():
    """Dense bf16 TFLOP/s of this pool's B200s (driver-measured burst figure), else the nominal."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        if "bf16_tflops" in d:
            return float(d["bf16_tflops"]), "MEASURED_PEAKS.json (bf16_tflops, burst, of measured)"
    return 2250.0, "nominal 2.25 PFLOP/s dense bf16 (no MEASURED_PEAKS.json)"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "MEASURED_PEAKS.json (of measured)"
    return 6650.0, "B200_PROFILING.md fallback (of fallback)"


class _SmiClockSampler(object):
    """nvidia-smi clocks/throttle reasons around the timed region (fallback of ClockSampler: its
    first sample arrives ~100 ms after start())."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "power_w_max": float(max(pw)), "samples": len(sm), "reasons": sorted(reasons)}


_NVML_POLL = r"""
import sys, time
import pynvml as n
n.nvmlInit()
h = n.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
out = sys.stdout
while True:
    t = time.time()
    try:
        line = "%.6f,%d,%d,%.1f,%d\n" % (t, n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM), mx,
                                         n.nvmlDeviceGetPowerUsage(h) / 1e3, get_reasons(h))
    except Exception:
        time.sleep(0.02)
        continue
    out.write(line)
    out.flush()
    time.sleep(0.004)
"""


class ClockSampler(object):
    """SM clock, board power and clock-event reasons DURING a timed region.

    One NVML poller per GPU index runs in its own process for the whole bench (a sample every
    ~5 ms, each stamped with time.time()); start() / stop() mark the window and stop() reports
    the samples that fall inside it - so a 40 ms loop is described by the clocks it ran at, not by
    what nvidia-smi printed 100 ms after it ended.  Falls back to nvidia-smi if NVML is missing."""
    REASONS = ((0x4, "sw_power_cap"), (0x8, "hw_slowdown"), (0x20, "sw_thermal_slowdown"),
               (0x40, "hw_thermal_slowdown"), (0x80, "hw_power_brake_slowdown"))
    _pollers = {}

    def __init__(self, index=0):
        self.index = index
        self.fallback = None
        self.t0 = None

    @classmethod
    def _poller(cls, index):
        p = cls._pollers.get(index)
        if p is None:
            p = {"lines": [], "ok": False}
            try:
                p["proc"] = subprocess.Popen([sys.executable, "-c", _NVML_POLL, str(index)], stdout=subprocess.PIPE,
                                             stderr=subprocess.DEVNULL, text=True)

                def pump():
                    for line in p["proc"].stdout:
                        p["lines"].append(line)
                threading.Thread(target=pump, daemon=True).start()
                t_end = time.time() + 5.0                      # first sample = NVML is up
                while time.time() < t_end and not p["lines"] and p["proc"].poll() is None:
                    time.sleep(0.01)
                p["ok"] = bool(p["lines"])
            except Exception:
                p["ok"] = False
            cls._pollers[index] = p
            import atexit
            atexit.register(cls._shutdown, index)
        return p

    @classmethod
    def _shutdown(cls, index):
        p = cls._pollers.get(index)
        if p and p.get("proc") is not None and p["proc"].poll() is None:
            p["proc"].terminate()

    def start(self):
        p = self._poller(self.index)
        if not p["ok"]:
            self.fallback = _SmiClockSampler(self.index).start()
            return self
        self.t0 = time.time()
        return self

    def stop(self):
        if self.fallback is not None:
            return self.fallback.stop()
        t1 = time.time()
        time.sleep(0.012)                                      # let the sample that covers t1 arrive
        rows = []
        for ln in list(self._pollers[self.index]["lines"]):
            f = ln.strip().split(",")
            if len(f) == 5:
                try:
                    rows.append((float(f[0]), float(f[1]), float(f[2]), float(f[3]), int(f[4])))
                except ValueError:
                    pass
        inside = [r for r in rows if self.t0 <= r[0] <= t1]
        window = "inside the timed region"
        if not inside:                                         # region shorter than one poll: the nearest samples
            inside = sorted(rows, key=lambda r: abs(r[0] - 0.5 * (self.t0 + t1)))[:2]
            window = "nearest to the timed region (shorter than one poll)"
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        bits = 0
        for r in inside:
            bits |= r[4]
        sm = [r[1] for r in inside]
        return {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)), "sm_max_mhz": float(max(r[2] for r in inside)),
                "power_w_max": float(max(r[3] for r in inside)), "samples": len(inside),
                "reasons": [name for bit, name in self.REASONS if bits & bit],
                "source": "NVML poller process, ~5 ms period, samples " + window,
                "window_ms": (t1 - self.t0) * 1e3}


class Dist(object):
    def __init__(self, gpus):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            import torch.distributed as dist
            torch.cuda.set_device(self.local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        else:
            torch.cuda.set_device(0)
            self.dist = None
        if gpus != self.world:
            raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)"
                             % (gpus, self.world))

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def timed_steps(D, step, steps, warmup, flush=None, inner=None):
    """W untimed + K timed steps.  Each step is bracketed by CUDA events on the
    launching stream (the L2 flush between steps is outside the events);
    barrier + synchronize on both sides; returns (sum of step ms, max over
    ranks), plus the summed ms of the `inner` event pairs the step records."""
    torch = D.torch
    for _ in range(warmup):
        if flush is not None:
            flush()
        step(None)
    D.barrier()
    evs = []
    for _ in range(steps):
        if flush is not None:
            flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        rec = []
        a.record()
        step(rec)
        b.record()
        evs.append((a, b, rec))
    D.barrier()
    total = sum(a.elapsed_time(b) for a, b, _ in evs)
    inner_ms = sum(x.elapsed_time(y) for _, _, rec in evs for x, y in rec)
    return D.max_over_ranks(total), D.max_over_ranks(inner_ms)


L2_NOTE = ("L2 flushed between timed steps: 256 MiB written, then 256 MiB of another buffer read so "
           "the flush's own dirty lines are written back before the timed kernel starts")


def make_flush(torch):
    """Flush = write a buffer larger than L2, then read a second one: after the
    write alone L2 is full of DIRTY lines whose write-back (126 MB, ~20 us of
    HBM time) would be charged to the next 40-100 us kernel."""
    from pybot_b200 import ops
    wbuf = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
    rbuf = torch.zeros(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
    sink = torch.zeros(4, dtype=torch.int32, device="cuda")

    def flush():
        wbuf.zero_()
        ops.hbm_probe(0, rbuf, None, sink)
    return flush


def copy_at_size(torch, nbytes_read, nbytes_written, reps=5):
    """HBM context for short kernels: best-of-N device time of a plain copy that
    moves the same number of bytes (read nbytes_read, write nbytes_written is
    approximated by a copy of (r+w)/2 bytes).  At 200-400 MB a copy reaches only
    ~80-87 % of the multi-GB figure in MEASURED_PEAKS.json (ramp + tail)."""
    from pybot_b200 import ops
    half = ((nbytes_read + nbytes_written) // 2 + 15) // 16 * 16
    a = torch.zeros(half, dtype=torch.uint8, device="cuda")
    b = torch.empty(half, dtype=torch.uint8, device="cuda")
    sink = torch.zeros(4, dtype=torch.int32, device="cuda")
    best = None
    for i in range(reps + 2):
        e0, e1 = ev_pair(torch)
        e0.record()
        b.copy_(a)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            t = e0.elapsed_time(e1)
            best = t if best is None else min(best, t)
    del a, b, sink
    return 2.0 * half / (best * 1e-3) / 1e9


def ev_pair(torch):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


# ------------------------------------------------------- popc pipe peak
def probe_peaks(D):
    """Measured POPC-pipe peak of this GPU, in this run: independent POPC chains, 8 CTAs of
    256 threads per SM, best of 3 launches of ~20 ms each (the pipe issues 16 POPC per clock
    per SM: 148 x 16 x 1.965 GHz = 4.65 Tpopc/s nominal)."""
    from pybot_b200 import ops
    torch = D.torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    best = 0.0
    for _ in range(3):
        a, b = ev_pair(torch)
        a.record()
        n = ops.popc_probe(mode=0, ctas=sms * 8, iters=4000)
        b.record()
        torch.cuda.synchronize()
        best = max(best, n / (a.elapsed_time(b) * 1e-3))
    tc_best = 0.0
    for _ in range(3):
        a, b = ev_pair(torch)
        a.record()
        n = ops.tc_peak_probe(iters=4000)
        b.record()
        torch.cuda.synchronize()
        tc_best = max(tc_best, n / (a.elapsed_time(b) * 1e-3))
    return {"popc_per_s": best, "sm_count": sms, "tc_mac_per_s": tc_best}


# ------------------------------------------------------------ workload C3
# POPC.32 the shipped kernel issues per descriptor pair (carry-save compression of the eight XOR
# words into four of weight 1, 1, 2, 4: profiles/r02_hamming_inner_loop.md counts them in the SASS)
POPC_ISSUED_PER_PAIR = 4


def run_c3(D, args):
    import torch
    from pybot_b200 import _ffi, ops
    from pybot_b200.distributed import ShardedKeyframeMatcher, shard_range
    from pybot_b200.vision.matcher import BFMatcher
    nq, per_kf, nkf = 2000, 2000, args.keyframes
    lo_kf, hi_kf = shard_range(nkf, D.rank, D.world)
    # every rank generates the same seeded database and keeps its keyframe range
    q_np, db_np = syn.c3_database(nq=nq, n_keyframes=nkf, per_kf=per_kf)
    shard = torch.from_numpy(db_np[lo_kf * per_kf:hi_kf * per_kf]).cuda()
    nt_total = nkf * per_kf
    q_pin = _ffi.pinned_empty((nq, 32), np.uint8)
    q_pin[:] = q_np
    q_dev = torch.from_numpy(q_np).cuda()
    m = ShardedKeyframeMatcher(shard, lo_kf * per_kf)
    flush = make_flush(torch)

    def step(rec):
        if D.world > 1:
            # sharded: both exchanges (top-2 keys, matched rows) are fused into the kernels
            # over NVLink peer memory; the events bracket the shard scan + its key exchange
            ev = ev_pair(torch) if rec is not None else None
            out = m.match(q_dev, 0.75, events=ev)
            if rec is not None:
                rec.append(ev)
            return out
        if rec is not None:
            a, b = ev_pair(torch)
            a.record()
            keys = m.local_keys(q_dev)
            b.record()
            rec.append((a, b))
        else:
            keys = m.local_keys(q_dev)
        rows = ops.hamming_gather_best(shard, keys, m.shard_base)
        rev = ops.hamming_knn2_keys(rows, q_dev, 0)
        return ops.hamming_ratio_mutual(keys, rev, 0.75)

    tc_fmt = ops.hamming_tc_format()
    tc_kernel = "hamming_tc4_kernel" if tc_fmt == 4 else "hamming_tc_kernel"
    tc_kind = "kind::mxf4 (e2m1 x e2m1 -> f32, unit block scales)" if tc_fmt == 4 else "kind::i8 (s8 x s8 -> s32)"
    engine = ("tcgen05 %s (%s)" % (tc_kind.split(" ")[0], tc_kernel)) if m.signs is not None else "integer pipe (hamming_knn2_kernel<2,4>)"
    sampler = ClockSampler(D.local).start() if D.rank == 0 else None
    total_ms, knn_ms = timed_steps(D, step, args.steps, args.warmup, flush)
    clocks = sampler.stop() if sampler else None
    result = [_ffi.to_host(x) for x in step(None)]          # idx, dist, ratio_ok, mutual_ok, valid
    m.check()
    # The end-to-end loops run HERE, directly after the device-resident loop and before the variants:
    # the scan is tensor-bound and follows the SM clock, which sags from ~1840 to ~1700 MHz over a
    # few seconds of sustained tcgen05 work (scripts/c3_scan_after_copy.py: the same resident step
    # 2.63 -> 2.99 ms while averaged board power climbs 290 -> 735 W).  Timed after ~0.5 s of
    # integer-pipe variants, `e2e` used to show 250 us of "host overhead" that was clock drift; the
    # copies and the synchronisation themselves cost ~40 us (scripts/c3_e2e_timeline.py).
    # end to end through the array API: pinned host queries in, numpy results out
    def e2e_step(rec):
        return m.match_host(torch.from_numpy(q_pin).to("cuda", non_blocking=True))
    sampler = ClockSampler(D.local).start() if D.rank == 0 else None
    e2e_ms, _ = timed_steps(D, e2e_step, args.steps, args.warmup, flush)
    e2e_clocks = sampler.stop() if sampler else None

    # ... and through the cv2-shaped facade: BFMatcher.knnMatch(host queries, k=2) over the
    # resident (sharded) collection -> rows of DMatch (imgIdx / trainIdx / distance)
    bf = BFMatcher.from_resident(shard, per_kf, shard_base=lo_kf * per_kf, n_images_total=nkf,
                                 group=True if D.world > 1 else None)
    if D.world > 1 and m._px is not None:
        bf._sharded[0].use_exchange(m._px)                  # same queries: share the peer buffers

    def e2e_knn_step(rec):
        return bf.knnMatch(q_pin, k=2)
    knn_api_ms, _ = timed_steps(D, e2e_knn_step, args.steps, args.warmup, flush)
    api_rows = e2e_knn_step(None)

    variants = {}
    # the integer-pipe kernel on the same shard (the north star's formulation; the shipped
    # default for databases below 16384 rows): scan + split merge, same events
    popc_ms = None
    if m.signs is not None:
        def step_popc(rec):
            a, b = ev_pair(torch)
            a.record()
            k = ops.hamming_knn2_keys(q_dev, shard, m.shard_base)
            b.record()
            if rec is not None:
                rec.append((a, b))
            return k
        _, popc_ms = timed_steps(D, step_popc, max(3, args.steps // 2), args.warmup, flush)
        popc_ms /= max(3, args.steps // 2)
        keys_tc = m.local_keys(q_dev)
        variants["integer_pipe_kernel_same_shard"] = {
            "Gpairs_per_s": float(nq) * float(shard.shape[0]) / (popc_ms * 1e-3) / 1e9, "kernel_ms": popc_ms,
            "keys_equal_tensor_core_keys": bool(torch.equal(keys_tc, step_popc(None)))}
        # ... and the OTHER tensor-core format (kind::i8 when the default kind::mxf4 is selected)
        other = 8 if tc_fmt == 4 else 4
        ops.hamming_tc_set_format(other)
        try:
            signs_o = ops.hamming_tc_expand(shard)

            def step_other(rec):
                a, b = ev_pair(torch)
                a.record()
                k = ops.hamming_knn2_keys_tc(q_dev, signs_o, int(shard.shape[0]), m.shard_base)
                b.record()
                if rec is not None:
                    rec.append((a, b))
                return k
            _, o_ms = timed_steps(D, step_other, max(3, args.steps // 2), args.warmup, flush)
            o_ms /= max(3, args.steps // 2)
            variants["tcgen05_%s_same_shard" % ("kind_i8" if other == 8 else "kind_mxf4")] = {
                "Gpairs_per_s": float(nq) * float(shard.shape[0]) / (o_ms * 1e-3) / 1e9, "kernel_ms": o_ms,
                "expanded_row_bytes": ops.hamming_tc_row_bytes(),
                "keys_equal_default_engine_keys": bool(torch.equal(keys_tc, step_other(None)))}
            del signs_o
        finally:
            ops.hamming_tc_set_format(tc_fmt)
    if D.world > 1:
        m_nccl = ShardedKeyframeMatcher(shard, lo_kf * per_kf, exchange="nccl")

        def step_nccl(rec):
            return m_nccl.match(q_dev, 0.75)
        ms_n, _ = timed_steps(D, step_nccl, args.steps, args.warmup, flush)
        variants["nccl_allgather_allreduce_between_kernels"] = {
            "Gpairs_per_s": float(nq) * nkf * per_kf * args.steps / (ms_n * 1e-3) / 1e9, "ms": ms_n / args.steps}

        # WEAK scaling of the same path (the headline is strong scaling of a 2.7 ms step, where ~60 us
        # of fixed cost per launch are 13 % of an 8-GPU shard): every rank holds a WHOLE copy of the
        # seeded database as its own keyframe range (global rows rank * nt_total ...), i.e. N x 20 M
        # rows in total, same kernels and fused exchanges.  The answer is known: the best row, then
        # either the tied row the single database already had or the best row's copy on rank 1.
        try:
            full = torch.from_numpy(db_np).cuda()
            m_w = ShardedKeyframeMatcher(full, D.rank * nt_total)
            if m._px is not None:
                m_w.use_exchange(m._px)

            def step_w(rec):
                return m_w.match(q_dev, 0.75)
            ms_w, _ = timed_steps(D, step_w, args.steps, args.warmup, flush)
            rw = [_ffi.to_host(x) for x in step_w(None)]
            m_w.check()
            tied = result[1][:, 1] == result[1][:, 0]
            want2 = np.where(tied, result[0][:, 1], result[0][:, 0] + nt_total)
            variants["weak_20M_rows_per_gpu"] = {
                "Gpairs_per_s": float(nq) * nt_total * D.world * args.steps / (ms_w * 1e-3) / 1e9,
                "ms": ms_w / args.steps, "rows_total": nt_total * D.world,
                "result_as_predicted": bool(np.array_equal(rw[0][:, 0], result[0][:, 0])
                                            and np.array_equal(rw[0][:, 1], want2)
                                            and np.array_equal(rw[1][:, 0], result[1][:, 0])
                                            and np.array_equal(rw[1][:, 1], result[1][:, 0])
                                            and not rw[2][~tied].any()),
                "note": "efficiency = this / (N x the N=1 headline value)"}
            del m_w, full
        except Exception as e:                                   # never lose the headline line to the variant
            variants["weak_20M_rows_per_gpu"] = {"error": repr(e)[:300]}

    pairs = float(nq) * nt_total
    launches =((6 if D.world == 1 else 9) + (1 if m.signs is not None else 0)) * args.steps
    res = {
        "metric": "hamming_knn_k2_ratio_mutual", "unit": "Gpairs/s",
        "value": pairs * args.steps / (total_ms * 1e-3) / 1e9,
        "ms_per_step": total_ms / args.steps, "scaling": "strong",
        "dtype": ("%s (tcgen05.mma), %s" % (tc_kind, "f32 min3" if tc_fmt == 4 else "s16x2 min/max"))
                 if m.signs is not None else "u32 (xor/popc/int-min)",
        "engine": engine,
        "config": {"workload": "C3 loop-closure: 2000 ORB queries vs %d keyframes x 2000 "
                               "(%d descriptors, %d MB), k=2 + ratio 0.75 + mutual; database sharded by "
                               "keyframe range over %d GPU(s)" % (nkf, nt_total, nt_total * 32 >> 20, D.world),
                   "nq": nq, "nt": nt_total, "shard_rows": int(shard.shape[0]),
                   "l2": L2_NOTE,
                   "state": "database resident in HBM (compact rows%s); queries are the per-step input"
                            % (" + the %d B/row expanded copy the tensor-core scan reads" % ops.hamming_tc_row_bytes() if m.signs is not None else "")},
        "e2e": {"value": pairs * args.steps / (e2e_ms * 1e-3) / 1e9, "unit": "Gpairs/s",
                "h2d_bytes_per_step": nq * 32,
                "d2h_bytes_per_step": nq * (2 * 8 + 2 * 4 + 3),
                "api": "ShardedKeyframeMatcher.match_host (array API: idx, dist, ratio_ok, mutual_ok, valid "
                       "as numpy; one packed copy)",
                "clocks": e2e_clocks,
                "order": "timed directly after the device-resident loop (same SM clock state), before the variants"},
        "e2e_knnMatch": {"value": pairs * args.steps / (knn_api_ms * 1e-3) / 1e9, "unit": "Gpairs/s",
                         "ms_per_step": knn_api_ms / args.steps,
                         "h2d_bytes_per_step": nq * 32, "d2h_bytes_per_step": nq * (2 * 8 + 2 * 4 + 3),
                         "api": "BFMatcher.knnMatch(host queries, k=2) over the resident collection -> "
                                "KnnMatches (rows of DMatch, built on access)"},
        "gpu_launches": launches,
        "clocks": clocks,
        "exchange": ("top-2 keys and matched rows pushed to every peer from inside the kernels over NVLink peer "
                     "memory (symmetric memory); no NCCL call in the step" if D.world > 1 else "none (1 GPU)"),
        "variants": variants,
    }
    # ------------------------------------------------------------ parity, in the record
    dg = digest(*result)
    parity = {"checked": "idx, dist, ratio_ok, mutual_ok, valid (all %d queries)" % nq}
    api_ok = (len(api_rows) == nq and all(
        [(d.imgIdx * per_kf + d.trainIdx, int(d.distance)) for d in api_rows[i]] ==
        [(int(result[0][i, j]), int(result[1][i, j])) for j in range(2)] for i in range(0, nq, 97)))
    parity["knnMatch_rows_equal_array_api"] = bool(api_ok)
    if D.world > 1:
        # every rank holds the same global result; rank 0 also scans the WHOLE database on its
        # own GPU (outside the timed region) and requires bit-equality with the sharded result
        all_dg = [None] * D.world
        D.dist.all_gather_object(all_dg, dg)
        parity["ranks_agree"] = bool(all(x == all_dg[0] for x in all_dg))
        if D.rank == 0:
            db_dev = torch.from_numpy(db_np).cuda()
            os.environ["PBK_HAMMING_ENGINE"] = "popc"          # the reference scan runs on the OTHER engine
            try:
                whole = [_ffi.to_host(x) for x in ops.knn_ratio_mutual(q_dev, db_dev, 0.75)]
            finally:
                del os.environ["PBK_HAMMING_ENGINE"]
            parity["vs_single_gpu_scan"] = bool(all(np.array_equal(a, b) for a, b in zip(result, whole)))
            parity["rows"] = nt_total
            parity["bit_exact"] = bool(parity["vs_single_gpu_scan"] and parity["ranks_agree"] and api_ok)
            parity["reference"] = ("single-GPU scan of the whole database on rank 0 with the integer-pipe kernel "
                                   "(the N=1 line checks the same result_digest against cv2.BFMatcher)")
            del db_dev
    if D.rank == 0:
        pk = probe_peaks(D)
        local_pairs = float(nq) * float(shard.shape[0])
        pair_rate = local_pairs * args.steps / (knn_ms * 1e-3)
        popc_roof = {
            "bound": "popc", "kernel": "hamming_knn2_kernel<2,4> (+ split merge)", "unit": "Tpopc/s",
            "peak": pk["popc_per_s"] / 1e12, "peak_source": "pbk_popc_peak_probe mode 0, measured in this run (burst)",
            "algorithmic": "frac = POPC.32 ISSUED (4 per pair: the eight XOR words are carry-save-compressed "
                           "into four of weight 1,1,2,4; SASS counts in profiles/r02_hamming_inner_loop.md) / "
                           "measured popc-pipe peak = utilisation of the binding (XU) pipe; frac_algorithmic "
                           "= the SURVEY 8(d) figure, 8 POPC per 256-bit pair / the same peak (> 1 because "
                           "the kernel issues half of them)",
            "traffic": ncu_traffic("hamming_knn2_kernel", scale=float(shard.shape[0]) / 2.5e6),
            "traffic_source": ncu_traffic_note("hamming_knn2_kernel", scale=float(shard.shape[0]) / 2.5e6),
            "traffic_stale": ncu_traffic_stale("hamming.cu")}
        if m.signs is None:
            res["roofline"] = dict(popc_roof, achieved=pair_rate * POPC_ISSUED_PER_PAIR / 1e12,
                                   frac=pair_rate * POPC_ISSUED_PER_PAIR / pk["popc_per_s"],
                                   frac_algorithmic=pair_rate * 8 / pk["popc_per_s"],
                                   kernel_ms=knn_ms / args.steps, kernel_share_of_step=knn_ms / total_ms,
                                   Gpairs_per_s_kernel=pair_rate / 1e9)
        else:
            bf16, bf16_src = measured_bf16()
            tc_peak = 2.0 * pk["tc_mac_per_s"] / 1e12    # TOP/s, measured in this run (burst, ~1 ms launches)
            tops = pair_rate * 256 * 2 / 1e12            # algorithmic: 256 multiply-adds per 256-bit pair
            nominal = 9000.0 if tc_fmt == 4 else 4500.0
            res["roofline"] = {
                "bound": "tensor", "kernel": "%s (tcgen05.mma %s, + query expansion and split merge)" % (tc_kernel, tc_kind.split(" ")[0]),
                "achieved": tops, "peak": tc_peak, "unit": "TOP/s", "frac": tops / tc_peak,
                "frac_issued": tops / tc_peak,
                "frac_of_nominal_%d" % int(nominal): tops / nominal,
                "frac_of_%dx_measured_bf16" % (4 if tc_fmt == 4 else 2): tops / ((4.0 if tc_fmt == 4 else 2.0) * bf16),
                "peak_source": "pbk_hamming_tc_peak_probe, measured in this run: back-to-back tcgen05.mma %s of the "
                               "kernel's shape (M 128, N 128, K %d) on one CTA per SM, no data movement (burst; nominal "
                               "dense %s = %.0f TOP/s; %d x the dense bf16 of %s = %.0f TOP/s)"
                               % (tc_kind.split(" ")[0], 64 if tc_fmt == 4 else 32, "fp4" if tc_fmt == 4 else "int8", nominal,
                                  4 if tc_fmt == 4 else 2, bf16_src, (4.0 if tc_fmt == 4 else 2.0) * bf16),
                "algorithmic": "256 multiply-adds (2 ops each) per descriptor pair: the dot product of the +-1 "
                               "expansions is 2 d - 256 (mxf4) / 128 d - 16384 (i8), and the kernel issues exactly those "
                               "(%s): frac_issued = frac" % ("4 K steps of 64" if tc_fmt == 4 else "8 K steps of 32"),
                "traffic": ncu_traffic(tc_kernel, scale=float(shard.shape[0]) / 2.5e6),
                "traffic_source": ncu_traffic_note(tc_kernel, scale=float(shard.shape[0]) / 2.5e6),
                "traffic_stale": ncu_traffic_stale("hamming_tc4.cu" if tc_fmt == 4 else "hamming_tc.cu"),
                "kernel_ms": knn_ms / args.steps, "kernel_share_of_step": knn_ms / total_ms,
                "Gpairs_per_s_kernel": pair_rate / 1e9,
                "integer_pipe_kernel": dict(
                    popc_roof, kernel_ms=popc_ms,
                    achieved=local_pairs / (popc_ms * 1e-3) * POPC_ISSUED_PER_PAIR / 1e12,
                    frac=local_pairs / (popc_ms * 1e-3) * POPC_ISSUED_PER_PAIR / pk["popc_per_s"],
                    frac_algorithmic=local_pairs / (popc_ms * 1e-3) * 8 / pk["popc_per_s"],
                    note="the north star's formulation, timed on the same shard in this run; the default "
                         "engine for databases below 16384 rows (per-frame matching)")}
        if D.world == 1:
            base, cpu_result = cpu_baseline_c3(q_np, db_np, per_kf, args.parity_keyframes)
            res["cpu_baseline"] = base
            kf = min(args.parity_keyframes, nkf)
            if kf == nkf:
                mine = result
            else:                           # the same rows on the GPU (outside the timed region)
                mine = [_ffi.to_host(x) for x in ops.knn_ratio_mutual(q_dev, shard[:kf * per_kf], 0.75)]
            same = [bool(np.array_equal(np.asarray(a).astype(np.int64), np.asarray(b).astype(np.int64)))
                    for a, b in zip(mine, cpu_result)]
            parity.update({"rows": kf * per_kf, "bit_exact": bool(all(same) and api_ok),
                           "reference": "cv2.BFMatcher(NORM_HAMMING).add(%d keyframes).knnMatch(q, k=2) + ratio + "
                                        "mutual (oracle.port) on the same rows: the cpu_baseline run" % kf,
                           "per_output": dict(zip(("idx", "dist", "ratio_ok", "mutual_ok", "valid"), same))})
        res["parity"] = parity
        res["result_digest"] = dg
    return res


def cpu_baseline_c3(q, db, per_kf, sample_kf=10_000):
    """cv2.BFMatcher k=2 + ratio + mutual (what a pybot user runs) on the first `sample_kf`
    keyframes of the same database (default: all of them), all host threads.  Returns the
    baseline record and the result arrays (the parity reference)."""
    from oracle import port
    try:
        import cv2
        threads = os.cpu_count()
        cv2.setNumThreads(threads)
        engine = "cv2.BFMatcher(NORM_HAMMING).knnMatch"
    except Exception:
        threads, engine = 1, "numpy model"
    sample_kf = min(sample_kf, len(db) // per_kf)
    kfs = [db[i * per_kf:(i + 1) * per_kf] for i in range(sample_kf)]
    port.knn_ratio_mutual(q[:64], kfs[:2], 0.75)
    t0 = time.perf_counter()
    out = port.knn_ratio_mutual(q, kfs, 0.75)
    dt = time.perf_counter() - t0
    return ({"value": len(q) * sample_kf * per_kf / dt / 1e9, "unit": "Gpairs/s", "cores": threads,
             "kind": "port",
             "sample": "%s k=2 + ratio + mutual via oracle.port, %d q x %d rows (%d keyframes), %.1f s"
                       % (engine, len(q), sample_kf * per_kf, sample_kf, dt)}, out)


# ------------------------------------------------------------ workload C1
def run_c1(D, args):
    """C1, the reference's own per-frame case: one KITTI-shaped stereo frame
    (1241x376 disparity -> XYZ) + 2000x2000 ORB match with ratio + mutual.  At this
    size the GPU is launch-latency-bound: what is reported is the per-frame
    LATENCY through the public API with host (pinned) arrays in and numpy out."""
    import torch
    from oracle import port
    from pybot_b200 import _ffi
    from pybot_b200.vision.camera_utils import StereoCamera
    from pybot_b200.vision.matcher import knn_ratio_mutual
    H, W = syn.KITTI_SHAPE
    cam = StereoCamera.from_calib_params(syn.KITTI_FX, syn.KITTI_FX, syn.KITTI_CX, syn.KITTI_CY,
                                         baseline_px=syn.KITTI_BASELINE_PX, shape=(H, W))
    disp_np = syn.c1_frame()
    q_np, t_np = syn.orb_pair()
    disp = _ffi.pinned_empty(disp_np.shape, np.float32)
    q = _ffi.pinned_empty(q_np.shape, np.uint8)
    t = _ffi.pinned_empty(t_np.shape, np.uint8)
    disp[:], q[:], t[:] = disp_np, q_np, t_np
    steps = max(args.steps, 20)

    def frame(rec):
        X = cam.reconstruct(disp)
        m = knn_ratio_mutual(q, t, 0.75)
        return X, m
    ms, _ = timed_steps(D, frame, steps, args.warmup)

    dq, dt_, dd = (torch.from_numpy(a).cuda() for a in (q_np, t_np, disp_np))

    def resident(rec):
        return cam.reconstruct(dd), knn_ratio_mutual(dq, dt_, 0.75)
    ms_dev, _ = timed_steps(D, resident, steps, args.warmup)

    # the same frame with a prepared, CUDA-graph-captured matcher (buffers allocated once)
    from pybot_b200 import ops
    pm = ops.PreparedMatcher(len(q_np), len(t_np), 0.75)

    def graphed(rec):
        X = cam.reconstruct(disp)
        idx, dist, r_ok, m_ok, valid = pm.run(q, t)
        return X, _ffi.to_host(idx), _ffi.to_host(dist), _ffi.to_host(valid)
    ms_graph, _ = timed_steps(D, graphed, steps, args.warmup)

    def graphed_resident(rec):
        return cam.reconstruct(dd), pm.run(dq, dt_)
    ms_graph_dev, _ = timed_steps(D, graphed_resident, steps, args.warmup)

    Q = cam.Q
    port.stereo_reconstruct(disp_np, Q)
    port.knn_ratio_mutual(q_np, [t_np], 0.75)
    t0 = time.perf_counter()
    for _ in range(5):
        port.stereo_reconstruct(disp_np, Q)
        port.knn_ratio_mutual(q_np, [t_np], 0.75)
    cpu_ms = (time.perf_counter() - t0) / 5 * 1e3
    return {"workload": "C1 one KITTI frame: %dx%d disparity -> XYZ + 2000x2000 ORB k=2 + ratio + mutual" % (W, H),
            "latency_ms_host_arrays": ms / steps, "latency_ms_device_resident": ms_dev / steps,
            "latency_ms_host_arrays_cuda_graph": ms_graph / steps,
            "latency_ms_device_resident_cuda_graph": ms_graph_dev / steps,
            "frames_per_s": 1e3 * steps / ms, "h2d_bytes": int(disp_np.nbytes + q_np.nbytes + t_np.nbytes),
            "d2h_bytes": int(H * W * 12 + 2000 * 25),
            "cpu_baseline": {"value": cpu_ms, "unit": "ms/frame", "cores": os.cpu_count(), "kind": "port",
                             "sample": "cv2.reprojectImageTo3D + cv2.BFMatcher k=2 + ratio + mutual (oracle.port), "
                                       "same frame, mean of 5"}}


# ------------------------------------------------------------ workload C2
def run_c2(D, args, n=10_000_000):
    import torch
    from pybot_b200 import _ffi, ops
    from pybot_b200.geometry import RigidTransform
    from pybot_b200.vision.camera_utils import Camera, CameraIntrinsic, CameraExtrinsic
    hbm, src = measured_peaks()
    X_np = syn.c2_points(n=n)
    X_pin = _ffi.pinned_empty((n, 3), np.float32)
    X_pin[:] = X_np
    X = torch.from_numpy(X_np).cuda()
    pose = RigidTransform.from_rpyxyz(*syn.C2_RPYXYZ)
    K = syn.kitti_K()
    cam = Camera.from_intrinsics_extrinsics(CameraIntrinsic(K, shape=syn.KITTI_SHAPE),
                                            CameraExtrinsic.from_rigid_transform(pose))
    Rp, t, Rz, tz = cam._projection_params()
    flush = make_flush(torch)
    out = {}

    pr_plain = ops.Projector(Rp, t, K, None, Rz=Rz, tz=tz)

    def plain(rec):
        return pr_plain.run(X)
    ms, _ = timed_steps(D, plain, args.steps, args.warmup, flush)
    out["project_plain"] = {"Mpts_per_s": n * args.steps / (ms * 1e-3) / 1e6, "ms": ms / args.steps,
                            "bytes_per_pt": 20, "GBps": n * 20 * args.steps / (ms * 1e-3) / 1e9}

    pr_mask = ops.Projector(Rp, t, K, None, Rz=Rz, tz=tz, shape=syn.KITTI_SHAPE, check_bounds=True,
                            check_depth=True, return_depth=True, return_valid=True)

    def masked(rec):
        return pr_mask.run(X)
    ms2, _ = timed_steps(D, masked, args.steps, args.warmup, flush)
    m_valid = int(masked(None)[3].item())
    bpp = 12 + 1 + 16.0 * m_valid / n
    out["project_masks_compact"] = {"Mpts_per_s": n * args.steps / (ms2 * 1e-3) / 1e6,
                                    "ms": ms2 / args.steps, "bytes_per_pt": bpp, "valid_frac": m_valid / n,
                                    "GBps": n * bpp * args.steps / (ms2 * 1e-3) / 1e9,
                                    "frac_hbm": n * bpp * args.steps / (ms2 * 1e-3) / 1e9 / hbm,
                                    "kernel": ("compact_classify_kernel + project_compacted_kernel (PBK_COMPACT_IMPL=2)"
                                               if os.environ.get("PBK_COMPACT_IMPL", "1")[:1] == "2" else
                                               "project_compact_kernel<float,false> (single pass, decoupled look-back)"),
                                    "traffic": ncu_traffic("project_compact_kernel<float, 0>"),
                                    "traffic_stale": ncu_traffic_stale("transform_project.cu")}
    pr_mask_nc = ops.Projector(Rp, t, K, None, Rz=Rz, tz=tz, shape=syn.KITTI_SHAPE, check_bounds=True,
                               check_depth=True, return_depth=True, return_valid=True, compact=False)

    def masked_nc(rec):
        return pr_mask_nc.run(X)
    ms2b, _ = timed_steps(D, masked_nc, args.steps, args.warmup, flush)
    out["project_masks_uncompacted"] = {"Mpts_per_s": n * args.steps / (ms2b * 1e-3) / 1e6, "ms": ms2b / args.steps,
                                        "bytes_per_pt": 29, "GBps": n * 29.0 * args.steps / (ms2b * 1e-3) / 1e9}

    Y64 = torch.empty((n, 3), dtype=torch.float64, device="cuda")

    def xform(rec):
        return ops.transform_points(X, Rp, t, 1.0, np.float64, out=Y64)
    ms3, _ = timed_steps(D, xform, args.steps, args.warmup, flush)
    out["transform_f64_out"] = {"Mpts_per_s": n * args.steps / (ms3 * 1e-3) / 1e6, "ms": ms3 / args.steps,
                                "bytes_per_pt": 36, "GBps": n * 36 * args.steps / (ms3 * 1e-3) / 1e9}

    def e2e(rec):
        return cam.project(X_pin)
    ems, _ = timed_steps(D, e2e, args.steps, args.warmup, flush)
    res = {
        "metric": "transform_project", "unit": "Mpts/s", "value": out["project_plain"]["Mpts_per_s"],
        "ms_per_step": out["project_plain"]["ms"], "dtype": "f64 registers, f32 in/out",
        "config": {"workload": "C2 batched SE(3) transform + pinhole Camera.project of %d points "
                               "(float32 AoS in, float32 pixels out)" % n,
                   "l2": L2_NOTE},
        "variants": out,
        "e2e": {"value": n * args.steps / (ems * 1e-3) / 1e6, "unit": "Mpts/s",
                "h2d_bytes_per_step": n * 12, "d2h_bytes_per_step": n * 8,
                "bound": "PCIe: 200 MB cross the host link per step (upload, kernel and download of 12 MB chunks "
                         "overlapped on three streams); the kernel itself is %.0fx faster"
                         % (out["project_plain"]["Mpts_per_s"] / (n * args.steps / (ems * 1e-3) / 1e6))},
        "gpu_launches": args.steps,
        "roofline": {"bound": "hbm", "kernel": "project_points_kernel<float,false,false>",
                     "achieved": out["project_plain"]["GBps"], "peak": hbm, "unit": "GB/s",
                     "frac": out["project_plain"]["GBps"] / hbm, "traffic": ncu_traffic("project_points_kernel<float, 0, 0>"),
                     "traffic_source": ncu_traffic_note("project_points_kernel<float, 0, 0>"),
                     "traffic_stale": ncu_traffic_stale("transform_project.cu"),
                     "peak_source": src, "algorithmic": "12 B in + 8 B out = 20 B/pt",
                     "copy_same_bytes_GBps": copy_at_size(torch, n * 12, n * 8)},
    }
    if D.rank == 0:
        res["cpu_baseline"] = cpu_baseline_c2(X_np[:1_000_000], cam)
    return res


def cpu_baseline_c2(X, cam):
    from oracle import port
    try:
        import cv2
        cv2.setNumThreads(os.cpu_count())
    except Exception:
        pass
    K = cam.K
    port.camera_project(K, np.zeros(5), cam.quat.q, cam.tvec, syn.KITTI_SHAPE, X[:1000])
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        port.camera_project(K, np.zeros(5), cam.quat.q, cam.tvec, syn.KITTI_SHAPE, X)
        best = min(best, time.perf_counter() - t0)
    return {"value": len(X) / best / 1e6, "unit": "Mpts/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "oracle.port.camera_project (cv2.Rodrigues + cv2.projectPoints as at "
                      "camera_utils.py:694-697) on %d of the points, best of 3" % len(X)}


# ------------------------------------------------------------ workload C4
def run_c4(D, args, weak=False):
    """C4.  Strong scaling: the 4 M observations are split into equal contiguous chunks over the
    ranks.  weak=True: every rank owns one whole C4 block (its own 1000 cameras' worth of 4 M
    observations; cameras and points replicated as always), i.e. 4 M observations PER GPU - the
    'large BA observation sets' case; the global cost is then world x the block's cost."""
    import torch
    from pybot_b200 import _ffi, ops
    from pybot_b200.distributed import PeerCostExchange, ShardedBA
    hbm, src = measured_peaks()
    prob = syn.c4_problem()
    rv, tv, K, pts, oc, op, uv = prob
    n_obs = len(oc) * (D.world if weak else 1)
    if weak and D.world > 1:
        sb = ShardedBA.__new__(ShardedBA)                  # this rank's block IS its chunk
        sb.dist, sb.group, sb.world, sb.rank = D.dist, None, D.world, D.rank
        sb.range = (D.rank * len(oc), (D.rank + 1) * len(oc))
        sb.problem = ops.BAProblemDevice(*prob)
        sb.exchange = PeerCostExchange(None)
    else:
        sb = ShardedBA(*prob)
    flush = make_flush(torch)

    def step(rec):
        # one pbk_ba_evaluate call: camera precompute + streaming kernel (launched with programmatic
        # stream serialization; world > 1: the cost all-reduce is fused into the kernel over NVLink
        # peer memory)
        return sb.evaluate()

    def step_graph(rec):
        return sb.evaluate(graph=True)

    def kernel_only(rec):
        a, b = ev_pair(torch)
        sb.problem.evaluate(events=(a, b), exchange=sb.exchange)   # events bracket the streaming kernel
        if rec is not None:
            rec.append((a, b))
        return sb.problem.cost
    sampler = ClockSampler(D.local).start() if D.rank == 0 else None
    ms, _ = timed_steps(D, step, args.steps, args.warmup, flush)
    ms_g, _ = timed_steps(D, step_graph, args.steps, args.warmup, flush)
    _, kms = timed_steps(D, kernel_only, args.steps, args.warmup, flush)
    clocks = sampler.stop() if sampler else None
    cost = sb.cost_host()
    variants = {"cuda_graph_replay": {"Mobs_per_s": n_obs * args.steps / (ms_g * 1e-3) / 1e6, "ms": ms_g / args.steps}}
    if D.world > 1 and not weak:
        sb_nccl = ShardedBA(*prob, exchange="nccl")

        def step_nccl(rec):
            return sb_nccl.evaluate()
        ms_n, _ = timed_steps(D, step_nccl, args.steps, args.warmup, flush)
        variants["nccl_allreduce_after_kernel"] = {"Mobs_per_s": n_obs * args.steps / (ms_n * 1e-3) / 1e6,
                                                   "ms": ms_n / args.steps}
        del sb_nccl

    # e2e: poses + points change every LM iteration -> H2D rvec/tvec/points, D2H cost
    rv_pin, tv_pin, pts_pin = (_ffi.pinned_empty(a.shape, np.float32) for a in (rv, tv, pts))
    rv_pin[:], tv_pin[:], pts_pin[:] = rv, tv, pts
    P = sb.problem

    def e2e(rec):
        P.rvec.copy_(torch.from_numpy(rv_pin), non_blocking=True)
        P.tvec.copy_(torch.from_numpy(tv_pin), non_blocking=True)
        P.points.copy_(torch.from_numpy(pts_pin), non_blocking=True)
        return float(sb.evaluate().item())
    ems, _ = timed_steps(D, e2e, args.steps, args.warmup, flush)
    local_obs = P.n_obs
    gbps = local_obs * 96.0 * args.steps / (kms * 1e-3) / 1e9
    step_ms = ms / args.steps
    hbm_ms = local_obs * 96.0 / (hbm * 1e9) * 1e3
    res = {
        "metric": "ba_residual_jacobian", "unit": "Mobs/s", "value": n_obs * args.steps / (ms * 1e-3) / 1e6,
        "ms_per_step": step_ms, "scaling": "weak" if weak else "strong", "dtype": "f64 registers, f32 in/out",
        "config": {"workload": "C4 BA residual + 2x6/2x3 Jacobians: 1000 cameras, 1M points, %d observations "
                               "%s over %d GPU(s)" % (n_obs, "(one 4M block per GPU)" if weak else "sharded", D.world),
                   "l2": L2_NOTE},
        "e2e": {"value": n_obs * args.steps / (ems * 1e-3) / 1e6, "unit": "Mobs/s",
                "h2d_bytes_per_step": int(rv.nbytes + tv.nbytes + pts.nbytes), "d2h_bytes_per_step": 8},
        "gpu_launches": 2 * args.steps, "clocks": clocks,
        "exchange": ("cost sum-all-reduce fused into the kernel over NVLink peer memory (symmetric memory), "
                     "no NCCL call in the step" if D.world > 1 else "none (1 GPU)"),
        "variants": variants,
        "limiter": {"hbm_ms_at_measured_peak": hbm_ms, "step_ms": step_ms, "fixed_ms": step_ms - hbm_ms,
                    "note": ("step = HBM time of this rank's %d observations x 96 B + launch / ramp / tail / last-CTA cost "
                             "reduction%s" % (local_obs, " + one NVLink crossing (self-validating words, no fence) and the wait for "
                                              "the slowest rank" if D.world > 1 else ""))},
        "roofline": {"bound": "hbm", "kernel": "ba_residual_jacobian_kernel<2,2> (behind ba_camera_precompute_kernel, "
                                               "whose ~3 us overlap its prologue)", "achieved": gbps,
                     "peak": hbm, "unit": "GB/s", "frac": gbps / hbm,
                     "frac_full_step": local_obs * 96.0 / (step_ms * 1e-3) / 1e9 / hbm,
                     "traffic": ncu_traffic("ba_residual_jacobian_kernel<2, 2>") if D.world == 1 else None,
                     "traffic_source": ncu_traffic_note("ba_residual_jacobian_kernel<2, 2>"),
                     "traffic_stale": ncu_traffic_stale("ba.cu"),
                     "peak_source": src,
                     "algorithmic": "16 B read + 8 + 48 + 24 B written = 96 B/obs",
                     "kernel_ms": kms / args.steps,
                     "copy_same_bytes_GBps": copy_at_size(torch, local_obs * 16, local_obs * 80),
                     "note": "frac: streaming kernel timed alone; frac_full_step: the whole pbk_ba_evaluate "
                             "call a caller gets (precompute + streaming kernel).  83 % of the traffic is "
                             "writes; a write-only stream of this size sustains ~5.3 TB/s on B200 HBM3e "
                             "(scripts/hbm_probe.py), below the copy figure"},
    }
    # ------------------------------------------------------------ parity, in the record
    whole = ops.BAProblemDevice(*prob).evaluate()             # the whole block on this GPU, outside the timed region
    a, b = (0, len(oc)) if weak else sb.range
    same = bool(torch.equal(P.r, whole.r[a:b]) and torch.equal(P.Jc, whole.Jc[a:b]) and torch.equal(P.Jp, whole.Jp[a:b]))
    ref_cost = float(whole.cost.item()) * (D.world if weak else 1)
    cost_ok = bool(abs(cost - ref_cost) <= 1e-12 * abs(ref_cost))
    parity = {"slices_equal_single_gpu": same, "cost_rel_err_vs_single_gpu": abs(cost - ref_cost) / abs(ref_cost)}
    if D.world > 1:
        flags = [None] * D.world
        D.dist.all_gather_object(flags, (same, cost))
        same = all(f[0] for f in flags)
        parity["ranks_agree_on_cost"] = bool(all(f[1] == flags[0][1] for f in flags))
        parity["slices_equal_single_gpu"] = bool(same)
        cost_ok = cost_ok and parity["ranks_agree_on_cost"]
    if D.rank == 0:
        res["result_digest"] = digest(whole.r, whole.Jc, whole.Jp)
        parity["obs"] = int(n_obs)
        if D.world == 1:
            base, tol = cpu_baseline_c4(prob, whole=whole)
            res["cpu_baseline"] = base
            parity.update(tol)
            parity["bit_exact"] = None
            parity["within_tolerance"] = bool(tol["r_within_1e-4px_plus_1e-5rel"] and tol["max_rel_J"] <= 1e-5
                                              and cost_ok and same)
            parity["tolerance"] = "north star: residual 1e-4 px, Jacobians 1e-5 relative (fp32 outputs)"
        else:
            parity["bit_exact"] = bool(same and cost_ok)
            parity["reference"] = "single-GPU evaluation of the whole problem (N=1 line: same result_digest)"
        res["parity"] = parity
    del whole
    return res


def cpu_baseline_c4(prob, n_cams=100, whole=None):
    """cv2.projectPoints with its Jacobian, looped per camera, on the observations of the first
    `n_cams` cameras.  With `whole` (the GPU evaluation) also returns the worst differences."""
    from oracle import port
    rv, tv, K, pts, oc, op, uv = prob
    sel = oc < n_cams
    port.ba_residual_jacobian(rv, tv, K, pts, oc[:100], op[:100], uv[:100])
    t0 = time.perf_counter()
    rr, rJc, rJp, rcost = port.ba_residual_jacobian(rv, tv, K, pts, oc[sel], op[sel], uv[sel])
    dt = time.perf_counter() - t0
    base = {"value": int(sel.sum()) / dt / 1e6, "unit": "Mobs/s", "cores": 1, "kind": "port",
            "sample": "cv2.projectPoints with Jacobian looped per camera (oracle.port), first %d "
                      "cameras = %d obs, %.1f s" % (n_cams, int(sel.sum()), dt)}
    if whole is None:
        return base
    idx = np.nonzero(sel)[0]
    assert idx[-1] == len(idx) - 1                        # sorted by camera: a prefix
    m = len(idx)
    r, Jc, Jp = (x[:m].cpu().numpy().astype(np.float64) for x in (whole.r, whole.Jc, whole.Jp))

    def rel(a, b):
        a, b = a.reshape(m, -1), np.asarray(b, np.float64).reshape(m, -1)
        return float((np.abs(a - b) / (np.abs(b).max(axis=1, keepdims=True) + 1e-30)).max())
    return base, {"rows_vs_cv2": int(m), "max_abs_r_px": float(np.abs(r - rr).max()),
                  "r_within_1e-4px_plus_1e-5rel": bool(np.allclose(r, rr, rtol=1e-5, atol=1e-4)),
                  "max_rel_J": max(rel(Jc, rJc), rel(Jp, rJp)),
                  "reference": "cv2.projectPoints (x, Jacobian) per camera on the first %d cameras' observations "
                               "(oracle.port): the cpu_baseline run" % n_cams}


# ------------------------------------------------------------ workload C5
def run_c5(D, args, batch=64):
    import torch
    from pybot_b200 import _ffi
    from pybot_b200.vision.camera_utils import StereoCamera
    hbm, src = measured_peaks()
    H, W = syn.UHD_SHAPE
    cam = StereoCamera.from_calib_params(syn.UHD_FX, syn.UHD_FX, syn.UHD_CX, syn.UHD_CY,
                                         baseline=syn.UHD_BASELINE, shape=syn.UHD_SHAPE)
    # same distribution as synthetic.c5_batch, generated on the device (3.7 GB)
    g = torch.Generator(device="cuda").manual_seed(1005)
    Z = torch.rand((batch, H, W), device="cuda", generator=g) * 76 + 4
    disp = (syn.UHD_FX * syn.UHD_BASELINE) / Z
    u = torch.rand((batch, H, W), device="cuda", generator=g)
    disp[u < 0.05] = 0.0
    disp[(u >= 0.05) & (u < 0.06)] = -1.0
    del Z, u
    im = torch.randint(0, 256, (batch, H, W, 3), device="cuda", dtype=torch.uint8, generator=g)
    npx = batch * H * W

    def step(rec):
        return cam.reconstruct_with_texture(disp, im)
    ms, _ = timed_steps(D, step, args.steps, args.warmup)      # 11.7 GB/step >> L2

    def plain(rec):
        return cam.reconstruct(disp)
    ms2, _ = timed_steps(D, plain, args.steps, args.warmup)

    # e2e on a bounded number of frames (pinned buffers for 64 4K frames = 12 GB)
    eb = 16
    d_pin = _ffi.pinned_empty((eb, H, W), np.float32)
    i_pin = _ffi.pinned_empty((eb, H, W, 3), np.uint8)
    d_pin[:] = disp[:eb].cpu().numpy()
    i_pin[:] = im[:eb].cpu().numpy()

    def e2e(rec):
        return cam.reconstruct_with_texture(d_pin, i_pin)
    ems, _ = timed_steps(D, e2e, max(2, args.steps // 4), 1)
    esteps = max(2, args.steps // 4)
    gbps = npx * 22.0 * args.steps / (ms * 1e-3) / 1e9
    res = {
        "metric": "disparity_to_colored_cloud", "unit": "Mpts/s", "value": npx * args.steps / (ms * 1e-3) / 1e6,
        "ms_per_step": ms / args.steps, "dtype": "f64 registers, f32 in/out",
        "config": {"workload": "C5 dense disparity -> coloured point cloud, %d frames of %dx%d per step"
                               % (batch, W, H), "l2": "inputs+outputs 11.7 GB per step >> L2"},
        "variants": {"uncoloured": {"Mpts_per_s": npx * args.steps / (ms2 * 1e-3) / 1e6,
                                    "GBps": npx * 16.0 * args.steps / (ms2 * 1e-3) / 1e9}},
        "e2e": {"value": eb * H * W * esteps / (ems * 1e-3) / 1e6, "unit": "Mpts/s",
                "h2d_bytes_per_step": eb * H * W * 7, "d2h_bytes_per_step": eb * H * W * 15,
                "bound": "PCIe: 22 B per point cross the host link (frame groups streamed on three streams)",
                "note": "%d frames per e2e step (pinned host buffers)" % eb},
        "gpu_launches": args.steps,
        "roofline": {"bound": "hbm", "kernel": "reproject_dense_kernel<float,true> (textured)",
                     "achieved": gbps, "peak": hbm, "unit": "GB/s", "frac": gbps / hbm,
                     "traffic": ncu_traffic("reproject_dense_kernel<float, 1>", scale=batch / 16.0),
                     "traffic_source": ncu_traffic_note("reproject_dense_kernel<float, 1>", scale=batch / 16.0),
                     "traffic_stale": ncu_traffic_stale("reproject.cu"),
                     "peak_source": src, "algorithmic": "4+3 B in, 12+3 B out = 22 B/pt"},
    }
    if D.rank == 0:
        res["cpu_baseline"] = cpu_baseline_c5(cam, disp[0].cpu().numpy(), im[0].cpu().numpy())
    return res


def cpu_baseline_c5(cam, disp, im):
    from oracle import port
    try:
        import cv2
        cv2.setNumThreads(os.cpu_count())
    except Exception:
        pass
    Q = cam.Q
    port.stereo_reconstruct_with_texture(disp, im, Q)
    t0 = time.perf_counter()
    n = 8
    for _ in range(n):
        port.stereo_reconstruct_with_texture(disp, im, Q)
    dt = time.perf_counter() - t0
    return {"value": n * disp.size / dt / 1e6, "unit": "Mpts/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "oracle.port.stereo_reconstruct_with_texture (cv2.reprojectImageTo3D + 2 np.copy, "
                      "camera_utils.py:987-989) on %d 4K frames, %.1f s" % (n, dt)}


# ------------------------------------------------- next rows
def run_next_rows(D, args):
    """DepthCamera.reconstruct on a batch of 4K depth frames and SceneFlow.process
    on one 4K frame (the callers either side of project / reconstruct)."""
    import torch
    from oracle import port
    from pybot_b200.vision.camera_utils import CameraIntrinsic, DepthCamera
    from pybot_b200.vision.optflow_utils import SceneFlow
    hbm, src = measured_peaks()
    H, W = syn.UHD_SHAPE
    K = np.array([[syn.UHD_FX, 0, syn.UHD_CX], [0, syn.UHD_FX, syn.UHD_CY], [0, 0, 1.0]])
    g = torch.Generator(device="cuda").manual_seed(1006)
    out = {}

    batch = 16
    depth = torch.rand((batch, H, W), device="cuda", generator=g) * 76 + 4
    dc = DepthCamera(K, shape=(H, W))

    def dstep(rec):
        return dc.reconstruct(depth)
    ms, _ = timed_steps(D, dstep, args.steps, args.warmup)              # 3.7 GB/step >> L2
    npx = batch * H * W
    d_np = depth[0].cpu().numpy()
    t0 = time.perf_counter()
    for _ in range(3):
        port.depth_reconstruct(K, (H, W), 1, d_np)
    cpu_dt = (time.perf_counter() - t0) / 3
    out["depth_camera_reconstruct"] = {
        "workload": "%d depth frames of %dx%d float32 -> organised float64 cloud per step" % (batch, W, H),
        "Mpts_per_s": npx * args.steps / (ms * 1e-3) / 1e6, "ms": ms / args.steps,
        "roofline": {"bound": "hbm", "kernel": "depth_reconstruct_dense_kernel<float>",
                     "achieved": npx * 28.0 * args.steps / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": npx * 28.0 * args.steps / (ms * 1e-3) / 1e9 / hbm, "peak_source": src,
                     "algorithmic": "4 B in + 24 B out = 28 B/pt (86 % writes)"},
        "cpu_baseline": {"value": H * W / cpu_dt / 1e6, "unit": "Mpts/s", "cores": 1, "kind": "port",
                         "sample": "oracle.port.depth_reconstruct (numpy, camera_utils.py:1032-1036) on one 4K frame"}}
    del depth

    Z = torch.rand((H, W), device="cuda", generator=g) * 56 + 4
    xs = (torch.arange(W, device="cuda", dtype=torch.float32) - syn.UHD_CX) / syn.UHD_FX
    ys = (torch.arange(H, device="cuda", dtype=torch.float32) - syn.UHD_CY) / syn.UHD_FX
    dX = torch.stack([xs[None, :] * Z + 0.35, ys[:, None] * Z - 0.05, Z - 0.8], dim=2).contiguous()
    sf = SceneFlow(CameraIntrinsic(K, shape=(H, W)))
    flush = make_flush(torch)

    def fstep(rec):
        return sf.process(dX)
    ms2, _ = timed_steps(D, fstep, args.steps, args.warmup, flush)
    Hk, Wk = syn.KITTI_SHAPE
    Kk = syn.kitti_K()
    dXk = dX[:Hk, :Wk].cpu().numpy()
    t0 = time.perf_counter()
    port.sceneflow_process(Kk, np.zeros(5), (Hk, Wk), dXk)
    cpu_dt = time.perf_counter() - t0
    out["sceneflow_process"] = {
        "workload": "one %dx%d frame of float32 scene points per step" % (W, H),
        "Mpts_per_s": H * W * args.steps / (ms2 * 1e-3) / 1e6, "ms": ms2 / args.steps,
        "algorithmic": "12 B in + 8 B flow out (+ 4 B winner map written and read) per pixel",
        "GBps": H * W * 28.0 * args.steps / (ms2 * 1e-3) / 1e9, "l2": L2_NOTE,
        "cpu_baseline": {"value": Hk * Wk / cpu_dt / 1e6, "unit": "Mpts/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": "oracle.port.sceneflow_process (cv2.projectPoints + numpy passes, "
                                   "optflow_utils.py:129-150) on one %dx%d frame" % (Wk, Hk)}}
    del dX, Z

    # CameraIntrinsic.reconstruct (undistort + ray * Z) on one 4K frame of pixels
    ci = CameraIntrinsic(K, D=np.float64([-0.3, 0.1, 0.001, -0.002, 0.01]), shape=(H, W))
    xyZ = torch.rand((H * W, 3), device="cuda", generator=g, dtype=torch.float32)
    xyZ[:, 0] *= W
    xyZ[:, 1] *= H
    xyZ[:, 2] = xyZ[:, 2] * 56 + 4

    def rstep(rec):
        return ci.reconstruct(xyZ)
    ms3, _ = timed_steps(D, rstep, args.steps, args.warmup, flush)
    sample = xyZ[:1_000_000].cpu().numpy()
    t0 = time.perf_counter()
    port.intrinsic_reconstruct(K, ci.D, sample)
    cpu_dt = time.perf_counter() - t0
    out["intrinsic_reconstruct"] = {
        "workload": "CameraIntrinsic.reconstruct (cv2.undistortPoints + ray * Z) of %d float32 [x, y, Z] rows" % (H * W),
        "Mpts_per_s": H * W * args.steps / (ms3 * 1e-3) / 1e6, "ms": ms3 / args.steps,
        "algorithmic": "12 B in + 24 B out per point", "GBps": H * W * 36.0 * args.steps / (ms3 * 1e-3) / 1e9,
        "l2": L2_NOTE,
        "cpu_baseline": {"value": len(sample) / cpu_dt / 1e6, "unit": "Mpts/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": "oracle.port.intrinsic_reconstruct (cv2.undistortPoints + numpy, "
                                   "camera_utils.py:499-538) on 1M rows"}}
    del xyZ

    # triangulate_points: 4M correspondences between two posed KITTI cameras
    from pybot_b200 import ops
    from pybot_b200.geometry import RigidTransform
    from pybot_b200.vision.camera_utils import Camera, CameraExtrinsic
    Kk = syn.kitti_K()
    mk = lambda *rp: Camera.from_intrinsics_extrinsics(                     # noqa: E731
        CameraIntrinsic(Kk, shape=syn.KITTI_SHAPE), CameraExtrinsic.from_rigid_transform(RigidTransform.from_rpyxyz(*rp)))
    c1, c2 = mk(0.01, -0.02, 0.03, 0.1, 0.0, 0.2), mk(-0.02, 0.05, 0.0, -0.4, 0.05, 0.9)
    nt = 4_000_000
    Xw = torch.rand((nt, 3), device="cuda", generator=g, dtype=torch.float64)
    Xw[:, 0] = Xw[:, 0] * 20 - 10
    Xw[:, 1] = Xw[:, 1] * 4 - 2
    Xw[:, 2] = Xw[:, 2] * 35 + 5
    p1 = c1.project(Xw) + torch.randn((nt, 2), device="cuda", generator=g, dtype=torch.float64) * 0.5
    p2 = c2.project(Xw) + torch.randn((nt, 2), device="cuda", generator=g, dtype=torch.float64) * 0.5
    P1, P2 = np.asarray(c1.P), np.asarray(c2.P)

    def tstep(rec):
        return ops.triangulate_points(P1, P2, p1, p2)
    ms4, _ = timed_steps(D, tstep, args.steps, args.warmup, flush)
    a, b = p1[:200_000].cpu().numpy(), p2[:200_000].cpu().numpy()
    t0 = time.perf_counter()
    port.triangulate_points(P1, P2, a, b)
    cpu_dt = time.perf_counter() - t0
    out["triangulate_points"] = {
        "workload": "%d float64 correspondences, two 3x4 cameras (per-point 4x4 Jacobi SVD)" % nt,
        "Mpts_per_s": nt * args.steps / (ms4 * 1e-3) / 1e6, "ms": ms4 / args.steps,
        "algorithmic": "32 B in + 24 B out per point; compute-bound (fp64 Jacobi sweeps)", "l2": L2_NOTE,
        "cpu_baseline": {"value": len(a) / cpu_dt / 1e6, "unit": "Mpts/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": "cv2.triangulatePoints + dehomogenisation (camera_utils.py:108-116) on 200k pairs"}}

    # knnMatch beyond the hot path: k = 8, with and without a per-pair mask (pbk_hamming_knnk)
    nqk, ntk, kk = 2000, 20_000, 8
    qk, tk = syn.orb_pair(seed=1007, nq=nqk, nt=ntk)
    mk8 = (np.random.default_rng(1007).random((nqk, ntk)) < 0.25).astype(np.uint8)
    qd, td, md = (torch.from_numpy(x).cuda() for x in (qk, tk, mk8))
    ms5, _ = timed_steps(D, lambda rec: ops.hamming_knnk(qd, td, kk), args.steps, args.warmup, flush)
    ms6, _ = timed_steps(D, lambda rec: ops.hamming_knnk(qd, td, kk, md), args.steps, args.warmup, flush)
    got = [x.cpu().numpy() for x in ops.hamming_knnk(qd, td, kk, md)]
    entry = {
        "workload": "BFMatcher.knnMatch(k=%d): %d queries x %d rows, without / with a 25 %% per-pair mask" % (kk, nqk, ntk),
        "Gpairs_per_s": nqk * ntk * args.steps / (ms5 * 1e-3) / 1e9, "ms": ms5 / args.steps,
        "masked_Gpairs_per_s": nqk * ntk * args.steps / (ms6 * 1e-3) / 1e9, "masked_ms": ms6 / args.steps,
        "algorithmic": "two passes of 8 POPC per pair (histogram -> threshold, then ordered collection); not the "
                       "hot path: k = 2 without a mask runs on the tensor-core scan", "l2": L2_NOTE}
    try:
        import cv2
        cv2.setNumThreads(os.cpu_count())
        t0 = time.perf_counter()
        res = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(qk, tk, k=kk, mask=mk8)
        cpu_dt = time.perf_counter() - t0
        ref_idx = np.array([[m.trainIdx for m in r] + [-1] * (kk - len(r)) for r in res], np.int64)
        ref_dist = np.array([[int(m.distance) for m in r] + [-1] * (kk - len(r)) for r in res], np.int32)
        entry["parity"] = {"bit_exact": bool(np.array_equal(got[0], ref_idx) and np.array_equal(got[1], ref_dist)),
                           "reference": "cv2.BFMatcher.knnMatch(q, t, k=%d, mask) on the same inputs" % kk}
        entry["cpu_baseline"] = {"value": nqk * ntk / cpu_dt / 1e9, "unit": "Gpairs/s", "cores": os.cpu_count(),
                                 "kind": "port", "sample": "the masked call above, whole workload, once"}
    except ImportError:
        entry["parity"] = {"bit_exact": None, "reference": "cv2 not importable on this box"}
    out["knnMatch_k8"] = entry
    return out


# ------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's CPU implementation of the headline path on the host
    cores: a pybot user's cv2.BFMatcher k=2 + ratio + mutual (oracle.port), a
    bounded sample of the C3 database per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    nq, per_kf = 2000, 2000
    sample_kf = 250
    q, db = syn.c3_database(nq=nq, n_keyframes=sample_kf, per_kf=per_kf)
    from oracle import port
    try:
        import cv2
        threads = os.cpu_count()
        cv2.setNumThreads(threads)
        engine = "cv2.BFMatcher(NORM_HAMMING).knnMatch"
    except Exception:
        threads, engine = 1, "numpy model"
    kfs = [db[i * per_kf:(i + 1) * per_kf] for i in range(sample_kf)]
    for _ in range(args.warmup):
        port.knn_ratio_mutual(q, kfs[:10], 0.75)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        port.knn_ratio_mutual(q, kfs, 0.75)
    dt = time.perf_counter() - t0
    v = nq * sample_kf * per_kf * args.steps / dt / 1e9
    sample = ("%s k=2 + ratio 0.75 + mutual, 2000 q x %d rows (%d of the 10000 keyframes) per step"
              % (engine, sample_kf * per_kf, sample_kf))
    return {"impl": "reference", "metric": "hamming_knn_k2_ratio_mutual", "value": v, "unit": "Gpairs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8/int (OpenCV batchDistance)", "data": "synthetic",
            "config": {"workload": "C3 loop-closure matching on the host CPU, bounded sample: " + sample},
            "cpu_baseline": {"value": v, "unit": "Gpairs/s", "cores": threads, "kind": "port",
                             "sample": sample},
            "e2e": {"value": v, "unit": "Gpairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


# --------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=["c3", "c2", "c4", "c5"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--keyframes", type=int, default=10_000)
    ap.add_argument("--parity-keyframes", type=int, default=10_000,
                    help="keyframes of the C3 database the cv2 baseline / parity reference scans (N=1)")
    ap.add_argument("--no-also", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        line = run_reference(args)
        if line is not None:
            print(json.dumps(line))
        return

    D = Dist(args.gpus)
    runners = {"c3": run_c3, "c2": run_c2, "c4": run_c4, "c5": run_c5}
    res = runners[args.workload](D, args)
    line = {"metric": res.pop("metric"), "value": res.pop("value"), "unit": res.pop("unit"),
            "n_gpus": D.world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res.pop("ms_per_step"), "higher_is_better": True,
            "scaling": res.pop("scaling", "weak"), "vs_baseline": None,
            "dtype": res.pop("dtype"), "data": "synthetic"}
    line.update(res)
    if not args.no_also and D.world == 1 and args.workload == "c3":
        also = {}
        sub = argparse.Namespace(**vars(args))
        sub.steps = min(args.steps, 10)
        for w in ("c2", "c4", "c5"):
            try:
                r = runners[w](D, sub)
                also[w] = {k: r[k] for k in ("metric", "value", "unit", "ms_per_step", "config", "e2e",
                                             "roofline", "cpu_baseline", "variants", "limiter", "parity",
                                             "result_digest") if k in r}
            except Exception as e:          # report, never hide
                also[w] = {"error": repr(e)}
            import torch
            torch.cuda.empty_cache()
        try:
            also["c1"] = run_c1(D, sub)
        except Exception as e:
            also["c1"] = {"error": repr(e)}
        try:
            also["next_rows"] = run_next_rows(D, sub)
        except Exception as e:
            also["next_rows"] = {"error": repr(e)}
        line["also"] = also
    if not args.no_also and D.world > 1 and args.workload == "c3":
        # the other sharded path of BASELINE.json (config 4) on the same ranks: strong scaling
        # (4 M observations split over the GPUs) and weak scaling (4 M observations per GPU)
        also = {}
        sub = argparse.Namespace(**vars(args))
        sub.steps = max(args.steps, 20)
        import torch
        for name, weak in (("c4_strong", False), ("c4_weak", True)):
            torch.cuda.empty_cache()
            try:
                r = run_c4(D, sub, weak=weak)
                also[name] = {k: r[k] for k in ("metric", "value", "unit", "ms_per_step", "scaling", "config", "e2e",
                                                "roofline", "variants", "limiter", "parity", "result_digest",
                                                "exchange") if k in r}
                also[name]["steps"] = sub.steps
            except Exception as e:          # report, never hide
                also[name] = {"error": repr(e)}
        line["also"] = also
    if D.rank == 0:
        print(json.dumps(line))
    D.close()


if __name__ == "__main__":
    main()
