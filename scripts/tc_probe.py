"""Timing of the tensor-core Hamming scan against the integer-pipe kernel (one GPU):
    python scripts/tc_probe.py [rows]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pybot_b200 import ops  # noqa: E402

nt = int(sys.argv[1]) if len(sys.argv) > 1 else 2_500_000
g = torch.Generator(device="cuda").manual_seed(1)
q = torch.randint(0, 256, (2000, 32), device="cuda", dtype=torch.uint8, generator=g)
t = torch.randint(0, 256, (nt, 32), device="cuda", dtype=torch.uint8, generator=g)
torch.cuda.synchronize()


def timed(name, fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print("%-22s %8.3f ms median  %8.3f min   %8.1f Gpairs/s" % (name, ts[len(ts) // 2], ts[0], 2000 * nt / ts[0] / 1e6), flush=True)


ref = ops.hamming_knn2_keys(q, t, 0)
if "--no-popc" not in sys.argv:
    timed("popc", lambda: ops.hamming_knn2_keys(q, t, 0))
for fmt in (8, 4):
    ops.hamming_tc_set_format(fmt)
    signs = ops.hamming_tc_expand(t)
    torch.cuda.synchronize()
    name = "tcgen05 %s" % ("kind::i8" if fmt == 8 else "kind::mxf4")
    timed(name, lambda: ops.hamming_knn2_keys_tc(q, signs, nt, 0))
    got = ops.hamming_knn2_keys_tc(q, signs, nt, 0)
    print("  equal to the integer-pipe keys:", bool(torch.equal(ref, got)), flush=True)
    # timing probes (results are wrong): 1 = no selection, 2 = no MMA
    modes = ((1, "no selection"), (2, "no MMA"))
    if fmt == 4 and "--more" in sys.argv:
        modes += ((3, "no MMA, no selection"), (7, "barrier handshake only"), (5, "MMA + handshake only"),
                  (15, "handshake, no tile loads"), (13, "MMA + handshake, no loads"), (10, "selection, no MMA, no loads"), (8, "all but tile loads"))
    for dbg, what in modes:
        os.environ["PBK_TC_DEBUG"] = str(dbg)
        timed("  %s" % what, lambda: ops.hamming_knn2_keys_tc(q, signs, nt, 0), reps=3)
    del os.environ["PBK_TC_DEBUG"]
    del signs
