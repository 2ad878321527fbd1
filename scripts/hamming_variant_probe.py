"""Matcher inner-loop ceilings (probe modes 2/3/4) and the database-scan kernel per variant.

    PBK_HAMMING_VARIANT=<0|1|2> python scripts/hamming_variant_probe.py [nt]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pybot_b200 import ops  # noqa: E402

nt = int(sys.argv[1]) if len(sys.argv) > 1 else 2_500_000
sms = torch.cuda.get_device_properties(0).multi_processor_count
for mode, iters in ((2, 100), (3, 200), (4, 200), (5, 200)):
    best = 0.0
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        n = ops.popc_probe(mode=mode, ctas=sms * 4, iters=iters)
        b.record()
        torch.cuda.synchronize()
        best = max(best, n / (a.elapsed_time(b) * 1e-3))
    ppp = 8.0 if mode == 2 else 4.0
    print("probe mode %d: %.1f Gpairs/s" % (mode, best / ppp / 1e9))
g = torch.Generator(device="cuda").manual_seed(1)
q = torch.randint(0, 256, (2000, 32), device="cuda", dtype=torch.uint8, generator=g)
db = torch.randint(0, 256, (nt, 32), device="cuda", dtype=torch.uint8, generator=g)
ops.hamming_knn2_keys(q, db, 0)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    keys = ops.hamming_knn2_keys(q, db, 0)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ts.sort()
print("variant %s: scan %d rows %.3f ms  %.1f Gpairs/s   keysum %d" % (
    os.environ.get("PBK_HAMMING_VARIANT", "default"), nt, ts[2], 2000 * nt / ts[2] / 1e6,
    int(keys.view(torch.int64).sum().item())))
