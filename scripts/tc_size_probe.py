"""Where the tensor-core scan overtakes the integer-pipe kernel (one GPU): 2000 queries against
databases of 4 k .. 256 k rows, expanded copy resident / expanded inside the call.
    python scripts/tc_size_probe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pybot_b200 import ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(1)
q = torch.randint(0, 256, (2000, 32), device="cuda", dtype=torch.uint8, generator=g)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


print("rows      popc us   tc resident us   tc + expansion us   (format %d)" % ops.hamming_tc_format())
for nt in (2000, 4096, 8192, 16384, 32768, 65536, 131072, 262144):
    t = torch.randint(0, 256, (nt, 32), device="cuda", dtype=torch.uint8, generator=g)
    signs = ops.hamming_tc_expand(t)
    a = timed(lambda: ops.hamming_knn2_keys(q, t, 0))
    b = timed(lambda: ops.hamming_knn2_keys_tc(q, signs, nt, 0))
    c = timed(lambda: ops.hamming_knn2_keys_tc(q, ops.hamming_tc_expand(t), nt, 0))
    ok = bool(torch.equal(ops.hamming_knn2_keys(q, t, 0), ops.hamming_knn2_keys_tc(q, signs, nt, 0)))
    print("%7d  %8.1f  %14.1f  %18.1f   equal %s" % (nt, a, b, c, ok), flush=True)
