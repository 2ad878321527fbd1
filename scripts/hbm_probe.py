"""Measures read-only / write-only / copy HBM bandwidth at several sizes (CUDA events, L2-sized and larger)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybot_b200 import ops

sink = torch.zeros(4, dtype=torch.int32, device="cuda")
for mb in (200, 400, 1024, 4096):
    n = mb << 20
    a = torch.empty(n, dtype=torch.uint8, device="cuda").zero_()
    b = torch.empty(n, dtype=torch.uint8, device="cuda")
    for mode, name, bytes_moved in ((0, "read", n), (1, "write", n), (2, "copy", 2 * n)):
        best = 1e9
        for it in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.hbm_probe(mode, a, b, sink)
            e1.record()
            torch.cuda.synchronize()
            if it >= 2:
                best = min(best, e0.elapsed_time(e1))
        print("%5d MiB %-5s %8.1f us  %7.1f GB/s" % (mb, name, best * 1e3, bytes_moved / best / 1e6))
    # torch reference points
    for name, fn, bytes_moved in (("t.zero", lambda: b.zero_(), n), ("t.copy", lambda: b.copy_(a), 2 * n)):
        best = 1e9
        for it in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            if it >= 2:
                best = min(best, e0.elapsed_time(e1))
        print("%5d MiB %-6s %8.1f us  %7.1f GB/s" % (mb, name, best * 1e3, bytes_moved / best / 1e6))
    del a, b
