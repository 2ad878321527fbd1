"""One tensor-core scan of 2000 x N rows in the selected format (for ncu): python scripts/tc4_once.py [rows] [format]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pybot_b200 import ops  # noqa: E402

nt = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
if len(sys.argv) > 2:
    ops.hamming_tc_set_format(int(sys.argv[2]))
g = torch.Generator(device="cuda").manual_seed(1)
q = torch.randint(0, 256, (2000, 32), device="cuda", dtype=torch.uint8, generator=g)
t = torch.randint(0, 256, (nt, 32), device="cuda", dtype=torch.uint8, generator=g)
signs = ops.hamming_tc_expand(t)
for _ in range(3):
    k = ops.hamming_knn2_keys_tc(q, signs, nt, 0)
torch.cuda.synchronize()
print("ok", int(k[0, 0]) >> 32)
