import os, sys, torch
sys.path.insert(0, os.getcwd())
from pybot_b200 import ops
g = torch.Generator(device="cuda").manual_seed(1)
q = torch.randint(0, 256, (2000, 32), device="cuda", dtype=torch.uint8, generator=g)
for tiles in (1, 2, 4, 8, 16, 32, 64):
    nt = 37 * 128 * tiles
    t = torch.randint(0, 256, (nt, 32), device="cuda", dtype=torch.uint8, generator=g)
    signs = ops.hamming_tc_expand(t)
    for _ in range(3):
        ops.hamming_knn2_keys_tc(q, signs, nt, 0)
    torch.cuda.synchronize()
