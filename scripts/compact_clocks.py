"""Phase clocks of project_compact_kernel.  Needs a diagnostic build of the library:
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
         -cudart static -DPBK_COMPACT_TIMING -I include -shared -o build/libpbk_timing.so pybot_b200/csrc/*.cu
"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["PYBOT_B200_LIB"] = os.environ.get("PYBOT_B200_LIB", os.path.join(ROOT, "scripts", "probes", "_bin", "libpbk_timing.so"))
sys.path.insert(0, ROOT)
import torch
from pybot_b200 import _ffi, ops, synthetic as syn
from pybot_b200.geometry import RigidTransform

n = 10_000_000
g = torch.Generator(device="cuda").manual_seed(1)
X = torch.rand((n, 3), device="cuda", generator=g)
X[:, 0] = X[:, 0] * 80 - 40; X[:, 1] = X[:, 1] * 6 - 3; X[:, 2] = X[:, 2] * 78 + 2
T = RigidTransform.from_rpyxyz(*syn.C2_RPYXYZ).to_matrix()
pc = ops.Projector(T[:3, :3], T[:3, 3], syn.kitti_K(), shape=syn.KITTI_SHAPE, check_bounds=True,
                   check_depth=True, return_depth=True, return_valid=True)
L = _ffi.lib()
out = (ctypes.c_ulonglong * 16)()
pc.run(X); torch.cuda.synchronize()
L.pbk_debug_compact_clocks(out, 1)
reps = 10
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    pc.run(X)
b.record(); torch.cuda.synchronize()
L.pbk_debug_compact_clocks(out, 1)
ms = a.elapsed_time(b) / reps
ntiles = (n + 1023) // 1024
names = ["-", "-", "scan:cnt_ready wait", "scan:scan+lookback",
         "cmp:in_full wait", "cmp:read+compute", "cmp:base_ready wait", "cmp:scatter"]
print("%.3f ms/iter, %d tiles" % (ms, ntiles))
if os.environ.get("PBK_COMPACT_IMPL") == "3":
    # per 1024-point tile: resolver phases once, compute-warp phases summed over the 8 warps
    for i, nm in ((8, "chain:load+scan /gen"), (9, "chain:carry wait /gen"), (10, "chain:store /gen"),
                  (15, "cmp:input wait (x8)"), (11, "cmp:classify (x8)"), (12, "cmp:read+project (x8)"),
                  (13, "cmp:offset wait (x8)"), (14, "cmp:scatter (x8)")):
        per = (ntiles + 294) // 295 if i <= 10 else ntiles
        print("  %-26s %8.0f clk" % (nm, out[i] / reps / per))
    sys.exit(0)
for i, nm in enumerate(names):
    print("  %-22s %8.0f clk/tile" % (nm, out[i] / reps / ntiles))

