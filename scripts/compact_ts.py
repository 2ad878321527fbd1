"""Per-generation timeline of project_compact_early_kernel (diagnostic build, see compact_clocks.py)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["PYBOT_B200_LIB"] = os.environ.get("PYBOT_B200_LIB", os.path.join(ROOT, "scripts", "probes", "_bin", "libpbk_timeline.so"))
os.environ["PBK_COMPACT_IMPL"] = "3"
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
L = ctypes.CDLL(os.environ["PYBOT_B200_LIB"])
out = (ctypes.c_ulonglong * 512)()
for _ in range(3):
    pc.run(X)
torch.cuda.synchronize()
L.pbk_debug_compact_ts(out, 1)
pc.run(X); torch.cuda.synchronize()
L.pbk_debug_compact_ts(out, 1)
t = [[out[i * 8 + c] for c in range(8)] for i in range(64)]
t0 = min(x for r in t for x in r if x)
print("gen | cta0 publish | last publish | chain sees | chain stored | cta0 starts wait | cta0 sees | last sees   (us)")
for i, r in enumerate(t):
    if not r[5]:
        break
    f = lambda x: "%8.2f" % ((x - t0) / 1e3) if x else "       -"
    print("%3d | %s | %s | %s | %s | %s | %s | %s" % (i, f(r[0]), f(r[5]), f(r[1]), f(r[2]), f(r[4]), f(r[3]), f(r[6])))

if os.environ.get("EA_PER_CTA"):
    o2 = (ctypes.c_ulonglong * 2048)()
    L.pbk_debug_compact_cta(o2)
    rows = [(o2[b], o2[512 + b], o2[1024 + b], o2[1536 + b], b) for b in range(512) if o2[b]]
    m = min(r[0] for r in rows)
    rows.sort()
    print("publish(20) us | publish(21) | sees(20) | sm | cta   (sorted by publish(20); every 8th + last 24)")
    for k, r in enumerate(rows):
        if k % 8 == 0 or k >= len(rows) - 24:
            print("%7.2f %7.2f %7.2f  sm %3d  cta %3d" % ((r[0] - m) / 1e3, (r[1] - m) / 1e3, (r[2] - m) / 1e3, r[3], r[4]))
