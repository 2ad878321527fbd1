"""profiles/<tag>_scaling.md from the committed bench lines (profiles/<tag>_bench_{1gpu_default,2gpu,4gpu,8gpu}.json):
    python scripts/scaling_table.py r02"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"


def line(name):
    p = os.path.join(ROOT, "profiles", "%s_bench_%s.json" % (tag, name))
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return json.loads(f.read().strip().splitlines()[-1])


runs = [(n, line(f)) for n, f in ((1, "1gpu_default"), (2, "2gpu"), (4, "4gpu"), (8, "8gpu"))]
runs = [(n, d) for n, d in runs if d is not None]
base = runs[0][1]
out = ["# %s: sharded paths at %s B200 (`bench.py --gpus N`)" % (tag, " / ".join(str(n) for n, _ in runs)), "",
       "C3 = 2000 ORB queries vs the 20 M-descriptor keyframe database, k=2 + ratio + mutual, database sharded by keyframe "
       "range; the shard scan runs on the FP4 tensor path (`hamming_tc4_kernel`), both exchanges fused into the kernels over "
       "NVLink peer memory.  Efficiency = value(N) / (N x value(1)): strong scaling of a 2.7 ms step.", "",
       "The weak columns (`variants.weak_20M_rows_per_gpu`, lines that carry it): every rank holds a whole 20 M-row copy of the "
       "database as its own keyframe range (N x 20 M rows in total), same kernels and fused exchanges; "
       "efficiency = weak value / (N x value(1)).", "",
       "| N | Gpairs/s | ms/step | shard scan ms | e2e Gpairs/s | e2e_knnMatch | efficiency vs N=1 | weak Gpairs/s | weak ms | weak efficiency | parity.bit_exact | result_digest |",
       "|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|---|"]
for n, d in runs:
    w = d.get("variants", {}).get("weak_20M_rows_per_gpu")
    if n == 1:
        weak = "%.0f | %.3f | 1.000" % (d["value"], d["ms_per_step"])
    elif w and "Gpairs_per_s" in w:
        weak = "%.0f | %.3f | %.3f" % (w["Gpairs_per_s"], w["ms"], w["Gpairs_per_s"] / (n * base["value"]))
    else:
        weak = "- | - | -"
    out.append("| %d | %.0f | %.3f | %.3f | %.0f | %.0f | %.3f | %s | %s | `%s...` |" % (
        n, d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["e2e_knnMatch"]["value"],
        d["value"] / (n * base["value"]), weak, d["parity"]["bit_exact"], d["result_digest"][:16]))
out += ["", "C4 = BA residual + Jacobians, 1000 cameras, 1 M points; strong: 4 M observations split over the ranks; weak: 4 M "
        "observations per rank.", "",
        "| N | strong Mobs/s | strong us/step | weak Mobs/s | weak us/step | weak efficiency | parity |", "|---:|---:|---:|---:|---:|---:|---|"]
c4_1 = base["also"]["c4"]
for n, d in runs:
    if n == 1:
        s = w = c4_1
    else:
        s, w = d["also"]["c4_strong"], d["also"]["c4_weak"]
    out.append("| %d | %.0f | %.1f | %.0f | %.1f | %.3f | %s / %s |" % (
        n, s["value"], s["ms_per_step"] * 1e3, w["value"], w["ms_per_step"] * 1e3, w["value"] / (n * c4_1["value"]),
        (s["parity"].get("bit_exact") or s["parity"].get("within_tolerance")), (w["parity"].get("bit_exact") or w["parity"].get("within_tolerance"))))
out += ["", "Strong scaling of both paths is bounded by fixed per-launch costs, not by bandwidth: a 2.5 M-row shard scan carries "
        "~15 us of launch / prologue / tail and the exact-path-heavy start of every split (DESIGN 4.1); one BA evaluation carries "
        "~13 us (`scripts/ba_size_probe.py`) plus one NVLink crossing."]
p = os.path.join(ROOT, "profiles", "%s_scaling.md" % tag)
with open(p, "w") as f:
    f.write("\n".join(out) + "\n")
print("\n".join(out))
