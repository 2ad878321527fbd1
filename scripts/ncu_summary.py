"""Turns the ncu outputs of scripts/profile_round.sh into the tracked summaries
under profiles/ (run here, no GPU needed):

    python scripts/ncu_summary.py <tag>

  profiles/<tag>_launches.csv          the launch list as captured (per-launch device time)
  profiles/<tag>_launches_summary.md   per-kernel count / total / share of the bench command
  profiles/<tag>_ncu_full_summary.md   one row per hot kernel from the --set full capture:
                                       duration, DRAM bytes, throughput %, pipe %, occupancy
  profiles/<tag>_bench_1gpu.json       the plain (un-profiled) bench line of the same command
"""
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

FULL_METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp_inst"),
]


PROBLEMS = {
    "hamming_knn2_kernel": "2000 queries x 2.5M rows (one 8-GPU shard of C3)",
    "hamming_tc_kernel": "2000 queries x 2.5M rows (one 8-GPU shard of C3), tcgen05 kind::i8 (format 8)",
    "hamming_tc4_kernel": "2000 queries x 2.5M rows (one 8-GPU shard of C3), tcgen05 kind::mxf4 (format 4, default)",
    "project_points_kernel": "C2: 10M float32 points",
    "project_compact_kernel": "C2: 10M float32 points, bounds+depth masks, 73.7 % valid",
    "transform_points_kernel": "C2: 10M float32 points -> float64",
    "ba_residual_jacobian_kernel": "C4: 4M observations",
    "reproject_dense_kernel": "16 frames of 3840x2160 (a quarter of a C5 step)",
    "depth_reconstruct_dense_kernel": "16 depth frames of 3840x2160 float32",
    "sceneflow_scatter_kernel": "one 3840x2160 frame of float32 scene points",
}


def short(name):
    name = name.replace("void ", "").replace("pbk::", "")
    cut = name.find("(")
    return name[:cut] if cut > 0 else name


def to_bytes(v, unit):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(v) * f


def to_us(v, unit):
    f = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)
    return float(v) * f


def launches(tag):
    src = os.path.join(OUT, tag + "_launches.csv")
    if not os.path.exists(src):
        return
    shutil.copy(src, os.path.join(PROF, tag + "_launches.csv"))
    rows = [r for r in csv.reader(open(src)) if len(r) >= 15 and r[0].isdigit()]
    agg = {}
    for r in rows:
        k = short(r[4])
        ns = float(r[14].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0, 1e30])
        a[0] += 1
        a[1] += ns
        a[2] = min(a[2], ns)
    total = sum(a[1] for a in agg.values())
    with open(os.path.join(PROF, tag + "_launches_summary.md"), "w") as f:
        f.write("# %s: ncu launch list of `python bench.py --steps 3 --warmup 3`\n\n" % tag)
        f.write("Per-launch times are cold-cache and serialised (ncu): compare SHARES, not absolutes.\n"
                "%d launches captured, %.2f ms of device time in total.\n\n" % (len(rows), total / 1e6))
        f.write("| kernel | launches | total us | min us | share |\n|---|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f | %.1f%% |\n" % (k[:90], a[0], a[1] / 1e3, a[2] / 1e3,
                                                                100 * a[1] / total))


def full(tag):
    src = os.path.join(OUT, tag + "_full_raw.csv")
    if not os.path.exists(src):
        return
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    seen = {}
    for r in rows[2:]:
        seen[short(r[col["Kernel Name"]])] = r          # keep the last (warm) launch of each kernel
    traffic = {}
    with open(os.path.join(PROF, tag + "_ncu_full_summary.md"), "w") as f:
        f.write("# %s: `ncu --set full --clock-control none` of scripts/prof_kernels.py all 1\n\n" % tag)
        f.write("One launch per hot kernel at its BASELINE.json size (hamming: one 8-GPU shard of C3 = "
                "2000 x 2.5M; reproject: 16 4K frames). Durations under ncu are not bench values.\n"
                "`traffic` = dram_rd + dram_wr per launch.\n\n")
        names = [m[1] for m in FULL_METRICS]
        f.write("| kernel | " + " | ".join(names) + " | traffic MB |\n|---|" + "---:|" * (len(names) + 1) + "\n")
        for k, r in seen.items():
            cells, rd, wr = [], 0.0, 0.0
            for m, n in FULL_METRICS:
                if m not in col:
                    cells.append("-")
                    continue
                v, u = r[col[m]].replace(",", ""), units[col[m]]
                try:
                    if n == "time":
                        cells.append("%.1f us" % to_us(v, u))
                    elif n in ("dram_rd", "dram_wr"):
                        b = to_bytes(v, u)
                        rd, wr = (b, wr) if n == "dram_rd" else (rd, b)
                        cells.append("%.1f MB" % (b / 1e6))
                    elif n.endswith("%"):
                        cells.append("%.1f" % float(v))
                    else:
                        cells.append("%d" % float(v))
                except ValueError:
                    cells.append(v)
            f.write("| `%s` | %s | %.1f |\n" % (k[:60], " | ".join(cells), (rd + wr) / 1e6))
            traffic[k] = {"dram_bytes": rd + wr, "problem": PROBLEMS.get(k.split("<")[0], "")}
    sha = {}
    sha_file = os.path.join(OUT, tag + "_source_sha.json")
    if os.path.exists(sha_file):
        sha = json.load(open(sha_file))
    json.dump({"tag": tag, "source_sha": sha, "kernels": traffic},
              open(os.path.join(PROF, "ncu_traffic.json"), "w"), indent=1)


def bench_line(tag):
    src = os.path.join(OUT, tag + "_bench_plain.json")
    if os.path.exists(src):
        try:
            d = json.loads(open(src).read().strip().splitlines()[-1])
            json.dump(d, open(os.path.join(PROF, tag + "_bench_1gpu.json"), "w"), indent=1)
        except Exception as e:
            print("bench line not parsed:", e)


if __name__ == "__main__":
    tag = sys.argv[1]
    os.makedirs(PROF, exist_ok=True)
    launches(tag)
    full(tag)
    bench_line(tag)
    print("wrote profiles/%s_*" % tag)
