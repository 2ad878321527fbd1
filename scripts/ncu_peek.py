"""Quick look at one ncu raw CSV: python scripts/ncu_peek.py <raw.csv> [kernel-substr]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
sub = sys.argv[2] if len(sys.argv) > 2 else ""
pat = re.compile(r"^(gpu__time_duration.sum|dram__bytes_(read|write).sum|gpu__dram_throughput.avg.pct|lts__throughput.avg.pct|"
                 r"l1tex__throughput.avg.pct_of_peak_sustained_active|l1tex__data_pipe_lsu_wavefronts(_mem_shared|_mem_lgds)?\.(avg|sum)$|"
                 r"sm__inst_executed_pipe_(fp64|alu|xu|lsu|fma|uniform|adu|cbu)\.avg\.pct_of_peak_sustained_active|"
                 r"smsp__issue_active.avg.pct|sm__warps_active.avg.pct|launch__registers|launch__grid_size|launch__occupancy_limit|"
                 r"smsp__inst_executed.sum$|smsp__average_warps?_issue_stalled_[a-z_]+_per_issue_active|smsp__average_warp_latency_issue_stalled|"
                 r"sm__cycles_elapsed.max|smsp__warps_eligible.avg.per_cycle_active|smsp__cycles_active.avg$|launch__waves)")
last = {}
for r in rows[2:]:
    n = r[hdr.index("Kernel Name")]
    if sub in n:
        last[n] = r
for n, r in last.items():
    print("==", n[:100])
    for i, h in enumerate(hdr):
        if pat.search(h):
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            if v != 0:
                print("   %-95s %14.2f %s" % (h[:95], v, units[i]))
