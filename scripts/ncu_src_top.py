"""Top stall sites from an ncu source-page CSV: python scripts/ncu_src_top.py <src.csv> [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
rows = rows[starts[-1]:]                      # last captured launch
print(rows[0][1][:100])
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[col['# Samples']]) for r in body)
stall_cols = [h for h in hdr if h.startswith('stall_')]
print("total samples", tot, "instructions", len(body))
agg = {h: sum(int(r[col[h]]) for r in body) for h in stall_cols}
print("by reason:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
idx = sorted(range(len(body)), key=lambda i: -int(body[i][col['# Samples']]))[:N]
for i in sorted(idx):
    r = body[i]
    reasons = {h[6:]: int(r[col[h]]) for h in stall_cols if int(r[col[h]]) > 0}
    top = sorted(reasons.items(), key=lambda kv: -kv[1])[:3]
    prev = body[i - 1][col['Source']].strip() if i else ''
    print("%4d %5.1f%% %-60s %s   [prev: %s]" % (i, 100.0 * int(r[col['# Samples']]) / tot, r[col['Source']].strip()[:60], top, prev[:40]))
