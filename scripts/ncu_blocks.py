"""Executed warp instructions per basic block from an ncu source-page CSV:
    python scripts/ncu_blocks.py <src.csv> <kernel-substring> <units> [min-per-unit]
`units` = what to normalise by (e.g. the number of 128-point tiles of the launch)."""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
sub, units = sys.argv[2], float(sys.argv[3])
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 5.0
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
ks = [k for k in range(len(starts) - 1) if sub in rows[starts[k]][1]]
k = ks[-1]
hdr = rows[starts[k] + 1]; col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[starts[k] + 2:starts[k + 1]] if len(r) == len(hdr)]
tot = sum(int(r[col['Instructions Executed']]) for r in body)
print(rows[starts[k]][1][:70], "| %.1f instructions per unit" % (tot / units))
ops = collections.Counter()
for r in body:
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[col['Source']].strip())
    ops[m.group(2) if m else '?'] += int(r[col['Instructions Executed']])
print("  ", [(o, round(n / units, 1)) for o, n in ops.most_common(30)])
blk = 0; b0 = 0
for i, r in enumerate(body):
    blk += int(r[col['Instructions Executed']])
    src = r[col['Source']].strip()
    if re.search(r'\b(BRA|EXIT|RET|CALL|BSYNC|WARPSYNC|SYNCS)\b', src) or i == len(body) - 1:
        if blk / units >= thr:
            n = [int(x[col['Instructions Executed']]) / units for x in body[b0:i + 1]]
            print("%5d-%5d %7.1f /unit (%d lines, max line %.1f) samples=%d ends: %s" % (
                b0, i, blk / units, i + 1 - b0, max(n), sum(int(x[col['# Samples']]) for x in body[b0:i + 1]), src[:50]))
        blk = 0; b0 = i + 1
