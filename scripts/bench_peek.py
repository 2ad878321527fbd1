"""Short view of a bench.py JSON line: python scripts/bench_peek.py <file>"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.1f %s  ms/step %.3f  e2e %.1f  e2e_knnMatch %s" % (
    d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d.get("e2e_knnMatch", {}).get("value")))
print("parity", d.get("parity"))
print("digest", d.get("result_digest"))
r = d.get("roofline", {})
print("roofline frac %s frac_alg %s stale %s" % (r.get("frac"), r.get("frac_algorithmic"), r.get("traffic_stale")))
print("cpu_baseline", d.get("cpu_baseline"))
print("clocks", d.get("clocks"))
for k, v in d.get("also", {}).items():
    if "error" in v:
        print(k, "ERROR", v)
        continue
    print("--", k, "value", v.get("value"), v.get("unit"), "ms", v.get("ms_per_step"), "frac", v.get("roofline", {}).get("frac"),
          "full", v.get("roofline", {}).get("frac_full_step"))
    for kk in ("parity", "limiter", "variants", "result_digest"):
        if kk in v:
            print("    ", kk, v[kk])
    if "e2e" in v:
        print("     e2e", v["e2e"].get("value"))
