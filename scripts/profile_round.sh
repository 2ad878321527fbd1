#!/bin/bash
# Profiling pass of one round (run under gpurun, one GPU):
#   scripts/profile_round.sh <tag>
# 1. bench.py plain (must exit 0), then its ncu launch list (per-launch device time)
# 2. scripts/prof_kernels.py plain, then one `ncu --set full` capture of every hot kernel
# Outputs land in gpurun_out/<tag>_*; scripts/ncu_summary.py turns them into profiles/.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
# the kernel sources this capture was made from (bench.py marks traffic figures of another
# version of a kernel as stale)
python - <<PY > $OUT/${TAG}_source_sha.json
import hashlib, json, os
d = "pybot_b200/csrc"
print(json.dumps({f: hashlib.sha256(open(os.path.join(d, f), "rb").read()).hexdigest()[:16]
                  for f in sorted(os.listdir(d)) if f.endswith((".cu", ".cuh"))}))
PY
BENCH="python bench.py --steps 3 --warmup 3"
$BENCH > $OUT/${TAG}_bench_plain.json 2> $OUT/${TAG}_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_bench_under_ncu.log 2>&1
echo "launch list rc=$?"
PROF="python scripts/prof_kernels.py all 1"
$PROF > $OUT/${TAG}_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:'hamming_knn2_kernel|hamming_tc4?_kernel|project_points|project_compact|transform_points|ba_residual|reproject_dense|depth_reconstruct_dense|sceneflow_scatter' \
    -c 24 -f -o $OUT/${TAG}_full $PROF > $OUT/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
ncu -i $OUT/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_full_raw.csv 2>/dev/null
tail -5 $OUT/${TAG}_prof_plain.log
