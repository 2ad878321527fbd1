#!/bin/bash
# A/B timing of libpybot_b200.so variants built by scripts/build_variants.py (run under gpurun):
#   scripts/run_tc4_variants.sh [--more] NAME...    ("base" = the in-tree library)
mkdir -p gpurun_out
MORE=""
if [ "$1" == "--more" ]; then MORE="--more"; shift; fi
(for v in base "$@"; do
   echo "== $v"
   if [ "$v" == "base" ]; then unset PYBOT_B200_LIB; else export PYBOT_B200_LIB=scripts/probes/_bin/libpbk_$v.so; fi
   for n in 2500000 20000000; do
     timeout 60 python scripts/tc_probe.py $n --no-popc $MORE | grep -A12 mxf4 | grep -v "TMEM load"
   done
 done) > gpurun_out/tc4_variants.log 2>&1
cat gpurun_out/tc4_variants.log
