#!/bin/bash
# One kernel-iteration round trip (under gpurun):  scripts/kernel_iter.sh <which> <kernel-regex> <tag> [pytest file]
#  parity test -> timing (20 reps, no profiler) -> one ncu --set full capture -> raw + source CSV
WHICH=$1; REGEX=$2; TAG=$3; TESTS=${4:-}
OUT=gpurun_out; mkdir -p $OUT
if [ -n "$TESTS" ]; then python -m pytest $TESTS -x -q -m gpu 2>&1 | tail -6; fi
python scripts/prof_kernels.py $WHICH 20
python scripts/prof_kernels.py $WHICH 1 > $OUT/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -c 2 -f -o $OUT/$TAG \
    python scripts/prof_kernels.py $WHICH 1 > $OUT/${TAG}_ncu.log 2>&1
ncu -i $OUT/$TAG.ncu-rep --page raw --csv > $OUT/${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/$TAG.ncu-rep --page source --csv > $OUT/${TAG}_src.csv 2>/dev/null
rm -f $OUT/$TAG.ncu-rep
