#!/bin/bash
# A/B of library variants (scripts/build_variants.py):  scripts/ab_variants.sh "<which kernels>" <reps> name1 name2 ...
WHICH=$1; REPS=$2; shift 2
for v in "$@"; do
  echo "== variant $v"
  for w in $WHICH; do
    PROF_FLUSH=1 PYBOT_B200_LIB=$PWD/scripts/probes/_bin/libpbk_$v.so python scripts/prof_kernels.py $w $REPS 2>&1 | grep -v Warning
  done
done
