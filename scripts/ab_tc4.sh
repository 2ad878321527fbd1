#!/bin/bash
# A/B of libpybot_b200.so variants (scripts/build_variants.py) on the short-scan and the C3 scan:
#   scripts/ab_tc4.sh NAME...      ("base" = the in-tree library is always run first)
for v in base "$@"; do
  echo "== $v"
  if [ "$v" == "base" ]; then unset PYBOT_B200_LIB; else export PYBOT_B200_LIB=scripts/probes/_bin/libpbk_$v.so; fi
  timeout 100 python scripts/tc_tiles_probe.py | tail -4
  timeout 60 python scripts/tc_probe.py 20000000 --no-popc | grep -A2 mxf4
done
