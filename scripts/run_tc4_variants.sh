mkdir -p gpurun_out
(echo == base; timeout 60 python scripts/tc_probe.py 20000000 --no-popc --more | grep -A9 mxf4 | grep -v "equal\|TMEM load\|no MMA  "
for v in "$@"; do echo == $v; PYBOT_B200_LIB=scripts/probes/_bin/libpbk_$v.so timeout 60 python scripts/tc_probe.py 20000000 --no-popc --more | grep -A9 mxf4 | grep -v "equal\|TMEM load"; done) > gpurun_out/tc4_variants.log 2>&1
cat gpurun_out/tc4_variants.log
