#!/bin/bash
# Same-box A/B of the gap-fill kernel: r01 library (if present), round-1 CTA shape (AUVI_FILL_CFG=1) and the default.
# usage: tools/ab_fill.sh <tag>   -> gpurun_out/fill_ab_<tag>.log
tag=${1:-x}
out=gpurun_out/fill_ab_${tag}.log
mkdir -p gpurun_out
: > $out
R01=auv-real-time-interpolation_b200/lib/libauvi_r01.so
run() { echo "## $*" >> $out; timeout 300 env "$@" >> $out 2>&1; }
for spec in "0.70 idw,nn,cubic,kriging,bilinear" "0.90 idw,kriging" "0.97 idw" "0.50 idw" "0.30 idw" "0.01 idw,bilinear"; do
  set -- $spec
  if [ -f $R01 ]; then run AUVI_LIB=$R01 python tools/run_fill.py 8192 $1 $2 10; fi
  run AUVI_FILL_CFG=1 python tools/run_fill.py 8192 $1 $2 10
  run AUVI_FILL_CFG=0 python tools/run_fill.py 8192 $1 $2 10
done
if [ -f $R01 ]; then run AUVI_LIB=$R01 python tools/run_fill.py 65536 0.70 idw 3; fi
run AUVI_FILL_CFG=0 python tools/run_fill.py 65536 0.70 idw,bilinear 3
tail -n 80 $out
