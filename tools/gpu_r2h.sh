#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/ab_r2h.log; : > $out
run() { echo "## $*" >> $out; timeout 300 env "$@" >> $out 2>&1; }
V=auv-real-time-interpolation_b200/lib/libauvi_bins.so; V2=auv-real-time-interpolation_b200/lib/libauvi_e2.so
for spec in "0.70 idw,nn,kriging" "0.90 idw,kriging" "0.97 idw" "0.50 idw"; do
  set -- $spec
  run python tools/run_fill.py 8192 $1 $2 10
  run AUVI_LIB=$V python tools/run_fill.py 8192 $1 $2 10
  run AUVI_LIB=$V2 python tools/run_fill.py 8192 $1 $2 10
done
cat $out
