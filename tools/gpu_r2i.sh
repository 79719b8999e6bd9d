#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/ab_r2j.log; : > $out
run() { echo "## $*" >> $out; timeout 300 env "$@" >> $out 2>&1; }
L=auv-real-time-interpolation_b200/lib
for lib in libauvi.so libauvi_k2.so libauvi_k3.so libauvi_k4.so; do
  run AUVI_LIB=$L/$lib python tools/run_fill.py 8192 0.70 kriging 10
  run AUVI_LIB=$L/$lib python tools/run_fill.py 8192 0.90 kriging 10
  run AUVI_LIB=$L/$lib python tools/run_fill.py 8192 0.30 kriging 10
  run AUVI_LIB=$L/$lib python tools/run_fill.py 8192 0.50 kriging 10
done
cat $out | sed 's/ nan_left.*sha1/ sha1/'
