#!/bin/bash
# Where does the persistent fill kernel lose against the round-1 build?  Timing matrix + three ncu captures.
out=gpurun_out/fill_diag_${1:-x}.log
mkdir -p gpurun_out; : > $out
R01=auv-real-time-interpolation_b200/lib/libauvi_r01.so
run() { echo "## $*" >> $out; timeout 300 env "$@" >> $out 2>&1; }
for frac in 0.70 0.01; do
  run AUVI_LIB=$R01 python tools/run_fill.py 8192 $frac idw 20
  for cfg in 0 1 2 3; do
    run AUVI_FILL_CFG=$cfg python tools/run_fill.py 8192 $frac idw 20
    run AUVI_FILL_CFG=$cfg AUVI_FILL_ONE_TILE=1 python tools/run_fill.py 8192 $frac idw 20
  done
done
cat $out
export AUVI_FILL_CFG=1
python tools/run_fill.py 8192 0.70 idw 1 > gpurun_out/plain_a.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fill_tiled -s 1 -c 1 -f -o gpurun_out/prof_r2_cfg1_070 python tools/run_fill.py 8192 0.70 idw 1 > gpurun_out/ncu_a.log 2>&1
python tools/run_fill.py 8192 0.01 idw 1 > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fill_tiled -s 1 -c 1 -f -o gpurun_out/prof_r2_cfg1_001 python tools/run_fill.py 8192 0.01 idw 1 > gpurun_out/ncu_b.log 2>&1
unset AUVI_FILL_CFG
AUVI_LIB=$R01 python tools/run_fill.py 8192 0.01 idw 1 > gpurun_out/plain_c.log 2>&1 && AUVI_LIB=$R01 ncu --set full --clock-control none --import-source on -k regex:fill_tiled -s 1 -c 1 -f -o gpurun_out/prof_r01_001 python tools/run_fill.py 8192 0.01 idw 1 > gpurun_out/ncu_c.log 2>&1
tail -3 gpurun_out/ncu_a.log gpurun_out/ncu_b.log gpurun_out/ncu_c.log
ls -la gpurun_out/*.ncu-rep
