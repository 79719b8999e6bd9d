#!/bin/bash
# Same-box A/B of two builds of libauvi.so (ab/libauvi_base.so = the previous commit's build): gap fill per method and mask
# fraction, the FP64 / small-factor upsamples.  Output: gpurun_out/ab_builds.txt
mkdir -p gpurun_out
O=gpurun_out/ab_builds.txt; : > $O
python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout=1500 -k "fill or lattice or config0 or points_match or tiny or gap" > gpurun_out/pytest_ab.log 2>&1; tail -3 gpurun_out/pytest_ab.log | tee -a $O
for rep in 1 2; do
for lib in ab/libauvi_base.so auv-real-time-interpolation_b200/lib/libauvi.so; do
  echo "== $lib (pass $rep)" >> $O
  AUVI_LIB=$PWD/$lib python tools/run_fill.py 8192 0.70 idw,kriging,nn,cubic 20 >> $O 2>&1
  AUVI_LIB=$PWD/$lib python tools/run_fill.py 8192 0.50 idw 20 >> $O 2>&1
  AUVI_LIB=$PWD/$lib python tools/run_fill.py 8192 0.30 idw 20 >> $O 2>&1
  AUVI_LIB=$PWD/$lib python tools/run_fill.py 8192 0.90 idw,kriging 10 >> $O 2>&1
  AUVI_LIB=$PWD/$lib python tools/run_fill.py 8192 0.97 idw 10 >> $O 2>&1
done; done
for lib in ab/libauvi_base.so auv-real-time-interpolation_b200/lib/libauvi.so; do
  echo "== $lib upsample" >> $O
  AUVI_LIB=$PWD/$lib python tools/run_upsample.py 8192 f64 2x2,2x1,4x4 >> $O 2>&1
done
cat $O
