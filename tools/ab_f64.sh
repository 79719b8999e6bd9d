#!/bin/bash
# Same-box A/B of FP64 bicubic variants (ab/*.so are git-ignored builds of csrc/upsample.cu with -D switches).
mkdir -p gpurun_out
O=gpurun_out/ab_f64.txt; : > $O
LIBS="${LIBS:-auv-real-time-interpolation_b200/lib/libauvi.so ab/libauvi_s5.so}"
for lib in $LIBS; do
  AUVI_LIB=$PWD/$lib python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout=900 -k "lattice or config0 or slab" 2>&1 | tail -2 | sed "s#^#$lib: #" >> $O
done
for rep in 1 2; do
for lib in $LIBS; do
  echo "== $lib (pass $rep)" >> $O
  AUVI_LIB=$PWD/$lib python tools/run_upsample.py 8192 f64 2x2,2x1,4x4,3x3,1x2 2>&1 | grep cubic >> $O
done; done
cat $O
