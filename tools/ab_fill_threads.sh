#!/bin/bash
# Same-box A/B of fill_tiled_kernel builds (CTA size / dispatch variants; ab/*.so are git-ignored builds with -D switches).
mkdir -p gpurun_out
O=gpurun_out/ab_fill_threads.txt; : > $O
LIBS="${LIBS:-ab/libauvi_base.so auv-real-time-interpolation_b200/lib/libauvi.so ab/libauvi_t320.so ab/libauvi_t288.so}"
for lib in $LIBS; do
  [ "$lib" = ab/libauvi_base.so ] && continue
  AUVI_LIB=$PWD/$lib python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout=900 -k "fill or gap or tiny" 2>&1 | tail -1 | sed "s#^#$lib: #" >> $O
done
for rep in 1 2; do for lib in $LIBS; do
  echo "== $lib (pass $rep)" >> $O
  AUVI_LIB=$PWD/$lib python tools/run_fill.py 8192 0.70 idw,kriging,nn,cubic 20 >> $O 2>&1
  AUVI_LIB=$PWD/$lib python tools/run_fill.py 8192 0.30 idw 20 >> $O 2>&1
  AUVI_LIB=$PWD/$lib python tools/run_fill.py 8192 0.90 idw,kriging 10 >> $O 2>&1
  AUVI_LIB=$PWD/$lib python tools/run_fill.py 8192 0.97 idw 10 >> $O 2>&1
done; done
cat $O
