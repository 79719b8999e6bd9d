#!/bin/bash
# Same-box A/B of two builds of libauvi.so through gpurun: ab/libauvi_base.so (an earlier commit's build, git-ignored) against
# the tree's own.  usage: tools/ab_builds.sh [fill] [upsample] [tests:<pytest -k expression>]
# Output: gpurun_out/ab_builds.txt
mkdir -p gpurun_out
O=gpurun_out/ab_builds.txt; : > $O
LIBS="${LIBS:-ab/libauvi_base.so auv-real-time-interpolation_b200/lib/libauvi.so}"
for what in "$@"; do
case "$what" in
tests:*)
  python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout=1500 -k "${what#tests:}" > gpurun_out/pytest_ab.log 2>&1; tail -3 gpurun_out/pytest_ab.log | tee -a $O ;;
fill)
  for rep in 1 2; do for lib in $LIBS; do
    echo "== $lib (pass $rep)" >> $O
    AUVI_LIB=$PWD/$lib python tools/run_fill.py 8192 0.70 idw,kriging,nn,cubic 20 >> $O 2>&1
    AUVI_LIB=$PWD/$lib python tools/run_fill.py 8192 0.50 idw 20 >> $O 2>&1
    AUVI_LIB=$PWD/$lib python tools/run_fill.py 8192 0.30 idw 20 >> $O 2>&1
    AUVI_LIB=$PWD/$lib python tools/run_fill.py 8192 0.90 idw,kriging 10 >> $O 2>&1
    AUVI_LIB=$PWD/$lib python tools/run_fill.py 8192 0.97 idw 10 >> $O 2>&1
  done; done ;;
upsample)
  for rep in 1 2; do for lib in $LIBS; do
    echo "== $lib upsample (pass $rep)" >> $O
    [ -z "$F64_ONLY" ] && AUVI_LIB=$PWD/$lib python tools/run_upsample.py 16384 f32 2x2,4x1,1x2,2x1,1x1,4x4 >> $O 2>&1
    AUVI_LIB=$PWD/$lib python tools/run_upsample.py 8192 f64 2x2,2x1,4x4 >> $O 2>&1
  done; done ;;
esac
done
cat $O
