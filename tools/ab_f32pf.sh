mkdir -p gpurun_out
O=gpurun_out/ab_f32pf.txt; : > $O
for rep in 1 2; do
for lib in auv-real-time-interpolation_b200/lib/libauvi.so ab/libauvi_f32pf.so; do
  echo "== $lib (pass $rep)" >> $O
  AUVI_LIB=$PWD/$lib python tools/run_upsample.py 16384 f32 2x2,4x1,1x2,2x1,1x1,3x2 2>&1 | grep cubic >> $O
done; done
cat $O
