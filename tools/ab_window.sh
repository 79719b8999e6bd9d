mkdir -p gpurun_out
O=gpurun_out/ab_window.txt; : > $O
for rep in 1 2; do
  echo "== window loads ON (pass $rep)" >> $O
  python tools/run_upsample.py 16384 f32 2x2,4x1,1x2,2x1,4x2,1x1 >> $O 2>&1
  echo "== AUVI_NO_WINDOW=1 (pass $rep)" >> $O
  AUVI_NO_WINDOW=1 python tools/run_upsample.py 16384 f32 2x2,4x1,1x2,2x1,4x2,1x1 >> $O 2>&1
done
grep -E "cubic|==" $O
