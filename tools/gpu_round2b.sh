#!/bin/bash
# Round-2 (second half) evidence run through gpurun: whole -m gpu suite, smoke(), ncu --set full of the small-factor bicubic launches.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=1500 > gpurun_out/pytest_verify.log 2>&1; tail -4 gpurun_out/pytest_verify.log
python __graft_entry__.py smoke 2>&1 | tail -1
for c in "8192 f64 2x2 f64_2x2" "16384 f32 2x2 f32_2x2" "16384 f32 4x1 f32_4x1"; do
  set -- $c
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:upsample_tiled_kernel --launch-skip 13 --launch-count 1 \
      -o gpurun_out/ncu_up_$4 -f python tools/run_upsample.py $1 $2 $3 > gpurun_out/ncu_up_$4.log 2>&1
  python tools/ncu_summary.py gpurun_out/ncu_up_$4.ncu-rep 1 > gpurun_out/ncu_up_$4.txt 2>&1
  cat gpurun_out/ncu_up_$4.txt
done
