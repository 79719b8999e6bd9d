#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=1500 > gpurun_out/pytest_r2f.log 2>&1; tail -12 gpurun_out/pytest_r2f.log
out=gpurun_out/ab_r2f.log; : > $out
run() { echo "## $*" >> $out; timeout 300 env "$@" >> $out 2>&1; }
for frac in 0.70 0.90 0.97; do
  run python tools/run_fill.py 8192 $frac idw 10
  run AUVI_FILL_COOP=1 python tools/run_fill.py 8192 $frac idw 10
done
for st in 0 1; do
  run AUVI_UPSAMPLE_STRIDED=$st python tools/run_upsample.py 16384 f32 2x2,4x1,1x1,3x3,4x4,1x4
  run AUVI_UPSAMPLE_STRIDED=$st python tools/run_upsample.py 8192 f64 1x1,2x2,4x4,2x1
done
run python tools/run_upsample.py 16384 f32 2x2,4x1,1x1,4x4
cat $out
python bench.py --steps 200 --warmup 5 > gpurun_out/bench_r2f_n1.json 2> gpurun_out/bench_r2f_n1.err; tail -c 600 gpurun_out/bench_r2f_n1.err; head -c 1500 gpurun_out/bench_r2f_n1.json; echo
python tools/run_fill.py 8192 0.70 idw,bilinear 1 > gpurun_out/plain_f.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fill -s 2 -c 1 -f -o gpurun_out/prof_r2_fill_idw_final python tools/run_fill.py 8192 0.70 idw 1 > gpurun_out/ncu_f.log 2>&1
python tools/run_fill.py 8192 0.70 bilinear 1 > gpurun_out/plain_g.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:bilinear_fill -s 1 -c 1 -f -o gpurun_out/prof_r2_bilinear_fill python tools/run_fill.py 8192 0.70 bilinear 1 > gpurun_out/ncu_g.log 2>&1
python tools/run_fill.py 8192 0.01 bilinear 1 > gpurun_out/plain_h.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:bilinear_fill -s 1 -c 1 -f -o gpurun_out/prof_r2_bilinear_fill_001 python tools/run_fill.py 8192 0.01 bilinear 1 > gpurun_out/ncu_h.log 2>&1
ls -la gpurun_out/prof_r2_fill_idw_final.ncu-rep gpurun_out/prof_r2_bilinear_fill*.ncu-rep
