#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=1200 > gpurun_out/pytest_r2c.log 2>&1; tail -15 gpurun_out/pytest_r2c.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2c_n1.json 2> gpurun_out/bench_r2c_n1.err; tail -c 3000 gpurun_out/bench_r2c_n1.err; head -c 6000 gpurun_out/bench_r2c_n1.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2c_ref.json 2>&1; head -c 600 gpurun_out/bench_r2c_ref.json; echo
python tools/run_fill.py 8192 0.70 idw,nn,cubic,kriging,bilinear 10 > gpurun_out/fill_r2c.log 2>&1; python tools/run_fill.py 8192 0.01 idw,bilinear 10 >> gpurun_out/fill_r2c.log 2>&1; cat gpurun_out/fill_r2c.log
python tools/run_points.py 5000000 5 > gpurun_out/points_r2c.log 2>&1; cat gpurun_out/points_r2c.log
python tools/run_upsample.py 8192 f64 2x2,4x4,1x1 > gpurun_out/ups_r2c.log 2>&1; python tools/run_upsample.py 16384 f32 2x2,4x1,4x4 >> gpurun_out/ups_r2c.log 2>&1; cat gpurun_out/ups_r2c.log
# ncu: points kernels, the f64 / small-factor upsample launches
python tools/run_points.py 5000000 1 > gpurun_out/plain_p.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:points_kernel -s 3 -c 3 -f -o gpurun_out/prof_r2_points python tools/run_points.py 5000000 1 > gpurun_out/ncu_p.log 2>&1
python tools/run_upsample.py 8192 f64 2x2 > gpurun_out/plain_u.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:upsample_tiled -s 12 -c 2 -f -o gpurun_out/prof_r2_ups_f64_2x2 python tools/run_upsample.py 8192 f64 2x2 > gpurun_out/ncu_u.log 2>&1
python tools/run_upsample.py 16384 f32 2x2 > gpurun_out/plain_v.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:upsample_tiled -s 12 -c 2 -f -o gpurun_out/prof_r2_ups_f32_2x2 python tools/run_upsample.py 16384 f32 2x2 > gpurun_out/ncu_v.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
