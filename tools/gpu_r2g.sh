#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=1500 > gpurun_out/pytest_r2g.log 2>&1; tail -6 gpurun_out/pytest_r2g.log
python __graft_entry__.py smoke 2>&1 | tail -2
python tools/run_fill.py 8192 0.70 idw,nn,cubic,kriging,bilinear 10 > gpurun_out/fill_r2g.log 2>&1; python tools/run_fill.py 8192 0.01 idw,bilinear 10 >> gpurun_out/fill_r2g.log 2>&1; python tools/run_fill.py 8192 0.90 idw,cubic,bilinear 10 >> gpurun_out/fill_r2g.log 2>&1; cat gpurun_out/fill_r2g.log
python bench.py --steps 200 --warmup 5 > gpurun_out/bench_r2g_n1.json 2> gpurun_out/bench_r2g_n1.err; tail -c 400 gpurun_out/bench_r2g_n1.err; head -c 700 gpurun_out/bench_r2g_n1.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2g_ref.json 2>/dev/null; head -c 300 gpurun_out/bench_r2g_ref.json; echo
# ncu: launch list of the bench command, then full captures of the two dominant kernels
python bench.py --steps 2 --warmup 1 --no-extra --no-cpu > gpurun_out/plain_l.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2g.csv python bench.py --steps 2 --warmup 1 --no-extra --no-cpu > gpurun_out/ncu_l.log 2>&1
python bench.py --steps 2 --warmup 1 --no-extra --no-cpu --no-e2e > gpurun_out/plain_m.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:upsample_tiled -s 1 -c 1 -f -o gpurun_out/prof_r2_upsample_cubic_f32 python bench.py --steps 2 --warmup 1 --no-extra --no-cpu --no-e2e > gpurun_out/ncu_m.log 2>&1
python tools/run_fill.py 8192 0.70 idw 1 > gpurun_out/plain_n.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fill_tiled -s 1 -c 1 -f -o gpurun_out/prof_r2_fill_idw_final python tools/run_fill.py 8192 0.70 idw 1 > gpurun_out/ncu_n.log 2>&1
ls -la gpurun_out/launches_r2g.csv gpurun_out/prof_r2_upsample_cubic_f32.ncu-rep gpurun_out/prof_r2_fill_idw_final.ncu-rep
