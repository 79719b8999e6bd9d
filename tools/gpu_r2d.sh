#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -q --timeout=1200 -k "multi or optin or drivers or points" > gpurun_out/pytest_r2d.log 2>&1; tail -25 gpurun_out/pytest_r2d.log
python tools/run_points.py 5000000 5 > gpurun_out/points_r2d.log 2>&1; cat gpurun_out/points_r2d.log
python tools/ref_gpu_compare.py 0 2>/dev/null | tail -1 > gpurun_out/refgpu_r2d.json; python -c "
import json; d=json.load(open('gpurun_out/refgpu_r2d.json'))
for c in ('grid_a_5M_random_points','mariana_50pct'):
    for m,v in d[c].items(): print(c, m, 'kernel ref %.4f ours %.4f (x%.2f)  e2e ref %.2f ours %.3f (x%.1f)' % (v['reference_gpu_kernel_ms'], v['ours_kernel_ms'], v['kernel_speedup'], v['reference_gpu_e2e_ms'], v['ours_e2e_ms'], v['e2e_speedup']))
"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r2d_n2.json 2> gpurun_out/bench_r2d_n2.err; tail -c 1500 gpurun_out/bench_r2d_n2.err; head -c 5000 gpurun_out/bench_r2d_n2.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_r2d_n2_ref.json 2>&1; tail -c 900 gpurun_out/bench_r2d_n2_ref.json; echo
