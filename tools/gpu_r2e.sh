#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | head -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_r2e_n8.json 2> gpurun_out/bench_r2e_n8.err; tail -c 1200 gpurun_out/bench_r2e_n8.err; head -c 7000 gpurun_out/bench_r2e_n8.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_r2e_n8_ref.json; head -c 300 gpurun_out/bench_r2e_n8_ref.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 3 --no-extra > gpurun_out/bench_r2e_n4.json 2> gpurun_out/bench_r2e_n4.err; head -c 2500 gpurun_out/bench_r2e_n4.json; echo
AUVI_GPUS=8 python - <<'PY' > gpurun_out/multi8_points.log 2>&1
import os, sys, time
import numpy as np
sys.path.insert(0, "auv-real-time-interpolation_b200/python"); sys.path.insert(0, ".")
import auvi
from oracle import binding as ob
B = (-180.0, -160.0, 20.0, 30.0)
z = ob.synth_grid(3200, 4000, csv_round=False)
rng = np.random.RandomState(1); n = 5_000_000
pts = np.zeros((n, 3)); pts[:, 0] = rng.uniform(B[0], B[1], n); pts[:, 1] = rng.uniform(B[2], B[3], n)
g = auvi.Grid(z, *B)
for k in (1, 2, 4, 8):
    m = auvi.MultiGrid(z, *B, n_gpus=k, replicate=True)
    for name, meth in (("bilinear", 0), ("kriging", 2)):
        m.interp_points(meth, pts)
        t0 = time.perf_counter()
        for _ in range(3): out = m.interp_points(meth, pts)
        dt = (time.perf_counter() - t0) / 3
        ref = g.interp_points(meth, pts)
        print(f"auvi_multi_interp_points x{k} {name}: {dt*1e3:7.3f} ms  {n/dt/1e6:8.1f} Mpts/s  equal={np.array_equal(np.nan_to_num(out,nan=7), np.nan_to_num(ref,nan=7))}")
    m.close()
PY
cat gpurun_out/multi8_points.log
