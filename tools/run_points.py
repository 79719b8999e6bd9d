"""Point-list mode (what GridD::batch* calls): Grid A 4000 x 3200 f64 with n random points, device-resident kernel time per
method, and the end-to-end call with host buffers.  usage: run_points.py [n_points] [reps]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "auv-real-time-interpolation_b200", "python")); sys.path.insert(0, ROOT)
import torch, auvi
from oracle import binding as ob   # synthetic field generator only
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
B = (-180.0, -160.0, 20.0, 30.0)
z = ob.synth_grid(3200, 4000, csv_round=False)
g = auvi.Grid(z, *B)
rng = np.random.RandomState(1)
pts = np.zeros((n, 3)); pts[:, 0] = rng.uniform(B[0], B[1], n); pts[:, 1] = rng.uniform(B[2], B[3], n)
d_pts = torch.from_numpy(pts).cuda(); d_out = torch.empty(n, dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for name, m in (("bilinear", 0), ("cubic", 1), ("kriging", 2)):
    fn = lambda: g.interp_points_device(m, d_pts.data_ptr(), n, 24, d_out.data_ptr(), None, None, st)
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    k_ms = e0.elapsed_time(e1) / reps
    g.interp_points(m, pts)
    t0 = time.perf_counter()
    for _ in range(reps): g.interp_points(m, pts)
    e_ms = (time.perf_counter() - t0) / reps * 1e3
    print(f"{name:8s} n={n}: kernel {k_ms:7.3f} ms ({n/k_ms/1e3:8.1f} Mpts/s, {32*n/k_ms/1e6:6.1f} GB/s of 24+8 B/pt)   e2e {e_ms:7.3f} ms ({n/e_ms/1e3:7.1f} Mpts/s)")
