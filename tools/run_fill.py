"""Gap-fill of an n x n FP32 grid at a given mask fraction (BASELINE config 5 scaled down); prints ms per method."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "auv-real-time-interpolation_b200", "python")); sys.path.insert(0, ROOT)
import torch, auvi
auvi.LIB_PATH = os.environ.get("AUVI_LIB", auvi.LIB_PATH)
if "AUVI_LIB" in os.environ:                                    # an older build: bind only what it exports
    import ctypes as _C; _l = _C.CDLL(auvi.LIB_PATH); auvi.SYMBOLS = {k: v for k, v in auvi.SYMBOLS.items() if hasattr(_l, k)}        # A/B runs of two builds on one box
n = int(sys.argv[1]); frac = float(sys.argv[2]); methods = sys.argv[3].split(","); reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = torch.device("cuda", 0)
i = torch.arange(n, device=dev, dtype=torch.float64) * (100.0 / (n - 1))
z = (-(10.0 + 2.0 * i)[None, :] + 100.0 * torch.exp(-(((i - 75.0) ** 2)[None, :] + ((i - 50.0) ** 2)[:, None]) / 450.0)).float()
gen = torch.Generator(device=dev); gen.manual_seed(42)
z[torch.rand((n, n), device=dev, generator=gen) < frac] = float("nan")
g = auvi.Grid(adopt=dict(ptr=z.data_ptr(), dtype=auvi.F32, n_lat=n, n_lon=n, ld=n, row0=0, rows=n, keep=z), min_lon=100.0, max_lon=110.0, min_lat=-10.0, max_lat=0.0)
out = torch.empty((n, n), dtype=torch.float32, device=dev)
st = torch.cuda.current_stream().cuda_stream
M = dict(bilinear=0, cubic=1, kriging=2, nn=3, idw=4)
for name in methods:
    fn = lambda: g.lattice_device(M[name], auvi.AXIS_NODES, 1, 1, 1, 0, n, out.data_ptr(), n, None, st)
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    import hashlib
    digest = hashlib.sha1(out.cpu().numpy().tobytes()).hexdigest()[:12] if n <= 8192 else "-"
    print(f"{name:8s} n={n} mask={frac:.2f}: {ms:8.3f} ms  {n*n/ms/1e3:10.1f} Mcells/s  nan_left={int(torch.isnan(out).sum())}  sha1={digest}")
