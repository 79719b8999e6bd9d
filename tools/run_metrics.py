"""Device error metrics of a gap fill (auvi_fill_metrics_device: MAE / RMSE / Max over the cells the masked grid lost) on an
n x n FP32 grid: ms per call (the call returns the numbers, so it is synchronous) and the fraction of the HBM copy peak on
3 x 4 bytes per cell."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "auv-real-time-interpolation_b200", "python")); sys.path.insert(0, ROOT)
import torch, auvi
n = int(sys.argv[1]); frac = float(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6455.9) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6455.9
dev = torch.device("cuda", 0)
i = torch.arange(n, device=dev, dtype=torch.float64) * (100.0 / (n - 1))
truth = (-(10.0 + 2.0 * i)[None, :] + 100.0 * torch.exp(-(((i - 75.0) ** 2)[None, :] + ((i - 50.0) ** 2)[:, None]) / 450.0)).float()
gen = torch.Generator(device=dev); gen.manual_seed(42)
z = truth.clone(); z[torch.rand((n, n), device=dev, generator=gen) < frac] = float("nan")
g = auvi.Grid(adopt=dict(ptr=z.data_ptr(), dtype=auvi.F32, n_lat=n, n_lon=n, ld=n, row0=0, rows=n, keep=z), min_lon=100.0, max_lon=110.0, min_lat=-10.0, max_lat=0.0)
out = torch.empty((n, n), dtype=torch.float32, device=dev)
st = torch.cuda.current_stream().cuda_stream
g.lattice_device(auvi.IDW, auvi.AXIS_NODES, 1, 1, 1, 0, n, out.data_ptr(), n, None, st)
torch.cuda.synchronize()
res = g.fill_metrics_device(out.data_ptr(), n, truth.data_ptr(), n, 0, n, st)
t0 = time.perf_counter()
for _ in range(reps): res = g.fill_metrics_device(out.data_ptr(), n, truth.data_ptr(), n, 0, n, st)
ms = (time.perf_counter() - t0) / reps * 1e3
print(f"fill metrics n={n} mask={frac:.2f}: {ms:8.3f} ms per call  {n*n*12/ms/1e6/peak:5.2f} of HBM peak (12 B/cell)  mae={res[0]:.6g} rmse={res[1]:.6g} max={res[2]:.6g} n_nan={res[3]} n={res[4]}")
res = auvi.error_metrics_device(truth.data_ptr(), out.data_ptr(), auvi.F32, n * n, st)
t0 = time.perf_counter()
for _ in range(reps): res = auvi.error_metrics_device(truth.data_ptr(), out.data_ptr(), auvi.F32, n * n, st)
ms = (time.perf_counter() - t0) / reps * 1e3
print(f"error metrics n={n}^2 elements: {ms:8.3f} ms per call  {n*n*8/ms/1e6/peak:5.2f} of HBM peak (8 B/element)  mae={res[0]:.6g} rmse={res[1]:.6g} max={res[2]:.6g} n_nan={res[3]}")
# the counter-hash mask (auvi_grid_mask_hash): writes NaN into `frac` of the cells, no reads
zz = truth.clone()
gm = auvi.Grid(adopt=dict(ptr=zz.data_ptr(), dtype=auvi.F32, n_lat=n, n_lon=n, ld=n, row0=0, rows=n, keep=zz), min_lon=100.0, max_lon=110.0, min_lat=-10.0, max_lat=0.0)
gm.mask_hash(frac, seed=42, count=False, stream=st); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): gm.mask_hash(frac, seed=42, count=False, stream=st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
import hashlib
print(f"mask_hash n={n} frac={frac:.2f}: {ms:8.3f} ms  {n*n/ms/1e6:8.1f} Gcells/s  masked={int(torch.isnan(zz).sum())}  sha1={hashlib.sha1(torch.isnan(zz).cpu().numpy().tobytes()).hexdigest()[:12] if n <= 8192 else '-'}")
