"""Lattice upsample of an n x n grid by (f_lat, f_lon): ms, Gcells/s and fraction of the measured HBM copy peak."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "auv-real-time-interpolation_b200", "python")); sys.path.insert(0, ROOT)
import torch, auvi
auvi.LIB_PATH = os.environ.get("AUVI_LIB", auvi.LIB_PATH)
if "AUVI_LIB" in os.environ:                                    # an older build: bind only what it exports
    import ctypes as _C; _l = _C.CDLL(auvi.LIB_PATH); auvi.SYMBOLS = {k: v for k, v in auvi.SYMBOLS.items() if hasattr(_l, k)}
n = int(sys.argv[1]); dt = sys.argv[2]; cases = [tuple(int(v) for v in c.split("x")) for c in sys.argv[3].split(",")]
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6455.9) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6455.9
tdt = torch.float32 if dt == "f32" else torch.float64
es = 4 if dt == "f32" else 8
z = torch.rand((n, n), dtype=tdt, device="cuda") * -5000.0
g = auvi.Grid(adopt=dict(ptr=z.data_ptr(), dtype=auvi.F32 if dt == "f32" else auvi.F64, n_lat=n, n_lon=n, ld=n, row0=0, rows=n, keep=z),
              min_lon=-180.0, max_lon=-160.0, min_lat=20.0, max_lat=30.0)
st = torch.cuda.current_stream().cuda_stream
for fl, fo in cases:
    rows, cols = fl * (n - 1) + 1, fo * (n - 1) + 1
    ld = (cols + 3) // 4 * 4
    out = torch.empty((rows, ld), dtype=tdt, device="cuda")
    for name, m in (("bilinear", auvi.BILINEAR), ("cubic", auvi.CUBIC)):
        fn = lambda: g.lattice_device(m, auvi.AXIS_EXPANDED, fl, fo, 0, 0, rows, out.data_ptr(), ld, None, st)
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        byt = rows * cols * (es + es / (fl * fo))
        print(f"{dt} n={n} f={fl}x{fo} {name:8s} {ms:8.3f} ms  {rows*cols/ms/1e6:8.1f} Gcells/s  {byt/ms/1e6/peak:5.2f} of HBM peak  tma={g.uses_tma}")
    del out
