"""The reference's own GPU code (kernels.cu + GridD.cu, unmodified, compiled for sm_100a: oracle/_ref_gpu) timed beside
libauvi on the same box and inputs: Grid A 5,000,000 random points and Mariana 50 % (BASELINE configs[1]).  Kernel only
(resident buffers, CUDA events) and end to end (what one GridD::batch* call costs its caller).  Prints one JSON line last;
bench.py runs this as a child process (the reference kernel printf()s from the device)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "auv-real-time-interpolation_b200", "python")); sys.path.insert(0, ROOT)
import torch, auvi
from oracle import binding as ob

BOUNDS = (-180.0, -160.0, 20.0, 30.0)
local = int(sys.argv[1]) if len(sys.argv) > 1 else 0
torch.cuda.set_device(local)
if not ob.ref_gpu_available():
    print(json.dumps({"unavailable": "oracle/_ref_gpu/libgridd_ref_sm100a.so not built (needs the reference checkout at build time)"}))
    sys.exit(0)
out = {"what": "reference kernels.cu:173-546 + GridD.cu:95-236 recompiled for sm_100a vs libauvi points_kernel / auvi_interp_points"}
rng = np.random.RandomState(1)
cases = []
za = ob.synth_grid(3200, 4000, csv_round=False)
pa = np.zeros((5_000_000, 3)); pa[:, 0] = rng.uniform(BOUNDS[0], BOUNDS[1], pa.shape[0]); pa[:, 1] = rng.uniform(BOUNDS[2], BOUNDS[3], pa.shape[0])
cases.append(("grid_a_5M_random_points", za, BOUNDS, pa))
mc = ob.masked_case("mariana", 0.5)
cases.append(("mariana_50pct", mc["z"], mc["bounds"], mc["pts"]))
stream = torch.cuda.current_stream().cuda_stream
# our GridD CLASS (host/GridD.cpp over libauvi) behind the same kind of C shim as the reference's: vector<Point> in and out
import ctypes as C
_dp = C.POINTER(C.c_double)
ours = C.CDLL(os.path.join(ROOT, "auv-real-time-interpolation_b200", "lib", "libgridd_c.so"))
ours.ourd_create.argtypes = [_dp, C.c_int, C.c_int] + [C.c_double] * 4; ours.ourd_create.restype = C.c_void_p
ours.ourd_destroy.argtypes = [C.c_void_p]
ours.ourd_batch.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int64, _dp]; ours.ourd_batch.restype = C.c_double
for tag, z, bounds, pts in cases:
    ref = ob.ReferenceGPU(z, *bounds)
    g = auvi.Grid(z, *bounds, device=local)
    zc = np.ascontiguousarray(z, dtype=np.float64)
    og = ours.ourd_create(zc.ctypes.data_as(_dp), zc.shape[0], zc.shape[1], *bounds)
    o_cls = np.empty(pts.shape[0])
    d_pts = torch.from_numpy(pts).cuda()
    d_out = torch.empty(pts.shape[0], dtype=torch.float64, device="cuda")
    tab = {}
    for name, meth in (("bilinear", 0), ("cubic", 1), ("kriging", 2)):
        ref.batch(meth, pts)
        r_ms = []
        for _ in range(3):
            r_out, ms = ref.batch(meth, pts)                         # end to end, as the drivers' chrono sees it
            r_ms.append(ms)
        k_out, k_ms = ref.kernel_ms(meth, pts, reps=5)
        g.interp_points(meth, pts)
        t0 = time.perf_counter()
        for _ in range(3):
            o_out = g.interp_points(meth, pts)
        o_ms = (time.perf_counter() - t0) / 3 * 1e3
        fn = lambda: g.interp_points_device(meth, d_pts.data_ptr(), pts.shape[0], 24, d_out.data_ptr(), None, None, stream)
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        ok_ms = e0.elapsed_time(e1) / 5
        ours.ourd_batch(og, meth, pts.ctypes.data_as(_dp), pts.shape[0], o_cls.ctypes.data_as(_dp))
        c_ms = float(np.mean([ours.ourd_batch(og, meth, pts.ctypes.data_as(_dp), pts.shape[0], o_cls.ctypes.data_as(_dp)) for _ in range(3)]))
        assert np.array_equal(np.nan_to_num(o_cls, nan=7.0), np.nan_to_num(o_out, nan=7.0))
        fin = ~np.isnan(r_out)
        tab[name] = {"reference_gpu_kernel_ms": k_ms, "ours_kernel_ms": ok_ms, "kernel_speedup": k_ms / ok_ms,
                     "reference_gpu_e2e_ms": float(np.mean(r_ms)), "ours_e2e_ms": o_ms, "e2e_speedup": float(np.mean(r_ms)) / o_ms,
                     "ours_gridd_class_e2e_ms": c_ms, "gridd_class_e2e_speedup": float(np.mean(r_ms)) / c_ms,
                     "same_nan_mask": bool(np.array_equal(np.isnan(r_out), np.isnan(o_out))),
                     "kernel_equals_batch": bool(np.array_equal(np.nan_to_num(k_out, nan=7.0), np.nan_to_num(r_out, nan=7.0))),
                     "max_abs_diff_m": float(np.max(np.abs(r_out[fin] - o_out[fin]))) if fin.any() else 0.0, "n": int(pts.shape[0])}
    out[tag] = tab
    g.close(); ref.close(); ours.ourd_destroy(og)
    del d_pts, d_out
sys.stdout.flush()
print("\n" + json.dumps(out))
