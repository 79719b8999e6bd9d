"""Where does the end-to-end (host buffer) lattice call spend its time?  Run on a GPU box."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "auv-real-time-interpolation_b200", "python")); sys.path.insert(0, ROOT)
import torch, auvi
lib = auvi.load()
n, f = 16384, 4
rows = cols = f * (n - 1) + 1
bounds = (-180.0, -160.0, 20.0, 30.0)
h_z = torch.rand((n, n), dtype=torch.float32).pin_memory()
h_out = torch.empty((rows, cols), dtype=torch.float32, pin_memory=True)
def t(fn, reps=2):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
h = C.c_void_p()
def create():
    global h
    assert lib.auvi_grid_create(h_z.data_ptr(), auvi.F32, n, n, *bounds, 0, C.byref(h)) == 0
def destroy(): lib.auvi_grid_destroy(h)
def create_destroy(): create(); destroy()
print("grid create+destroy %.1f ms" % t(create_destroy))
create()
def lat(ptr, r0=0, r1=rows): assert lib.auvi_lattice(h, auvi.CUBIC, auvi.AXIS_EXPANDED, f, f, 0, r0, r1, ptr) == 0, lib.auvi_last_error()
print("auvi_lattice -> torch pinned      %.1f ms  (%.1f GB/s)" % ((ms := t(lambda: lat(h_out.data_ptr()))), rows * cols * 4 / ms / 1e6))
d = torch.empty((rows, 65536), dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
print("device-only lattice               %.2f ms" % t(lambda: lib.auvi_lattice_device(h, auvi.CUBIC, auvi.AXIS_EXPANDED, f, f, 0, 0, rows, d.data_ptr(), 65536, None, st)))
print("torch 2D copy device->pinned      %.1f ms" % t(lambda: h_out.copy_(d[:, :cols], non_blocking=True)))
dd = torch.empty((rows, cols), dtype=torch.float32, device="cuda")
print("torch dense copy device->pinned   %.1f ms" % (ms := t(lambda: h_out.copy_(dd, non_blocking=True))), "(%.1f GB/s)" % (rows * cols * 4 / ms / 1e6))
part = 8192
pageable = np.empty((part, cols), dtype=np.float32)
print("auvi_lattice -> pageable (8192 rows) %.1f ms (%.1f GB/s)" % ((ms := t(lambda: lat(pageable.ctypes.data, 0, part))), part * cols * 4 / ms / 1e6))
destroy()
del d, dd
torch.cuda.empty_cache()

# ---- the exact sequence bench.py times, piece by piece ---------------------------------------------------
print("--- bench-like e2e steps (create / lattice / destroy), ms")
for it in range(4):
    t0 = time.perf_counter(); create(); t1 = time.perf_counter()
    lat(h_out.data_ptr()); t2 = time.perf_counter()
    destroy(); t3 = time.perf_counter()
    print("step %d: create %.1f  lattice %.1f  destroy %.1f  total %.1f" % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3))
