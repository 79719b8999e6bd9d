"""Grid upload from pinned vs pageable host memory (auvi_grid_create), GB/s.  Run on a GPU box."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "auv-real-time-interpolation_b200", "python")); sys.path.insert(0, ROOT)
import torch, auvi
lib = auvi.load()
n = 16384
pin = torch.rand((n, n), dtype=torch.float32).pin_memory()
page = pin.numpy().copy()
def t(ptr, reps=3):
    h = C.c_void_p()
    assert lib.auvi_grid_create(ptr, auvi.F32, n, n, 0.0, 1.0, 0.0, 1.0, 0, C.byref(h)) == 0; lib.auvi_grid_destroy(h)
    t0 = time.perf_counter()
    for _ in range(reps):
        assert lib.auvi_grid_create(ptr, auvi.F32, n, n, 0.0, 1.0, 0.0, 1.0, 0, C.byref(h)) == 0
        lib.auvi_grid_destroy(h)
    return (time.perf_counter() - t0) / reps
for name, ptr in (("pinned", pin.data_ptr()), ("pageable", page.ctypes.data)):
    dt = t(ptr)
    print(f"{name:9s} {dt*1e3:8.1f} ms  {n*n*4/dt/1e9:6.1f} GB/s")
g = auvi.Grid(page[:2000, :], 0.0, 1.0, 0.0, 1.0)          # correctness of the bounce path (131 MB, pitched source view)
assert np.array_equal(g.read(), page[:2000, :]); g.close(); print("bounce upload matches")
