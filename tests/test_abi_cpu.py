"""CPU suite, part 2: the drop-in boundary.

* libauvi.so loads and exports exactly the symbols include/auvi.h declares (no compute calls: there
  is no GPU in the build container);
* without a CUDA device every compute entry point fails loudly -- there is no CPU fallback;
* nothing in the product tree (package, include/) references oracle/.
"""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "auv-real-time-interpolation_b200")
sys.path.insert(0, os.path.join(PKG, "python"))

import auvi  # noqa: E402


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "auvi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(auvi_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built():
    assert os.path.exists(auvi.LIB_PATH), "run `python -c 'import __graft_entry__ as g; g.build()'` first"


def test_header_symbols_all_exported_and_bound():
    declared = _declared_symbols()
    assert len(declared) >= 14
    lib = auvi.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/auvi.h but not exported"
    assert sorted(auvi.SYMBOLS) == declared, "python binding and header disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", auvi.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (auvi_[a-z0-9_]+)", out)))
    assert exported == declared, "exported auvi_* symbols differ from the header"


def test_no_cpu_fallback_without_a_device():
    lib = auvi.load()
    if lib.auvi_device_count() > 0:
        pytest.skip("a CUDA device is present")
    z = np.zeros((4, 4))
    with pytest.raises(auvi.AuviError, match="no CUDA device"):
        auvi.Grid(z, 0.0, 1.0, 0.0, 1.0)
    h = C.c_void_p()
    rc = lib.auvi_grid_adopt(C.c_void_p(16), auvi.F32, 4, 4, 4, 0, 4, 0.0, 1.0, 0.0, 1.0, 0, C.byref(h))
    assert rc != 0 and b"no CUDA device" in lib.auvi_last_error()
    out3 = (C.c_double * 3)()
    rc = lib.auvi_error_metrics_device(C.c_void_p(16), C.c_void_p(16), auvi.F64, 4, out3, None, None)
    assert rc != 0 and b"no CUDA device" in lib.auvi_last_error()
    assert lib.auvi_grid_destroy(None) == 0          # idempotent on NULL
    assert lib.auvi_version() >= 100


def test_argument_validation_messages():
    lib = auvi.load()
    h = C.c_void_p()
    assert lib.auvi_grid_create(None, auvi.F64, 4, 4, 0.0, 1.0, 0.0, 1.0, 0, C.byref(h)) != 0
    assert b"null host grid" in lib.auvi_last_error()
    assert lib.auvi_interp_points(None, auvi.BILINEAR, None, 1, 24, None, 8) != 0
    assert b"null grid handle" in lib.auvi_last_error()


def test_peer_and_prep_entries_validate_arguments():
    lib = auvi.load()
    buf = (C.c_ubyte * 72)()
    assert lib.auvi_peer_export(None, buf) != 0 and b"null argument" in lib.auvi_last_error()
    out = C.c_void_p()
    assert lib.auvi_peer_open(None, C.byref(out)) != 0 and b"null argument" in lib.auvi_last_error()
    assert lib.auvi_peer_close(None) == 0
    h = C.c_void_p()
    assert lib.auvi_grid_create_raw(None, 3, 1, 1, 1.0, 0.0, auvi.F64, 4, 4, 0.0, 1.0, 0.0, 1.0, 0, C.byref(h)) != 0
    assert b"null host buffer" in lib.auvi_last_error()
    assert lib.auvi_grid_create_raw(buf, 9, 1, 1, 1.0, 0.0, auvi.F64, 4, 4, 0.0, 1.0, 0.0, 1.0, 0, C.byref(h)) != 0
    assert b"raw element type" in lib.auvi_last_error()
    assert lib.auvi_grid_mask_cells(None, None, 1, None) != 0 and b"null grid handle" in lib.auvi_last_error()
    if lib.auvi_device_count() == 0:
        assert lib.auvi_grid_create_csv(b"1,2\n3,4\n", 8, auvi.F64, 0.0, 1.0, 0.0, 1.0, 0, C.byref(h)) != 0
        assert b"no CUDA device" in lib.auvi_last_error()


def test_product_tree_never_touches_the_oracle():
    bad = []
    for base in (PKG, os.path.join(ROOT, "include")):
        for dirpath, _, files in os.walk(base):
            if os.path.basename(dirpath) in ("build", "lib", "bin", "__pycache__"):
                continue
            for fn in files:
                if not fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".c")) and fn != "Makefile":
                    continue
                text = open(os.path.join(dirpath, fn), errors="replace").read()
                for m in re.finditer(r"^.*(liboracle|libgridh_ref|interp_oracle|from oracle|import oracle|oracle/).*$",
                                     text, flags=re.M):
                    line = m.group(0).strip()
                    if line.startswith(("//", "#", "*", "/*")) or "never imports" in line:
                        continue
                    bad.append((fn, line))
    assert not bad, bad


def test_gridd_header_matches_reference_layout():
    """Our GridD must keep the reference's member layout (the drivers are compiled against the
    reference's header): same member order and types."""
    text = open(os.path.join(PKG, "host", "include", "GridD.h")).read()
    priv = text[text.index("private:"):text.index("public:")]
    members = re.findall(r"^\s*(double\*|int|double|bool)\s+([a-z_, ]+);", priv, flags=re.M)
    flat = [(t, n.strip()) for t, names in members for n in names.split(",")]
    assert flat == [("double*", "d_grid"), ("int", "num_lon"), ("int", "num_lat"), ("double", "min_lon"),
                    ("double", "max_lon"), ("double", "min_lat"), ("double", "max_lat"), ("double", "lon_step"),
                    ("double", "lat_step"), ("bool", "initialized")]


def test_multi_gpu_row_plan_covers_the_lattice_and_its_halo():
    """auvi_multi_plan (host only): the shards' lattice rows partition the lattice exactly, and every shard holds the grid
    rows its kernels can touch -- stencil -1/+2 and a radius-10 ring search around a centre that FP64 noise may move by one
    row (the checks launch_tiled / launch_fill make on the device side: base - 12 .. base + 13)."""
    import auvi
    for n_lat in (2, 5, 97, 301, 65536):
        for n_gpus in (1, 2, 3, 8, 64):
            for f in (1, 2, 4):
                total = f * (n_lat - 1) + 1
                nxt = 0
                for k in range(n_gpus):
                    p = auvi.multi_plan(n_lat, n_gpus, k, f)
                    assert p["row_lo"] == nxt and p["row_hi"] >= p["row_lo"]
                    nxt = p["row_hi"]
                    if p["row_hi"] > p["row_lo"]:
                        need_lo = max(0, p["row_lo"] // f - 12)
                        need_hi = min(n_lat - 1, (p["row_hi"] - 1) // f + 13)
                        assert p["in_lo"] <= need_lo and need_hi < p["in_hi"], (n_lat, n_gpus, k, f, p)
                    r = auvi.multi_plan(n_lat, n_gpus, k, f, replicate=True)
                    assert (r["in_lo"], r["in_hi"]) == (0, n_lat) and (r["row_lo"], r["row_hi"]) == (p["row_lo"], p["row_hi"])
                assert nxt == total
    with pytest.raises(auvi.AuviError):
        auvi.multi_plan(10, 2, 2)
