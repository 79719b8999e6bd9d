"""oracle/binding.py -- ctypes loaders for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  Nothing under auv-real-time-interpolation_b200/ does (tests/test_no_cpu_fallback.py
greps for it).

  Oracle      -> oracle/liboracle.so          plain-C restatement (oracle/interp_oracle.c)
  Reference   -> oracle/_ref/libgridh_ref.so  the unmodified reference GridH.cpp behind a C shim

Also here: the fixture helpers that restate the reference's *data preparation* (not its hot path):
  load_tile()      tests/golden/tiles/*.i16.xz  (GEBCO int16 tiles, delta + xz packed)
  make_mask()      code/subset_bathymetry.py:32-39 (legacy np.random.seed(42) + choice)
  node_queries()   code/test_gebco.cpp:72-81,150-160 (row,col -> lon,lat)
  lattice_queries() code/test_interpolation.cpp:91-109 ((2n-1)x(2n-1) expanded lattice)
"""
from __future__ import annotations

import ctypes as C
import json
import lzma
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")

BILINEAR, CUBIC, KRIGING, NN, IDW = 0, 1, 2, 3, 4
METHOD_NAMES = {BILINEAR: "bilinear", CUBIC: "cubic", KRIGING: "kriging", NN: "nn", IDW: "idw"}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def _ptr(a, ty):
    return a.ctypes.data_as(ty) if a is not None else None


def build(ref: bool = True) -> None:
    """Compile liboracle.so (always) and _ref/libgridh_ref.so (when /root/reference exists)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "all" if ref else os.path.join(HERE, "liboracle.so")])


class _OrcGrid(C.Structure):
    _fields_ = [("z", _dp), ("n_lat", C.c_int), ("n_lon", C.c_int),
                ("min_lon", C.c_double), ("max_lon", C.c_double),
                ("min_lat", C.c_double), ("max_lat", C.c_double),
                ("lon_step", C.c_double), ("lat_step", C.c_double)]


def _as_points(pts) -> np.ndarray:
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    assert pts.ndim == 2 and pts.shape[1] == 3, "points are n x {lon,lat,elev} (Point.h:9-13)"
    return pts


class Oracle:
    """Plain-C restatement of GridH (oracle/interp_oracle.c)."""

    def __init__(self, z, min_lon, max_lon, min_lat, max_lat):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = C.CDLL(path)
        self.z = np.ascontiguousarray(z, dtype=np.float64)
        self.n_lat, self.n_lon = self.z.shape
        self.g = _OrcGrid()
        self.lib.orc_grid_init.argtypes = [C.POINTER(_OrcGrid), _dp, C.c_int, C.c_int] + [C.c_double] * 4
        self.lib.orc_grid_init(C.byref(self.g), _ptr(self.z, _dp), self.n_lat, self.n_lon,
                               min_lon, max_lon, min_lat, max_lat)
        self.lib.orc_batch.argtypes = [C.POINTER(_OrcGrid), C.c_int, _dp, C.c_int64, _dp, _ip, _ip]
        self.lib.orc_batch.restype = C.c_int
        for name in ("orc_mae", "orc_rmse", "orc_maxerr"):
            f = getattr(self.lib, name)
            f.argtypes = [_dp, _dp, C.c_int64]
            f.restype = C.c_double

    def batch(self, method, pts, want_sel=False):
        pts = _as_points(pts)
        n = pts.shape[0]
        out = np.empty(n, dtype=np.float64)
        sel = np.full((n, 4, 2), -1, dtype=np.int32) if want_sel else None
        found = np.full(n, -1, dtype=np.int32) if want_sel else None
        rc = self.lib.orc_batch(C.byref(self.g), method, _ptr(pts, _dp), n, _ptr(out, _dp),
                                _ptr(sel, _ip), _ptr(found, _ip))
        assert rc == 0
        return (out, sel, found) if want_sel else out

    def metrics(self, truth, est):
        t = np.ascontiguousarray(truth, dtype=np.float64)
        e = np.ascontiguousarray(est, dtype=np.float64)
        n = t.shape[0]
        return tuple(getattr(self.lib, f)(_ptr(t, _dp), _ptr(e, _dp), n)
                     for f in ("orc_mae", "orc_rmse", "orc_maxerr"))

    # ---- the opt-in methods (no reference code: checkers of our own definitions) ----
    def batch_optin(self, method, pts, params=None, want_sel=False):
        """method 6 = IDW over the true four nearest (no early break), 7 = kriging with params = (c0, c1, range)."""
        pts = _as_points(pts)
        n = pts.shape[0]
        out = np.empty(n, dtype=np.float64)
        sel = np.full((n, 4, 2), -1, dtype=np.int32) if want_sel else None
        found = np.full(n, -1, dtype=np.int32) if want_sel else None
        par = np.ascontiguousarray(params if params is not None else (0.0, 1.0, 1.0), dtype=np.float64)
        self.lib.orc_batch_optin.argtypes = [C.POINTER(_OrcGrid), C.c_int, _dp, _dp, C.c_int64, _dp, _ip, _ip]
        self.lib.orc_batch_optin.restype = C.c_int
        rc = self.lib.orc_batch_optin(C.byref(self.g), method, _ptr(par, _dp), _ptr(pts, _dp), n, _ptr(out, _dp),
                                      _ptr(sel, _ip), _ptr(found, _ip))
        assert rc == 0
        return (out, sel, found) if want_sel else out

    def variogram_sums(self):
        out = np.empty(16, dtype=np.float64)
        self.lib.orc_variogram_sums.argtypes = [C.POINTER(_OrcGrid), _dp]
        self.lib.orc_variogram_sums(C.byref(self.g), _ptr(out, _dp))
        return out

    def variogram_fit(self, sums=None):
        """-> (c0, c1, range) of the exponential model fitted to this grid (orc_variogram_fit)."""
        sums = self.variogram_sums() if sums is None else np.ascontiguousarray(sums, dtype=np.float64)
        out = np.empty(3, dtype=np.float64)
        self.lib.orc_variogram_fit.argtypes = [_dp, C.c_double, C.c_double, _dp]
        self.lib.orc_variogram_fit.restype = C.c_int
        rc = self.lib.orc_variogram_fit(_ptr(sums, _dp), self.g.lon_step, self.g.lat_step, _ptr(out, _dp))
        assert rc == 0, "variogram fit failed"
        return tuple(out)


def node_axis(lo, hi, n):
    lib = C.CDLL(os.path.join(HERE, "liboracle.so"))
    out = np.empty(n, dtype=np.float64)
    lib.orc_node_axis.argtypes = [C.c_double, C.c_double, C.c_int, _dp]
    lib.orc_node_axis(lo, hi, n, _ptr(out, _dp))
    return out


def lattice_axis(lo, hi, new_n):
    lib = C.CDLL(os.path.join(HERE, "liboracle.so"))
    out = np.empty(new_n, dtype=np.float64)
    lib.orc_lattice_axis.argtypes = [C.c_double, C.c_double, C.c_int, _dp]
    lib.orc_lattice_axis(lo, hi, new_n, _ptr(out, _dp))
    return out


def synth_grid(n_lat, n_lon, csv_round=True):
    """generate_csv_grids.cpp:32-70; csv_round pushes values through the 6-significant-digit
    text form the CSV writer (:73-88, default ostream precision) imposes."""
    lib = C.CDLL(os.path.join(HERE, "liboracle.so"))
    z = np.empty((n_lat, n_lon), dtype=np.float64)
    lib.orc_synth_grid.argtypes = [C.c_int, C.c_int, _dp]
    lib.orc_synth_grid(n_lat, n_lon, _ptr(z, _dp))
    if csv_round:
        if z.size <= 1 << 20:      # exact: the same %g text the CSV writer emits, parsed back
            z = np.array([float("%g" % v) for v in z.ravel()], dtype=np.float64).reshape(n_lat, n_lon)
        else:
            z = _round6(z)
    return z


def _round6(z):
    # vectorised %.6g: scale to 6 significant decimal digits, round half-even like printf does on
    # the exact binary value is NOT guaranteed here, so only use it for large synthetic grids where
    # the exact CSV text is not part of a golden comparison.
    mag = np.floor(np.log10(np.maximum(np.abs(z), 1e-300)))
    s = 10.0 ** (5 - mag)
    return np.round(z * s) / s


def ref_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libgridh_ref.so"))


class Reference:
    """The unmodified reference GridH behind oracle/ref_shim.cpp."""

    def __init__(self, z, min_lon, max_lon, min_lat, max_lat):
        self.lib = C.CDLL(os.path.join(HERE, "_ref", "libgridh_ref.so"))
        z = np.ascontiguousarray(z, dtype=np.float64)
        self.n_lat, self.n_lon = z.shape
        self.lib.refh_create.argtypes = [_dp, C.c_int, C.c_int] + [C.c_double] * 4
        self.lib.refh_create.restype = C.c_void_p
        self.lib.refh_destroy.argtypes = [C.c_void_p]
        self.lib.refh_batch.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int64, _dp, C.c_int]
        self.lib.refh_batch.restype = C.c_int
        self.lib.refh_select4_batch.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int64, _ip, _ip]
        self.lib.refh_metric.argtypes = [C.c_int, _dp, _dp, C.c_int64]
        self.lib.refh_metric.restype = C.c_double
        self.h = self.lib.refh_create(_ptr(z, _dp), self.n_lat, self.n_lon,
                                      min_lon, max_lon, min_lat, max_lat)

    def __del__(self):
        try:
            self.lib.refh_destroy(self.h)
        except Exception:
            pass

    def batch(self, method, pts, threads=1):
        pts = _as_points(pts)
        out = np.empty(pts.shape[0], dtype=np.float64)
        rc = self.lib.refh_batch(self.h, method, _ptr(pts, _dp), pts.shape[0], _ptr(out, _dp), threads)
        assert rc == 0
        return out

    def select4(self, centre_rule, pts):
        pts = _as_points(pts)
        n = pts.shape[0]
        found = np.empty(n, dtype=np.int32)
        sel = np.empty((n, 4, 2), dtype=np.int32)
        self.lib.refh_select4_batch(self.h, centre_rule, _ptr(pts, _dp), n, _ptr(found, _ip), _ptr(sel, _ip))
        return found, sel

    def metrics(self, truth, est):
        t = np.ascontiguousarray(truth, dtype=np.float64)
        e = np.ascontiguousarray(est, dtype=np.float64)
        return tuple(self.lib.refh_metric(w, _ptr(t, _dp), _ptr(e, _dp), t.shape[0]) for w in (0, 1, 2))


# --------------------------------------------------------------------------------------------
# fixtures
# --------------------------------------------------------------------------------------------
def tile_manifest():
    with open(os.path.join(GOLDEN, "tiles", "tiles.json")) as f:
        return json.load(f)


def load_tile(name):
    """-> (z float64 [n_lat][n_lon] in the row order the reference's mask tool writes, bounds dict)."""
    m = tile_manifest()[name]
    with open(os.path.join(GOLDEN, "tiles", m["file"]), "rb") as f:
        raw = lzma.decompress(f.read())
    d = np.frombuffer(raw, dtype="<i2").reshape(m["n_lat"], m["n_lon"])
    z = np.cumsum(d.astype(np.int64), axis=1).astype(np.float64)
    return z, m


def make_mask(n_lat, n_lon, fraction, seed=42):
    """Flat indices removed by subset_bathymetry.py:32-39, in the order it writes them."""
    total = n_lat * n_lon
    n_remove = int(total * fraction)
    np.random.seed(seed)
    return np.random.choice(total, size=n_remove, replace=False)


def masked_case(name, fraction):
    """-> dict(z_masked, truth, rows, cols, pts (n x 3), bounds...) for one Grid-B run."""
    z, m = load_tile(name)
    flat = make_mask(m["n_lat"], m["n_lon"], fraction)
    rows, cols = flat // m["n_lon"], flat % m["n_lon"]
    truth = z[rows, cols].copy()
    zm = z.copy()
    zm[rows, cols] = np.nan
    pts = node_queries(rows, cols, m)
    return dict(z=zm, truth=truth, rows=rows, cols=cols, pts=pts, meta=m,
                bounds=(m["min_lon"], m["max_lon"], m["min_lat"], m["max_lat"]))


def node_queries(rows, cols, m):
    """test_gebco.cpp:72-81: lat = min_lat + row*lat_step, lon = min_lon + col*lon_step."""
    lat_ax = node_axis(m["min_lat"], m["max_lat"], m["n_lat"])
    lon_ax = node_axis(m["min_lon"], m["max_lon"], m["n_lon"])
    pts = np.zeros((len(rows), 3), dtype=np.float64)
    pts[:, 0] = lon_ax[cols]
    pts[:, 1] = lat_ax[rows]
    return pts


def lattice_queries(n_lat, n_lon, min_lon, max_lon, min_lat, max_lat, f_lat=2, f_lon=2):
    """test_interpolation.cpp:91-109 generalised to new_n = f*(n-1)+1 (f=2 gives 2n-1)."""
    nn_lat, nn_lon = f_lat * (n_lat - 1) + 1, f_lon * (n_lon - 1) + 1
    lat_ax = lattice_axis(min_lat, max_lat, nn_lat)
    lon_ax = lattice_axis(min_lon, max_lon, nn_lon)
    pts = np.zeros((nn_lat * nn_lon, 3), dtype=np.float64)
    pts[:, 0] = np.tile(lon_ax, nn_lat)
    pts[:, 1] = np.repeat(lat_ax, nn_lon)
    return pts, nn_lat, nn_lon


def ref_gpu_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref_gpu", "libgridd_ref_sm100a.so"))


class ReferenceGPU:
    """The unmodified reference GPU code (kernels.cu + GridD.cu) compiled for sm_100a behind oracle/ref_gpu_shim.cu:
    the secondary comparator (needs a CUDA device)."""

    def __init__(self, z, min_lon, max_lon, min_lat, max_lat):
        self.lib = C.CDLL(os.path.join(HERE, "_ref_gpu", "libgridd_ref_sm100a.so"))
        z = np.ascontiguousarray(z, dtype=np.float64)
        self.n_lat, self.n_lon = z.shape
        self.lib.refd_create.argtypes = [_dp, C.c_int, C.c_int] + [C.c_double] * 4
        self.lib.refd_create.restype = C.c_void_p
        self.lib.refd_destroy.argtypes = [C.c_void_p]
        self.lib.refd_batch.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int64, _dp]
        self.lib.refd_batch.restype = C.c_double
        self.lib.refd_kernel_ms.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int64, C.c_int, _dp]
        self.lib.refd_kernel_ms.restype = C.c_double
        self.h = self.lib.refd_create(_ptr(z, _dp), self.n_lat, self.n_lon, min_lon, max_lon, min_lat, max_lat)
        if not self.h:
            raise RuntimeError("reference GridD could not be created")

    def close(self):
        if self.h:
            self.lib.refd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def batch(self, method, pts):
        """GridD::batch* end to end -> (depths, milliseconds as the drivers' chrono sees them)."""
        pts = _as_points(pts)
        out = np.empty(pts.shape[0], dtype=np.float64)
        ms = self.lib.refd_batch(self.h, method, _ptr(pts, _dp), pts.shape[0], _ptr(out, _dp))
        return out, ms

    def kernel_ms(self, method, pts, reps=5):
        """The reference kernel alone on resident buffers -> (depths, ms per launch)."""
        pts = _as_points(pts)
        out = np.empty(pts.shape[0], dtype=np.float64)
        ms = self.lib.refd_kernel_ms(self.h, method, _ptr(pts, _dp), pts.shape[0], reps, _ptr(out, _dp))
        return out, ms


def hash_mask(row_lo, row_hi, n_lon, fraction, seed=42):
    """Rows [row_lo,row_hi) of the counter-hash mask auvi_grid_mask_hash draws (csrc/ingest.cu mask_hash_kernel): cell
    (r,c) is removed iff splitmix64(flat ^ seed*K) >> 11 < fraction * 2^53, flat = r*n_lon + c.  -> bool array."""
    r = np.arange(row_lo, row_hi, dtype=np.uint64)[:, None]
    c = np.arange(n_lon, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        x = (r * np.uint64(n_lon) + c) ^ (np.uint64(seed) * np.uint64(0xD1342543DE82EF95))
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    f = min(max(float(fraction), 0.0), 1.0)
    return (x >> np.uint64(11)) < np.uint64(int(f * 9007199254740992.0))


def synth_rows(n_lat_global, n_lon, row_lo, row_hi):
    """Rows [row_lo,row_hi) of the seamount field (generate_csv_grids.cpp:32-70) of an n_lat_global x n_lon grid, float64."""
    i = np.arange(n_lon, dtype=np.float64) * (100.0 / (n_lon - 1))
    j = np.arange(row_lo, row_hi, dtype=np.float64) * (100.0 / (n_lat_global - 1))
    return -(10.0 + 2.0 * i)[None, :] + 100.0 * np.exp(-(((i - 75.0) ** 2)[None, :] + ((j - 50.0) ** 2)[:, None]) / 450.0)
