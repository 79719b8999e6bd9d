"""ctypes view of libauvi.so (include/auvi.h) for tests, bench.py and __graft_entry__.py.

This is plumbing, not product: the product is the C-ABI library and the C++ class GridD on top of
it (host/GridD.cpp).  Nothing here computes; every call goes to the CUDA library, and loading
fails loudly when the library is missing -- there is no CPU fallback (tests/test_abi_cpu.py checks
that this module never imports anything from oracle/).

    g = auvi.Grid(z, min_lon, max_lon, min_lat, max_lat)          # uploads, like GridD::GridD
    out = g.interp_points(auvi.KRIGING, pts)                        # host Point list -> host depths
    lat = g.lattice(auvi.CUBIC, auvi.AXIS_EXPANDED, 2, 2)           # (2n-1)x(2n-1) host array
    g.lattice_device(...), g.interp_points_device(...)              # raw device pointers (torch .data_ptr())
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB_PATH = os.environ.get("AUVI_LIB") or os.path.join(PKG, "lib", "libauvi.so")   # AUVI_LIB: another build of the same library (A/B runs)

BILINEAR, CUBIC, KRIGING, NN, IDW, BILINEAR_SEARCH, IDW_KNN, KRIGING_FITTED = 0, 1, 2, 3, 4, 5, 6, 7
METHOD_NAMES = {BILINEAR: "bilinear", CUBIC: "cubic", KRIGING: "kriging", NN: "nn", IDW: "idw",
                BILINEAR_SEARCH: "bilinear_search", IDW_KNN: "idw_knn", KRIGING_FITTED: "kriging_fitted"}
F64, F32 = 0, 1
AXIS_EXPANDED, AXIS_NODES = 0, 1

# every symbol include/auvi.h declares: (name, restype, argtypes)
_vp, _i64, _i32, _dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
SYMBOLS = {
    "auvi_grid_create": (_i32, [_vp, _i32, _i64, _i64, _dbl, _dbl, _dbl, _dbl, _i32, C.POINTER(_vp)]),
    "auvi_grid_create_slab": (_i32, [_vp, _i32, _i64, _i64, _i64, _i64, _dbl, _dbl, _dbl, _dbl, _i32, C.POINTER(_vp)]),
    "auvi_grid_adopt": (_i32, [_vp, _i32, _i64, _i64, _i64, _i64, _i64, _dbl, _dbl, _dbl, _dbl, _i32, C.POINTER(_vp)]),
    "auvi_grid_destroy": (_i32, [_vp]),
    "auvi_trim": (_i32, []),
    "auvi_interp_points": (_i32, [_vp, _i32, _vp, _i64, _i64, _vp, _i64]),
    "auvi_interp_points_device": (_i32, [_vp, _i32, _vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    "auvi_lattice_dims": (_i32, [_vp, _i32, _i32, _i32, C.POINTER(_i64), C.POINTER(_i64)]),
    "auvi_lattice_device": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _i64, _i64, _vp, _i64, _vp, _vp]),
    "auvi_lattice": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _i64, _i64, _vp]),
    "auvi_error_metrics_device": (_i32, [_vp, _vp, _i32, _i64, C.POINTER(_dbl), C.POINTER(_i64), _vp]),
    "auvi_fill_metrics_device": (_i32, [_vp, _vp, _i64, _vp, _i64, _i64, _i64, C.POINTER(_dbl), C.POINTER(_i64),
                                        C.POINTER(_i64), _vp]),
    "auvi_netcdf3_find": (_i32, [_vp, _i64, C.c_char_p, _vp]),
    "auvi_netcdf3_read_f64": (_i32, [_vp, _i64, C.c_char_p, _vp, _i64]),
    "auvi_legacy_choice": (_i32, [_i64, _i64, C.c_uint32, _vp]),
    "auvi_grid_create_raw": (_i32, [_vp, _i32, _i32, _i32, _dbl, _dbl, _i32, _i64, _i64, _dbl, _dbl, _dbl, _dbl, _i32,
                                    C.POINTER(_vp)]),
    "auvi_csv_dims": (_i32, [C.c_char_p, _i64, C.POINTER(_i64), C.POINTER(_i64)]),
    "auvi_grid_create_csv": (_i32, [C.c_char_p, _i64, _i32, _dbl, _dbl, _dbl, _dbl, _i32, C.POINTER(_vp)]),
    "auvi_grid_mask_cells": (_i32, [_vp, _vp, _i64, _vp]),
    "auvi_grid_mask_hash": (_i32, [_vp, _dbl, C.c_uint64, C.POINTER(_i64), _vp]),
    "auvi_grid_read": (_i32, [_vp, _i64, _i64, _vp]),
    "auvi_grid_fit_variogram": (_i32, [_vp, C.POINTER(_dbl)]),
    "auvi_grid_set_variogram": (_i32, [_vp, _dbl, _dbl, _dbl]),
    "auvi_variogram_fit_from_sums": (_i32, [C.POINTER(_dbl), _dbl, _dbl, C.POINTER(_dbl)]),
    "auvi_peer_export": (_i32, [_vp, _vp]),
    "auvi_peer_open": (_i32, [_vp, C.POINTER(_vp)]),
    "auvi_peer_close": (_i32, [_vp]),
    "auvi_multi_create": (_i32, [_vp, _i32, _i64, _i64, _dbl, _dbl, _dbl, _dbl, _i32, _vp, _i32, C.POINTER(_vp)]),
    "auvi_multi_destroy": (_i32, [_vp]),
    "auvi_multi_plan": (_i32, [_i64, _i32, _i32, _i32, _i32, C.POINTER(_i64)]),
    "auvi_multi_count": (_i32, [_vp]),
    "auvi_multi_shard": (_i32, [_vp, _i32, _i32, C.POINTER(_i32), C.POINTER(_vp), C.POINTER(_i64), C.POINTER(_i64)]),
    "auvi_multi_mask_hash": (_i32, [_vp, _dbl, C.c_uint64]),
    "auvi_multi_lattice": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "auvi_multi_lattice_device": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _i64]),
    "auvi_multi_enable_peer": (_i32, [_vp, _i32]),
    "auvi_multi_sync": (_i32, [_vp]),
    "auvi_multi_last_kernel_ms": (C.c_float, [_vp]),
    "auvi_multi_interp_points": (_i32, [_vp, _i32, _vp, _i64, _i64, _vp, _i64]),
    "auvi_host_prefault": (_i32, [_vp, _i64]),
    "auvi_last_error": (C.c_char_p, []),
    "auvi_last_kernel_ms": (C.c_float, [_vp]),
    "auvi_launch_count": (_i64, []),
    "auvi_uses_tma": (_i32, [_vp]),
    "auvi_uses_window": (_i32, [_vp]),
    "auvi_device_count": (_i32, []),
    "auvi_version": (_i32, []),
}

_lib = None


class AuviError(RuntimeError):
    pass


def load():
    """Load libauvi.so and bind every declared symbol.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AuviError(f"{LIB_PATH} is missing: build it (python -c 'import __graft_entry__ as g; g.build()'); "
                        "there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        raise AuviError(load().auvi_last_error().decode())


def launch_count() -> int:
    return int(load().auvi_launch_count())


def device_count() -> int:
    return int(load().auvi_device_count())


def _np_dtype(dtype):
    return np.float64 if dtype == F64 else np.float32


class NcVar(C.Structure):
    """struct auvi_nc_var (include/auvi.h)."""
    _fields_ = [("nc_type", C.c_int32), ("elem_bytes", C.c_int32), ("ndims", C.c_int32), ("has_fill", C.c_int32),
                ("shape", C.c_int64 * 4), ("n_elems", C.c_int64), ("data_offset", C.c_int64),
                ("scale_factor", C.c_double), ("add_offset", C.c_double), ("fill_value", C.c_double)]


def netcdf3_find(image: bytes, name: str) -> NcVar:
    v = NcVar()
    _check(load().auvi_netcdf3_find(image, len(image), name.encode(), C.byref(v)))
    return v


def netcdf3_read_f64(image: bytes, name: str) -> np.ndarray:
    v = netcdf3_find(image, name)
    out = np.empty(v.n_elems, dtype=np.float64)
    _check(load().auvi_netcdf3_read_f64(image, len(image), name.encode(), out.ctypes.data, out.size))
    return out.reshape([v.shape[k] for k in range(v.ndims)])


def legacy_choice(total: int, n: int, seed: int = 42) -> np.ndarray:
    """numpy.random.seed(seed); numpy.random.choice(total, n, replace=False) -- restated in C++ (prep.cpp)."""
    out = np.empty(n, dtype=np.int64)
    _check(load().auvi_legacy_choice(total, n, seed, out.ctypes.data))
    return out


class Grid:
    """A depth grid resident on one GPU (handle owner)."""

    @classmethod
    def from_netcdf(cls, image: bytes, var: str, bounds, dtype=F64, flip_rows=True, device=0):
        """A NetCDF-3 file image -> device grid: the variable's big-endian elements are uploaded as they lie in the
        file and decoded on the GPU (subset_bathymetry.py:8-18 without netCDF4 / pandas)."""
        v = netcdf3_find(image, var)
        assert v.ndims == 2, "a 2-D variable is expected"
        self = cls.__new__(cls)
        self._h = _vp()
        self.bounds = tuple(bounds)
        self.dtype = dtype
        self.n_lat, self.n_lon = int(v.shape[0]), int(v.shape[1])
        raw = (C.c_char * (v.n_elems * v.elem_bytes)).from_buffer_copy(image, v.data_offset)
        _check(load().auvi_grid_create_raw(raw, v.nc_type, 1, int(flip_rows), v.scale_factor, v.add_offset, dtype,
                                           self.n_lat, self.n_lon, *bounds, device, C.byref(self._h)))
        return self

    @classmethod
    def from_csv(cls, text: bytes, bounds, dtype=F64, device=0):
        """CSV matrix text (readGridCSV's input, test_gebco.cpp:19-40) -> device grid, parsed on the GPU."""
        self = cls.__new__(cls)
        self._h = _vp()
        self.bounds = tuple(bounds)
        self.dtype = dtype
        r, c = _i64(), _i64()
        _check(load().auvi_csv_dims(text, len(text), C.byref(r), C.byref(c)))
        self.n_lat, self.n_lon = r.value, c.value
        _check(load().auvi_grid_create_csv(text, len(text), dtype, *bounds, device, C.byref(self._h)))
        return self

    def mask_cells(self, flat_idx, want_truth=True):
        """Set the listed cells (row*n_lon+col) to NaN on the device; -> their former values in list order."""
        flat_idx = np.ascontiguousarray(flat_idx, dtype=np.int64)
        truth = np.empty(flat_idx.size, dtype=_np_dtype(self.dtype)) if want_truth else None
        _check(load().auvi_grid_mask_cells(self._h, flat_idx.ctypes.data, flat_idx.size,
                                           truth.ctypes.data if want_truth else None))
        return truth

    def mask_hash(self, fraction, seed=42, count=True, stream=None):
        n = _i64()
        _check(load().auvi_grid_mask_hash(self._h, fraction, seed, C.byref(n) if count else None, stream))
        return n.value if count else None

    def read(self, row_begin=0, row_end=None):
        row_end = self.n_lat if row_end is None else row_end
        out = np.empty((row_end - row_begin, self.n_lon), dtype=_np_dtype(self.dtype))
        _check(load().auvi_grid_read(self._h, row_begin, row_end, out.ctypes.data))
        return out

    def fit_variogram(self):
        """-> (c0, c1, range) of the exponential variogram fitted to this grid (opt-in KRIGING_FITTED)."""
        out3 = (_dbl * 3)()
        _check(load().auvi_grid_fit_variogram(self._h, out3))
        return out3[0], out3[1], out3[2]

    def set_variogram(self, c0, c1, rng):
        _check(load().auvi_grid_set_variogram(self._h, c0, c1, rng))

    def fill_metrics_device(self, filled_ptr, filled_ld, truth_ptr, truth_ld, row_begin, row_end, stream=None):
        """-> (mae, rmse, max, n_nan, n) over the cells that are NaN in this (masked) grid."""
        out3 = (_dbl * 3)()
        nn, cnt = _i64(), _i64()
        _check(load().auvi_fill_metrics_device(self._h, filled_ptr, filled_ld, truth_ptr, truth_ld, row_begin, row_end,
                                               out3, C.byref(nn), C.byref(cnt), stream))
        return out3[0], out3[1], out3[2], nn.value, cnt.value

    def __init__(self, z=None, min_lon=0.0, max_lon=0.0, min_lat=0.0, max_lat=0.0, device=0, dtype=None,
                 adopt=None):
        lib = load()
        self._h = _vp()
        self.bounds = (min_lon, max_lon, min_lat, max_lat)
        if adopt is not None:
            # adopt = dict(ptr, dtype, n_lat, n_lon, ld, row0, rows): borrow device memory (e.g. a torch tensor)
            self.dtype = adopt["dtype"]
            self.n_lat, self.n_lon = adopt["n_lat"], adopt["n_lon"]
            _check(lib.auvi_grid_adopt(adopt["ptr"], self.dtype, self.n_lat, self.n_lon, adopt["ld"], adopt["row0"],
                                       adopt["rows"], min_lon, max_lon, min_lat, max_lat, device, C.byref(self._h)))
            self._keep = adopt.get("keep")
        else:
            if dtype is None:
                dtype = F32 if np.asarray(z).dtype == np.float32 else F64
            self.dtype = dtype
            z = np.ascontiguousarray(z, dtype=_np_dtype(dtype))
            self.n_lat, self.n_lon = z.shape
            _check(lib.auvi_grid_create(z.ctypes.data, dtype, self.n_lat, self.n_lon, min_lon, max_lon, min_lat,
                                        max_lat, device, C.byref(self._h)))

    def close(self):
        if self._h:
            load().auvi_grid_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- point list ------------------------------------------------------------------------------
    def interp_points(self, method, pts):
        """pts: n x 3 float64 {lon,lat,elev} (Point.h:9-13) on the host -> n float64 depths."""
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        assert pts.ndim == 2 and pts.shape[1] >= 2
        out = np.empty(pts.shape[0], dtype=np.float64)
        _check(load().auvi_interp_points(self._h, method, pts.ctypes.data, pts.shape[0], pts.strides[0],
                                         out.ctypes.data, 8))
        return out

    def interp_points_device(self, method, pts_ptr, n, stride_bytes, out_ptr, sel_ptr=None, found_ptr=None, stream=None):
        _check(load().auvi_interp_points_device(self._h, method, pts_ptr, n, stride_bytes, out_ptr, sel_ptr,
                                                found_ptr, stream))

    # ---- lattice ---------------------------------------------------------------------------------
    def lattice_dims(self, axis_kind, f_lat=1, f_lon=1):
        r, c = _i64(), _i64()
        _check(load().auvi_lattice_dims(self._h, axis_kind, f_lat, f_lon, C.byref(r), C.byref(c)))
        return r.value, c.value

    def lattice(self, method, axis_kind, f_lat=1, f_lon=1, fill=0, row_begin=0, row_end=None, out=None):
        rows, cols = self.lattice_dims(axis_kind, f_lat, f_lon)
        if row_end is None:
            row_end = rows
        if out is None:
            out = np.empty((row_end - row_begin, cols), dtype=_np_dtype(self.dtype))
        _check(load().auvi_lattice(self._h, method, axis_kind, f_lat, f_lon, fill, row_begin, row_end, out.ctypes.data))
        return out

    def lattice_into(self, method, axis_kind, f_lat, f_lon, fill, row_begin, row_end, host_ptr):
        _check(load().auvi_lattice(self._h, method, axis_kind, f_lat, f_lon, fill, row_begin, row_end, host_ptr))

    def lattice_device(self, method, axis_kind, f_lat, f_lon, fill, row_begin, row_end, out_ptr, out_ld,
                       sel9_ptr=None, stream=None):
        _check(load().auvi_lattice_device(self._h, method, axis_kind, f_lat, f_lon, fill, row_begin, row_end,
                                          out_ptr, out_ld, sel9_ptr, stream))

    # ---- diagnostics -----------------------------------------------------------------------------
    @property
    def last_kernel_ms(self):
        return float(load().auvi_last_kernel_ms(self._h))

    @property
    def uses_tma(self):
        return bool(load().auvi_uses_tma(self._h))

    @property
    def uses_window(self):
        return int(load().auvi_uses_window(self._h))


def error_metrics_device(truth_ptr, est_ptr, dtype, n, stream=None):
    """-> (mae, rmse, max, n_nan) with the reference's NaN convention (error_calculator.cpp:5-45)."""
    out3 = (_dbl * 3)()
    nn = _i64()
    _check(load().auvi_error_metrics_device(truth_ptr, est_ptr, dtype, n, out3, C.byref(nn), stream))
    return out3[0], out3[1], out3[2], nn.value


def multi_plan(n_lat, n_gpus, k, f_lat=1, replicate=False):
    """-> dict(own_lo, own_hi, in_lo, in_hi, row_lo, row_hi): the row plan of shard k (host only)."""
    out = (_i64 * 6)()
    _check(load().auvi_multi_plan(n_lat, n_gpus, k, f_lat, 1 if replicate else 0, out))
    return dict(zip(("own_lo", "own_hi", "in_lo", "in_hi", "row_lo", "row_hi"), list(out)))


class MultiGrid:
    """One host grid on several GPUs of this process (auvi_multi_*): row slabs + halo, or replicated."""

    def __init__(self, z, min_lon, max_lon, min_lat, max_lat, n_gpus, devices=None, replicate=False, dtype=None):
        lib = load()
        if dtype is None:
            dtype = F32 if np.asarray(z).dtype == np.float32 else F64
        self.dtype = dtype
        z = np.ascontiguousarray(z, dtype=_np_dtype(dtype))
        self.n_lat, self.n_lon = z.shape
        self._h = _vp()
        dev = (C.c_int * n_gpus)(*devices) if devices is not None else None
        _check(lib.auvi_multi_create(z.ctypes.data, dtype, self.n_lat, self.n_lon, min_lon, max_lon, min_lat, max_lat,
                                     n_gpus, dev, 1 if replicate else 0, C.byref(self._h)))
        self.n = n_gpus

    def close(self):
        if self._h:
            load().auvi_multi_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def shard(self, k, f_lat=1):
        """-> (device, row_lo, row_hi) of shard k at lattice factor f_lat."""
        dev, lo, hi, g = _i32(), _i64(), _i64(), _vp()
        _check(load().auvi_multi_shard(self._h, k, f_lat, C.byref(dev), C.byref(g), C.byref(lo), C.byref(hi)))
        return dev.value, lo.value, hi.value

    def mask_hash(self, fraction, seed=42):
        _check(load().auvi_multi_mask_hash(self._h, fraction, seed))

    def lattice(self, method, axis_kind, f_lat=1, f_lon=1, fill=0, out=None):
        rows = f_lat * (self.n_lat - 1) + 1
        cols = f_lon * (self.n_lon - 1) + 1
        if out is None:
            out = np.empty((rows, cols), dtype=_np_dtype(self.dtype))
        _check(load().auvi_multi_lattice(self._h, method, axis_kind, f_lat, f_lon, fill, out.ctypes.data))
        return out

    def lattice_device(self, method, axis_kind, f_lat, f_lon, fill, out_ptrs, out_ld):
        arr = (_vp * self.n)(*out_ptrs)
        _check(load().auvi_multi_lattice_device(self._h, method, axis_kind, f_lat, f_lon, fill, arr, out_ld))

    def enable_peer(self, root_shard=0):
        _check(load().auvi_multi_enable_peer(self._h, root_shard))

    def sync(self):
        _check(load().auvi_multi_sync(self._h))

    @property
    def last_kernel_ms(self):
        return float(load().auvi_multi_last_kernel_ms(self._h))

    def interp_points(self, method, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        out = np.empty(pts.shape[0], dtype=np.float64)
        _check(load().auvi_multi_interp_points(self._h, method, pts.ctypes.data, pts.shape[0], pts.strides[0],
                                               out.ctypes.data, 8))
        return out
