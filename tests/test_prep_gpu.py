"""GPU suite: Grid-B data preparation on the device against the fixture pipeline.

A GEBCO tile goes NetCDF-3 image -> auvi_grid_create_raw (decode + row flip on the GPU) -> auvi_legacy_choice +
auvi_grid_mask_cells (the reference's seeded removal) -> gap fill -> auvi_fill_metrics_device, and must reproduce
(1) the masked grid and truth column the numpy restatement of subset_bathymetry.py builds, bit for bit,
(2) the fill of the host-uploaded grid, bit for bit, (3) the golden MAE / RMSE / Max / NaN rows.
"""
import json
import os
import sys

import numpy as np
import pytest
from scipy.io import netcdf_file

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "auv-real-time-interpolation_b200", "python"))
sys.path.insert(0, ROOT)
from conftest import bits_equal  # noqa: E402
from oracle import binding as ob  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def auvi():
    import auvi as m
    m.load()
    if m.device_count() == 0:
        pytest.fail("no CUDA device: the gpu suite must run on a GPU box")
    return m


@pytest.fixture(scope="module")
def torch():
    import torch as t
    return t


def _tile_as_netcdf(tmp_path, name):
    """The fixture tile written the way GEBCO ships it: int16 'elevation'[lat][lon], row order before the tool's flip."""
    z, m = ob.load_tile(name)
    path = str(tmp_path / (name + ".nc"))
    f = netcdf_file(path, "w", version=1)
    f.createDimension("lat", m["n_lat"]); f.createDimension("lon", m["n_lon"])
    v = f.createVariable("lat", "d", ("lat",)); v[:] = np.linspace(m["min_lat"], m["max_lat"], m["n_lat"])
    v = f.createVariable("lon", "d", ("lon",)); v[:] = np.linspace(m["min_lon"], m["max_lon"], m["n_lon"])
    v = f.createVariable("elevation", "h", ("lat", "lon")); v[:] = z[::-1].astype(np.int16)
    f.close()
    return open(path, "rb").read(), z, m


@pytest.mark.parametrize("name,frac", [("mid_atlantic", 0.5), ("mariana", 0.1)])
@pytest.mark.parametrize("dtype_name", ["f64", "f32"])
def test_netcdf_to_masked_grid_to_metrics(auvi, torch, tmp_path, name, frac, dtype_name):
    img, z, m = _tile_as_netcdf(tmp_path, name)
    case = ob.masked_case(name, frac)
    dtype = auvi.F64 if dtype_name == "f64" else auvi.F32
    npdt = np.float64 if dtype_name == "f64" else np.float32
    bounds = case["bounds"]
    g = auvi.Grid.from_netcdf(img, "elevation", bounds, dtype=dtype, flip_rows=True)
    assert (g.n_lat, g.n_lon) == z.shape
    assert bits_equal(g.read(), z)                                          # decode + flip == the tool's DataFrame
    flat = auvi.legacy_choice(z.size, int(z.size * frac), 42)
    assert np.array_equal(flat // m["n_lon"], case["rows"]) and np.array_equal(flat % m["n_lon"], case["cols"])
    truth = g.mask_cells(flat)
    assert bits_equal(truth, case["truth"])                                  # reference_missing.csv, third column
    assert bits_equal(g.read(), case["z"])                                   # reduced_data.csv
    # gap fill of the device-prepared grid == gap fill of the host-uploaded fixture grid
    g_host = auvi.Grid(case["z"].astype(npdt), *bounds)
    d_truth = torch.from_numpy(z.astype(npdt)).cuda()
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_metrics.json")))["computed"].get(f"{name}@{frac:.2f}")
    for meth in (auvi.CUBIC, auvi.KRIGING, auvi.NN, auvi.IDW):
        a = g.lattice(meth, auvi.AXIS_NODES, 1, 1, fill=1)
        b = g_host.lattice(meth, auvi.AXIS_NODES, 1, 1, fill=1)
        assert bits_equal(a, b), auvi.METHOD_NAMES[meth]
        d_fill = torch.from_numpy(a).cuda()
        mae, rmse, mx, n_nan, n = g.fill_metrics_device(d_fill.data_ptr(), g.n_lon, d_truth.data_ptr(), g.n_lon, 0, g.n_lat)
        assert n == flat.size
        est = a[case["rows"], case["cols"]].astype(np.float64)
        d = np.abs(case["truth"] - est)
        ok = ~np.isnan(est)
        np.testing.assert_allclose([mae, rmse, mx], [d[ok].sum() / n, np.sqrt((d[ok] ** 2).sum() / n), np.nanmax(d)], rtol=1e-12)
        assert n_nan == int((~ok).sum())
        if gold and dtype_name == "f64" and auvi.METHOD_NAMES[meth] in gold:
            gm = gold[auvi.METHOD_NAMES[meth]]
            np.testing.assert_allclose([mae, rmse, mx], [gm["mae"], gm["rmse"], gm["max"]], rtol=1e-9)
            assert n_nan == gm["n_nan"]
    g.close(); g_host.close()


def test_mask_hash_is_slab_invariant(auvi, torch):
    """Two ranks holding row slabs (with halo) of one grid draw the same mask as one rank holding all of it."""
    n_lat, n_lon = 300, 257
    base = torch.arange(n_lat * n_lon, dtype=torch.float32, device="cuda").reshape(n_lat, n_lon).contiguous()
    def grid_of(t, row0):
        return auvi.Grid(adopt=dict(ptr=t.data_ptr(), dtype=auvi.F32, n_lat=n_lat, n_lon=n_lon, ld=n_lon, row0=row0,
                                    rows=t.shape[0], keep=t), min_lon=0.0, max_lon=1.0, min_lat=0.0, max_lat=1.0)
    whole = base.clone()
    g = grid_of(whole, 0)
    n_masked = g.mask_hash(0.7, seed=42)
    g.close()
    frac = float(torch.isnan(whole).float().mean())
    assert n_masked == int(torch.isnan(whole).sum()) and abs(frac - 0.7) < 0.01
    for lo, hi in ((0, 160), (140, 300)):
        slab = base[lo:hi].clone()
        gs = grid_of(slab, lo)
        gs.mask_hash(0.7, seed=42, count=False)
        torch.cuda.synchronize()
        gs.close()
        assert torch.equal(torch.isnan(slab), torch.isnan(whole[lo:hi]))
    other = base.clone()
    g = grid_of(other, 0); g.mask_hash(0.7, seed=43); g.close()
    assert not torch.equal(torch.isnan(other), torch.isnan(whole))


def _csv_text(z, fmt):
    return ("\n".join(",".join("nan" if np.isnan(v) else fmt(v) for v in row) for row in z) + "\n").encode()


@pytest.mark.parametrize("style", ["int", "g6", "repr", "exp", "crlf_noeol", "long"])
def test_csv_parsed_on_device_equals_stod(auvi, style):
    """auvi_grid_create_csv against Python's float() (correctly rounded, as glibc strtod / std::stod) cell by cell:
    GEBCO integers, the 6-significant-digit text of generate_csv_grids.cpp, 17-digit repr (beyond the exact fast
    path for most cells: host strtod patch), exponent forms, CRLF without a final newline, >19-digit fields."""
    rng = np.random.RandomState(5)
    n_lat, n_lon = 211, 173
    z = rng.uniform(-11000.0, 500.0, size=(n_lat, n_lon))
    z.ravel()[rng.choice(z.size, 4000, replace=False)] = np.nan
    fmt = {"int": lambda v: str(int(round(v))), "g6": lambda v: "%.6g" % v, "repr": lambda v: repr(float(v)),
           "exp": lambda v: "%.9e" % (v * 1e-12), "crlf_noeol": lambda v: "%.4f" % v,
           "long": lambda v: "%.25f" % v}[style]
    text = _csv_text(z, fmt)
    if style == "crlf_noeol":
        text = text.replace(b"\n", b"\r\n")[:-2]
    want = np.array([[float(c) for c in line.split(",")] for line in text.decode().replace("\r", "").strip().split("\n")])
    g = auvi.Grid.from_csv(text, (0.0, 1.0, 0.0, 1.0))
    assert (g.n_lat, g.n_lon) == (n_lat, n_lon)
    assert bits_equal(g.read(), want), style
    g.close()
    g32 = auvi.Grid.from_csv(text, (0.0, 1.0, 0.0, 1.0), dtype=auvi.F32)
    assert bits_equal(g32.read().astype(np.float64), want.astype(np.float32).astype(np.float64))
    g32.close()


def test_csv_grid_runs_the_grid_b_case(auvi):
    """reduced_data.csv of a fixture case as text -> device grid: identical to the host-uploaded grid."""
    case = ob.masked_case("mid_atlantic", 0.5)
    text = _csv_text(case["z"], lambda v: repr(float(v)))                    # pandas to_csv writes -5559.0
    g = auvi.Grid.from_csv(text, case["bounds"])
    assert bits_equal(g.read(), case["z"])
    h = auvi.Grid(case["z"], *case["bounds"])
    for meth in (auvi.CUBIC, auvi.KRIGING):
        assert bits_equal(g.lattice(meth, auvi.AXIS_NODES, 1, 1, fill=1), h.lattice(meth, auvi.AXIS_NODES, 1, 1, fill=1))
    g.close(); h.close()


def test_csv_errors(auvi):
    with pytest.raises(auvi.AuviError, match="same number of fields"):
        auvi.Grid.from_csv(b"1,2,3\n4,5\n6,7,8\n", (0.0, 1.0, 0.0, 1.0))
    with pytest.raises(auvi.AuviError, match="same number of fields"):
        auvi.Grid.from_csv(b"1,2,3\n4,5,6,7\n8,9\n", (0.0, 1.0, 0.0, 1.0))     # right total, wrong rows
    with pytest.raises(auvi.AuviError, match="not a number"):
        auvi.Grid.from_csv(b"1,2\n3,abc\n", (0.0, 1.0, 0.0, 1.0))
    with pytest.raises(auvi.AuviError, match="not a number"):
        auvi.Grid.from_csv(b"1,,2\n3,4,5\n", (0.0, 1.0, 0.0, 1.0))             # empty field


def test_hash_mask_restatement_equals_the_device_mask():
    """oracle/binding.hash_mask (numpy) is what bench.py's CPU arm masks its window with: it must be the mask
    auvi_grid_mask_hash draws on the device, for a whole grid and for a row slab of it (global flat index)."""
    import auvi
    from oracle import binding as ob
    n_lat, n_lon = 300, 517
    z = np.ones((n_lat, n_lon), dtype=np.float32)
    g = auvi.Grid(z, 0.0, 1.0, 0.0, 1.0)
    n_masked = g.mask_hash(0.7, seed=42)
    dev = np.isnan(g.read())
    want = ob.hash_mask(0, n_lat, n_lon, 0.7, 42)
    assert np.array_equal(dev, want) and n_masked == int(want.sum())
    g.close()
    lib = auvi.load()
    import ctypes as C
    h = C.c_void_p()
    slab = np.ones((50, n_lon), dtype=np.float32)
    assert lib.auvi_grid_create_slab(slab.ctypes.data, auvi.F32, n_lat, n_lon, 100, 50, 0.0, 1.0, 0.0, 1.0, 0, C.byref(h)) == 0
    assert lib.auvi_grid_mask_hash(h, 0.3, 7, None, None) == 0
    got = np.empty((50, n_lon), dtype=np.float32)
    assert lib.auvi_grid_read(h, 100, 150, got.ctypes.data) == 0
    lib.auvi_grid_destroy(h)
    assert np.array_equal(np.isnan(got), ob.hash_mask(100, 150, n_lon, 0.3, 7))
