"""GPU suite: the CUDA path, called through the C-ABI (libauvi.so), against the CPU oracle.

Bars (BASELINE.json north_star):
  * NaN masks and neighbour selections: bit-exact;
  * bilinear / bicubic on FP64 grids: bit-exact (same operation order, no contraction);
  * kriging (exp() differs by <=1 ulp between CUDA and glibc), IDW (FP32 weights), FP32 grids:
    |got - want| <= 1e-3 m + 1e-5 * |want|  (written out below as ATOL / RTOL);
  * MAE / RMSE / Max against the unmasked truth reproduce the reference's published rows.
"""
import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "auv-real-time-interpolation_b200", "python"))

from oracle import binding as ob  # noqa: E402
from conftest import bits_equal  # noqa: E402

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-3          # north_star tolerance for the floating-point methods
TIGHT = 1e-6                     # what FP64 kriging actually achieves (exp ulp differences only)

GOLD = ob.GOLDEN
METHODS = (ob.BILINEAR, ob.CUBIC, ob.KRIGING, ob.NN, ob.IDW)


@pytest.fixture(scope="module")
def auvi():
    import auvi as m
    m.load()
    assert m.device_count() > 0, "GPU tests need a CUDA device"
    return m


@pytest.fixture(scope="module")
def torch():
    import torch as t
    assert t.cuda.is_available()
    return t


def _close(got, want, atol=ATOL, rtol=RTOL):
    assert np.array_equal(np.isnan(got), np.isnan(want)), "NaN mask differs"
    np.testing.assert_allclose(got, want, rtol=rtol, atol=atol, equal_nan=True)


def _device_points(torch, g, method, pts):
    """Device-buffer form with the selection dump."""
    n = pts.shape[0]
    d_pts = torch.from_numpy(np.ascontiguousarray(pts)).cuda()
    d_out = torch.empty(n, dtype=torch.float64, device="cuda")
    d_sel = torch.empty((n, 4, 2), dtype=torch.int32, device="cuda")
    d_found = torch.empty(n, dtype=torch.int32, device="cuda")
    g.interp_points_device(method, d_pts.data_ptr(), n, 24, d_out.data_ptr(), d_sel.data_ptr(), d_found.data_ptr(),
                           torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return d_out.cpu().numpy(), d_sel.cpu().numpy(), d_found.cpu().numpy()


# ---- point-list mode against the golden fixtures of the UNMODIFIED reference -------------------------
POINT_FILES = sorted(glob.glob(os.path.join(GOLD, "points_*.npz")))


@pytest.mark.parametrize("fn", POINT_FILES, ids=[os.path.basename(f)[7:-4] for f in POINT_FILES])
def test_points_match_reference_golden(auvi, torch, fn):
    name, pct = os.path.basename(fn)[7:-4].rsplit("_", 1)
    case = ob.masked_case(name, int(pct) / 100.0)
    rec = np.load(fn)
    g = auvi.Grid(case["z"], *case["bounds"])
    pts = rec["pts"]
    assert bits_equal(g.interp_points(auvi.BILINEAR, pts), rec["bilinear"])
    assert bits_equal(g.interp_points(auvi.CUBIC, pts), rec["cubic"])
    _close(g.interp_points(auvi.KRIGING, pts), rec["kriging"], atol=TIGHT, rtol=0)
    # neighbour selection: bit-exact and in the reference's order
    _, sel_c, found_c = _device_points(torch, g, auvi.CUBIC, pts)
    searched = found_c >= 0
    assert np.array_equal(found_c[searched], rec["found_floor"][searched])
    assert np.array_equal(sel_c[searched], rec["sel_floor"][searched])
    for meth in (auvi.KRIGING, auvi.NN, auvi.IDW):
        _, sel, found = _device_points(torch, g, meth, pts)
        assert np.array_equal(found, rec["found_round"])
        assert np.array_equal(sel, rec["sel_round"])
    g.close()


# ---- point-list mode against the live oracle: whole cases, all five methods ----------------------------
@pytest.mark.parametrize("name,frac", [("mini", 0.3), ("mid_atlantic", 0.1), ("mid_atlantic", 0.5),
                                       ("mid_atlantic", 0.9), ("mid_atlantic", 0.97), ("mariana", 0.5)])
def test_points_match_oracle_full_case(auvi, torch, name, frac):
    case = ob.masked_case(name, frac)
    orc = ob.Oracle(case["z"], *case["bounds"])
    g = auvi.Grid(case["z"], *case["bounds"])
    rng = np.random.RandomState(5)
    lo_lon, hi_lon, lo_lat, hi_lat = case["bounds"]
    n = 30000
    rnd = np.zeros((n, 3))
    rnd[:, 0] = rng.uniform(lo_lon - 0.01, hi_lon + 0.01, n)     # some out of bounds -> NaN
    rnd[:, 1] = rng.uniform(lo_lat - 0.01, hi_lat + 0.01, n)
    rnd[0, :2] = (lo_lon, lo_lat)
    rnd[1, :2] = (hi_lon, hi_lat)                                 # query exactly on the max bound
    rnd[2, :2] = (lo_lon, hi_lat)
    for pts in (case["pts"], rnd):
        for meth in METHODS:
            want, sel_w, found_w = orc.batch(meth, pts, want_sel=True)
            got = g.interp_points(meth, pts)
            if meth in (ob.BILINEAR, ob.CUBIC, ob.NN):
                assert bits_equal(got, want), ob.METHOD_NAMES[meth]
            elif meth == ob.KRIGING:
                _close(got, want, atol=TIGHT, rtol=0)
            else:
                _close(got, want)
            if meth != ob.BILINEAR:
                got_d, sel, found = _device_points(torch, g, meth, pts[:20000])
                assert bits_equal(got_d, got[:20000])
                assert np.array_equal(found, found_w[:20000]), ob.METHOD_NAMES[meth]
                ok = found >= 0
                assert np.array_equal(sel[ok], sel_w[:20000][ok]), ob.METHOD_NAMES[meth]
    g.close()


def test_points_edge_cases(auvi):
    """Empty input, one point, strides, all-NaN neighbourhoods, found < 4."""
    z = np.full((40, 50), np.nan)
    z[3, 4] = -100.0
    z[30, 45] = -200.0
    z[31, 45] = -300.0
    bounds = (10.0, 11.0, -5.0, -4.0)
    orc = ob.Oracle(z, *bounds)
    g = auvi.Grid(z, *bounds)
    assert g.interp_points(auvi.KRIGING, np.zeros((0, 3))).shape == (0,)
    rng = np.random.RandomState(1)
    pts = np.zeros((5000, 3))
    pts[:, 0] = rng.uniform(10.0, 11.0, 5000)
    pts[:, 1] = rng.uniform(-5.0, -4.0, 5000)
    for meth in METHODS:
        want = orc.batch(meth, pts)
        got = g.interp_points(meth, pts)
        if meth == ob.IDW:
            _close(got, want)
        else:
            assert bits_equal(got, want), ob.METHOD_NAMES[meth]      # found<4 -> plain means, exact
        assert bits_equal(g.interp_points(meth, pts[:1]), want[:1])
    # wide records: stride 40 bytes
    wide = np.zeros((5000, 5))
    wide[:, :2] = pts[:, :2]
    assert bits_equal(g.interp_points(auvi.CUBIC, wide), orc.batch(ob.CUBIC, pts))
    g.close()


def test_points_large_batch_pipeline(auvi):
    """More than two pipeline chunks (3.3 M points): chunk seams and buffer reuse."""
    z = ob.synth_grid(200, 160)
    bounds = (-180.0, -160.0, 20.0, 30.0)
    g = auvi.Grid(z, *bounds)
    orc = ob.Oracle(z, *bounds)
    rng = np.random.RandomState(2)
    n = 3_300_000
    pts = np.zeros((n, 3))
    pts[:, 0] = rng.uniform(-180.0, -160.0, n)
    pts[:, 1] = rng.uniform(20.0, 30.0, n)
    got = g.interp_points(auvi.BILINEAR, pts)
    idx = np.concatenate([np.arange(0, 5000), np.arange((1 << 20) - 2500, (1 << 20) + 2500),
                          np.arange((2 << 20) - 2500, (2 << 20) + 2500), np.arange(n - 5000, n)])
    assert bits_equal(got[idx], orc.batch(ob.BILINEAR, pts[idx]))
    full = orc.batch(ob.BILINEAR, pts[::7])
    assert bits_equal(got[::7], full)
    g.close()


# ---- lattice mode --------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["small", "holes"])
def test_lattice_matches_reference_golden(auvi, tag):
    rec = np.load(os.path.join(GOLD, f"lattice_{tag}.npz"))
    z = rec["z"]
    bounds = (-180.0, -160.0, 20.0, 30.0)
    nn_lat, nn_lon = rec["nn"]
    g = auvi.Grid(z, *bounds)
    assert g.lattice_dims(auvi.AXIS_EXPANDED, 2, 2) == (nn_lat, nn_lon)
    assert bits_equal(g.lattice(auvi.BILINEAR, auvi.AXIS_EXPANDED, 2, 2).ravel(), rec["bilinear"])
    assert bits_equal(g.lattice(auvi.CUBIC, auvi.AXIS_EXPANDED, 2, 2).ravel(), rec["cubic"])
    _close(g.lattice(auvi.KRIGING, auvi.AXIS_EXPANDED, 2, 2).ravel(), rec["kriging"], atol=TIGHT, rtol=0)
    g.close()


@pytest.mark.parametrize("n_lat,n_lon,f_lat,f_lon,holes", [(320, 400, 2, 2, 0), (97, 131, 4, 4, 0), (97, 131, 4, 1, 25),
                                                            (300, 517, 3, 5, 400), (64, 1030, 2, 2, 50)])
def test_lattice_f64_matches_oracle(auvi, n_lat, n_lon, f_lat, f_lon, holes):
    z = ob.synth_grid(n_lat, n_lon)
    if holes:
        z.ravel()[np.random.RandomState(9).choice(z.size, holes, replace=False)] = np.nan
    bounds = (-180.0, -160.0, 20.0, 30.0)
    pts, nn_lat, nn_lon = ob.lattice_queries(n_lat, n_lon, *bounds, f_lat=f_lat, f_lon=f_lon)
    orc = ob.Oracle(z, *bounds)
    g = auvi.Grid(z, *bounds)
    for meth in METHODS:
        got = g.lattice(meth, auvi.AXIS_EXPANDED, f_lat, f_lon)
        want = orc.batch(meth, pts).reshape(nn_lat, nn_lon)
        if meth in (ob.BILINEAR, ob.CUBIC, ob.NN):
            assert bits_equal(got, want), ob.METHOD_NAMES[meth]
        elif meth == ob.KRIGING:
            _close(got, want, atol=TIGHT, rtol=0)
        else:
            _close(got, want)
    # a row range (the unit of multi-GPU sharding) equals the same rows of the full result
    part = g.lattice(auvi.CUBIC, auvi.AXIS_EXPANDED, f_lat, f_lon, row_begin=7, row_end=nn_lat - 5)
    assert bits_equal(part, g.lattice(auvi.CUBIC, auvi.AXIS_EXPANDED, f_lat, f_lon)[7:nn_lat - 5])
    g.close()


@pytest.mark.parametrize("f", [2, 4])
def test_lattice_f32_within_tolerance(auvi, f):
    z = ob.synth_grid(256, 300).astype(np.float32)
    z.ravel()[np.random.RandomState(4).choice(z.size, 60, replace=False)] = np.nan
    bounds = (-180.0, -160.0, 20.0, 30.0)
    pts, nn_lat, nn_lon = ob.lattice_queries(256, 300, *bounds, f_lat=f, f_lon=f)
    orc = ob.Oracle(z.astype(np.float64), *bounds)
    g = auvi.Grid(z, *bounds)
    assert g.dtype == auvi.F32
    for meth in METHODS:
        got = g.lattice(meth, auvi.AXIS_EXPANDED, f, f)
        assert got.dtype == np.float32
        want = orc.batch(meth, pts).reshape(nn_lat, nn_lon)
        _close(got.astype(np.float64), want)
    g.close()


@pytest.mark.parametrize("n_lat,n_lon,f_lat,f_lon,holes,bounds", [
    (200, 700, 2, 2, 0, (-180.0, -160.0, 20.0, 30.0)),     # noisy node columns: floor() one below on ~half of them
    (200, 701, 2, 2, 40, (-180.0, -160.0, 20.0, 30.0)),    # NaN holes: zero-weight window words that are NaN -> exact re-evaluation
    (90, 1500, 4, 1, 0, (-180.0, -160.0, 20.0, 30.0)),     # longitude factor 1: eight-word windows, three tiles wide
    (90, 1027, 4, 1, 30, (0.0, 1.0, 0.0, 1.0)),            # noise-free axis, ragged last thread
    (150, 515, 1, 2, 10, (100.0, 110.0, -10.0, 0.0)),
    (64, 300, 3, 2, 5, (-30.9967, -29.4993, -0.5035, 1.0071)),
])
def test_lattice_f32_bicubic_window_loads(auvi, n_lat, n_lon, f_lat, f_lon, holes, bounds):
    """The window-load form of the FP32 bicubic kernel (longitude factors 1 and 2: vector shared-memory loads + five shifted
    weights per column, csrc/upsample.cu) against the FP64 oracle on every cell, and that it is the form that ran; the same
    grids through a longitude factor the form does not cover take the generic kernel."""
    z = ob.synth_grid(n_lat, n_lon).astype(np.float32)
    if holes:
        z.ravel()[np.random.RandomState(11).choice(z.size, holes, replace=False)] = np.nan
    pts, nn_lat, nn_lon = ob.lattice_queries(n_lat, n_lon, *bounds, f_lat=f_lat, f_lon=f_lon)
    orc = ob.Oracle(z.astype(np.float64), *bounds)
    g = auvi.Grid(z, *bounds)
    got = g.lattice(auvi.CUBIC, auvi.AXIS_EXPANDED, f_lat, f_lon)
    assert g.uses_window == (1 if f_lon == 1 else 2), "expected the window-load kernel"
    want = orc.batch(ob.CUBIC, pts).reshape(nn_lat, nn_lon)
    _close(got.astype(np.float64), want)
    g.lattice(auvi.CUBIC, auvi.AXIS_EXPANDED, f_lat, 4)
    assert g.uses_window == 0
    g.lattice(auvi.BILINEAR, auvi.AXIS_EXPANDED, f_lat, f_lon)
    assert g.uses_window == 0
    g.close()


@pytest.mark.parametrize("name,frac", [("mid_atlantic", 0.5), ("mariana", 0.5), ("mid_atlantic", 0.9)])
def test_gap_fill_nodes_matches_oracle(auvi, torch, name, frac):
    """Grid-B as a full-grid fill: every masked cell gets method(node query); valid cells pass through."""
    case = ob.masked_case(name, frac)
    m = case["meta"]
    orc = ob.Oracle(case["z"], *case["bounds"])
    g = auvi.Grid(case["z"], *case["bounds"])
    rows, cols = case["rows"], case["cols"]
    for meth in METHODS:
        filled = g.lattice(meth, auvi.AXIS_NODES, 1, 1, fill=1)
        want = orc.batch(meth, case["pts"])
        got = filled[rows, cols]
        if meth in (ob.BILINEAR, ob.CUBIC, ob.NN):
            assert bits_equal(got, want), ob.METHOD_NAMES[meth]
        elif meth == ob.KRIGING:
            _close(got, want, atol=TIGHT, rtol=0)
        else:
            _close(got, want)
        keep = ~np.isnan(case["z"])
        assert bits_equal(filled[keep], case["z"][keep])
    # selection dump of the fill kernel
    n_cells = m["n_lat"] * m["n_lon"]
    d_out = torch.empty((m["n_lat"], m["n_lon"]), dtype=torch.float64, device="cuda")
    d_sel = torch.empty((n_cells, 9), dtype=torch.int32, device="cuda")
    g.lattice_device(auvi.IDW, auvi.AXIS_NODES, 1, 1, 1, 0, m["n_lat"], d_out.data_ptr(), m["n_lon"],
                     d_sel.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    sel = d_sel.cpu().numpy().reshape(m["n_lat"], m["n_lon"], 9)[rows, cols]
    _, sel_w, found_w = orc.batch(ob.IDW, case["pts"], want_sel=True)
    assert np.array_equal(sel[:, 0], found_w)
    assert np.array_equal(sel[:, 1:].reshape(-1, 4, 2), sel_w)
    g.close()


@pytest.mark.parametrize("ld", [260, 261])
def test_row_slabs_equal_whole_grid(auvi, torch, ld):
    """Two row slabs with halos (what two ranks hold) reproduce the single-GPU result bit for bit.  ld = 261 gives the
    slabs a row pitch TMA cannot address: the plain-load staging path, whose boxes must stay inside the slab."""
    z = ob.synth_grid(301, 260).astype(np.float32)
    z.ravel()[np.random.RandomState(8).choice(z.size, 3000, replace=False)] = np.nan
    bounds = (0.0, 2.0, 40.0, 43.0)
    whole = auvi.Grid(z, *bounds)
    dz = torch.full((301, ld), float("nan"), dtype=torch.float32, device="cuda")
    dz[:, :260] = torch.from_numpy(z).cuda()
    halo = 12
    for meth, kind, f, fill in ((auvi.CUBIC, auvi.AXIS_EXPANDED, 4, 0), (auvi.BILINEAR, auvi.AXIS_EXPANDED, 4, 0),
                                (auvi.CUBIC, auvi.AXIS_EXPANDED, 2, 0),      # the window-load kernel on slabs
                                (auvi.IDW, auvi.AXIS_NODES, 1, 1), (auvi.KRIGING, auvi.AXIS_NODES, 1, 1)):
        ref = whole.lattice(meth, kind, f, f, fill=fill)
        out_rows = ref.shape[0]
        cut = out_rows // 2
        parts = []
        for (lo, hi) in ((0, cut), (cut, out_rows)):
            in_lo = max(0, lo // f - halo)
            in_hi = min(301, (hi - 1) // f + 1 + halo + 1)
            slab = dz[in_lo:in_hi]
            g = auvi.Grid(adopt=dict(ptr=slab.data_ptr(), dtype=auvi.F32, n_lat=301, n_lon=260, ld=ld, row0=in_lo,
                                     rows=in_hi - in_lo, keep=slab), min_lon=bounds[0], max_lon=bounds[1],
                          min_lat=bounds[2], max_lat=bounds[3])
            parts.append(g.lattice(meth, kind, f, f, fill=fill, row_begin=lo, row_end=hi))
            assert bool(g.uses_tma) == (ld % 4 == 0)
            g.close()
        got = np.concatenate(parts, axis=0)
        assert bits_equal(got.astype(np.float64), ref.astype(np.float64)), auvi.METHOD_NAMES[meth]
    # a slab without enough halo is refused, not silently wrong
    slab = dz[100:200]
    g = auvi.Grid(adopt=dict(ptr=slab.data_ptr(), dtype=auvi.F32, n_lat=301, n_lon=260, ld=ld, row0=100, rows=100,
                             keep=slab), min_lon=bounds[0], max_lon=bounds[1], min_lat=bounds[2], max_lat=bounds[3])
    with pytest.raises(auvi.AuviError, match="halo"):
        g.lattice(auvi.IDW, auvi.AXIS_NODES, 1, 1, fill=1, row_begin=100, row_end=200)
    g.close()
    whole.close()


# ---- metrics ---------------------------------------------------------------------------------------------
with open(os.path.join(GOLD, "golden_metrics.json")) as _f:
    _G = json.load(_f)


@pytest.mark.parametrize("case", sorted(k for k in _G["published"] if not k.startswith("us_east@0.1") and
                                        not k.startswith("us_east@0.2")))
def test_published_metrics_reproduced_on_device(auvi, torch, case):
    """The reference author's MAE/RMSE/Max rows (results/TestingResults1.csv), computed entirely on the
    GPU: GridD-style batch -> device error metrics, printed to the same 6 significant digits."""
    name, frac = case.split("@")
    c = ob.masked_case(name, float(frac))
    g = auvi.Grid(c["z"], *c["bounds"])
    d_truth = torch.from_numpy(c["truth"]).cuda()
    for meth in (ob.BILINEAR, ob.CUBIC, ob.KRIGING):
        est = g.interp_points(meth, c["pts"])
        d_est = torch.from_numpy(est).cuda()
        mae, rmse, mx, n_nan = auvi.error_metrics_device(d_truth.data_ptr(), d_est.data_ptr(), auvi.F64, est.size)
        got = ["%g" % v for v in (mae, rmse, mx)]
        for machine, want in _G["published"][case][ob.METHOD_NAMES[meth]].items():
            assert got == want, (case, ob.METHOD_NAMES[meth], machine, got, want)
        assert n_nan == int(np.isnan(est).sum())
    g.close()


@pytest.mark.parametrize("case", ["mariana@0.50", "mid_atlantic@0.90", "east_pacific@0.10"])
def test_computed_metrics_reproduced_on_device(auvi, torch, case):
    name, frac = case.split("@")
    c = ob.masked_case(name, float(frac))
    g = auvi.Grid(c["z"], *c["bounds"])
    d_truth = torch.from_numpy(c["truth"]).cuda()
    for meth in (ob.BILINEAR, ob.CUBIC, ob.KRIGING):
        est = g.interp_points(meth, c["pts"])
        d_est = torch.from_numpy(est).cuda()
        mae, rmse, mx, n_nan = auvi.error_metrics_device(d_truth.data_ptr(), d_est.data_ptr(), auvi.F64, est.size)
        want = _G["computed"][case][ob.METHOD_NAMES[meth]]
        assert n_nan == want["n_nan"]
        np.testing.assert_allclose([mae, rmse, mx], [want["mae"], want["rmse"], want["max"]], rtol=1e-9)
    g.close()


# ---- BASELINE sizes: size-independent properties ------------------------------------------------------------
def test_config4_upsample_properties_at_full_size(auvi, torch):
    """16384 x 16384 FP32, 4x in both axes (BASELINE config 4).  Properties: (1) both stencils
    interpolate, so every 4th output equals its input node; (2) random windows agree with the oracle;
    (3) bilinear is linear: upsample(a*z + b) == a*upsample(z) + b within FP32 rounding."""
    n = 16384
    f = 4
    ii = torch.arange(n, device="cuda", dtype=torch.float32) * (100.0 / (n - 1))
    z = -(10.0 + 2.0 * ii)[None, :] + 100.0 * torch.exp(-(((ii - 75.0) ** 2)[None, :] + ((ii - 50.0) ** 2)[:, None]) / 450.0)
    z = z.contiguous()
    bounds = (-180.0, -160.0, 20.0, 30.0)
    g = auvi.Grid(adopt=dict(ptr=z.data_ptr(), dtype=auvi.F32, n_lat=n, n_lon=n, ld=n, row0=0, rows=n, keep=z),
                  min_lon=bounds[0], max_lon=bounds[1], min_lat=bounds[2], max_lat=bounds[3])
    rows, cols = g.lattice_dims(auvi.AXIS_EXPANDED, f, f)
    assert (rows, cols) == (65533, 65533)
    ld = 65536
    out = torch.empty((rows, ld), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    lat_ax = ob.lattice_axis(bounds[2], bounds[3], rows)
    lon_ax = ob.lattice_axis(bounds[0], bounds[1], cols)
    assert lat_ax.max() <= bounds[3] and lon_ax.max() <= bounds[1]      # no lattice node falls out of bounds
    orc = ob.Oracle(z.cpu().numpy().astype(np.float64), *bounds)          # FP32 grid widened to FP64
    for meth in (auvi.BILINEAR, auvi.CUBIC):
        out.fill_(7.0)
        g.lattice_device(meth, auvi.AXIS_EXPANDED, f, f, 0, 0, rows, out.data_ptr(), ld, None, st)
        torch.cuda.synchronize()
        assert g.uses_tma, "the tiled upsample kernel should stage its input by TMA here"
        assert not torch.isnan(out[:, :cols]).any()
        assert bool((out[:, cols:] == 7.0).all()), "padding columns were written"
        err = (out[::f, :cols:f] - z).abs().max().item()
        assert err <= 1e-3, err
        rng = np.random.RandomState(3)
        for r0, c0 in [(0, 0), (rows - 96, cols - 96)] + [(int(rng.randint(0, rows - 96)), int(rng.randint(0, cols - 96)))
                                                          for _ in range(3)]:
            J, I = np.arange(r0, r0 + 96), np.arange(c0, c0 + 96)
            pts = np.zeros((J.size * I.size, 3))
            pts[:, 0] = np.tile(lon_ax[I], J.size)
            pts[:, 1] = np.repeat(lat_ax[J], I.size)
            want = orc.batch(meth, pts).reshape(J.size, I.size)
            got = out[r0:r0 + 96, c0:c0 + 96].cpu().numpy().astype(np.float64)
            _close(got, want)
    # linearity of bilinear
    z2 = (z * 0.5 + 3.0).contiguous()
    g2 = auvi.Grid(adopt=dict(ptr=z2.data_ptr(), dtype=auvi.F32, n_lat=n, n_lon=n, ld=n, row0=0, rows=n, keep=z2),
                   min_lon=bounds[0], max_lon=bounds[1], min_lat=bounds[2], max_lat=bounds[3])
    sub = 2048
    a = torch.empty((sub, ld), dtype=torch.float32, device="cuda")
    b = torch.empty((sub, ld), dtype=torch.float32, device="cuda")
    g.lattice_device(auvi.BILINEAR, auvi.AXIS_EXPANDED, f, f, 0, 30000, 30000 + sub, a.data_ptr(), ld, None, st)
    g2.lattice_device(auvi.BILINEAR, auvi.AXIS_EXPANDED, f, f, 0, 30000, 30000 + sub, b.data_ptr(), ld, None, st)
    torch.cuda.synchronize()
    assert ((a[:, :cols] * 0.5 + 3.0) - b[:, :cols]).abs().max().item() <= 1e-3
    g.close()
    g2.close()


def test_config5_gap_fill_properties_at_scale(auvi, torch):
    """A 8192 x 8192 FP32 slice of BASELINE config 5 (70 % mask): (1) valid cells pass through
    untouched; (2) no NaN remains (found < 4 never happens at 70 %); (3) idempotence: filling the filled
    grid changes nothing; (4) NN output values all occur among the valid inputs of their window;
    (5) random windows agree with the oracle."""
    n = 8192
    gen = torch.Generator(device="cuda")
    gen.manual_seed(42)
    ii = torch.arange(n, device="cuda", dtype=torch.float32) * (100.0 / (n - 1))
    z = -(10.0 + 2.0 * ii)[None, :] + 100.0 * torch.exp(-(((ii - 75.0) ** 2)[None, :] + ((ii - 50.0) ** 2)[:, None]) / 450.0)
    mask = torch.rand((n, n), device="cuda", generator=gen) < 0.70
    zm = torch.where(mask, torch.full_like(z, float("nan")), z).contiguous()
    bounds = (100.0, 110.0, -10.0, 0.0)
    g = auvi.Grid(adopt=dict(ptr=zm.data_ptr(), dtype=auvi.F32, n_lat=n, n_lon=n, ld=n, row0=0, rows=n, keep=zm),
                  min_lon=bounds[0], max_lon=bounds[1], min_lat=bounds[2], max_lat=bounds[3])
    st = torch.cuda.current_stream().cuda_stream
    out = torch.empty((n, n), dtype=torch.float32, device="cuda")
    zh = zm.cpu().numpy().astype(np.float64)
    orc = ob.Oracle(zh, *bounds)
    lat_ax = ob.node_axis(bounds[2], bounds[3], n)
    lon_ax = ob.node_axis(bounds[0], bounds[1], n)
    for meth in (auvi.IDW, auvi.NN, auvi.KRIGING, auvi.CUBIC):
        g.lattice_device(meth, auvi.AXIS_NODES, 1, 1, 1, 0, n, out.data_ptr(), n, None, st)
        torch.cuda.synchronize()
        assert bool((out[~mask] == zm[~mask]).all())
        # lattice rows/cols whose node coordinate rounds past the max bound are NaN in the reference too
        inb = torch.ones((n, n), dtype=torch.bool, device="cuda")
        inb[torch.from_numpy(lat_ax > bounds[3]).cuda(), :] = False
        inb[:, torch.from_numpy(lon_ax > bounds[1]).cuda()] = False
        assert not torch.isnan(out[inb]).any()
        rng = np.random.RandomState(6)
        for _ in range(2):
            r0, c0 = int(rng.randint(0, n - 64)), int(rng.randint(0, n - 64))
            J, I = np.arange(r0, r0 + 64), np.arange(c0, c0 + 64)
            pts = np.zeros((64 * 64, 3))
            pts[:, 0] = np.tile(lon_ax[I], 64)
            pts[:, 1] = np.repeat(lat_ax[J], 64)
            want = orc.batch(meth, pts).reshape(64, 64)
            win = zh[r0:r0 + 64, c0:c0 + 64]
            want = np.where(np.isnan(win), want, win)
            _close(out[r0:r0 + 64, c0:c0 + 64].cpu().numpy().astype(np.float64), want)
        if meth == auvi.NN:
            g_f = auvi.Grid(adopt=dict(ptr=out.data_ptr(), dtype=auvi.F32, n_lat=n, n_lon=n, ld=n, row0=0, rows=n,
                                       keep=out), min_lon=bounds[0], max_lon=bounds[1], min_lat=bounds[2],
                            max_lat=bounds[3])
            again = torch.empty_like(out)
            g_f.lattice_device(meth, auvi.AXIS_NODES, 1, 1, 1, 0, n, again.data_ptr(), n, None, st)
            torch.cuda.synchronize()
            same = (again == out) | (torch.isnan(again) & torch.isnan(out))
            assert bool(same.all()), "gap fill is not idempotent"
            g_f.close()
    g.close()


# ---- BASELINE configs 0/2 at their full sizes ---------------------------------------------------------------
@pytest.mark.parametrize("name", ["mariana", "east_pacific", "mid_atlantic"])
@pytest.mark.parametrize("frac", [0.10, 0.50, 0.90])
def test_config2_all_regions_every_method(auvi, torch, name, frac):
    """BASELINE config 2: every bundled region (Kerguelen's tile is missing from the checkout) at 10/50/90 %
    removal, every method, as a full-grid gap fill: per-point parity with the oracle on EVERY removed cell
    and the MAE/RMSE/Max/NaN-count rows computed with the unmodified reference (golden_metrics.json)."""
    case = ob.masked_case(name, frac)
    m = case["meta"]
    orc = ob.Oracle(case["z"], *case["bounds"])
    g = auvi.Grid(case["z"], *case["bounds"])
    rows, cols = case["rows"], case["cols"]
    d_truth = torch.from_numpy(case["truth"]).cuda()
    key = f"{name}@{frac:.2f}"
    for meth in METHODS:
        filled = g.lattice(meth, auvi.AXIS_NODES, 1, 1, fill=1)
        got = filled[rows, cols]
        want = orc.batch(meth, case["pts"])
        if meth in (ob.BILINEAR, ob.CUBIC, ob.NN):
            assert bits_equal(got, want), (key, ob.METHOD_NAMES[meth])
        elif meth == ob.KRIGING:
            _close(got, want, atol=TIGHT, rtol=0)
        else:
            _close(got, want)
        if meth in (ob.BILINEAR, ob.CUBIC, ob.KRIGING):
            d_est = torch.from_numpy(np.ascontiguousarray(got)).cuda()
            mae, rmse, mx, n_nan = auvi.error_metrics_device(d_truth.data_ptr(), d_est.data_ptr(), auvi.F64, got.size)
            gold = _G["computed"][key][ob.METHOD_NAMES[meth]]
            assert n_nan == gold["n_nan"] and got.size == gold["n"]
            np.testing.assert_allclose([mae, rmse, mx], [gold["mae"], gold["rmse"], gold["max"]], rtol=1e-9)
    g.close()


def test_config0_grid_a_full_size(auvi):
    """BASELINE config 0 at the reference generator's shipped size: the 4000 x 3200 synthetic Grid A
    (generate_csv_grids.cpp:99-104, values through the 6-digit CSV), 2x expanded lattice 7999 x 6399 =
    51.2 M queries (test_interpolation.cpp:283-297).  Bilinear and bicubic: bit-identical to the oracle on
    every cell; kriging / NN / IDW on a band of rows."""
    n_lat, n_lon = 3200, 4000
    z = ob.synth_grid(n_lat, n_lon)
    bounds = (-180.0, -160.0, 20.0, 30.0)
    g = auvi.Grid(z, *bounds)
    orc = ob.Oracle(z, *bounds)
    nn_lat, nn_lon = g.lattice_dims(auvi.AXIS_EXPANDED, 2, 2)
    assert (nn_lat, nn_lon) == (6399, 7999)
    lat_ax = ob.lattice_axis(bounds[2], bounds[3], nn_lat)
    lon_ax = ob.lattice_axis(bounds[0], bounds[1], nn_lon)
    for meth in (ob.BILINEAR, ob.CUBIC):
        got = g.lattice(meth, auvi.AXIS_EXPANDED, 2, 2)
        for r0 in range(0, nn_lat, 800):                       # oracle in row blocks to bound host memory
            r1 = min(nn_lat, r0 + 800)
            pts = np.zeros(((r1 - r0) * nn_lon, 3))
            pts[:, 0] = np.tile(lon_ax, r1 - r0)
            pts[:, 1] = np.repeat(lat_ax[r0:r1], nn_lon)
            assert bits_equal(got[r0:r1].ravel(), orc.batch(meth, pts)), (ob.METHOD_NAMES[meth], r0)
    r0, r1 = 3000, 3060
    pts = np.zeros(((r1 - r0) * nn_lon, 3))
    pts[:, 0] = np.tile(lon_ax, r1 - r0)
    pts[:, 1] = np.repeat(lat_ax[r0:r1], nn_lon)
    for meth in (ob.KRIGING, ob.NN, ob.IDW):
        got = g.lattice(meth, auvi.AXIS_EXPANDED, 2, 2, row_begin=r0, row_end=r1).ravel()
        want = orc.batch(meth, pts)
        if meth == ob.NN:
            assert bits_equal(got, want)
        else:
            _close(got, want, atol=TIGHT if meth == ob.KRIGING else ATOL, rtol=0 if meth == ob.KRIGING else RTOL)
    g.close()


# ---- the tiled gap-fill kernel against the per-query exact path on whole grids ---------------------------------------
@pytest.mark.parametrize("dtype_name", ["f32", "f64"])
@pytest.mark.parametrize("frac", [0.2, 0.5, 0.7, 0.9, 0.97, 0.997])
def test_fill_tiled_equals_exact_path_on_whole_grids(auvi, torch, frac, dtype_name):
    """Every path of fill_tiled_kernel (near path, sqrt replay, general path, fewer-than-four, literal hand-backs)
    against lattice_exact_kernel -- the literal restatement checked against the oracle elsewhere -- on every cell:
    masks from 20 % (many candidates) to 99.7 % (searches run out of rings).  Also run-to-run determinism."""
    n_lat, n_lon = 1100, 1300
    tdt = torch.float32 if dtype_name == "f32" else torch.float64
    dt = auvi.F32 if dtype_name == "f32" else auvi.F64
    jj = torch.arange(n_lat, device="cuda", dtype=torch.float64)[:, None]
    ii = torch.arange(n_lon, device="cuda", dtype=torch.float64)[None, :]
    z = (-4000.0 + 900.0 * torch.sin(ii * 0.013) * torch.cos(jj * 0.017) + 0.37 * ii - 0.21 * jj).to(tdt).contiguous()
    g = auvi.Grid(adopt=dict(ptr=z.data_ptr(), dtype=dt, n_lat=n_lat, n_lon=n_lon, ld=n_lon, row0=0, rows=n_lat, keep=z),
                  min_lon=-30.9967, max_lon=-29.4993, min_lat=-0.5035, max_lat=1.0071)
    g.mask_hash(frac, seed=7, count=False)
    st = torch.cuda.current_stream().cuda_stream
    a = torch.empty((n_lat, n_lon), dtype=tdt, device="cuda")
    b = torch.empty_like(a)
    c = torch.empty_like(a)
    sel = torch.empty((n_lat * n_lon, 9), dtype=torch.int32, device="cuda")
    for meth in (auvi.BILINEAR, auvi.NN, auvi.CUBIC, auvi.IDW, auvi.KRIGING):
        g.lattice_device(meth, auvi.AXIS_NODES, 1, 1, 1, 0, n_lat, a.data_ptr(), n_lon, None, st)
        g.lattice_device(meth, auvi.AXIS_NODES, 1, 1, 1, 0, n_lat, c.data_ptr(), n_lon, None, st)
        g.lattice_device(meth, auvi.AXIS_NODES, 1, 1, 1, 0, n_lat, b.data_ptr(), n_lon, sel.data_ptr(), st)   # exact path
        torch.cuda.synchronize()
        same_run = (a == c) | (torch.isnan(a) & torch.isnan(c))
        assert bool(same_run.all()), "tiled fill is not deterministic"
        assert torch.equal(torch.isnan(a), torch.isnan(b)), auvi.METHOD_NAMES[meth]
        ok = ~torch.isnan(a)
        if meth in (auvi.BILINEAR, auvi.NN, auvi.CUBIC):
            assert torch.equal(a[ok], b[ok]), auvi.METHOD_NAMES[meth]
        else:
            tol = (1e-3 if dtype_name == "f32" or meth == auvi.IDW else 1e-6)
            err = (a[ok].double() - b[ok].double()).abs() - 1e-5 * b[ok].double().abs() * (meth == auvi.IDW or dtype_name == "f32")
            assert float(err.max()) <= tol, (auvi.METHOD_NAMES[meth], float(err.max()))
    g.close()


def test_config4_full_size_strips_against_exact_path(auvi, torch):
    """BASELINE config 4 at its stated size -- 65536 x 65536 FP32, 70 % mask drawn by the device hash -- filled by the
    tiled kernel in ONE launch (4.29 G cells: every flat index beyond 2^31 is exercised), then row strips spread over the
    grid (first rows, around 2^31 / n_lon, last rows) are recomputed by the per-query exact path and compared:
    NN bit for bit, IDW within the north-star tolerance; valid cells pass through; no NaN remains."""
    n = 65536
    free, _ = torch.cuda.mem_get_info()
    if free < 60 * (1 << 30):
        pytest.skip("needs ~40 GB of device memory")
    z = torch.empty((n, n), dtype=torch.float32, device="cuda")
    i = torch.arange(n, device="cuda", dtype=torch.float64) * (100.0 / (n - 1))
    base = -(10.0 + 2.0 * i)[None, :]
    gx = ((i - 75.0) ** 2)[None, :]
    for r in range(0, n, 2048):
        j = i[r:r + 2048]
        z[r:r + 2048] = (base + 100.0 * torch.exp(-(gx + ((j - 50.0) ** 2)[:, None]) / 450.0)).float()
    truth_strips = {}
    strips = [0, 4096 - 32, 32768 - 7, 32768 + 1000, n - 64]
    for r in strips:
        truth_strips[r] = z[r:r + 64].clone()
    g = auvi.Grid(adopt=dict(ptr=z.data_ptr(), dtype=auvi.F32, n_lat=n, n_lon=n, ld=n, row0=0, rows=n, keep=z),
                  min_lon=100.0, max_lon=110.0, min_lat=-10.0, max_lat=0.0)
    n_masked = g.mask_hash(0.70, seed=42)
    assert abs(n_masked / (n * n) - 0.70) < 1e-3
    st = torch.cuda.current_stream().cuda_stream
    out = torch.empty((n, n), dtype=torch.float32, device="cuda")
    ref = torch.empty((64, n), dtype=torch.float32, device="cuda")
    sel = torch.empty((64 * n, 9), dtype=torch.int32, device="cuda")
    lon_ok = torch.from_numpy(ob.node_axis(100.0, 110.0, n) <= 110.0).cuda()
    lat_ax_ok = ob.node_axis(-10.0, 0.0, n) <= 0.0
    for meth in (auvi.NN, auvi.IDW):
        g.lattice_device(meth, auvi.AXIS_NODES, 1, 1, 1, 0, n, out.data_ptr(), n, None, st)
        torch.cuda.synchronize()
        for r in strips:
            g.lattice_device(meth, auvi.AXIS_NODES, 1, 1, 1, r, r + 64, ref.data_ptr(), n, sel.data_ptr(), st)
            torch.cuda.synchronize()
            got = out[r:r + 64]
            masked = torch.isnan(z[r:r + 64])
            assert torch.equal(got[~masked], truth_strips[r][~masked])          # pass-through
            assert torch.equal(torch.isnan(got), torch.isnan(ref))
            inb = torch.from_numpy(lat_ax_ok[r:r + 64]).cuda()[:, None] & lon_ok[None, :]
            assert not bool(torch.isnan(got[inb]).any())
            ok = ~torch.isnan(got)
            if meth == auvi.NN:
                assert torch.equal(got[ok], ref[ok])
            else:
                err = (got[ok].double() - ref[ok].double()).abs() - 1e-5 * ref[ok].double().abs()
                assert float(err.max()) <= 1e-3
    g.close()


# ---- opt-in method: bilinear with the bicubic method's search fallback --------------------------------
def _bilinear_search_expected(orc, pts):
    """AUVI_BILINEAR_SEARCH composed from two reference methods: bilinear, and -- exactly where that is NaN for an in-bounds
    query, i.e. all four corners missing, which puts NaN into the bicubic 4x4 stencil too -- the bicubic method's
    floor-centred 4-nearest mean."""
    b = orc.batch(ob.BILINEAR, pts)
    c = orc.batch(ob.CUBIC, pts)
    return np.where(np.isnan(b), c, b)


@pytest.mark.parametrize("name,frac", [("mid_atlantic", 0.5), ("mid_atlantic", 0.9), ("mariana", 0.5)])
def test_bilinear_search_opt_in(auvi, torch, name, frac):
    case = ob.masked_case(name, frac)
    orc = ob.Oracle(case["z"], *case["bounds"])
    g = auvi.Grid(case["z"], *case["bounds"])
    want = _bilinear_search_expected(orc, case["pts"])
    assert np.isnan(orc.batch(ob.BILINEAR, case["pts"])).sum() > 0 and not np.isnan(want).any()
    # point list (exact path), full-grid fill (tiled kernel), and a 2x lattice with holes (exact path)
    assert bits_equal(g.interp_points(auvi.BILINEAR_SEARCH, case["pts"]), want)
    filled = g.lattice(auvi.BILINEAR_SEARCH, auvi.AXIS_NODES, 1, 1, fill=1)
    assert bits_equal(filled[case["rows"], case["cols"]], want)
    keep = ~np.isnan(case["z"])
    assert bits_equal(filled[keep], case["z"][keep])
    m = case["meta"]
    q, nn_lat, nn_lon = ob.lattice_queries(m["n_lat"], m["n_lon"], *case["bounds"])
    sub = slice(0, 40 * nn_lon)                                       # the first 40 lattice rows
    got = g.lattice(auvi.BILINEAR_SEARCH, auvi.AXIS_EXPANDED, 2, 2, row_begin=0, row_end=40)
    assert bits_equal(got.ravel(), _bilinear_search_expected(orc, q[sub]))
    g.close()


@pytest.mark.parametrize("shape", [(2, 2), (3, 5), (7, 130), (17, 70), (65, 33), (100, 3), (33, 257)])
def test_tiny_and_ragged_grids_every_path(auvi, shape):
    """Grids smaller than a tile (down to 2 x 2, the smallest a GridD can hold), narrower than the TMA box, with row
    pitches that need padding: gap fill, 2x / 3x lattices and point lists of every method against the oracle."""
    n_lat, n_lon = shape
    rng = np.random.RandomState(n_lat * 1000 + n_lon)
    z = np.round(rng.uniform(-6000.0, -100.0, size=shape))
    for frac in (0.0, 0.4, 0.85):
        zm = z.copy()
        zm[rng.rand(*shape) < frac] = np.nan
        if frac > 0 and not np.isnan(zm).any():
            zm[0, 0] = np.nan
        bounds = (10.0, 10.0 + 0.01 * (n_lon - 1), -5.0, -5.0 + 0.013 * (n_lat - 1))
        orc = ob.Oracle(zm, *bounds)
        g = auvi.Grid(zm, *bounds)
        meta = dict(n_lat=n_lat, n_lon=n_lon, min_lon=bounds[0], max_lon=bounds[1], min_lat=bounds[2], max_lat=bounds[3])
        rows, cols = np.nonzero(np.isnan(zm))
        node_pts = ob.node_queries(rows, cols, meta) if rows.size else np.zeros((0, 3))
        for meth in METHODS:
            tight = meth in (ob.BILINEAR, ob.CUBIC, ob.NN)
            if rows.size:
                filled = g.lattice(meth, auvi.AXIS_NODES, 1, 1, fill=1)
                want = orc.batch(meth, node_pts)
                got = filled[rows, cols]
                if tight: assert bits_equal(got, want), (shape, frac, ob.METHOD_NAMES[meth])
                else: _close(got, want, atol=TIGHT if meth == ob.KRIGING else 1e-3, rtol=0 if meth == ob.KRIGING else 1e-5)
                assert bits_equal(filled[~np.isnan(zm)], zm[~np.isnan(zm)])
            for f in (2, 3):
                q, nn_lat, nn_lon = ob.lattice_queries(n_lat, n_lon, *bounds, f_lat=f, f_lon=f)
                got = g.lattice(meth, auvi.AXIS_EXPANDED, f, f).ravel()
                want = orc.batch(meth, q)
                if tight: assert bits_equal(got, want), (shape, frac, f, ob.METHOD_NAMES[meth])
                else: _close(got, want, atol=TIGHT if meth == ob.KRIGING else 1e-3, rtol=0 if meth == ob.KRIGING else 1e-5)
        g.close()


@pytest.mark.parametrize("dtype_name", ["f32", "f64"])
def test_adopted_grid_with_unaligned_pitch_takes_the_plain_load_path(auvi, torch, dtype_name):
    """An adopted grid whose row pitch is not a multiple of 16 bytes cannot be described by a TMA tensor map: both tiled
    kernels stage their blocks with plain coalesced loads instead.  Same results as the exact per-query path."""
    n_lat, n_lon = 333, 1001                                        # 4004 / 8008 bytes per row
    tdt = torch.float32 if dtype_name == "f32" else torch.float64
    dt = auvi.F32 if dtype_name == "f32" else auvi.F64
    jj = torch.arange(n_lat, device="cuda", dtype=torch.float64)[:, None]
    ii = torch.arange(n_lon, device="cuda", dtype=torch.float64)[None, :]
    z = (-3000.0 + 700.0 * torch.sin(ii * 0.02) * torch.cos(jj * 0.03) + 0.5 * ii).to(tdt).contiguous()
    g = auvi.Grid(adopt=dict(ptr=z.data_ptr(), dtype=dt, n_lat=n_lat, n_lon=n_lon, ld=n_lon, row0=0, rows=n_lat, keep=z),
                  min_lon=10.0, max_lon=12.0, min_lat=50.0, max_lat=51.0)
    g.mask_hash(0.6, seed=3, count=False)
    st = torch.cuda.current_stream().cuda_stream
    a = torch.empty((n_lat, n_lon), dtype=tdt, device="cuda"); b = torch.empty_like(a)
    sel = torch.empty((n_lat * n_lon, 9), dtype=torch.int32, device="cuda")
    for meth in (auvi.BILINEAR, auvi.NN, auvi.CUBIC):
        g.lattice_device(meth, auvi.AXIS_NODES, 1, 1, 1, 0, n_lat, a.data_ptr(), n_lon, None, st)
        assert not g.uses_tma
        g.lattice_device(meth, auvi.AXIS_NODES, 1, 1, 1, 0, n_lat, b.data_ptr(), n_lon, sel.data_ptr(), st)
        torch.cuda.synchronize()
        assert torch.equal(torch.isnan(a), torch.isnan(b))
        ok = ~torch.isnan(a)
        assert torch.equal(a[ok], b[ok]), auvi.METHOD_NAMES[meth]
    # upsampling lattice, padded output pitch
    rows, cols = 2 * (n_lat - 1) + 1, 2 * (n_lon - 1) + 1
    ld = cols + 7
    u = torch.full((rows, ld), 7.0, dtype=tdt, device="cuda"); v = torch.full_like(u, 7.0)
    sel2 = torch.empty((rows * cols, 9), dtype=torch.int32, device="cuda")
    for meth in (auvi.BILINEAR, auvi.CUBIC, auvi.NN):
        g.lattice_device(meth, auvi.AXIS_EXPANDED, 2, 2, 0, 0, rows, u.data_ptr(), ld, None, st)
        g.lattice_device(meth, auvi.AXIS_EXPANDED, 2, 2, 0, 0, rows, v.data_ptr(), ld, sel2.data_ptr(), st)
        torch.cuda.synchronize()
        same = (u == v) | (torch.isnan(u) & torch.isnan(v))
        if dtype_name == "f64" or meth == auvi.NN:
            assert bool(same.all()), auvi.METHOD_NAMES[meth]
        else:                                                       # FP32 tap weights in the tiled stencils
            assert torch.equal(torch.isnan(u), torch.isnan(v))
            ok = ~torch.isnan(u)
            assert float((u[ok].double() - v[ok].double()).abs().max()) <= 1e-3 + 1e-5 * 4000
        assert bool((u[:, cols:] == 7.0).all())                     # the padding of the output rows is untouched
    g.close()


def test_fill_patched_tile_variant_equals_the_default():
    """fill_tiled_kernel has a second way out for its results: patched into the staged tile and stored as whole 16-byte rows
    (picked automatically when the output is peer memory: the gather fused into the kernel).  AUVI_FILL_PATCH=1 forces it
    (read once per process, hence the child): every method, both dtypes, aligned and unaligned output pitches, must produce
    the bits of the default variant's output -- recorded here by a first child without the variable."""
    import hashlib
    import subprocess
    code = r'''
import sys, hashlib, numpy as np, torch
sys.path.insert(0, "auv-real-time-interpolation_b200/python"); sys.path.insert(0, ".")
import auvi
for dt, tdt in ((auvi.F32, torch.float32), (auvi.F64, torch.float64)):
    n_lat, n_lon = 517, 1301
    jj = torch.arange(n_lat, device="cuda", dtype=torch.float64)[:, None]; ii = torch.arange(n_lon, device="cuda", dtype=torch.float64)[None, :]
    z = (-4000.0 + 900.0 * torch.sin(ii * 0.013) * torch.cos(jj * 0.017) + 0.37 * ii - 0.21 * jj).to(tdt).contiguous()
    g = auvi.Grid(adopt=dict(ptr=z.data_ptr(), dtype=dt, n_lat=n_lat, n_lon=n_lon, ld=n_lon, row0=0, rows=n_lat, keep=z), min_lon=-30.9967, max_lon=-29.4993, min_lat=-0.5035, max_lat=1.0071)
    for frac in (0.3, 0.7, 0.95):
        z2 = z.clone(); g2 = auvi.Grid(adopt=dict(ptr=z2.data_ptr(), dtype=dt, n_lat=n_lat, n_lon=n_lon, ld=n_lon, row0=0, rows=n_lat, keep=z2), min_lon=-30.9967, max_lon=-29.4993, min_lat=-0.5035, max_lat=1.0071)
        g2.mask_hash(frac, seed=5, count=False)
        for ld in (n_lon, n_lon + 3):
            for meth in (auvi.NN, auvi.CUBIC, auvi.IDW, auvi.KRIGING, auvi.BILINEAR):
                out = torch.full((n_lat, ld), 7.0, dtype=tdt, device="cuda")
                g2.lattice_device(meth, auvi.AXIS_NODES, 1, 1, 1, 0, n_lat, out.data_ptr(), ld, None, torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                assert bool((out[:, n_lon:] == 7.0).all())
                print(dt, frac, ld, meth, hashlib.sha1(out.cpu().numpy().tobytes()).hexdigest())
        g2.close()
    g.close()
'''
    outs = []
    for patch in (None, "1"):
        env = dict(os.environ)
        env.pop("AUVI_FILL_PATCH", None)
        if patch:
            env["AUVI_FILL_PATCH"] = patch
        r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
        outs.append(r.stdout)
    assert len(outs[0].splitlines()) == 2 * 3 * 2 * 5 and outs[0] == outs[1]
