"""GPU suite: the opt-in methods of SURVEY.md section 8(f) N4 that have no reference code -- AUVI_IDW_KNN (IDW over the
TRUE four nearest valid cells: the reference's enumeration without its two early breaks) and AUVI_KRIGING_FITTED
(ordinary kriging on the reference's picks with the variogram fitted to the grid) -- against their CPU oracles in
oracle/interp_oracle.c, through the Point-list API and the full-grid fill, and their depth RMSE beside the reference
methods on Mariana at 10 / 50 / 90 % removal."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "auv-real-time-interpolation_b200", "python"))
sys.path.insert(0, ROOT)

from oracle import binding as ob  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def auvi():
    import auvi as m
    m.load()
    if m.device_count() == 0:
        pytest.skip("no CUDA device")
    return m


@pytest.fixture(scope="module")
def torch():
    import torch as t
    return t


def _close(got, want, atol, rtol):
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    err = np.abs(got[ok] - want[ok]) - rtol * np.abs(want[ok])
    assert err.max() <= atol, float(err.max())


@pytest.mark.parametrize("name,frac", [("mid_atlantic", 0.5), ("mid_atlantic", 0.9), ("mariana", 0.5)])
def test_idw_knn_matches_its_oracle(auvi, torch, name, frac):
    case = ob.masked_case(name, frac)
    orc = ob.Oracle(case["z"], *case["bounds"])
    g = auvi.Grid(case["z"], *case["bounds"])
    pts = case["pts"]
    want, sel_w, found_w = orc.batch_optin(auvi.IDW_KNN, pts, want_sel=True)
    got = g.interp_points(auvi.IDW_KNN, pts)
    _close(got, want, 1e-3, 1e-5)                                   # FP32 weights, as IDW
    # the picks, bit for bit: device selection dump against the oracle's stable four nearest
    n = pts.shape[0]
    d_pts = torch.from_numpy(pts).cuda()
    d_out = torch.empty(n, dtype=torch.float64, device="cuda")
    d_sel = torch.empty((n, 8), dtype=torch.int32, device="cuda")
    d_found = torch.empty(n, dtype=torch.int32, device="cuda")
    g.interp_points_device(auvi.IDW_KNN, d_pts.data_ptr(), n, 24, d_out.data_ptr(), d_sel.data_ptr(), d_found.data_ptr(),
                           torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    sel = d_sel.cpu().numpy().reshape(n, 4, 2)
    assert np.array_equal(sel, sel_w)
    # full-grid fill: the same values at the removed cells, valid cells pass through
    filled = g.lattice(auvi.IDW_KNN, auvi.AXIS_NODES, 1, 1, fill=1)
    _close(filled[case["rows"], case["cols"]], want, 1e-3, 1e-5)
    keep = ~np.isnan(case["z"])
    assert np.array_equal(filled[keep], case["z"][keep])
    g.close()


@pytest.mark.parametrize("name,frac", [("mid_atlantic", 0.5), ("mariana", 0.5), ("mid_atlantic", 0.9)])
def test_kriging_fitted_matches_its_oracle(auvi, name, frac):
    case = ob.masked_case(name, frac)
    orc = ob.Oracle(case["z"], *case["bounds"])
    g = auvi.Grid(case["z"], *case["bounds"])
    c0, c1, rng = g.fit_variogram()                                  # device reduction + host fit
    w0, w1, wr = orc.variogram_fit()                                 # plain-C sums + the restated fit
    assert rng == wr                                                 # the same rung of the range ladder
    np.testing.assert_allclose([c0, c1], [w0, w1], rtol=1e-9, atol=1e-9)
    want = orc.batch_optin(auvi.KRIGING_FITTED, case["pts"], params=(c0, c1, rng))
    got = g.interp_points(auvi.KRIGING_FITTED, case["pts"])
    _close(got, want, 1e-6, 0.0)
    filled = g.lattice(auvi.KRIGING_FITTED, auvi.AXIS_NODES, 1, 1, fill=1)      # the tiled kernel with the fitted model
    _close(filled[case["rows"], case["cols"]], want, 1e-6, 0.0)
    # the reference method is untouched by the fit
    ref = orc.batch(ob.KRIGING, case["pts"])
    _close(g.interp_points(auvi.KRIGING, case["pts"]), ref, 1e-6, 0.0)
    # parameters handed in instead of fitted (what a row slab does)
    g2 = auvi.Grid(case["z"], *case["bounds"])
    g2.set_variogram(2.0, 500.0, 0.05)
    _close(g2.interp_points(auvi.KRIGING_FITTED, case["pts"][:5000]),
           orc.batch_optin(auvi.KRIGING_FITTED, case["pts"][:5000], params=(2.0, 500.0, 0.05)), 1e-6, 0.0)
    with pytest.raises(auvi.AuviError, match="c1 > 0"):
        g2.set_variogram(0.0, -1.0, 1.0)
    g2.close()
    g.close()


def test_optin_rmse_beside_the_reference_methods(auvi, capsys):
    """Depth RMSE against the unmasked GEBCO truth on Mariana at 10 / 50 / 90 % removal, the opt-ins beside the methods whose
    semantics they relax.  Reported, not asserted as an improvement: measured on B200 (profiles/r02_optin_rmse.txt) the
    true four nearest are WORSE than the reference's early-terminated four at 10 % removal (53.4 m vs 47.6 m: with the
    reference's +0.5 cell-centre convention the four nearest cells all lie on one side of a node query), and the fitted
    variogram changes a four-point kriging estimate by centimetres."""
    rows = []
    for frac in (0.1, 0.5, 0.9):
        case = ob.masked_case("mariana", frac)
        g = auvi.Grid(case["z"], *case["bounds"])
        r = {}
        for tag, meth in (("idw", auvi.IDW), ("idw_knn", auvi.IDW_KNN), ("kriging", auvi.KRIGING), ("kriging_fitted", auvi.KRIGING_FITTED)):
            est = g.interp_points(meth, case["pts"])
            assert not np.isnan(est).any()
            r[tag] = float(np.sqrt(np.mean((est - case["truth"]) ** 2)))
        rows.append((frac, r))
        g.close()
    with capsys.disabled():
        for frac, r in rows:
            print(f"\nmariana @{frac:.0%}: RMSE m  " + "  ".join(f"{k} {v:.3f}" for k, v in r.items()), end="")
        print()
    for frac, r in rows:
        assert all(np.isfinite(v) and 0 < v < 200 for v in r.values()), (frac, r)
