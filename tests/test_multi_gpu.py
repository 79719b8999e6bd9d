"""GPU suite: several devices behind the C ABI in one process (auvi_multi_*, csrc/multi.cu).

The shards of a multi-device job must reproduce the single-device result bit for bit (same kernels, same inputs:
SURVEY.md section 8(e) "Check").  Every test runs with two shards on device 0 (what a one-GPU box can do: the
sharding, halo and assembly logic is the same) and, when the box has them, on two different devices -- there also
with the gather fused into the kernels (peer stores into device 0's buffer)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "auv-real-time-interpolation_b200", "python"))
sys.path.insert(0, ROOT)

from oracle import binding as ob  # noqa: E402  (fixture generator only)
from conftest import bits_equal  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def auvi():
    import auvi as m
    m.load()
    if m.device_count() == 0:
        pytest.skip("no CUDA device")
    return m


def _device_sets(auvi):
    sets = [[0, 0], [0, 0, 0]]
    if auvi.device_count() >= 2:
        sets.append([0, 1])
    if auvi.device_count() >= 4:
        sets.append([0, 1, 2, 3])
    return sets


def _grid(dtype):
    z = ob.synth_grid(301, 260).astype(dtype)
    z.ravel()[np.random.RandomState(8).choice(z.size, 9000, replace=False)] = np.nan
    return z, (0.0, 2.0, 40.0, 43.0)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_multi_lattice_equals_single_device(auvi, dtype):
    z, bounds = _grid(dtype)
    whole = auvi.Grid(z, *bounds)
    jobs = ((auvi.CUBIC, auvi.AXIS_EXPANDED, 4, 0), (auvi.BILINEAR, auvi.AXIS_EXPANDED, 3, 0), (auvi.KRIGING, auvi.AXIS_EXPANDED, 2, 0),
            (auvi.IDW, auvi.AXIS_NODES, 1, 1), (auvi.NN, auvi.AXIS_NODES, 1, 1), (auvi.KRIGING, auvi.AXIS_NODES, 1, 1),
            (auvi.BILINEAR, auvi.AXIS_NODES, 1, 1), (auvi.CUBIC, auvi.AXIS_NODES, 1, 1))
    want = {j: whole.lattice(j[0], j[1], j[2], j[2], fill=j[3]) for j in jobs}
    for devs in _device_sets(auvi):
        for replicate in (False, True):
            m = auvi.MultiGrid(z, *bounds, n_gpus=len(devs), devices=devs, replicate=replicate)
            rows_seen = 0
            for k in range(len(devs)):
                dev, lo, hi = m.shard(k, 4)
                assert dev == devs[k] and lo == rows_seen
                rows_seen = hi
            assert rows_seen == 4 * 300 + 1
            for j in jobs:
                got = m.lattice(j[0], j[1], j[2], j[2], fill=j[3])
                assert bits_equal(got.astype(np.float64), want[j].astype(np.float64)), (devs, replicate, j)
            assert m.last_kernel_ms > 0
            m.close()
    whole.close()


def test_multi_device_resident_and_gathered(auvi):
    """Device form: results left sharded (one buffer per shard) and gathered into ONE buffer on shard 0's device by the
    kernels' own stores (peer access when the shards sit on different devices)."""
    import torch
    z, bounds = _grid(np.float32)
    whole = auvi.Grid(z, *bounds)
    f = 4
    want_up = torch.from_numpy(whole.lattice(auvi.CUBIC, auvi.AXIS_EXPANDED, f, f))
    want_fill = torch.from_numpy(whole.lattice(auvi.IDW, auvi.AXIS_NODES, 1, 1, fill=1))
    for devs in _device_sets(auvi):
        m = auvi.MultiGrid(z, *bounds, n_gpus=len(devs), devices=devs)
        for want, meth, kind, ff, fill in ((want_up, auvi.CUBIC, auvi.AXIS_EXPANDED, f, 0), (want_fill, auvi.IDW, auvi.AXIS_NODES, 1, 1)):
            rows, cols = want.shape
            ld = (cols + 3) // 4 * 4
            # sharded
            bufs, ptrs = [], []
            for k in range(len(devs)):
                dev, lo, hi = m.shard(k, ff)
                b = torch.full((max(hi - lo, 1), ld), 7.0, dtype=torch.float32, device=f"cuda:{dev}")
                bufs.append((b, lo, hi)); ptrs.append(b.data_ptr())
            m.lattice_device(meth, kind, ff, ff, fill, ptrs, ld)
            m.sync()
            assert m.last_kernel_ms > 0
            for b, lo, hi in bufs:
                assert torch.equal(b[:hi - lo, :cols].cpu(), want[lo:hi])
            # gathered on shard 0's device
            m.enable_peer(0)
            root = torch.full((rows, ld), 7.0, dtype=torch.float32, device=f"cuda:{devs[0]}")
            ptrs = [root.data_ptr() + m.shard(k, ff)[1] * ld * 4 for k in range(len(devs))]
            m.lattice_device(meth, kind, ff, ff, fill, ptrs, ld)
            m.sync()
            assert torch.equal(root[:, :cols].cpu(), want)
        m.close()
    whole.close()


def test_multi_points_and_mask(auvi):
    case = ob.masked_case("mid_atlantic", 0.5)
    whole = auvi.Grid(case["z"], *case["bounds"])
    rng = np.random.RandomState(3)
    pts = np.tile(case["pts"], (4, 1))[:200_000]
    pts[:, 0] += rng.uniform(-1e-3, 1e-3, pts.shape[0])
    os.environ["AUVI_MULTI_MIN_POINTS"] = "30000"                 # the library spreads only > 4 Mi points per device: lower it
    try:
        _points_on_device_sets(auvi, case, whole, pts)
    finally:
        del os.environ["AUVI_MULTI_MIN_POINTS"]
    whole.close()
    _mask_on_device_sets(auvi)


def _points_on_device_sets(auvi, case, whole, pts):
    for devs in _device_sets(auvi):
        m = auvi.MultiGrid(case["z"], *case["bounds"], n_gpus=len(devs), devices=devs, replicate=True)
        for meth in (auvi.BILINEAR, auvi.CUBIC, auvi.KRIGING, auvi.NN, auvi.IDW):
            assert bits_equal(m.interp_points(meth, pts), whole.interp_points(meth, pts)), (devs, meth)
            assert bits_equal(m.interp_points(meth, pts[:100]), whole.interp_points(meth, pts[:100]))
        m.close()
        sl = auvi.MultiGrid(case["z"], *case["bounds"], n_gpus=len(devs), devices=devs)
        with pytest.raises(auvi.AuviError, match="replicated"):
            sl.interp_points(auvi.NN, pts[:10])
        sl.close()


def _mask_on_device_sets(auvi):
    # one global mask drawn shard by shard equals the mask drawn on the whole grid
    z = ob.synth_grid(200, 300).astype(np.float32)
    bounds = (0.0, 1.0, 0.0, 1.0)
    one = auvi.Grid(z, *bounds)
    one.mask_hash(0.6, seed=11, count=False)
    want = one.lattice(auvi.NN, auvi.AXIS_NODES, 1, 1, fill=1)
    for devs in _device_sets(auvi):
        m = auvi.MultiGrid(z, *bounds, n_gpus=len(devs), devices=devs)
        m.mask_hash(0.6, seed=11)
        assert bits_equal(m.lattice(auvi.NN, auvi.AXIS_NODES, 1, 1, fill=1).astype(np.float64), want.astype(np.float64))
        m.close()
    one.close()


def test_multi_rejects_bad_arguments(auvi):
    z, bounds = _grid(np.float32)
    with pytest.raises(auvi.AuviError, match="n_gpus"):
        auvi.MultiGrid(z, *bounds, n_gpus=0)
    with pytest.raises(auvi.AuviError, match="ordinal"):
        auvi.MultiGrid(z, *bounds, n_gpus=1, devices=[99])
    # more shards than the halo logic can make useful still works (idle shards)
    tiny = ob.synth_grid(5, 40)
    m = auvi.MultiGrid(tiny, *bounds, n_gpus=8, devices=[0] * 8)
    one = auvi.Grid(tiny, *bounds)
    assert bits_equal(m.lattice(auvi.CUBIC, auvi.AXIS_EXPANDED, 2, 2), one.lattice(auvi.CUBIC, auvi.AXIS_EXPANDED, 2, 2))
    m.close(); one.close()
