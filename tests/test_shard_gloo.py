"""CPU suite, part 3: the N>1 host logic over torch.distributed (gloo, world_size 2).

Each rank plans its output-row block (python/shard.py, the planner bench.py uses), computes ONLY those
rows -- here with the CPU oracle standing in for the device, restricted to the grid rows the plan says
the rank holds -- and the blocks are all-gathered.  The assembled lattice must equal the single-process
result bit for bit, for a stencil method and for a ring-search method, on a grid with holes."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "auv-real-time-interpolation_b200", "python"))
sys.path.insert(0, ROOT)

import shard  # noqa: E402

N_LAT, N_LON, F = 61, 47, 4
BOUNDS = (-180.0, -160.0, 20.0, 30.0)


def _grid():
    from oracle import binding as ob
    z = ob.synth_grid(N_LAT, N_LON)
    z.ravel()[np.random.RandomState(12).choice(z.size, 500, replace=False)] = np.nan
    return z


def _worker(rank, world, port, q):
    from oracle import binding as ob
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z = _grid()
    plan = shard.plan_rows(N_LAT, F, world, rank)
    # the rank only "holds" its slab: everything outside is poisoned, so a too-small halo changes results
    held = np.full_like(z, -9.0e9)
    held[plan.in_lo:plan.in_hi] = z[plan.in_lo:plan.in_hi]
    pts, nn_lat, nn_lon = ob.lattice_queries(N_LAT, N_LON, *BOUNDS, f_lat=F, f_lon=F)
    mine = pts.reshape(nn_lat, nn_lon, 3)[plan.row_lo:plan.row_hi].reshape(-1, 3)
    orc = ob.Oracle(held, *BOUNDS)
    per = -(-nn_lat // world)
    res = {}
    for meth in (ob.CUBIC, ob.KRIGING):
        block = torch.full((per, nn_lon), float("nan"), dtype=torch.float64)
        block[:plan.out_rows] = torch.from_numpy(orc.batch(meth, mine).reshape(plan.out_rows, nn_lon))
        parts = [torch.empty_like(block) for _ in range(world)]
        dist.all_gather(parts, block)
        res[meth] = torch.cat(parts)[:nn_lat].numpy()
    t = torch.tensor([plan.row_lo, plan.row_hi, plan.in_lo, plan.in_hi])
    plans = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(plans, t)
    if rank == 0:
        q.put((res, [p.tolist() for p in plans]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_row_sharding_matches_single_process():
    from oracle import binding as ob
    from conftest import bits_equal
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res, plans = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    z = _grid()
    pts, nn_lat, nn_lon = ob.lattice_queries(N_LAT, N_LON, *BOUNDS, f_lat=F, f_lon=F)
    orc = ob.Oracle(z, *BOUNDS)
    for meth in (ob.CUBIC, ob.KRIGING):
        want = orc.batch(meth, pts).reshape(nn_lat, nn_lon)
        assert bits_equal(res[meth], want), ob.METHOD_NAMES[meth]
    # blocks tile the lattice exactly
    assert plans[0][0] == 0 and plans[-1][1] == nn_lat
    assert all(plans[k][1] == plans[k + 1][0] for k in range(world - 1))


@pytest.mark.parametrize("n_lat,f,world", [(16384, 4, 1), (16384 * 8, 4, 8), (65536, 1, 8), (963, 1, 4), (10, 2, 8), (5, 1, 8)])
def test_plan_properties(n_lat, f, world):
    out_rows = shard.lattice_rows(n_lat, f)
    seen = 0
    for r in range(world):
        p = shard.plan_rows(n_lat, f, world, r)
        assert p.row_lo == seen and p.row_hi >= p.row_lo
        seen = p.row_hi
        if p.out_rows:
            # every grid row within reach of the block is held: (base-1)-1-11 .. base+2+11, where FP64 noise may
            # put base one below row//f (clamped to the grid)
            assert p.in_lo <= max(0, p.row_lo // f - 1 - 1 - 11)
            assert p.in_hi >= min(n_lat, (p.row_hi - 1) // f + 1 + 2 + 11)
            assert 0 <= p.in_lo < p.in_hi <= n_lat
    assert seen == out_rows


def test_e2e_row_budget():
    assert shard.e2e_row_budget(65533, 262132, 196 << 30, 1) == 65533
    assert shard.e2e_row_budget(65533, 262132, 64 << 30, 8) == (int((64 << 30) * 0.4 / 8)) // 262132
    assert shard.e2e_row_budget(10, 1 << 40, 1 << 20, 8) == 1
