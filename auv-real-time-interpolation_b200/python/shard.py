"""Row sharding of a lattice job across ranks.  Host logic only: no CUDA, no oracle.

Every output cell is independent, so rank r owns a contiguous block of OUTPUT rows and needs the grid
rows its stencils and ring searches can touch:  [floor(lo/f) - HALO, floor((hi-1)/f) + 2 + HALO).
HALO = 14 covers the radius-10 ring search (GridH.cpp:275,339) around a floor/round centre that FP64
noise may put one row below the nominal one, plus the bicubic stencil's
-1/+2 and one row of margin.  Halos are replicated when the slab is loaded, never exchanged.
"""
from dataclasses import dataclass

HALO = 14


@dataclass(frozen=True)
class RowShard:
    rank: int
    world: int
    out_rows_global: int
    row_lo: int      # first output row of this rank
    row_hi: int      # one past the last output row
    in_lo: int       # first grid row the rank must hold
    in_hi: int       # one past the last grid row the rank must hold

    @property
    def out_rows(self):
        return self.row_hi - self.row_lo

    @property
    def in_rows(self):
        return self.in_hi - self.in_lo


def lattice_rows(n_lat: int, factor: int) -> int:
    """Output rows of an axis of n_lat nodes at an integer factor: f*(n-1)+1 (2n-1 for f=2,
    test_interpolation.cpp:94-95)."""
    return factor * (n_lat - 1) + 1


def plan_rows(n_lat_global: int, factor: int, world: int, rank: int, halo: int = HALO) -> RowShard:
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    out_rows = lattice_rows(n_lat_global, factor)
    per = -(-out_rows // world)
    lo, hi = min(out_rows, rank * per), min(out_rows, (rank + 1) * per)
    if hi <= lo:
        return RowShard(rank, world, out_rows, lo, lo, 0, 0)
    in_lo = max(0, lo // factor - halo)
    in_hi = min(n_lat_global, (hi - 1) // factor + 2 + halo)
    return RowShard(rank, world, out_rows, lo, hi, in_lo, in_hi)


def e2e_row_budget(shard_rows: int, row_bytes: int, avail_bytes: int, local_world: int, frac: float = 0.4) -> int:
    """Rows of host-pinned output a rank may allocate: at most `frac` of the available host memory split
    over the ranks of the box."""
    cap = int(avail_bytes * frac / max(1, local_world)) // max(1, row_bytes)
    return max(1, min(shard_rows, cap))
