"""Cartesian block decomposition of the global lattice.

Mirrors cavity_opt2.py:225-258: ``Create_cart((ndx, ndy))`` numbers ranks
row-major (rank = px*ndy + py); every block gets ``n // nd`` cells and the last
block along an axis takes the remainder.  Unlike the reference
(``periods=(False, False)``) the neighbour rings are PERIODIC in both
directions: the reference's single-rank run wraps ``np.roll`` around the whole
lattice (the cavity's lid term observes that wrap at the two top corners), and
walls are predicates on global coordinates, so periodic rings are what makes
the N-block result bit-identical to the 1-block result.
"""
from collections import namedtuple

from ._lib import DIRS

Block = namedtuple("Block", "rank px py x0 y0 lnx lny")


def axis_extents(n, nd):
    """[(offset, length)] of the nd blocks along an axis of n cells (cavity_opt2.py:231-250)."""
    if nd < 1 or n < nd:
        raise ValueError("cannot split %d cells into %d blocks" % (n, nd))
    base = n // nd
    out = [(i * base, base) for i in range(nd - 1)]
    out.append(((nd - 1) * base, n - base * (nd - 1)))
    return out


class Decomposition:
    def __init__(self, nx, ny, ndx=1, ndy=1):
        self.nx, self.ny, self.ndx, self.ndy = int(nx), int(ny), int(ndx), int(ndy)
        self.xs = axis_extents(self.nx, self.ndx)
        self.ys = axis_extents(self.ny, self.ndy)

    @property
    def size(self):
        return self.ndx * self.ndy

    def coords(self, rank):
        return divmod(rank, self.ndy)          # Create_cart row-major: rank = px*ndy + py

    def rank_of(self, px, py):
        return (px % self.ndx) * self.ndy + (py % self.ndy)

    def block(self, rank):
        px, py = self.coords(rank)
        (x0, lnx), (y0, lny) = self.xs[px], self.ys[py]
        return Block(rank, px, py, x0, y0, lnx, lny)

    def blocks(self):
        return [self.block(r) for r in range(self.size)]

    def neighbour(self, rank, d):
        """Rank in direction slot d (periodic rings)."""
        px, py = self.coords(rank)
        dx, dy = DIRS[d]
        return self.rank_of(px + dx, py + dy)

    def neighbours(self, rank):
        return [self.neighbour(rank, d) for d in range(len(DIRS))]

    def slices(self, rank):
        b = self.block(rank)
        return slice(b.x0, b.x0 + b.lnx), slice(b.y0, b.y0 + b.lny)

    def scatter(self, g, rank):
        """Local part (copy) of a global (..., nx, ny) array."""
        sx, sy = self.slices(rank)
        return g[..., sx, sy].copy()

    def gather_into(self, g, rank, local):
        sx, sy = self.slices(rank)
        g[..., sx, sy] = local


# ---- temporal blocking is a collective decision (csrc/temporal.cuh) -------------------------------------
T2_MIN_EXTENT = 16            # a block needs at least 16 x 16 cells
T2_TILE_COLS = 254            # output columns of a fused tile
T2_TILE_ROWS = 16             # rows of the smallest fused tile the library picks
T2_AUTO_MIN_TILES = 512       # automatic mode: fused tiles a block must offer to fill the GPU (about 1400^2 cells)


def temporal_mode(blocks, boundary, requested=None):
    """The stepping mode every block of a decomposition must use: 2 (two time steps per pass over HBM) or 1
    (single-step kernel).  A block waits for its neighbours' level-(n+1) frame ghosts, so the mode cannot be
    mixed.  requested: None / 0 = automatic (temporal blocking only if the SMALLEST block still fills the
    GPU), 1 = single-step, 2 = temporal blocking whenever every block is eligible."""
    eligible = boundary in ("periodic", "cavity", "cavity_xperiodic") and all(
        b.lnx >= T2_MIN_EXTENT and b.lny >= T2_MIN_EXTENT for b in blocks)
    if requested == 1 or not eligible:
        return 1
    if requested == 2:
        return 2
    tiles = min(-(-(b.lny - 4) // T2_TILE_COLS) * -(-(b.lnx - 4) // T2_TILE_ROWS) for b in blocks)
    return 2 if tiles >= T2_AUTO_MIN_TILES else 1
