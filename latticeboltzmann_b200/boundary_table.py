"""General per-cell boundary tables.

The simple_flows family of the reference applies its walls as a SEQUENCE of numpy slice assignments on the
streamed array -- ``grid[3, -2, :] = grid[1, -1, :]``, ``grid[7, :, -2] = grid[5, :, -1] - 1/6*uw`` ... -- whose
ranges overlap at corners, so later lines read what earlier lines wrote (slidingLidMPI.py:180-204 with its full
ranges, the rectangular obstacle of experimantal_flows/obstacle_canal.py:413-458).  Every such line is "copy a
streamed population, optionally shifted by a constant".  ``SymbolicGrid`` replays a sequence of that kind on an
INDEX array instead of on values: after ``stream()`` entry (i, k, l) refers to pre-stream element
(i, k - cx_i, l - cy_i); each assignment moves references (and accumulates the constant) with exactly numpy's
slicing and ordering semantics.  What is left is, for every cell the walls touched, the pre-stream source of each
of its nine populations and an additive constant -- a table the kernel applies as a pure gather:

    post[i, cell] = pre[src[i, cell]] + add[i, cell]            then moments + collision as everywhere else.

The table lives on the device (``Lattice.set_boundary_table``); cells not in it stream periodically.  The step order
is the family's: stream -> walls -> moments -> collide (slidingLidMPI.py:264-268, obstacle_canal.py:270-275), the
arithmetic is the simple_flows equilibrium / collision (bit-identical to numpy, fp64).
"""
import numpy as np

CX = np.array([0, 1, 0, -1, 0, 1, -1, -1, 1])
CY = np.array([0, 0, 1, 0, -1, 1, 1, -1, -1])


class _Ref:
    """A view of the symbolic array: source indices + constants, as returned by SymbolicGrid[...]."""

    def __init__(self, src, add):
        self.src, self.add = src, add

    def _shift(self, c):
        # (x + c1) + c2 is not x + (c1 + c2) in floating point and the table holds ONE constant per entry
        if np.any(self.add != 0):
            raise ValueError("a population that already carries a constant is shifted again")
        return _Ref(self.src, self.add + c)

    def __add__(self, c):
        return self._shift(float(c))

    def __sub__(self, c):
        return self._shift(-float(c))


class SymbolicGrid:
    """``grid`` of shape (9, nx, ny) whose entries are references into the pre-stream state."""

    def __init__(self, nx, ny):
        self.nx, self.ny = int(nx), int(ny)
        self.src = np.arange(9 * nx * ny, dtype=np.int64).reshape(9, nx, ny)
        self.add = np.zeros((9, nx, ny))
        self._plain = None

    def stream(self):
        """np.roll of every channel by its velocity (PyLB/Streaming.py:45-46, slidingLidMPI.py:123-125)."""
        for i in range(1, 9):
            self.src[i] = np.roll(self.src[i], (CX[i], CY[i]), axis=(0, 1))
        self._plain = self.src.copy()
        return self

    def __getitem__(self, key):
        return _Ref(self.src[key].copy(), self.add[key].copy())

    def __setitem__(self, key, ref):
        if not isinstance(ref, _Ref):
            raise TypeError("only copies of (optionally shifted) populations can be assigned")
        self.src[key] = ref.src
        self.add[key] = ref.add

    def table(self):
        """-> (cells, src, add): flat cell indices k*ny + l of the cells whose gather differs from the plain
        periodic pull, and for each of them the nine flat pre-stream element indices i*nx*ny + k*ny + l and the
        nine additive constants."""
        if self._plain is None:
            raise RuntimeError("call stream() before applying the wall assignments")
        touched = np.any((self.src != self._plain) | (self.add != 0), axis=0)
        cells = np.flatnonzero(touched.reshape(-1)).astype(np.int64)
        src = self.src.reshape(9, -1)[:, cells].T.copy()
        add = self.add.reshape(9, -1)[:, cells].T.copy()
        return cells, np.ascontiguousarray(src), np.ascontiguousarray(add)


def sliding_lid_mpi_table(nx, ny, uw, right=True, left=True, bottom=True, top=True):
    """slidingLidMPI.py:180-204 ``bounce_back_choosen`` (single rank: all four walls apply) on an (nx, ny) array
    that includes the wall layers: FULL index ranges, fixed order right, left, bottom, top."""
    g = SymbolicGrid(nx, ny).stream()
    if right:                                   # :183-187
        g[3, -2, :] = g[1, -1, :]
        g[6, -2, :] = g[8, -1, :]
        g[7, -2, :] = g[5, -1, :]
    if left:                                    # :188-192
        g[1, 1, :] = g[3, 0, :]
        g[5, 1, :] = g[7, 0, :]
        g[8, 1, :] = g[6, 0, :]
    if bottom:                                  # :195-199
        g[2, :, 1] = g[4, :, 0]
        g[5, :, 1] = g[7, :, 0]
        g[6, :, 1] = g[8, :, 0]
    if top:                                     # :200-204
        g[4, :, -2] = g[2, :, -1]
        g[7, :, -2] = g[5, :, -1] - 1 / 6 * uw
        g[8, :, -2] = g[6, :, -1] + 1 / 6 * uw
    return g.table()


def obstacle_channel_table(nx, ny, x0, x1, y0, y1, uw=0.0):
    """A channel with bounce-back walls at the bottom and the top (obstacle_canal.py:320-329 as in
    slidingLidMPI.py:195-204, x periodic) and a rectangular bounce-back obstacle with corners (x0, y0), (x1, y1),
    applied as obstacle_canal.py:413-458 ``apply_obstacle`` writes it (left, right, top, bottom faces in that
    order, its index ranges and channel pairs taken literally)."""
    g = SymbolicGrid(nx, ny).stream()
    g[2, :, 1] = g[4, :, 0]
    g[5, :, 1] = g[7, :, 0]
    g[6, :, 1] = g[8, :, 0]
    g[4, :, -2] = g[2, :, -1]
    g[7, :, -2] = g[5, :, -1] - 1 / 6 * uw
    g[8, :, -2] = g[6, :, -1] + 1 / 6 * uw
    ys, xs = slice(y0, y1), slice(x0, x1)
    g[3, x0 - 1, ys] = g[1, x0, ys]             # :419-427  left face
    g[7, x0 - 1, ys] = g[5, x0, ys]
    g[6, x0 - 1, ys] = g[8, x0, ys]
    g[1, x1 + 1, ys] = g[3, x1, ys]             # :428-437  right face
    g[5, x1 + 1, ys] = g[7, x1, ys]
    g[6, x1 + 1, ys] = g[8, x1, ys]
    g[2, xs, y1 + 1] = g[4, xs, y1]             # :438-447  top face
    g[5, xs, y1 + 1] = g[7, xs, y1]
    g[6, xs, y1 + 1] = g[8, xs, y1]
    g[4, xs, y0 - 1] = g[2, xs, y0]             # :448-457  bottom face
    g[7, xs, y0 - 1] = g[5, xs, y0]
    g[8, xs, y0 - 1] = g[6, xs, y0]
    return g.table()
