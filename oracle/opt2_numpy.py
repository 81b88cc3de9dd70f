"""TEST INFRASTRUCTURE ONLY -- the reference's opt2 step in its own *structure*:
numpy ``np.roll`` streaming + numpy slice stores for the walls (Python side,
cavity_opt2.py:94-177) followed by the compiled collide (c/d2q9.h:121-131, here
the C oracle).  This is what ``bench.py`` times as the CPU baseline / reference
arm: the per-step cost profile (8 rolls with temporaries, ~30 strided slice
assignments, one serial C loop) is the reference's, which a fused C loop would
not represent.  Pinned bit-exactly against the golden vectors produced by the
reference's own function (tests/test_oracle_golden.py).
"""
import numpy as np

from . import oracle as orc

# (destination channel, source channel) pairs of the half-way bounce-back per wall
# (cavity_opt2.py:134-136, 143-145, 150-157); names E=1 N=2 W=3 S=4 NE=5 NW=6 SW=7 SE=8.
_BOTTOM = ((2, 4), (5, 7), (6, 8))
_LEFT = ((1, 3), (5, 7), (8, 6))
_RIGHT = ((3, 1), (6, 8), (7, 5))


def stream(f):
    """PyLB/Streaming.py:45-46."""
    for i in range(1, 9):
        f[i] = np.roll(f[i], orc.C_IC[i], axis=(0, 1))


def stream_and_bounce_back(f, u0=0.1, walls_lr=True):
    """cavity_opt2.py:109-177 (the four redundant corner blocks :160-177 repeat
    values the wall lines already stored and are therefore omitted)."""
    w_se = f.dtype.type(1 / 36)
    bottom = f[:, :, 0].copy()
    top = f[:, :, -1].copy()
    left = f[:, 0, :].copy()
    right = f[:, -1, :].copy()
    stream(f)
    for dst, src in _BOTTOM:
        f[dst, :, 0] = bottom[src]
    rho = top[6] + top[2] + top[5] + f[6, :, -1] + f[2, :, -1] + f[5, :, -1] + f[3, :, -1] + f[0, :, -1] + f[1, :, -1]
    f[4, :, -1] = top[2]
    f[8, :, -1] = top[6] + 6 * w_se * rho * u0
    f[7, :, -1] = top[5] - 6 * w_se * rho * u0
    if walls_lr:
        for dst, src in _LEFT:
            f[dst, 0, :] = left[src]
        for dst, src in _RIGHT:
            f[dst, -1, :] = right[src]


def cavity_step(f, omega, u0=0.1):
    """cavity_opt2.py:275-277 on one rank."""
    stream_and_bounce_back(f, u0)
    orc.collide(f.reshape(9, -1), omega)


def _worker(args):
    import time
    nx, ny, omega, warmup, steps = args
    f = orc.init_equilibrium(nx, ny)
    for _ in range(warmup):
        cavity_step(f, omega)
    t0 = time.perf_counter()
    for _ in range(steps):
        cavity_step(f, omega)
    return time.perf_counter() - t0, float(f.sum())


def run_independent_blocks(nproc, nx, ny, omega, warmup, steps):
    """nproc processes, each advancing its own (nx, ny) single-rank opt2 cavity block
    (the reference's one-MPI-rank-per-block model with the halo
    exchange left out -- mpirun/mpi4py are not installed).  Returns the slowest
    process's seconds for `steps` steps."""
    import multiprocessing as mp
    orc.build()
    ctx = mp.get_context("fork")
    with ctx.Pool(nproc) as pool:
        res = pool.map(_worker, [(nx, ny, omega, warmup, steps)] * nproc)
    return max(r[0] for r in res)



# ---- the reference's decomposed run: one process per block, ghost exchange every step -------------------------
def _local_layout(n, nd, p):
    """cavity_opt2.py:231-258 along one axis: (global offset, real cells, ghost below?, ghost above?)."""
    base = n // nd
    real = base if p < nd - 1 else n - base * (nd - 1)
    return p * base, real, p > 0, p < nd - 1


def _decomposed_worker(rank, ndx, ndy, nx, ny, omega, warmup, steps, box, barrier, times, out):
    import time
    px, py = divmod(rank, ndy)                       # Create_cart row-major (cavity_opt2.py:225)
    x0, rx, gxl, gxr = _local_layout(nx, ndx, px)
    y0, ry, gyb, gyt = _local_layout(ny, ndy, py)
    lnx, lny = rx + gxl + gxr, ry + gyb + gyt
    f = orc.init_equilibrium(lnx, lny)               # :265-269, ghosts included
    view = lambda r, name, shape: np.frombuffer(box[r][name], dtype=np.float64).reshape(shape)   # noqa: E731

    def communicate():
        # :191-199 -- x direction: real column 1 -> left neighbour's ghost -1, real column -2 -> right neighbour's ghost 0
        if gxl:
            view(rank, "to_left", (9, lny))[...] = f[:, 1, :]
        if gxr:
            view(rank, "to_right", (9, lny))[...] = f[:, -2, :]
        barrier.wait()
        if gxr:
            f[:, -1, :] = view(rank + ndy, "to_left", (9, lny))
        if gxl:
            f[:, 0, :] = view(rank - ndy, "to_right", (9, lny))
        barrier.wait()
        # :201-210 -- y direction, x ghosts included (diagonals arrive in two hops)
        if gyb:
            view(rank, "to_bottom", (9, lnx))[...] = f[:, :, 1]
        if gyt:
            view(rank, "to_top", (9, lnx))[...] = f[:, :, -2]
        barrier.wait()
        if gyt:
            f[:, :, -1] = view(rank + 1, "to_bottom", (9, lnx))
        if gyb:
            f[:, :, 0] = view(rank - 1, "to_top", (9, lnx))
        barrier.wait()

    def step():
        communicate()
        cavity_step(f, omega)                        # :276-277 on the local array, ghosts included, like the reference

    for _ in range(warmup):
        step()
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    barrier.wait()
    times[rank] = time.perf_counter() - t0
    if out is not None:
        g = np.frombuffer(out, dtype=np.float64).reshape(9, nx, ny)
        g[:, x0:x0 + rx, y0:y0 + ry] = f[:, int(gxl):int(gxl) + rx, int(gyb):int(gyb) + ry]


def run_decomposed(ndx, ndy, nx, ny, omega, warmup, steps, gather=False):
    """The reference's parallel cavity as it runs under mpirun (cavity_opt2.py:214-277): ndx x ndy processes, one
    block each (non-periodic Cartesian topology, ghost layers only towards existing neighbours, the last block of
    an axis takes the remainder), per step communicate() -> stream_and_bounce_back() -> collide().  mpirun / mpi4py
    are not installed, so the four Sendrecv of communicate() are copies through shared-memory mailboxes ordered by
    process barriers.  Returns (seconds of the slowest process for `steps` steps, gathered field or None)."""
    import multiprocessing as mp
    orc.build()
    ctx = mp.get_context("fork")
    n = ndx * ndy
    box = []
    for r in range(n):
        px, py = divmod(r, ndy)
        _, rx, gxl, gxr = _local_layout(nx, ndx, px)
        _, ry, gyb, gyt = _local_layout(ny, ndy, py)
        lnx, lny = rx + gxl + gxr, ry + gyb + gyt
        box.append({"to_left": ctx.RawArray("d", 9 * lny), "to_right": ctx.RawArray("d", 9 * lny),
                    "to_bottom": ctx.RawArray("d", 9 * lnx), "to_top": ctx.RawArray("d", 9 * lnx)})
    barrier = ctx.Barrier(n)
    times = ctx.RawArray("d", n)
    out = ctx.RawArray("d", 9 * nx * ny) if gather else None
    procs = [ctx.Process(target=_decomposed_worker, args=(r, ndx, ndy, nx, ny, omega, warmup, steps, box, barrier, times, out))
             for r in range(n)]
    for p in procs:
        p.start()
    for p in procs:
        p.join()
        if p.exitcode != 0:
            raise RuntimeError("a worker of the decomposed CPU run failed (exit code %s)" % p.exitcode)
    g = np.frombuffer(out, dtype=np.float64).reshape(9, nx, ny).copy() if gather else None
    return max(times), g
