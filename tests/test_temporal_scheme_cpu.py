"""CPU proof of the temporal-blocking partition used by csrc/temporal.cuh (K1 / K2 / K3), with the oracle
as the single-step operator and NaN poisoning as the dependency tracker:

  K1  level n+1 on the FRAME (distance < 3) needs level-n cells of distance < 4 (+ the periodic ring);
  K2  level n+2 on the DEEP INTERIOR (distance >= 2) needs level n+1 on distance >= 1, which needs only real
      level-n cells and never a wall rule;
  K3  level n+2 on distance < 2 needs level n+1 on distance < 3 only.

If a region's result is NaN-free although everything outside its claimed sources was poisoned, and equal to the
plain two-step oracle result, the three kernels together reproduce two single steps exactly."""
import numpy as np
import pytest

from oracle import oracle as orc


def dist_map(nx, ny):
    k = np.arange(nx)[:, None]
    l = np.arange(ny)[None, :]
    return np.minimum(np.minimum(k, nx - 1 - k), np.minimum(l, ny - 1 - l))


def step(f, boundary, omega=1.7):
    g = np.empty_like(f)
    if boundary == "periodic":
        orc.periodic_step_pull(f, g, omega)
    else:
        orc.cavity_step_pull(f, g, omega, 0.1, walls_lr=(boundary == "cavity"))
    return g


@pytest.mark.parametrize("boundary", ["periodic", "cavity", "cavity_xperiodic"])
@pytest.mark.parametrize("shape", [(23, 19), (16, 16), (40, 17)])
def test_double_step_partition_is_closed_and_exact(boundary, shape):
    nx, ny = shape
    d = dist_map(nx, ny)
    f0 = orc.perturbed_state(nx, ny, seed=nx + ny)
    l1 = step(f0, boundary)
    l2 = step(l1, boundary)

    # K1: the frame at level n+1 from level-n cells closer than 4 to the perimeter
    p = f0.copy()
    p[:, d >= 4] = np.nan
    k1 = step(p, boundary)
    assert not np.isnan(k1[:, d < 3]).any()
    assert np.array_equal(k1[:, d < 3], l1[:, d < 3])

    # K3: distance < 2 at level n+2 from the frame at level n+1
    p = l1.copy()
    p[:, d >= 3] = np.nan
    k3 = step(p, boundary)
    assert not np.isnan(k3[:, d < 2]).any()
    assert np.array_equal(k3[:, d < 2], l2[:, d < 2])

    # K2: the deep interior at level n+2 from level n+1 on distance >= 1, itself from real level-n cells only
    # (poisoning the perimeter of level n+1 proves no wall rule and no ghost is involved)
    p = l1.copy()
    p[:, d < 1] = np.nan
    k2 = step(p, boundary)
    assert not np.isnan(k2[:, d >= 2]).any()
    assert np.array_equal(k2[:, d >= 2], l2[:, d >= 2])

    # the two level-(n+2) regions tile the block
    assert ((d < 2) | (d >= 2)).all()


# ---- the fused tile's ring (csrc/temporal.cuh: t2_tile) ---------------------------------------------------------
# A model of the DATA MOVEMENT of one fused tile with the kernel's own index expressions (ring_slots / ring_base /
# mirror_k / LBM_SLOT / lag registers), on tagged values instead of numbers: level-(n+1) value (i, row, col) must
# arrive in the level-(n+2) pull of cell (row + cx_i, col + cy_i).  The ring is sized for ONE barrier per row: a
# warp that is already past the barrier writes the next row while slower warps still read, so the model performs
# the writes of iteration s+1 BEFORE the reads of iteration s (the worst interleaving the barrier allows).  rev=True
# is the mirrored (top-down) walk with E/W roles swapped: measured on the GPU, bit-identical, not faster, not
# shipped (profiles/r02_t2_order_sweep.log) -- kept here because it pins the ring sizing argument from both sides.
CX = [0, 1, 0, -1, 0, 1, -1, -1, 1]
CY = [0, 0, 1, 0, -1, 1, 1, -1, -1]
Q0, QE, QN, QW, QS, QNE, QNW, QSW, QSE = range(9)


def _ring_slots(i):
    return 2 if i in (QNW, QSW) else 3 if i in (QN, QS) else 4


def _ring_base(i):
    order = [QNW, QSW, QN, QS, QNE, QSE]
    return sum(_ring_slots(j) for j in order[:order.index(i)])


def _mirror_k(i):
    return {QE: QW, QW: QE, QNE: QNW, QNW: QNE, QSW: QSE, QSE: QSW}.get(i, i)


@pytest.mark.parametrize("rev", [False, True])
@pytest.mark.parametrize("rows", [1, 2, 5, 16])
def test_fused_tile_ring_walk(rev, rows):
    width = 12                                   # threads of the model tile
    k0, k1 = 2, 2 + rows                         # emits rows k0 .. k1-1 from level-(n+1) rows k0-1 .. k1
    nit = k1 - k0 + 2
    direction = -1 if rev else 1
    jfirst = k1 if rev else k0 - 1
    qlag, qlead = (QW, QE) if rev else (QE, QW)
    rs = lambda i: _ring_slots(_mirror_k(i) if rev else i)
    rb = lambda i: _ring_base(_mirror_k(i) if rev else i)
    ring = {}
    shifted = (QN, QS, QNE, QNW, QSW, QSE)

    def slot(i, s, d):
        n = rs(i)
        return (s - d) % n                       # (s - d) & 1, (s3 + 3 - d) % 3 with s3 = s % 3, (s - d) & 3

    def write(s):
        row = jfirst + direction * s
        for t in range(width):
            for i in shifted:
                ring[(rb(i) + slot(i, s, 0), t)] = (i, row, t)

    emitted = {}
    regs = {t: {"rest_m1": None, "lag_m1": None, "lag_m2": None} for t in range(width)}
    write(0)
    for s in range(nit):
        row = jfirst + direction * s
        if s + 1 < nit:
            write(s + 1)                         # a fast warp is already one row ahead
        for t in range(width):
            r = regs[t]
            if s >= 2 and 1 <= t < width - 1:
                g = {Q0: r["rest_m1"], qlag: r["lag_m2"], qlead: (qlead, row, t)}
                for i in shifted:
                    g[i] = ring[(rb(i) + slot(i, s, 1 + direction * CX[i]), t - CY[i])]
                emitted[(row - direction, t)] = g
            r["lag_m2"], r["lag_m1"], r["rest_m1"] = r["lag_m1"], (qlag, row, t), (Q0, row, t)
    assert len({rb(i) + n for i in shifted for n in range(rs(i))}) == 18      # the six populations tile the 18 ring rows
    assert sorted({k for k, _ in emitted}) == list(range(k0, k1))
    for (row, t), g in emitted.items():
        for i in range(9):
            assert g[i] == (i, row - CX[i], t - CY[i]), (rev, row, t, i, g[i])
