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
