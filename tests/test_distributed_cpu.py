"""world_size-2 gloo tests (CPU) of the multi-rank host logic: rendezvous, export
exchange, periodic neighbour rings, scatter/gather -- checked bitwise against the
single-rank oracle (the same gate the GPU path has to meet: 1 rank == N ranks)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from latticeboltzmann_b200.decomposition import Decomposition, axis_extents
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def run_world(tmp_path, boundary, ndx, ndy, nx, ny, nsteps):
    world = ndx * ndy
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "dist_worker.py"), str(tmp_path), boundary,
                                       str(ndx), str(ndy), str(nx), str(ny), str(nsteps)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out
    return np.load(os.path.join(tmp_path, "gathered.npy"))


@pytest.mark.parametrize("boundary,ndx,ndy", [("cavity", 2, 1), ("cavity", 1, 2), ("periodic", 2, 1)])
def test_two_ranks_equal_single_rank_oracle(tmp_path, boundary, ndx, ndy):
    nx, ny, nsteps = 13, 10, 6
    got = run_world(tmp_path, boundary, ndx, ndy, nx, ny, nsteps)
    ref = orc.perturbed_state(nx, ny, seed=21)
    if boundary == "periodic":
        orc.periodic_run(ref, 1.7, nsteps)
    else:
        orc.cavity_run(ref, 1.7, nsteps)
    assert np.array_equal(got, ref)


def test_axis_extents_follow_reference_remainder_rule():
    # cavity_opt2.py:231-239: base = n // nd, the last block takes the remainder
    assert axis_extents(10, 3) == [(0, 3), (3, 3), (6, 4)]
    assert axis_extents(32768, 8)[-1] == (28672, 4096)
    assert axis_extents(7, 1) == [(0, 7)]
    with pytest.raises(ValueError):
        axis_extents(2, 3)


def test_decomposition_rank_layout_and_rings():
    d = Decomposition(100, 60, 4, 2)
    # Create_cart row-major numbering (cavity_opt2.py:225): rank = px*ndy + py
    assert d.coords(5) == (2, 1) and d.rank_of(2, 1) == 5
    # periodic rings: the left neighbour of px=0 is px=ndx-1
    assert d.neighbour(0, 0) == d.rank_of(3, 0)
    assert d.neighbour(7, 1) == d.rank_of(0, 1)
    assert d.neighbour(0, 4) == d.rank_of(3, 1)      # (-1,-1) wraps in both directions
    cover = np.zeros((100, 60), int)
    for r in range(d.size):
        sx, sy = d.slices(r)
        cover[sx, sy] += 1
    assert (cover == 1).all()
    one = Decomposition(9, 9, 1, 1)
    assert one.neighbours(0) == [0] * 8


def test_temporal_mode_decision():
    from latticeboltzmann_b200.decomposition import temporal_mode
    big = Decomposition(16384 * 4, 16384 * 2, 4, 2).blocks()       # bench weak scaling at 8 GPUs
    assert temporal_mode(big, "cavity") == 2 and temporal_mode(big, "cavity", 1) == 1
    assert temporal_mode(Decomposition(4096, 4096).blocks(), "periodic") == 2
    small = Decomposition(512, 512).blocks()
    assert temporal_mode(small, "cavity") == 1                      # automatic: too few fused tiles to fill the GPU
    assert temporal_mode(small, "cavity", 2) == 2                   # forced
    assert temporal_mode(small, "sf_couette", 2) == 1               # simple_flows boundaries: single-step kernel
    assert temporal_mode(Decomposition(53, 47, 8, 1).blocks(), "cavity", 2) == 1      # 6-row slabs are ineligible
    # the smallest block decides (fused tiles counted at 16 rows x 254 columns, 512 needed): 65536 x 2900 split in 2
    # columns of 1450 -> 6 x 4096 tiles each, fine; split 100 ways in x -> blocks of 655 rows x 2900: 41 x 12 = 492
    # tiles -> single steps for everybody
    assert temporal_mode(Decomposition(65536, 2900, 1, 2).blocks(), "cavity") == 2
    assert temporal_mode(Decomposition(65536, 2900, 100, 1).blocks(), "cavity") == 1
    assert temporal_mode(Decomposition(1536, 1536).blocks(), "cavity") == 2       # measured crossover: 1536^2 wins, 1024^2 does not
    assert temporal_mode(Decomposition(1024, 1024).blocks(), "cavity") == 1


def test_temporal_blocking_eligibility_is_collective():
    """A decomposition uses temporal blocking only if EVERY block is at least 16 x 16 (host-side rule that
    DistributedLattice / Lattice apply before stepping; a mixed world would dead-wait on frame-ghost flags)."""
    ok = Decomposition(64, 64, 2, 2)
    assert all(b.lnx >= 16 and b.lny >= 16 for b in ok.blocks())
    mixed = Decomposition(53, 47, 8, 1)          # 6- and 11-row slabs
    assert not all(b.lnx >= 16 and b.lny >= 16 for b in mixed.blocks())
    rem = Decomposition(47, 64, 3, 1)            # 15, 15, 17: the remainder block alone would be eligible
    assert [b.lnx for b in rem.blocks()] == [15, 15, 17]
    assert not all(b.lnx >= 16 for b in rem.blocks())


def test_cpu_baseline_decomposed_run_is_the_references():
    """oracle/opt2_numpy.run_decomposed (what bench.py's reference arm times): the reference's own decomposition
    and ghost exchange.  As in the reference y-splits reproduce the single-rank run bit for
    bit, x-splits differ at the lid corners only (ghost column read instead of the opposite wall)."""
    from oracle import opt2_numpy
    nx, ny, steps = 23, 18, 12
    ref = orc.init_equilibrium(nx, ny)
    orc.cavity_run(ref, 1.7, steps)
    for ndx, ndy in ((1, 1), (1, 2), (1, 3)):
        _, g = opt2_numpy.run_decomposed(ndx, ndy, nx, ny, 1.7, 0, steps, gather=True)
        assert np.array_equal(g, ref), (ndx, ndy)
    _, g = opt2_numpy.run_decomposed(2, 2, nx, ny, 1.7, 0, steps, gather=True)
    assert 0 < np.abs(g - ref).max() < 1e-3
