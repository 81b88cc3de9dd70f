"""The re-hosted drivers (latticeboltzmann_b200/simulators) against the oracle."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from gpu_util import require_gpu

pytestmark = pytest.mark.gpu


def test_shear_wave_driver(tmp_path, monkeypatch):
    """shear_wave_opt2.py with its own parameters (300 x 300, omega = 0.3, 1000 steps, a0 = 1):
    amplitude series == oracle; SURVEY.md §8c known answer a_1000/a_0 = 0.661877475747."""
    require_gpu()
    from latticeboltzmann_b200.simulators import shear_wave
    monkeypatch.chdir(tmp_path)
    shear_wave.main([])
    ampl = np.loadtxt("amplitudes_opt2.out")
    f, uy_k = orc.shear_wave_init(300, 300)
    ref = orc.periodic_run(f, 0.3, 1000, uy_k)
    assert ampl.shape == (1000,)
    assert np.abs(ampl - ref).max() < 1e-13
    a0 = (uy_k * uy_k).sum() * 2 / 300
    assert abs(ampl[-1] / a0 - 0.661877475747) < 1e-9
    nu = shear_wave.viscosity_from_decay(ampl, 300, a0)
    assert abs(nu - 0.943332) < 1e-4                       # SURVEY.md §8c (analytic 0.944444, -1.2e-3 discretisation)


def test_cavity_driver_single_gpu(tmp_path):
    """cavity_opt2.py's loop and dump schedule (dump after steps 0, dump_freq, ... and the final state as
    ux/uy_{nsteps-1}.npy, :279-286), 1 x 1 decomposition."""
    require_gpu()
    from latticeboltzmann_b200.simulators import cavity
    nx, ny, nsteps, dump = 48, 40, 25, 10
    files = cavity.run(1, 1, nx, ny, np.float64, nsteps, dump, 1.7, outdir=str(tmp_path), verbose=False)
    assert sorted(os.path.basename(f) for f in files) == sorted(
        ["%s_%d.npy" % (n, i) for i in (0, 10, 20, 24) for n in ("ux", "uy")])
    f = orc.init_equilibrium(nx, ny)
    done = 0
    for i in (0, 10, 20, 24):
        orc.cavity_run(f, 1.7, i + 1 - done)
        done = i + 1
        rho, ux, uy = orc.moments(f)
        assert np.array_equal(np.load(tmp_path / ("ux_%d.npy" % i)), ux)
        assert np.array_equal(np.load(tmp_path / ("uy_%d.npy" % i)), uy)


def test_cavity_driver_checkpoint_restart_is_bit_identical(tmp_path):
    """An interrupted run continued from its f checkpoint == the uninterrupted run, bitwise (the reference has no
    restart at all: cavity_opt2.py:279-288 dumps velocities only)."""
    require_gpu()
    from latticeboltzmann_b200.simulators import cavity
    nx, ny, nsteps = 64, 48, 37
    a, b = tmp_path / "a", tmp_path / "b"
    a.mkdir()
    b.mkdir()
    cavity.run(1, 1, nx, ny, np.float64, nsteps, 10, 1.7, outdir=str(a), verbose=False)
    cavity.run(1, 1, nx, ny, np.float64, 15, 10, 1.7, outdir=str(b), verbose=False, checkpoint_freq=15)       # "killed" after 15 steps ...
    files = cavity.run(1, 1, nx, ny, np.float64, 15, 10, 1.7, outdir=str(b), verbose=False, checkpoint_freq=5)
    assert any(f.endswith("f_10.npy") for f in files)
    cavity.run(1, 1, nx, ny, np.float64, nsteps, 10, 1.7, outdir=str(b), verbose=False, restart=str(b / "f_10.npy"))   # ... continued from step 10
    for name in ("ux_20.npy", "uy_30.npy", "ux_36.npy", "uy_36.npy"):
        assert np.array_equal(np.load(a / name), np.load(b / name)), name


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_checkpoint_restart_under_temporal_blocking(tmp_path, dt):
    """Forced two-steps-per-pass mode, odd step counts on both sides of the interruption, a different block
    decomposition after the restart: interrupted == uninterrupted == oracle, bit for bit."""
    lb = require_gpu()
    nx, ny = 70, 530
    f0 = orc.perturbed_state(nx, ny, dt, seed=8)
    ref = f0.copy()
    orc.cavity_run(ref, 1.7, 15 + 23)
    lat = lb.Lattice(nx, ny, "cavity", omega=1.7, dtype=dt, temporal=2)
    lat.upload(f0)
    lat.step(15)
    fn = str(tmp_path / "f_15.npy")
    lat.save_checkpoint(fn, omega=1.7)
    lat.close()
    assert np.load(fn).shape == (9, nx, ny) and np.load(fn).dtype == dt
    lat = lb.Lattice(nx, ny, "cavity", omega=1.7, dtype=dt, ndx=2, ndy=2, temporal=2)
    meta = lat.load_checkpoint(fn)
    assert meta["steps_done"] == 15 and meta["omega"] == 1.7
    lat.step(23)
    assert np.array_equal(lat.download(), ref)
    lat.health()
    lat.close()


def test_sliding_lid_mpi_driver():
    """slidingLidMPI.py on one rank (base 40, 300 steps, its Re / uw / relaxation formula) == the literal numpy loop."""
    require_gpu()
    from latticeboltzmann_b200.simulators import sliding_lid_mpi
    from oracle import simple_flows as sf
    base, steps, re, uw = 40, 300, 1000.0, 0.1
    ux, uy, _ = sliding_lid_mpi.run(base, steps, re, uw, verbose=False)
    n = base + 2
    f = sf.feq(np.ones((n, n)), np.zeros((n, n)), np.zeros((n, n)))
    assert np.array_equal(f, sliding_lid_mpi.initial_state(n))
    omega = (2 * re) / (6 * base * uw + re)
    for _ in range(steps):
        sf.sliding_lid_mpi_step(f, omega, uw)
    _, rux, ruy = sf.moments(f)
    assert np.array_equal(ux, rux[1:-1, 1:-1]) and np.array_equal(uy, ruy[1:-1, 1:-1])
    assert np.abs(ux).max() > 1e-3          # the lid drags the fluid
