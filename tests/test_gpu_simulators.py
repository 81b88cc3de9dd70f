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
    """cavity_opt2.py's loop and dump schedule (dump after steps 0, dump_freq, ...), 1 x 1 decomposition."""
    require_gpu()
    from latticeboltzmann_b200.simulators import cavity
    nx, ny, nsteps, dump = 48, 40, 25, 10
    files = cavity.run(1, 1, nx, ny, np.float64, nsteps, dump, 1.7, outdir=str(tmp_path), verbose=False)
    assert sorted(os.path.basename(f) for f in files) == sorted(
        ["%s_%d.npy" % (n, i) for i in (0, 10, 20) for n in ("ux", "uy")])
    f = orc.init_equilibrium(nx, ny)
    done = 0
    for i in (0, 10, 20):
        orc.cavity_run(f, 1.7, i + 1 - done)
        done = i + 1
        rho, ux, uy = orc.moments(f)
        assert np.array_equal(np.load(tmp_path / ("ux_%d.npy" % i)), ux)
        assert np.array_equal(np.load(tmp_path / ("uy_%d.npy" % i)), uy)
