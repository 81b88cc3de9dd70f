"""simulators/serial_shear_wave/Python/shear_wave_opt2.py, re-hosted on the device-resident lattice.

    python -m latticeboltzmann_b200.simulators.shear_wave [nx ny nsteps omega dtype]

Defaults are the reference's module constants (:42-52: 300 x 300, 1000 steps, omega = 0.3, float64).
Initialises rho = 1, ux = 0, uy(k, l) = sin(2 pi k / nx) (:79-88), advances `stream; collide`
(:95-97) as one fused kernel per step, records the Fourier amplitude of :99 after every step ON
THE DEVICE, and writes ``amplitudes_opt2.out`` with np.savetxt (:103).
"""
import sys

import numpy as np

from .. import Lattice


def run(nx=300, ny=300, nsteps=1000, omega=0.3, dtype=np.float64, arith="exact", a0=1.0, device=0):
    dtype = np.dtype(dtype)
    uy_k = (a0 * np.sin(2 * np.pi / nx * np.arange(nx))).astype(dtype)
    lat = Lattice(nx, ny, "periodic", omega=omega, dtype=dtype, arith=arith, devices=device)
    lat.init_equilibrium(uy=np.resize(uy_k, (ny, nx)).T)
    lat.probe_shear_enable(uy_k, nsteps)
    lat.step(nsteps)
    ampl = lat.probe_shear_read(nsteps)
    lat.health()
    lat.close()
    return ampl


def viscosity_from_decay(ampl, nx, a_init=None):
    """nu from a(t) = a0 exp(-nu k^2 t); analytic value (1/omega - 1/2)/3 (shear_wave_decay.py:162)."""
    t = np.arange(1, len(ampl) + 1)
    a_init = ampl[0] if a_init is None else a_init
    slope = np.polyfit(t, np.log(np.asarray(ampl, float) / a_init), 1)[0]
    return -slope / (2 * np.pi / nx) ** 2


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    nx = int(argv[0]) if len(argv) > 0 else 300
    ny = int(argv[1]) if len(argv) > 1 else 300
    nsteps = int(argv[2]) if len(argv) > 2 else 1000
    omega = float(argv[3]) if len(argv) > 3 else 0.3
    dtype = np.dtype(argv[4]) if len(argv) > 4 else np.float64
    ampl = run(nx, ny, nsteps, omega, dtype)
    np.savetxt("amplitudes_opt2.out", ampl)
    nu = viscosity_from_decay(ampl, nx)
    print("viscosity from decay: %.6f   analytic (1/omega - 1/2)/3: %.6f" % (nu, (1 / omega - 0.5) / 3))


if __name__ == "__main__":
    main()
