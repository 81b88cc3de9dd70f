"""simulators/simple_flows/slidingLidMPI.py, re-hosted on one GPU.

    python -m latticeboltzmann_b200.simulators.sliding_lid_mpi [base_length steps re uw]

The reference's parameters (``call()``, :310-327): base_lenght = 300, steps = 1 000 000, Re = 1000, uw = 0.1,
relaxation = 2 Re / (6 * base_lenght * uw + Re).  On one rank its loop (:264-268) is stream -> bounce_back_choosen
(all four walls, full index ranges, :180-204) -> moments -> collision on an array of (base + 2)^2 nodes that
includes the wall layers; ``comunicate`` is a no-op.  Here the wall sequence is a per-cell boundary table
(boundary_table.sliding_lid_mpi_table) and all steps run inside the resident multi-step kernel; results are
bit-identical to the numpy loop.  The only performance numbers in the reference tree are this script's wall times
on bwUniCluster (amdahldataviewer.py:40-49: 4.5 MLUPS on 1 rank ... 237 MLUPS on 400 ranks).
"""
import sys
import time

import numpy as np

from .. import boundary_table
from ..lattice import Lattice


def initial_state(n):
    """equilibrium(rho = 1, u = 0) as slidingLidMPI.py:127-145 evaluates it: (2 rho/9)(2 - uu), (rho/18)(2 ...), (rho/36)(1 ...)."""
    rho = 1.0
    w = np.array([(2 * rho / 9) * 2.0] + [(rho / 18) * 2.0] * 4 + [(rho / 36) * 1.0] * 4)
    return np.ascontiguousarray(np.broadcast_to(w[:, None, None], (9, n, n)))


def run(base_length=300, steps=1000000, re=1000.0, uw=0.1, device=0, verbose=True):
    """-> (ux, uy) of the base_length^2 fluid nodes (what the reference's plotter recomputes, :280), seconds."""
    relaxation = (2 * re) / (6 * base_length * uw + re)            # :316
    n = base_length + 2                                            # :100-101 on one rank
    lat = Lattice(n, n, "sf_table", omega=relaxation, u_wall=uw, devices=device)
    lat.set_boundary_table(*boundary_table.sliding_lid_mpi_table(n, n, uw))
    lat.upload(initial_state(n))
    lat.sync()
    t0 = time.perf_counter()
    lat.step(steps)
    lat.sync()
    dt = time.perf_counter() - t0
    lat.health()
    _, ux, uy = lat.moments()
    lat.close()
    if verbose:
        print("sliding lid %dx%d, %d steps, relaxation %.6f: %.2f s, %.0f MLUPS (fluid nodes)" %
              (base_length, base_length, steps, relaxation, dt, base_length * base_length * steps / dt / 1e6))
    return ux[1:-1, 1:-1], uy[1:-1, 1:-1], dt


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    base = int(argv[0]) if len(argv) > 0 else 300
    steps = int(argv[1]) if len(argv) > 1 else 1000000
    re = float(argv[2]) if len(argv) > 2 else 1000.0
    uw = float(argv[3]) if len(argv) > 3 else 0.1
    run(base, steps, re, uw)


if __name__ == "__main__":
    main()
