"""Runs the reference's OWN simulators/parallel_lid_drive_cavity/cavity_opt2.py (read from /root/reference at run
time, not copied) under the drop-in modules: `PyLB` / `PyLB.IO` from latticeboltzmann_b200/dropin and `mpi4py`
from latticeboltzmann_b200/mpi_shim (RANK / WORLD_SIZE / MASTER_* describe the world, as under torchrun).

    ref_script_worker.py <script> <backend: gpu|cpu-checker> <nsteps> <dump_freq> ndx ndy nx ny dtype

Only the two run-length constants of the script are substituted (nsteps = 100000, dump_freq = 10000,
cavity_opt2.py:60-63) -- everything else executes verbatim.  backend "cpu-checker" (CPU-only test of the host
side) pre-registers a `_lbkernels` module backed by the test suite's CPU checker; "gpu" uses the real drop-in."""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cpu_checker_lbkernels():
    from oracle import oracle as orc
    m = types.ModuleType("_lbkernels")

    def equilibrium(rho, ux, uy, f):
        orc.equilibrium(rho, ux, uy, f)

    def collide(f, omega):
        orc.collide(f, float(omega))
    m.equilibrium, m.collide = equilibrium, collide
    return m


def main():
    script, backend, nsteps, dump_freq = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    if backend == "cpu-checker":
        sys.modules["_lbkernels"] = cpu_checker_lbkernels()
    from latticeboltzmann_b200 import dropin
    dropin.activate()
    src = open(script).read()
    assert src.count("nsteps = 100000") == 1 and src.count("dump_freq = 10000") == 1
    src = src.replace("nsteps = 100000", "nsteps = %d" % nsteps).replace("dump_freq = 10000", "dump_freq = %d" % dump_freq)
    sys.argv = [script] + sys.argv[5:]
    exec(compile(src, script, "exec"), {"__name__": "__main__", "__file__": script})


if __name__ == "__main__":
    main()
