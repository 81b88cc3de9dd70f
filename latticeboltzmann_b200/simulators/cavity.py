"""simulators/parallel_lid_drive_cavity/cavity_opt2.py, re-hosted: one process per GPU.

    torchrun --nproc-per-node NDX*NDY -m latticeboltzmann_b200.simulators.cavity ndx ndy nx ny dtype [nsteps dump_freq omega]
    python -m latticeboltzmann_b200.simulators.cavity 1 1 nx ny float64          (single GPU)

Same five positional arguments as the reference (:49-57); nsteps = 100000, dump_freq = 10000 and
omega = 1.7 are its hard-coded constants (:60-66) and may be overridden by three more arguments.
Per step the reference does communicate(); stream_and_bounce_back(); collide() (:275-277) -- here
that is one fused kernel launch per step with the halo exchange inside.  Every dump_freq steps
(``i % dump_freq == 0``, :279) the velocities are written to ``ux_{i}.npy`` / ``uy_{i}.npy`` (:282-283)
as single .npy files (each rank stores its rows; format identical to save_mpiio).
"""
import os
import sys

import numpy as np

from .. import distributed as D
from .. import npyio


def run(ndx, ndy, nx, ny, dtype=np.float64, nsteps=100000, dump_freq=10000, omega=1.7, u0=0.1, arith="exact",
        outdir=".", verbose=True):
    rank, world, local = D.init_process_group()
    if rank == 0 and verbose:
        print("Running in parallel on {} processes (one per GPU).".format(world))
        print("Domain decomposition: {} x {} blocks.".format(ndx, ndy))
        print("Global grid has size {}x{}.".format(nx, ny))
        print("Using {} floating point data type.".format(np.dtype(dtype)))
    lat = D.DistributedLattice(nx, ny, ndx, ndy, "cavity", omega=float(omega), u_wall=u0, dtype=dtype,
                               arith=arith, device=local)
    lat.init_equilibrium()                                        # :265-269
    import torch.distributed as dist
    i = 0
    written = []
    while i < nsteps:
        lat.step(1)                                               # step i
        if i % dump_freq == 0:                                    # :279 (dumps after step 0, dump_freq, ...)
            _, ux, uy = lat.block.moments()                       # :280-281
            for name, field in (("ux", ux), ("uy", uy)):
                fn = os.path.join(outdir, "{}_{}.npy".format(name, i))
                npyio.save_field(fn, field, lat.decomp, rank, dist.barrier)
                written.append(fn)
        nxt = min(nsteps, (i // dump_freq + 1) * dump_freq)       # run to the next dump step in one go
        if nxt - (i + 1) > 0:
            lat.step(nxt - (i + 1))
        i = nxt
        if rank == 0 and verbose:
            sys.stdout.write("=== Step {}/{} ===\r".format(i, nsteps))
    lat.health()
    lat.close()
    return written


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    ndx, ndy, nx, ny = (int(a) for a in argv[:4])
    dtype = np.dtype(argv[4]) if len(argv) > 4 else np.float64
    nsteps = int(argv[5]) if len(argv) > 5 else 100000
    dump_freq = int(argv[6]) if len(argv) > 6 else 10000
    omega = float(argv[7]) if len(argv) > 7 else 1.7
    run(ndx, ndy, nx, ny, dtype, nsteps, dump_freq, omega)


if __name__ == "__main__":
    main()
