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
        outdir=".", verbose=True, checkpoint_freq=0, restart=None):
    """Steps are numbered like the reference's loop variable (`for i in range(nsteps)`, :272).  After step i with
    i % dump_freq == 0 the velocities are dumped (:279-283); after the loop they are dumped once more as
    ux/uy_{nsteps-1}.npy (:285-286 -- from the FINAL state here; the reference re-writes the moments of its last
    periodic dump).  Additions the reference lacks: every `checkpoint_freq` completed steps the populations go to
    f_{steps}.npy (+ .json), and `restart=<that file>` continues a run bit-identically."""
    rank, world, local = D.init_process_group()
    if rank == 0 and verbose:
        print("Running in parallel on {} processes (one per GPU).".format(world))
        print("Domain decomposition: {} x {} blocks.".format(ndx, ndy))
        print("Global grid has size {}x{}.".format(nx, ny))
        print("Using {} floating point data type.".format(np.dtype(dtype)))
    lat = D.DistributedLattice(nx, ny, ndx, ndy, "cavity", omega=float(omega), u_wall=u0, dtype=dtype,
                               arith=arith, device=local)
    import torch.distributed as dist
    done = 0                                                      # completed steps = next value of the loop variable
    if restart:
        done = int(lat.load_checkpoint(restart).get("steps_completed", 0))
    else:
        lat.init_equilibrium()                                    # :265-269
    written = []

    def dump(i):
        _, ux, uy = lat.block.moments()                           # :280-281
        for name, field in (("ux", ux), ("uy", uy)):
            fn = os.path.join(outdir, "{}_{}.npy".format(name, i))
            npyio.save_field(fn, field, lat.decomp, rank, dist.barrier)
            written.append(fn)

    while done < nsteps:
        # run to the next step index that is dumped (i % dump_freq == 0) or checkpointed, in one go
        stops = [nsteps, (done + dump_freq - 1) // dump_freq * dump_freq + 1]
        if checkpoint_freq:
            stops.append((done // checkpoint_freq + 1) * checkpoint_freq)
        nxt = min(s for s in stops if s > done)
        lat.step(nxt - done)
        done = nxt
        if (done - 1) % dump_freq == 0:                           # :279, i = done - 1
            dump(done - 1)
        if checkpoint_freq and done % checkpoint_freq == 0 and done < nsteps:
            fn = os.path.join(outdir, "f_{}.npy".format(done))
            lat.save_checkpoint(fn, steps_completed=done, omega=float(omega), u0=u0, nx=nx, ny=ny)
            written.append(fn)
        if rank == 0 and verbose:
            sys.stdout.write("=== Step {}/{} ===\r".format(done, nsteps))
    if nsteps > 0 and (nsteps - 1) % dump_freq != 0:
        dump(nsteps - 1)                                          # :285-286
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
