"""torchrun worker for the multi-GPU parity test: one rank per GPU, halo pushes over
NVLink peer memory (CUDA IPC).  Rank 0 compares the gathered field with the
single-rank CPU oracle BITWISE and writes a verdict file."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from latticeboltzmann_b200 import distributed as D     # noqa: E402
from oracle import oracle as orc                       # noqa: E402


def main():
    out, boundary, ndx, ndy, nx, ny, nsteps = sys.argv[1], sys.argv[2], *map(int, sys.argv[3:8])
    # LBM_TEST_SHARE_GPUS=<n>: more ranks than GPUs -- ranks share devices (rank % n), the halo path is then CUDA IPC
    # between PROCESSES on one device (system-scope flags, time-sliced contexts); NCCL refuses duplicate devices, so
    # the rendezvous runs over gloo.
    share = int(os.environ.get("LBM_TEST_SHARE_GPUS", "0"))
    rank, world, local = D.init_process_group("gloo" if share else "nccl")
    if share:
        local = rank % share
    f0 = orc.perturbed_state(nx, ny, seed=33)
    # temporal=2: force the two-steps-per-pass mode (the automatic mode would pick the single-step kernel for
    # blocks this small), so that the level-(n+1) frame-ghost exchange crosses NVLink as well
    lat = D.DistributedLattice(nx, ny, ndx, ndy, boundary, omega=1.7, u_wall=0.1, arith="exact", device=local,
                               temporal=2)
    lat.upload_global(f0)
    lat.step(nsteps)
    got = lat.gather_f()
    lat.health()
    verdict = None
    if rank == 0:
        ref = f0.copy()
        if boundary == "periodic":
            orc.periodic_run(ref, 1.7, nsteps)
        else:
            orc.cavity_run(ref, 1.7, nsteps, 0.1, walls_lr=(boundary == "cavity"))
        verdict = {"bit_exact": bool(np.array_equal(got, ref)), "max_abs": float(np.abs(got - ref).max()),
                   "world": world, "boundary": boundary, "ndx": ndx, "ndy": ndy}
        with open(out, "w") as fh:
            json.dump(verdict, fh)
    lat.close()
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
