"""Drop-in for PyLB/IO.py: ``save_mpiio(comm, fn, g_kl)`` (PyLB/IO.py:27-80).

Writes a global two-dimensional array to a single .npy file readable with ``numpy.load``.
`comm` may be a real mpi4py Cartesian communicator, the single-process shim in
``latticeboltzmann_b200/dropin/mpi4py``, or ``None`` (one process).  The reference computes the
global shape with ``Allreduce`` over the sub-communicators and the offsets with ``Exscan``
(:50-70); the same calls are used here, then each rank stores its rows at their global offsets
(the reference's MPI-IO vector file view, :72-78).  Unlike the reference this works on numpy >= 1.23
(the original calls the removed ``np.asscalar``, :57).
"""
import numpy as np

from latticeboltzmann_b200 import npyio


def save_mpiio(comm, fn, g_kl):
    g_kl = np.ascontiguousarray(g_kl)
    local_nx, local_ny = g_kl.shape
    if comm is None:
        return npyio.write_block(fn, g_kl, 0, 0, local_nx, local_ny, True)
    nx = np.zeros(1, dtype=np.int64)
    ny = np.zeros(1, dtype=np.int64)
    commx = comm.Sub((True, False))
    commy = comm.Sub((False, True))
    commx.Allreduce(np.array([local_nx], dtype=np.int64), nx)
    commy.Allreduce(np.array([local_ny], dtype=np.int64), ny)
    offx = np.zeros(1, dtype=np.int64)
    offy = np.zeros(1, dtype=np.int64)
    commx.Exscan(np.array([local_nx], dtype=np.int64), offx)
    commy.Exscan(np.array([local_ny], dtype=np.int64), offy)
    rank = comm.Get_rank()
    if rank == 0:
        import os
        if os.path.exists(fn):
            os.remove(fn)
    comm.Barrier()
    npyio.write_block(fn, g_kl, int(offx[0]), int(offy[0]), int(nx[0]), int(ny[0]), rank == 0)
    comm.Barrier()
