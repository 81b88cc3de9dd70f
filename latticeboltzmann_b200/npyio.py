"""Single-file .npy output of a decomposed 2-D field, byte-compatible with the reference's
``save_mpiio`` (PyLB/IO.py:27-80): magic ``\\x93NUMPY\\x01\\x00``, little-endian int16 header
length, the dict literal padded with spaces so that the payload starts 16-byte aligned,
C-order ``(nx, ny)`` payload.  Each writer stores its block's rows at their global offsets
(what the reference does with an MPI-IO vector file view, PyLB/IO.py:72-78), so no gather
is needed.  ``np.load`` reads the result.

Checkpoint / restart (the reference has none: cavity_opt2.py:279-288 only dumps velocities, from which a run
cannot be continued): ``save_checkpoint`` stores the populations ``f`` as one ``(9, nx, ny)`` .npy file in the same
format -- every rank writes its block's rows at their global offsets -- plus a small JSON sidecar (step count,
lattice parameters); ``load_checkpoint_block`` memory-maps the file and returns one block's part.  A run continued
from a checkpoint is bit-identical to the uninterrupted run (tests/test_gpu_simulators.py).
"""
import os

import numpy as np

MAGIC = b"\x93NUMPY\x01\x00"


def npy_header(shape, dtype):
    """Header bytes exactly as PyLB/IO.py:56-62 builds them."""
    from numpy.lib.format import dtype_to_descr
    d = str({"descr": dtype_to_descr(np.dtype(dtype)), "fortran_order": False,
             "shape": tuple(int(s) for s in shape)})
    while (len(d) + len(MAGIC) + 2) % 16 != 15:
        d += " "
    d += "\n"
    return MAGIC + np.int16(len(d)).tobytes() + d.encode("latin-1")


def write_block(fn, g_kl, x0, y0, nx, ny, write_header):
    """Store the local block g_kl (lnx, lny) of a global (nx, ny) array into file `fn`."""
    g_kl = np.ascontiguousarray(g_kl)
    hdr = npy_header((nx, ny), g_kl.dtype)
    item = g_kl.dtype.itemsize
    fd = os.open(fn, os.O_CREAT | os.O_WRONLY, 0o644)
    try:
        if write_header:
            os.pwrite(fd, hdr, 0)
        lnx, lny = g_kl.shape
        if lny == ny:      # full rows: one contiguous write
            os.pwrite(fd, g_kl.tobytes(), len(hdr) + (x0 * ny + y0) * item)
        else:
            for k in range(lnx):
                os.pwrite(fd, g_kl[k].tobytes(), len(hdr) + ((x0 + k) * ny + y0) * item)
    finally:
        os.close(fd)


def save_field(fn, g_kl, decomp=None, rank=0, barrier=None):
    """Collective write of a decomposed field: every rank calls this with its local block."""
    if decomp is None:
        nx, ny = g_kl.shape
        return write_block(fn, g_kl, 0, 0, nx, ny, True)
    b = decomp.block(rank)
    if rank == 0 and os.path.exists(fn):
        os.remove(fn)
    if barrier:
        barrier()
    write_block(fn, g_kl, b.x0, b.y0, decomp.nx, decomp.ny, rank == 0)
    if barrier:
        barrier()


load_field = np.load


def write_f_block(fn, f_ikl, x0, y0, nx, ny, write_header):
    """Store the local populations f_ikl (9, lnx, lny) of a global (9, nx, ny) array into file `fn`."""
    f_ikl = np.ascontiguousarray(f_ikl)
    hdr = npy_header((9, nx, ny), f_ikl.dtype)
    item = f_ikl.dtype.itemsize
    fd = os.open(fn, os.O_CREAT | os.O_WRONLY, 0o644)
    try:
        if write_header:
            os.pwrite(fd, hdr, 0)
        _, lnx, lny = f_ikl.shape
        for i in range(9):
            base = len(hdr) + i * nx * ny * item
            if lny == ny:
                os.pwrite(fd, f_ikl[i].tobytes(), base + (x0 * ny) * item)
            else:
                for k in range(lnx):
                    os.pwrite(fd, f_ikl[i, k].tobytes(), base + ((x0 + k) * ny + y0) * item)
    finally:
        os.close(fd)


def save_checkpoint(fn, f_local, decomp=None, rank=0, barrier=None, meta=None):
    """Collective: every rank passes its block's populations (9, lnx, lny).  Rank 0 also writes `fn + ".json"`."""
    import json
    if decomp is None:
        _, nx, ny = f_local.shape
        x0 = y0 = 0
    else:
        b = decomp.block(rank)
        nx, ny, x0, y0 = decomp.nx, decomp.ny, b.x0, b.y0
    if rank == 0:
        if os.path.exists(fn):
            os.remove(fn)
        with open(fn + ".json", "w") as fh:
            json.dump(dict(meta or {}, shape=[9, nx, ny], dtype=str(np.dtype(f_local.dtype))), fh)
    if barrier:
        barrier()
    write_f_block(fn, f_local, x0, y0, nx, ny, rank == 0)
    if barrier:
        barrier()


def load_checkpoint_block(fn, decomp=None, rank=0):
    """-> (this rank's populations (9, lnx, lny) as a fresh contiguous array, metadata dict)."""
    import json
    f = np.load(fn, mmap_mode="r")
    meta = {}
    if os.path.exists(fn + ".json"):
        with open(fn + ".json") as fh:
            meta = json.load(fh)
    if decomp is None:
        return np.ascontiguousarray(f), meta
    if f.shape != (9, decomp.nx, decomp.ny):
        raise ValueError("checkpoint %s holds %s, the lattice is (9, %d, %d)" % (fn, f.shape, decomp.nx, decomp.ny))
    sx, sy = decomp.slices(rank)
    return np.ascontiguousarray(f[:, sx, sy]), meta
