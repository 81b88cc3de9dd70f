"""Single-file .npy output of a decomposed 2-D field, byte-compatible with the reference's
``save_mpiio`` (PyLB/IO.py:27-80): magic ``\\x93NUMPY\\x01\\x00``, little-endian int16 header
length, the dict literal padded with spaces so that the payload starts 16-byte aligned,
C-order ``(nx, ny)`` payload.  Each writer stores its block's rows at their global offsets
(what the reference does with an MPI-IO vector file view, PyLB/IO.py:72-78), so no gather
is needed.  ``np.load`` reads the result -- that is also the restart reader the reference lacks.
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
