"""The subset of ``mpi4py.MPI`` that cavity_opt2.py (:214-229, :191-210) and PyLB/IO.py (:50-80) touch,
for a world of exactly one process: every Cartesian neighbour is PROC_NULL (the reference creates
its topology with periods=(False, False)), so Sendrecv is a no-op, reductions are copies and
exclusive scans are zero."""
import os

import numpy as np

PROC_NULL = -1
MODE_CREATE, MODE_WRONLY, MODE_RDONLY = 1, 4, 2


class _Datatype:
    def __init__(self, np_dtype, count=1, blocklength=1, stride=1):
        self.np_dtype = np.dtype(np_dtype)
        self.count, self.blocklength, self.stride = count, blocklength, stride

    def Get_size(self):
        return self.np_dtype.itemsize

    def Create_vector(self, count, blocklength, stride):
        return _Datatype(self.np_dtype, int(count), int(blocklength), int(stride))

    def Commit(self):
        return self

    def Free(self):
        pass


_typedict = {c: _Datatype(np.dtype(c)) for c in "fdiIlLqQhHbB"}


class _Comm:
    def __init__(self, dims=(1,), periods=(False,)):
        self.dims, self.periods = tuple(dims), tuple(periods)

    def Get_size(self):
        return 1

    def Get_rank(self):
        return 0

    def Barrier(self):
        pass

    def Create_cart(self, dims, periods=None, reorder=False):
        dims = tuple(int(d) for d in dims)
        if int(np.prod(dims)) != 1:
            raise RuntimeError("the single-process mpi4py stand-in cannot create a %s topology; use "
                               "latticeboltzmann_b200.distributed (one process per GPU) for decomposed runs" % (dims,))
        return _Comm(dims, periods or (False,) * len(dims))

    def Get_coords(self, rank):
        return [0] * len(self.dims)

    def Shift(self, direction, disp):
        if self.periods[direction]:
            return 0, 0
        return PROC_NULL, PROC_NULL

    def Sub(self, remain_dims):
        keep = [i for i, r in enumerate(remain_dims) if r]
        return _Comm([self.dims[i] for i in keep] or (1,), [self.periods[i] for i in keep] or (False,))

    def Sendrecv(self, sendbuf, dest, sendtag=0, recvbuf=None, source=PROC_NULL, recvtag=0, status=None):
        if dest == PROC_NULL or source == PROC_NULL:
            return                      # MPI semantics: communication with PROC_NULL does nothing
        np.copyto(np.asarray(recvbuf), np.asarray(sendbuf))      # periodic self-neighbour

    def Allreduce(self, sendbuf, recvbuf, op=None):
        np.copyto(np.asarray(recvbuf), np.asarray(sendbuf))

    def Exscan(self, sendbuf, recvbuf, op=None):
        pass                            # rank 0's receive buffer is left untouched (zeros in the callers)


COMM_WORLD = _Comm()


class File:
    """MPI.File over a plain file descriptor: Write at the individual pointer, Set_view + Write_all
    for the strided block the reference writes (PyLB/IO.py:72-78)."""

    def __init__(self, fd):
        self.fd, self.pos, self.disp, self.ftype = fd, 0, 0, None

    @classmethod
    def Open(cls, comm, filename, amode=MODE_RDONLY, info=None):
        flags = os.O_WRONLY if amode & MODE_WRONLY else os.O_RDONLY
        if amode & MODE_CREATE:
            flags |= os.O_CREAT
        return cls(os.open(filename, flags, 0o644))

    def Write(self, buf):
        data = buf if isinstance(buf, (bytes, bytearray)) else np.asarray(buf).tobytes()
        os.pwrite(self.fd, data, self.disp + self.pos)
        self.pos += len(data)

    def Set_view(self, disp=0, etype=None, filetype=None, datarep="native", info=None):
        self.disp, self.pos, self.ftype = int(disp), 0, filetype

    def Write_all(self, buf):
        a = np.ascontiguousarray(buf)
        t = self.ftype
        if t is None or t.count <= 1 or t.blocklength == t.stride:
            return self.Write(a)
        flat = a.reshape(-1)
        item = a.dtype.itemsize
        for i in range(t.count):
            os.pwrite(self.fd, flat[i * t.blocklength:(i + 1) * t.blocklength].tobytes(), self.disp + i * t.stride * item)

    def Close(self):
        os.close(self.fd)
