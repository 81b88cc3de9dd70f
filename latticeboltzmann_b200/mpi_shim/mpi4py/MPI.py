"""The subset of ``mpi4py.MPI`` that cavity_opt2.py (:214-229, :191-210, :282-288) and PyLB/IO.py
(:50-80) touch, for one OR several processes.

The real mpi4py / mpirun are not installed on the B200 image.  A world of one process needs nothing
else; a world of several processes is the ``torch.distributed`` world the launcher describes
(RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT, i.e. ``torchrun`` or
``python -m torch.distributed.run``): point-to-point messages and the small collectives go over a
``gloo`` group (host buffers, exactly what the reference's ``Sendrecv`` moves).

Semantics kept from MPI where the reference relies on them:
  * ``Create_cart`` numbers ranks row-major (rank = px*ndy + py, cavity_opt2.py:225);
  * ``Shift(direction, disp)`` returns ``(source, dest)`` with ``PROC_NULL`` (< 0: the script tests
    ``right_dst < 0``, :237) beyond a non-periodic edge;
  * ``Sendrecv`` with ``PROC_NULL`` on either side skips that half;
  * ``Exscan`` leaves rank 0's receive buffer untouched (the callers pre-zero it, PyLB/IO.py:64-67);
  * ``File.Set_view`` + ``Write_all`` with a ``Create_vector`` file type store strided rows (:72-78).
"""
import os

import numpy as np

PROC_NULL = -2
MODE_RDONLY, MODE_WRONLY, MODE_CREATE = 2, 4, 1
SUM = "sum"


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


# ---- transport: nothing for one process, a gloo group of torch.distributed otherwise ---------------------
class _World:
    def __init__(self):
        self.size = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0")) if self.size > 1 else 0
        self._group = None

    def group(self):
        if self.size == 1:
            return None
        if self._group is None:
            import torch.distributed as dist
            if not dist.is_initialized():
                os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
                os.environ.setdefault("MASTER_PORT", "29533")
                dist.init_process_group("gloo", rank=self.rank, world_size=self.size)
            if dist.get_backend() == "gloo":
                self._group = dist.group.WORLD
            else:       # an NCCL world (one process per GPU) already exists: host messages get their own gloo group
                self._group = dist.new_group(backend="gloo")
        return self._group

    def all_gather(self, obj):
        """One python object per world rank (rank order)."""
        if self.size == 1:
            return [obj]
        import torch.distributed as dist
        out = [None] * self.size
        dist.all_gather_object(out, obj, group=self.group())
        return out

    def barrier(self):
        if self.size > 1:
            import torch.distributed as dist
            dist.barrier(group=self.group())

    def sendrecv(self, send, dest, recv, source, tag):
        import torch
        import torch.distributed as dist
        g = self.group()
        reqs = []
        rbuf = None
        if source >= 0:
            rbuf = torch.empty(recv.shape, dtype=torch.from_numpy(np.empty(0, recv.dtype)).dtype)
            reqs.append(dist.irecv(rbuf, src=source, group=g, tag=tag))
        if dest >= 0:
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(send)), dst=dest, group=g, tag=tag))
        for r in reqs:
            r.wait()
        if rbuf is not None:
            np.copyto(recv, rbuf.numpy())


_world = _World()


def _as_array(buf):
    a = np.asarray(buf)
    return a


class _Comm:
    """A communicator over `members` (world ranks, in communicator-rank order), optionally Cartesian."""

    def __init__(self, members, dims=None, periods=None):
        self.members = list(members)
        self.dims = tuple(dims) if dims is not None else None
        self.periods = tuple(periods) if periods is not None else None

    # -- basics ---------------------------------------------------------------
    def Get_size(self):
        return len(self.members)

    def Get_rank(self):
        return self.members.index(_world.rank)

    def Barrier(self):
        _world.barrier()        # a world barrier is a barrier of every sub-communicator too

    # -- Cartesian topology ---------------------------------------------------
    def Create_cart(self, dims, periods=None, reorder=False):
        dims = tuple(int(d) for d in dims)
        if int(np.prod(dims)) != len(self.members):
            raise ValueError("Create_cart%s needs %d processes, the communicator has %d"
                             % (dims, int(np.prod(dims)), len(self.members)))
        return _Comm(self.members, dims, tuple(bool(p) for p in (periods or (False,) * len(dims))))

    def _coords_of(self, r):
        c = []
        for d in reversed(self.dims):
            c.append(r % d)
            r //= d
        return list(reversed(c))

    def _rank_of(self, coords):
        r = 0
        for c, d in zip(coords, self.dims):
            r = r * d + c
        return r

    def Get_coords(self, rank):
        return self._coords_of(rank)

    def Shift(self, direction, disp):
        me = self._coords_of(self.Get_rank())

        def at(off):
            c = list(me)
            c[direction] += off
            if self.periods[direction]:
                c[direction] %= self.dims[direction]
            elif not 0 <= c[direction] < self.dims[direction]:
                return PROC_NULL
            return self._rank_of(c)
        return at(-disp), at(disp)

    def Sub(self, remain_dims):
        me = self._coords_of(self.Get_rank())
        keep = [i for i, r in enumerate(remain_dims) if r]
        members = []
        for r in range(len(self.members)):
            c = self._coords_of(r)
            if all(c[i] == me[i] for i in range(len(self.dims)) if i not in keep):
                members.append(self.members[r])
        return _Comm(members, [self.dims[i] for i in keep] or (1,), [self.periods[i] for i in keep] or (False,))

    # -- communication --------------------------------------------------------
    def Sendrecv(self, sendbuf, dest, sendtag=0, recvbuf=None, source=PROC_NULL, recvtag=0, status=None):
        dest = self.members[dest] if dest >= 0 else PROC_NULL
        source = self.members[source] if source >= 0 else PROC_NULL
        if dest < 0 and source < 0:
            return                      # MPI semantics: communication with PROC_NULL does nothing
        recv = _as_array(recvbuf) if recvbuf is not None else None
        if dest == _world.rank and source == _world.rank:        # periodic ring closed on this rank
            np.copyto(recv, _as_array(sendbuf))
            return
        _world.sendrecv(_as_array(sendbuf), dest, recv, source, sendtag)

    def _gathered(self, sendbuf):
        """The send buffers of this communicator's members, in communicator-rank order.  Collectives of
        all sub-communicators run at the same time in the callers (PyLB/IO.py:52-53,65-67), so one world
        all-gather serves them all."""
        parts = _world.all_gather(np.array(_as_array(sendbuf)))
        return [parts[m] for m in self.members]

    def Allreduce(self, sendbuf, recvbuf, op=SUM):
        parts = self._gathered(sendbuf)
        total = parts[0].copy()
        for p in parts[1:]:
            total = total + p
        np.copyto(_as_array(recvbuf), total)

    def Exscan(self, sendbuf, recvbuf, op=SUM):
        parts = self._gathered(sendbuf)
        me = self.Get_rank()
        if me == 0:
            return                      # rank 0's receive buffer is left untouched (zeros in the callers)
        total = parts[0].copy()
        for p in parts[1:me]:
            total = total + p
        np.copyto(_as_array(recvbuf), total)


COMM_WORLD = _Comm(range(_world.size))


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
        _world.barrier()                # collective like MPI_File_close: the file is complete afterwards
