"""One process per GPU: the decomposed lattice over a torch.distributed world.

Replaces the reference's MPI plumbing (cavity_opt2.py:214-261: COMM_WORLD,
Create_cart, Shift, local extents) and its per-step ``communicate()``
(cavity_opt2.py:179-210).  torch.distributed is used for rendezvous only:
every rank exports its block (CUDA IPC handle + geometry), the exports are
all-gathered, and each rank maps its 8 neighbours' allocations.  After that
there is NO per-step collective call: the step kernel's rim CTAs store the
outgoing ghost populations straight into the neighbours' memory over
NVLink / NVSwitch and order themselves with device-side flags.
"""
import os

import numpy as np

from ._lib import LbExport
from .decomposition import Decomposition, temporal_mode
from .lattice import Block


def _dist():
    import torch.distributed as dist
    return dist


def init_process_group(backend=None):
    """Rendezvous from torchrun's env (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT).
    Returns (rank, world, local_rank)."""
    import torch
    dist = _dist()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local_rank


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPUs NVML reports as local to its GPU, so that pinned host buffers
    (first touch) and the launch thread sit on the GPU's NUMA node.  Best effort: returns the CPU
    set applied, or None if NVML / the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, mask in enumerate(words) for b in range(64) if (mask >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


def exchange_blobs(blob, group=None):
    """All-gather one bytes object per rank (rank order)."""
    dist = _dist()
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, bytes(blob), group=group)
    return out


def gather_blocks(local, decomp, dst=0, group=None):
    """Assemble the global (..., nx, ny) array on rank `dst` from every rank's local
    (..., lnx, lny) part (replaces save_mpiio's file view, PyLB/IO.py:66-80)."""
    dist = _dist()
    rank = dist.get_rank(group)
    parts = [None] * dist.get_world_size(group) if rank == dst else None
    dist.gather_object(np.ascontiguousarray(local), parts, dst=dst, group=group)
    if rank != dst:
        return None
    g = np.empty(local.shape[:-2] + (decomp.nx, decomp.ny), local.dtype)
    for r, p in enumerate(parts):
        decomp.gather_into(g, r, p)
    return g


def max_over_ranks(x, group=None):
    import torch
    dist = _dist()
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


class DistributedLattice:
    """This rank's block of an (nx, ny) lattice split ndx x ndy over the world
    (rank = px*ndy + py, as Create_cart numbers them)."""

    def __init__(self, nx, ny, ndx, ndy, boundary="cavity", omega=1.0, u_wall=0.1, dtype=np.float64,
                 arith="exact", device=None, group=None, rows_per_tile=None, temporal=None):
        dist = _dist()
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if ndx * ndy != self.world:
            raise ValueError("ndx*ndy = %d but the world has %d ranks" % (ndx * ndy, self.world))   # cavity_opt2.py:218
        self.decomp = Decomposition(nx, ny, ndx, ndy)
        b = self.decomp.block(self.rank)
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self.blockinfo = b
        self.block = Block(nx, ny, b.x0, b.y0, b.lnx, b.lny, boundary, omega, u_wall, dtype, arith, device)
        if rows_per_tile:
            self.block.set_rows_per_tile(rows_per_tile)
        # Temporal blocking is a collective property (a rank that advances two steps per pass waits for its
        # neighbours' level-(n+1) frame ghosts): decide it for the whole decomposition and set it explicitly.
        self.block.set_temporal(temporal_mode(self.decomp.blocks(), boundary, temporal))
        blobs = exchange_blobs(bytes(self.block.export()), group)
        self.exports = [LbExport.from_buffer_copy(x) for x in blobs]
        for d, nb in enumerate(self.decomp.neighbours(self.rank)):
            self.block.connect(d, self.exports[nb])
        dist.barrier(group)

    def close(self):
        self.block.sync()
        _dist().barrier(self.group)      # nobody unmaps while a neighbour may still push
        self.block.close()

    def _refresh(self):
        dist = _dist()
        self.block.sync()
        dist.barrier(self.group)
        self.block.halo_refresh()
        self.block.sync()
        dist.barrier(self.group)

    def init_equilibrium(self, rho=None, ux=None, uy=None):
        """Local (lnx, lny) arrays or None (rho=1, u=0; cavity_opt2.py:265-269)."""
        self.block.init_equilibrium(rho, ux, uy)
        self._refresh()

    def upload_global(self, f):
        self.block.upload(self.decomp.scatter(f, self.rank))
        self._refresh()

    def step(self, n=1):
        self.block.step(n)

    def step_host(self, f_in, f_out=None, nslabs=64):
        """One step with this rank's block in HOST memory (pinned (9, lnx, lny) arrays; in place by default): the
        host-side schedule that replaces communicate() (cavity_opt2.py:179-210) for a state that lives on the
        host -- rim up, barrier, rim pushed into the neighbours' ghosts over NVLink, barrier, then the block's
        slab pipeline (H2D / compute / D2H overlapped on three streams)."""
        dist = _dist()
        f_out = f_in if f_out is None else f_out
        if self.world == 1:
            return self.block.step_host(f_in, f_out, nslabs)
        self.block.step_host_begin(f_in)
        dist.barrier(self.group)
        self.block.halo_refresh()
        self.block.sync()
        dist.barrier(self.group)
        self.block.step_host(f_in, f_out, nslabs)

    def step_timed(self, n):
        """Barrier + sync, n steps timed with CUDA events on each rank's stream, max over ranks (ms)."""
        dist = _dist()
        self.block.sync()
        dist.barrier(self.group)
        ms = self.block.step_timed(n)
        return max_over_ranks(ms, self.group)

    def sync(self):
        self.block.sync()

    def health(self):
        self.block.health()

    def checksum(self):
        """Digest of the GLOBAL state: the per-block digests added modulo 2^64 (same value on every rank)."""
        parts = [None] * self.world
        _dist().all_gather_object(parts, self.block.checksum(), group=self.group)
        return sum(parts) % (1 << 64)

    def save_checkpoint(self, fn, **meta):
        """Populations of the whole lattice into ONE (9, nx, ny) .npy file (every rank writes its rows) + JSON sidecar."""
        from . import npyio
        dist = _dist()
        npyio.save_checkpoint(fn, self.block.download(), self.decomp, self.rank, lambda: dist.barrier(self.group),
                              dict(meta, steps_done=int(self.block.steps_done)))

    def load_checkpoint(self, fn):
        """Continue from a checkpoint written by any decomposition of the same lattice; returns its metadata."""
        from . import npyio
        f, meta = npyio.load_checkpoint_block(fn, self.decomp, self.rank)
        self.block.upload(f.astype(self.block.dtype, copy=False))
        self._refresh()
        return meta

    def gather_f(self, dst=0):
        return gather_blocks(self.block.download(), self.decomp, dst, self.group)

    def gather_moments(self, dst=0):
        rho, ux, uy = self.block.moments()
        return gather_blocks(np.stack([rho, ux, uy]), self.decomp, dst, self.group)
