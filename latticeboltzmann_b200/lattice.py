"""Device-resident D2Q9 lattice objects (host side of include/lbm_b200.h group 2).

``Block``        one block of the global lattice on one GPU (thin ctypes wrapper).
``Lattice``      a whole lattice decomposed into ndx x ndy blocks driven by THIS
                 process: one block (the common single-GPU case, ring closed on
                 itself), several blocks on one GPU (decomposition tests) or one
                 block per visible GPU with direct peer stores.
For one process per GPU (torchrun / NCCL world) see ``distributed.py``.

The time step the blocks run is the reference's opt2 loop body
(cavity_opt2.py:272-277 / shear_wave_opt2.py:95-97) as ONE fused CUDA kernel.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import LbConfig, LbExport, check, np_ptr
from .decomposition import Decomposition, temporal_mode


class Block:
    def __init__(self, gnx, gny, x0=0, y0=0, lnx=None, lny=None, boundary="periodic", omega=1.0, u_wall=0.1,
                 dtype=np.float64, arith="exact", device=0, rho_in=1.0, rho_out=1.0, inplace=False):
        self.lib = _lib.load()
        _lib.require_device()
        self.dtype = np.dtype(dtype)
        self.lnx = int(gnx if lnx is None else lnx)
        self.lny = int(gny if lny is None else lny)
        cfg = LbConfig(device=device, dtype=_lib.dtype_code(dtype), boundary=_lib.BOUNDARY[boundary],
                       arith=_lib.ARITH[arith], gnx=gnx, gny=gny, x0=x0, y0=y0, lnx=self.lnx, lny=self.lny,
                       omega=float(omega), u_wall=float(u_wall), rho_in=float(rho_in), rho_out=float(rho_out))
        self.cfg = cfg
        h = ctypes.c_void_p()
        check(self.lib.lb_create_ex(ctypes.byref(cfg), 1 if inplace else 0, ctypes.byref(h)))      # 1 = LB_CREATE_INPLACE
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.lb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- wiring -------------------------------------------------------------
    def export(self):
        e = LbExport()
        check(self.lib.lb_get_export(self.h, ctypes.byref(e)))
        return e

    def connect(self, d, export):
        check(self.lib.lb_connect(self.h, d, ctypes.byref(export)))

    def connect_self(self):
        e = self.export()
        for d in range(_lib.NUM_DIRS):
            self.connect(d, e)

    def set_stream(self, cuda_stream):
        check(self.lib.lb_set_stream(self.h, ctypes.c_void_p(cuda_stream)))

    def set_rows_per_tile(self, rows):
        check(self.lib.lb_set_rows_per_tile(self.h, rows))

    def set_halo_timeout_ms(self, ms):
        check(self.lib.lb_set_halo_timeout_ms(self.h, int(ms)))

    def set_temporal(self, steps_per_pass, rows_per_tile=0):
        """1: single-step kernel; 2: temporal blocking, two time steps per pass over HBM (bit-identical)."""
        check(self.lib.lb_set_temporal(self.h, int(steps_per_pass), int(rows_per_tile)))

    def double_step_phase(self, phase):
        check(self.lib.lb_double_step_phase(self.h, int(phase)))

    @property
    def temporal_active(self):
        return bool(self.lib.lb_temporal_active(self.h))

    @property
    def temporal_rows(self):
        """Rows per fused tile of the two-steps-per-pass kernel (tiles start at row 2 + m * temporal_rows)."""
        return int(self.lib.lb_temporal_rows(self.h))

    def set_boundary_table(self, cells, src, add):
        """Per-cell boundary table of an "sf_table" lattice (boundary_table.py: cells (n,), src (n, 9), add (n, 9))."""
        cells = np.ascontiguousarray(cells, dtype=np.int64)
        src = np.ascontiguousarray(src, dtype=np.int64)
        add = np.ascontiguousarray(add, dtype=np.float64)
        n = cells.shape[0]
        if src.shape != (n, 9) or add.shape != (n, 9):
            raise ValueError("expected src and add of shape (%d, 9)" % n)
        check(self.lib.lb_set_boundary_table(self.h, n, np_ptr(cells), np_ptr(src), np_ptr(add)))

    def set_resident(self, on):
        """Allow / forbid the resident multi-step kernel (L2-resident single blocks, one launch for many steps)."""
        check(self.lib.lb_set_resident(self.h, int(bool(on))))

    def set_use_graph(self, on):
        check(self.lib.lb_set_use_graph(self.h, int(bool(on))))

    # -- state ----------------------------------------------------------------
    def _host(self, a, shape):
        a = np.ascontiguousarray(a, dtype=self.dtype)
        if a.shape != shape:
            raise ValueError("expected shape %s, got %s" % (shape, a.shape))
        return a

    def upload(self, f):
        f = self._host(f, (9, self.lnx, self.lny))
        check(self.lib.lb_upload_f(self.h, np_ptr(f)))

    def download(self, out=None):
        if out is None:
            out = np.empty((9, self.lnx, self.lny), self.dtype)
        assert out.flags.c_contiguous and out.dtype == self.dtype and out.shape == (9, self.lnx, self.lny)
        check(self.lib.lb_download_f(self.h, np_ptr(out)))
        return out

    def download_rows(self, k_lo, k_hi, out=None):
        """Rows [k_lo, k_hi) of the current state as a (9, k_hi - k_lo, lny) array."""
        if out is None:
            out = np.empty((9, k_hi - k_lo, self.lny), self.dtype)
        assert out.flags.c_contiguous and out.dtype == self.dtype and out.shape == (9, k_hi - k_lo, self.lny)
        check(self.lib.lb_download_rows(self.h, int(k_lo), int(k_hi), np_ptr(out)))
        return out

    def upload_rows(self, k_lo, k_hi, rows):
        rows = self._host(rows, (9, k_hi - k_lo, self.lny))
        check(self.lib.lb_upload_rows(self.h, int(k_lo), int(k_hi), np_ptr(rows)))

    def checksum(self):
        """64-bit digest of the current state (sum over cells of mix(bits, GLOBAL index) mod 2^64): the digests
        of the blocks of any decomposition add up to the digest of the undecomposed lattice."""
        d = ctypes.c_uint64()
        check(self.lib.lb_checksum(self.h, ctypes.byref(d)))
        return int(d.value)

    def init_equilibrium(self, rho=None, ux=None, uy=None):
        arrs = [None if a is None else self._host(np.broadcast_to(a, (self.lnx, self.lny)), (self.lnx, self.lny))
                for a in (rho, ux, uy)]
        check(self.lib.lb_init_equilibrium(self.h, *[None if a is None else np_ptr(a) for a in arrs]))

    def halo_refresh(self):
        check(self.lib.lb_halo_refresh(self.h))

    # -- stepping -------------------------------------------------------------
    def step(self, n=1):
        check(self.lib.lb_step(self.h, n))

    def stream_only(self, n=1):
        check(self.lib.lb_stream_only(self.h, n))

    def step_timed(self, n):
        ms = ctypes.c_float()
        check(self.lib.lb_step_timed(self.h, n, ctypes.byref(ms)))
        return ms.value

    def step_host(self, f_in, f_out=None, nslabs=16):
        """One step with HOST input/output (the reference's stateless convention), H2D / compute / D2H
        overlapped slab by slab.  f_out defaults to f_in (in place).  Arrays: C-contiguous (9, lnx, lny)."""
        f_out = f_in if f_out is None else f_out
        for a in (f_in, f_out):
            if not (isinstance(a, np.ndarray) and a.flags.c_contiguous and a.dtype == self.dtype
                    and a.shape == (9, self.lnx, self.lny)):
                raise TypeError("step_host needs C-contiguous %s arrays of shape (9, %d, %d)" % (self.dtype, self.lnx, self.lny))
        check(self.lib.lb_step_host(self.h, np_ptr(f_in), np_ptr(f_out), int(nslabs)))

    def step_host_begin(self, f_in):
        """Phase 1 of a host step on a decomposed lattice: this block's rim of `f_in` goes to the device (then:
        barrier, halo_refresh + sync, barrier, step_host)."""
        if not (isinstance(f_in, np.ndarray) and f_in.flags.c_contiguous and f_in.dtype == self.dtype
                and f_in.shape == (9, self.lnx, self.lny)):
            raise TypeError("step_host_begin needs a C-contiguous %s array of shape (9, %d, %d)" % (self.dtype, self.lnx, self.lny))
        check(self.lib.lb_step_host_begin(self.h, np_ptr(f_in)))

    def sync(self):
        check(self.lib.lb_sync(self.h))

    def health(self):
        check(self.lib.lb_health(self.h))

    @property
    def steps_done(self):
        return self.lib.lb_steps_done(self.h)

    @property
    def kernel_launches(self):
        c = ctypes.c_int64()
        check(self.lib.lb_kernel_launches(self.h, ctypes.byref(c)))
        return c.value

    # -- observables ------------------------------------------------------------
    def moments(self):
        rho = np.empty((self.lnx, self.lny), self.dtype)
        ux = np.empty_like(rho)
        uy = np.empty_like(rho)
        check(self.lib.lb_moments(self.h, np_ptr(rho), np_ptr(ux), np_ptr(uy)))
        return rho, ux, uy

    def probe_shear_enable(self, l_probe_global, uy_k_local, capacity):
        uy_k_local = self._host(uy_k_local, (self.lnx,))
        check(self.lib.lb_probe_shear_enable(self.h, l_probe_global, np_ptr(uy_k_local), capacity))

    def probe_shear_read(self, n):
        out = np.empty(n, self.dtype)
        check(self.lib.lb_probe_shear_read(self.h, np_ptr(out), n))
        return out


class Lattice:
    """A global nx x ny lattice split into ndx x ndy blocks, all driven by this process.

    ``devices``: list of CUDA ordinals, one per block (rank order, rank = px*ndy+py),
    or a single ordinal for all blocks.  Blocks that share a device also share one
    stream, so their launches are ordered on the device; blocks on different
    devices run concurrently and order themselves with the halo flags.
    """

    def __init__(self, nx, ny, boundary="periodic", omega=1.0, u_wall=0.1, dtype=np.float64, arith="exact",
                 ndx=1, ndy=1, devices=0, rho_in=1.0, rho_out=1.0, rows_per_tile=None, temporal=None, inplace=False):
        """inplace=True: ONE copy of the populations advanced with the AA pattern (half the memory, same traffic,
        bit-identical; a single block with periodic / cavity boundaries)."""
        self.nx, self.ny = int(nx), int(ny)
        self.dtype = np.dtype(dtype)
        self.boundary = boundary
        self.omega = omega
        self.decomp = Decomposition(nx, ny, ndx, ndy)
        n = self.decomp.size
        if isinstance(devices, int):
            devices = [devices] * n
        if len(devices) != n:
            raise ValueError("need one device per block")
        self.devices = list(devices)
        self.blocks = []
        for b in self.decomp.blocks():
            self.blocks.append(Block(nx, ny, b.x0, b.y0, b.lnx, b.lny, boundary, omega, u_wall, dtype, arith,
                                     self.devices[b.rank], rho_in, rho_out, inplace=inplace))
        if rows_per_tile:
            for blk in self.blocks:
                blk.set_rows_per_tile(rows_per_tile)
        # Stepping mode (1: single-step kernel, 2: two time steps per pass over HBM), decided for the whole
        # decomposition: blocks that exchange halos cannot mix modes (decomposition.temporal_mode).
        mode = temporal_mode(self.decomp.blocks(), boundary, temporal)
        for blk in self.blocks:
            blk.set_temporal(mode)
        exports = [blk.export() for blk in self.blocks]
        for r, blk in enumerate(self.blocks):
            for d, nb in enumerate(self.decomp.neighbours(r)):
                blk.connect(d, exports[nb])
        # one stream per device
        lead = {}
        for r, blk in enumerate(self.blocks):
            dev = self.devices[r]
            if dev in lead:
                blk.set_stream(lead[dev])
            else:
                lead[dev] = self._own_stream(blk)
        self._probe = None

    @staticmethod
    def _own_stream(blk):
        return blk.lib.lb_get_stream(blk.h)

    def close(self):
        for b in self.blocks:
            if b.h:
                b.sync()
        for b in self.blocks:
            if b.h:
                b.set_stream(0)      # stop borrowing the lead block's stream before it is destroyed
        for b in self.blocks:
            b.close()
        self.blocks = []

    # -- state ----------------------------------------------------------------
    def _refresh(self):
        for b in self.blocks:
            b.sync()
        for b in self.blocks:
            b.halo_refresh()
        for b in self.blocks:
            b.sync()

    def upload(self, f):
        f = np.asarray(f)
        if f.shape != (9, self.nx, self.ny):
            raise ValueError("expected (9, %d, %d)" % (self.nx, self.ny))
        for r, b in enumerate(self.blocks):
            b.upload(self.decomp.scatter(f, r))
        self._refresh()

    def download(self):
        g = np.empty((9, self.nx, self.ny), self.dtype)
        for r, b in enumerate(self.blocks):
            self.decomp.gather_into(g, r, b.download())
        return g

    def save_checkpoint(self, fn, **meta):
        from . import npyio
        for r, b in enumerate(self.blocks):
            npyio.save_checkpoint(fn, b.download(), self.decomp, r, None, dict(meta, steps_done=int(b.steps_done)))

    def load_checkpoint(self, fn):
        from . import npyio
        meta = {}
        for r, b in enumerate(self.blocks):
            f, meta = npyio.load_checkpoint_block(fn, self.decomp, r)
            b.upload(f.astype(self.dtype, copy=False))
        self._refresh()
        return meta

    def init_equilibrium(self, rho=None, ux=None, uy=None):
        """f = feq(rho, ux, uy) (c/d2q9.h:59-81); defaults rho=1, u=0 (cavity_opt2.py:265-269)."""
        for r, b in enumerate(self.blocks):
            loc = [None if a is None else self.decomp.scatter(np.broadcast_to(np.asarray(a, self.dtype), (self.nx, self.ny)), r)
                   for a in (rho, ux, uy)]
            b.init_equilibrium(*loc)
        self._refresh()

    # -- stepping -------------------------------------------------------------
    def step(self, n=1):
        if len(self.blocks) == 1:
            self.blocks[0].step(n)
        else:
            # Blocks that share a device share a stream, so launches are interleaved such that every flag a
            # kernel waits for is posted by a launch already queued: double steps phase by phase (phase 3 of a
            # block waits for phase 1 of its neighbours), single steps block by block.
            pairs = n // 2 if all(b.temporal_active for b in self.blocks) else 0
            for _ in range(pairs):
                for phase in (1, 2, 3):
                    for b in self.blocks:
                        b.double_step_phase(phase)
            for _ in range(n - 2 * pairs):
                for b in self.blocks:
                    b.step(1)

    def stream_only(self, n=1):
        for _ in range(n):
            for b in self.blocks:
                b.stream_only(1)

    def step_host(self, f, nslabs=8):
        """One step with the state in HOST memory (global (9, nx, ny) array, updated in place): per block the rim
        goes up first, every block pushes it into its neighbours' ghosts, then each block runs its slab pipeline
        (H2D / compute / D2H overlapped)."""
        if len(self.blocks) == 1:
            self.blocks[0].step_host(f, f, nslabs)
            return f
        parts = [np.ascontiguousarray(self.decomp.scatter(f, r)) for r in range(len(self.blocks))]
        for b, p in zip(self.blocks, parts):
            b.step_host_begin(p)
        self._refresh()
        for b, p in zip(self.blocks, parts):
            b.step_host(p, p, nslabs)
        for r, p in enumerate(parts):
            self.decomp.gather_into(f, r, p)
        return f

    def step_timed(self, n):
        """Milliseconds for n steps (CUDA events on the launching stream; single block only)."""
        if len(self.blocks) != 1:
            raise ValueError("step_timed is defined for a single block; use distributed timing otherwise")
        return self.blocks[0].step_timed(n)

    def sync(self):
        for b in self.blocks:
            b.sync()

    def health(self):
        for b in self.blocks:
            b.health()

    def checksum(self):
        return sum(b.checksum() for b in self.blocks) % (1 << 64)

    @property
    def kernel_launches(self):
        return sum(b.kernel_launches for b in self.blocks)

    # -- observables ------------------------------------------------------------
    def moments(self):
        rho = np.empty((self.nx, self.ny), self.dtype)
        ux = np.empty_like(rho)
        uy = np.empty_like(rho)
        for r, b in enumerate(self.blocks):
            lr, lx, ly = b.moments()
            self.decomp.gather_into(rho, r, lr)
            self.decomp.gather_into(ux, r, lx)
            self.decomp.gather_into(uy, r, ly)
        return rho, ux, uy

    def set_boundary_table(self, cells, src, add):
        """Walls of the simple_flows family as a per-cell gather table (boundary "sf_table", one block)."""
        if len(self.blocks) != 1:
            raise ValueError("boundary tables run on a single block")
        self.blocks[0].set_boundary_table(cells, src, add)

    def probe_shear_enable(self, uy_k, capacity, l_probe=None):
        """Record shear_wave_opt2.py:99's amplitude after every step, on the device."""
        l_probe = self.ny // 2 if l_probe is None else l_probe
        uy_k = np.asarray(uy_k, self.dtype)
        self._probe = []
        for b in self.blocks:           # the probe runs after every single step: the whole decomposition switches mode together
            b.set_temporal(1)
        for r, b in enumerate(self.blocks):
            blk = self.decomp.block(r)
            if blk.y0 <= l_probe < blk.y0 + blk.lny:
                b.probe_shear_enable(l_probe, uy_k[blk.x0:blk.x0 + blk.lnx], capacity)
                self._probe.append(b)

    def probe_shear_read(self, n):
        return sum(b.probe_shear_read(n) for b in self._probe)
