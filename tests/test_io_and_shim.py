"""CPU tests of the output format (N2) and the mpi4py stand-in (N3): one process, and a two-process gloo world
that runs the reference's own cavity_opt2.py."""
import os

import numpy as np
import pytest

from latticeboltzmann_b200 import dropin, npyio
from latticeboltzmann_b200.decomposition import Decomposition

dropin.activate()


def reference_header(shape, dtype):
    """PyLB/IO.py:47-62 restated: magic(1, 0) + int16 length + dict padded so that
    (len + len(magic) + 2) % 16 == 15, then a newline."""
    from numpy.lib.format import dtype_to_descr, magic
    m = magic(1, 0)
    d = str({'descr': dtype_to_descr(np.dtype(dtype)), 'fortran_order': False, 'shape': tuple(shape)})
    while (len(d) + len(m) + 2) % 16 != 15:
        d += ' '
    d += '\n'
    return m + np.int16(len(d)).tobytes() + d.encode('latin-1')


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_header_is_byte_identical_to_save_mpiio(dt):
    for shape in [(300, 300), (4096, 4096), (7, 123456)]:
        h = npyio.npy_header(shape, dt)
        assert h == reference_header(shape, dt)
        assert len(h) % 16 == 0


@pytest.mark.parametrize("ndx,ndy", [(1, 1), (3, 2), (1, 4), (5, 1)])
def test_blockwise_write_roundtrip(tmp_path, ndx, ndy):
    g = np.random.default_rng(0).random((37, 22))
    dec = Decomposition(37, 22, ndx, ndy)
    fn = str(tmp_path / "ux_0.npy")
    for r in reversed(range(dec.size)):          # any order: every rank writes its own rows
        npyio.write_block(fn, dec.scatter(g, r), dec.block(r).x0, dec.block(r).y0, 37, 22, r == 0)
    assert np.array_equal(np.load(fn), g)
    assert np.array_equal(npyio.load_field(fn), g)


def test_mpi4py_standin_runs_the_reference_io_sequence(tmp_path):
    """The call sequence of PyLB/IO.py:50-80 against the single-process stand-in."""
    from mpi4py import MPI
    comm = MPI.COMM_WORLD.Create_cart((1, 1), periods=(False, False))         # cavity_opt2.py:225
    assert comm.Get_size() == 1 and comm.Get_rank() == 0
    assert comm.Shift(0, -1) == (MPI.PROC_NULL, MPI.PROC_NULL)                 # :226-229: walls all around
    assert comm.Get_coords(0) == [0, 0]
    with pytest.raises(ValueError):
        MPI.COMM_WORLD.Create_cart((2, 1), periods=(False, False))       # 2 processes needed, the world has 1
    g_kl = np.random.default_rng(1).random((6, 5))
    recv = np.full(5, 7.0)
    comm.Sendrecv(g_kl[0].copy(), MPI.PROC_NULL, recvbuf=recv, source=MPI.PROC_NULL)   # :192-194: no-op
    assert (recv == 7.0).all()
    local_nx, local_ny = g_kl.shape
    nx, ny = np.empty_like(local_nx), np.empty_like(local_ny)
    commx, commy = comm.Sub((True, False)), comm.Sub((False, True))
    commx.Allreduce(np.asarray(local_nx), nx)
    commy.Allreduce(np.asarray(local_ny), ny)
    assert (int(nx), int(ny)) == (6, 5)
    offsetx, offsety = np.zeros_like(local_nx), np.zeros_like(local_ny)
    commx.Exscan(np.asarray(ny * local_nx), offsetx)
    commy.Exscan(np.asarray(local_ny), offsety)
    hdr = reference_header((int(nx), int(ny)), g_kl.dtype)
    fn = str(tmp_path / "f.npy")
    f = MPI.File.Open(comm, fn, MPI.MODE_CREATE | MPI.MODE_WRONLY)
    f.Write(hdr[:8])
    f.Write(np.int16(len(hdr) - 10))
    f.Write(hdr[10:])
    mpitype = MPI._typedict[g_kl.dtype.char]
    filetype = mpitype.Create_vector(g_kl.shape[0], g_kl.shape[1], int(ny))
    filetype.Commit()
    f.Set_view(len(hdr) + int(offsety + offsetx) * mpitype.Get_size(), filetype=filetype)
    f.Write_all(g_kl.copy())
    filetype.Free()
    f.Close()
    assert np.array_equal(np.load(fn), g_kl)


def test_dropin_save_mpiio(tmp_path):
    from mpi4py import MPI
    from PyLB.IO import save_mpiio
    g = np.random.default_rng(2).random((9, 4)).astype(np.float32)
    comm = MPI.COMM_WORLD.Create_cart((1, 1), periods=(False, False))
    save_mpiio(comm, str(tmp_path / "a.npy"), g)
    save_mpiio(None, str(tmp_path / "b.npy"), g[1:-1, 1:])       # non-contiguous slices, as cavity_opt2.py:282 passes
    assert np.array_equal(np.load(tmp_path / "a.npy"), g)
    assert np.array_equal(np.load(tmp_path / "b.npy"), g[1:-1, 1:])
    assert open(tmp_path / "a.npy", "rb").read(8) == npyio.MAGIC


REF_SCRIPT = "/root/reference/simulators/parallel_lid_drive_cavity/cavity_opt2.py"


def run_reference_script(tmp_path, backend, ndx, ndy, nx, ny, nsteps, dump_freq):
    import socket
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    world = ndx * ndy
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(here, "ref_script_worker.py"), REF_SCRIPT, backend, str(nsteps), str(dump_freq),
                                       str(ndx), str(ndy), str(nx), str(ny), "float64"], cwd=str(tmp_path), env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=300)
        assert p.returncode == 0, out[-3000:]


def expected_velocity_dumps(nx, ny, nsteps, dump_freq):
    """Single-rank result with the checker: after step i (i % dump_freq == 0) the script dumps
    u = (f^T . c)/rho (cavity_opt2.py:279-283)."""
    from oracle import oracle as orc
    f = orc.init_equilibrium(nx, ny)
    out = {}
    for i in range(nsteps):
        orc.cavity_run(f, 1.7, 1)
        if i % dump_freq == 0:
            rho = np.sum(f, axis=0)
            out[i] = np.dot(f.T, orc.C_IC).T / rho
    return out


@pytest.mark.skipif(not os.path.exists(REF_SCRIPT), reason="the reference tree is not present on this machine")
@pytest.mark.parametrize("ndx,ndy", [(1, 1), (1, 2)])
def test_reference_cavity_script_runs_unmodified_under_the_shim(tmp_path, ndx, ndy):
    """cavity_opt2.py itself (Create_cart, Shift, four Sendrecv per step, save_mpiio with Sub / Allreduce /
    Exscan / MPI.File) on 1 and on 2 gloo ranks split in y -- the split for which the reference's decomposed
    run equals its single-rank run -- against the single-rank checker."""
    nx, ny, nsteps, dump_freq = 23, 18, 12, 5
    run_reference_script(tmp_path, "cpu-checker", ndx, ndy, nx, ny, nsteps, dump_freq)
    want = expected_velocity_dumps(nx, ny, nsteps, dump_freq)
    for i, u in want.items():
        for c, name in enumerate(("ux", "uy")):
            got = np.load(tmp_path / ("%s_%d.npy" % (name, i)))
            assert got.shape == (nx, ny)
            assert np.max(np.abs(got - u[c])) < 1e-15, (name, i)
    last = max(want)
    assert np.array_equal(np.load(tmp_path / ("ux_%d.npy" % (nsteps - 1))), np.load(tmp_path / ("ux_%d.npy" % last)))   # :285-286 re-dumps the last moments
