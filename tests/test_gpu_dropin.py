"""The stateless reference-API entry points (lbk_*, HOST buffers) against the oracle."""
import ctypes
import os

import numpy as np
import pytest

from oracle import oracle as orc
from gpu_util import rel_err, require_gpu

pytestmark = pytest.mark.gpu


def _lib():
    lb = require_gpu()
    from latticeboltzmann_b200 import _lib as L
    return lb, L, L.load()


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_lbk_collide_equilibrium_stream_bitexact(dt):
    lb, L, lib = _lib()
    suf = "f64" if dt == np.float64 else "f32"
    rng = np.random.default_rng(4)
    n = 1000
    rho = (1 + 0.05 * rng.standard_normal(n)).astype(dt)
    ux = (0.1 * rng.standard_normal(n)).astype(dt)
    uy = (0.1 * rng.standard_normal(n)).astype(dt)
    f = np.zeros((9, n), dt)
    L.check(getattr(lib, "lbk_equilibriumn_" + suf)(L.np_ptr(rho), L.np_ptr(ux), L.np_ptr(uy), L.np_ptr(f), n))
    ref = np.zeros((9, n), dt)
    orc.equilibrium(rho, ux, uy, ref)
    assert np.array_equal(f, ref)
    e1 = np.zeros(9, dt)
    L.check(getattr(lib, "lbk_equilibrium1_" + suf)(dt(1.1), dt(0.05), dt(-0.02), L.np_ptr(e1)))
    assert np.array_equal(e1, orc.equilibrium1(1.1, 0.05, -0.02, dt))
    for omega in (0.5, 1.7):
        c = f.copy()
        r = f.copy()
        L.check(getattr(lib, "lbk_collide_" + suf)(L.np_ptr(c), n, dt(omega)))
        orc.collide(r, omega)
        assert np.array_equal(c, r)
    g = rng.random((9, 17, 23)).astype(dt)
    r = g.copy()
    L.check(getattr(lib, "lbk_stream_" + suf)(L.np_ptr(g), 17, 23))
    orc.stream(r)
    assert np.array_equal(g, r)


def test_lbk_against_reference_test_golden(golden_dir):
    """tests/02-CollideTest.py:94-111 restated: same shapes, reference tolerance 1e-7."""
    lb, L, lib = _lib()
    g = np.load(os.path.join(golden_dir, "collide_test_ref.npz"))
    rho, ux, uy = (np.ascontiguousarray(g[k].reshape(-1)) for k in ("eq_rho", "eq_ux", "eq_uy"))
    e1 = np.zeros((9, rho.size))
    L.check(lib.lbk_equilibriumn_f64(L.np_ptr(rho), L.np_ptr(ux), L.np_ptr(uy), L.np_ptr(e1), rho.size))
    assert np.abs(e1.reshape(g["eq_out"].shape) - g["eq_out"]).max() < 1e-7
    for omega in (0.5, 1.7):
        c = g["col_in"].copy()
        L.check(lib.lbk_collide_f64(L.np_ptr(c), 16, omega))
        assert np.abs(c - g["col_out_%s" % omega]).max() < 1e-7


def test_lbk_step_host_matches_oracle():
    lb, L, lib = _lib()
    f = orc.perturbed_state(40, 36, seed=8)
    r = f.copy()
    L.check(lib.lbk_step_host_f64(L.np_ptr(f), 40, 36, L.BOUNDARY["cavity"], 1.7, 0.1, 12))
    orc.cavity_run(r, 1.7, 12)
    assert np.array_equal(f, r)


def test_empty_inputs_are_noops():
    lb, L, lib = _lib()
    z = np.zeros((9, 0))
    L.check(lib.lbk_collide_f64(L.np_ptr(z), 0, 1.0))
    L.check(lib.lbk_stream_f64(L.np_ptr(z), 0, 0))


@pytest.mark.parametrize("boundary", ["periodic", "cavity", "cavity_xperiodic"])
def test_pipelined_host_step_bitexact(boundary):
    """lb_step_host: slab-pipelined H2D / compute / D2H == the oracle, for any slab count, in place and
    out of place, and the resident state stays consistent for further resident steps."""
    lb, L, lib = _lib()
    nx, ny = 37, 300
    for nslabs in (1, 2, 5, 37, 64):
        f = orc.perturbed_state(nx, ny, seed=nslabs)
        ref = f.copy()
        blk = lb.Block(nx, ny, boundary=boundary, omega=1.7, u_wall=0.1)
        blk.connect_self()
        out = np.empty_like(f)
        blk.step_host(f, out, nslabs)           # step 1: out of place
        blk.step_host(out, None, nslabs)        # step 2: in place
        blk.step(3)                             # resident steps continue from the device state
        got = blk.download()
        blk.health()
        blk.close()
        if boundary == "periodic":
            orc.periodic_run(ref, 1.7, 2)
        else:
            orc.cavity_run(ref, 1.7, 2, 0.1, walls_lr=(boundary == "cavity"))
        assert np.array_equal(out, ref), nslabs
        if boundary == "periodic":
            orc.periodic_run(ref, 1.7, 3)
        else:
            orc.cavity_run(ref, 1.7, 3, 0.1, walls_lr=(boundary == "cavity"))
        assert np.array_equal(got, ref), nslabs


def test_plain_c_client_runs_on_the_gpu(tmp_path):
    """examples/c_abi_cavity.c (plain C over the ABI): A/B run, then the same run on an in-place lattice -- equal digests."""
    import subprocess
    require_gpu()
    from latticeboltzmann_b200.build import build_native
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so_dir = os.path.dirname(build_native())
    exe = str(tmp_path / "c_abi_cavity")
    subprocess.run(["gcc", "-std=c11", "-O2", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                    os.path.join(root, "examples", "c_abi_cavity.c"), "-o", exe, "-L", so_dir, "-llbm_b200",
                    "-Wl,-rpath," + so_dir], check=True)
    r = subprocess.run([exe, "200", "150", "301"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "MLUPS" in r.stdout and "identical" in r.stdout
