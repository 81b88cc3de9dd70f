"""Target for compute-sanitizer (memcheck / racecheck / synccheck): every stepping path once, on small lattices,
each checked against the CPU checker.  Developer tool (GPU box):

    compute-sanitizer --tool memcheck python tools/sanitize_target.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latticeboltzmann_b200 as lb
from latticeboltzmann_b200 import boundary_table as bt
from oracle import oracle as orc
from oracle import simple_flows as sf

only = set(sys.argv[1:])


def case(name, fn):
    if only and name not in only:
        return
    fn()
    print("ok", name, flush=True)


def fused(temporal, resident, ndx=1, ndy=1, inplace=False, nx=70, ny=530, steps=5, dt=np.float64):
    os.environ["LBM_RESIDENT"] = "1" if resident else "0"
    f0 = orc.perturbed_state(nx, ny, dt, seed=3)
    ref = f0.copy()
    orc.cavity_run(ref, 1.7, steps)
    lat = lb.Lattice(nx, ny, "cavity", omega=1.7, dtype=dt, temporal=temporal, ndx=ndx, ndy=ndy, inplace=inplace)
    lat.upload(f0)
    lat.step(steps)
    got = lat.download()
    lat.checksum()
    lat.moments()
    lat.health()
    lat.close()
    assert np.array_equal(got, ref)


def host_step():
    f = orc.perturbed_state(40, 300, seed=4)
    ref = f.copy()
    orc.cavity_run(ref, 1.6, 2)
    lat = lb.Lattice(40, 300, "cavity", omega=1.6, ndx=2, ndy=2)
    lat.upload(f)
    lat.step_host(f, nslabs=3)
    lat.step_host(f, nslabs=3)
    assert np.array_equal(f, ref)
    lat.close()


def table():
    os.environ["LBM_RESIDENT"] = "1"
    n = 40
    f0 = sf.feq(np.ones((n, n)), np.zeros((n, n)), np.zeros((n, n))) * (1 + 0.01 * np.random.default_rng(1).standard_normal((9, n, n)))
    for resident in (1, 0):
        os.environ["LBM_RESIDENT"] = str(resident)
        lat = lb.Lattice(n, n, "sf_table", omega=1.1)
        lat.set_boundary_table(*bt.sliding_lid_mpi_table(n, n, 0.1))
        lat.upload(f0)
        lat.step(4)
        ref = f0.copy()
        for _ in range(4):
            sf.sliding_lid_mpi_step(ref, 1.1, 0.1)
        assert np.array_equal(lat.download(), ref)
        lat.close()


def poiseuille():
    for resident in (1, 0):
        os.environ["LBM_RESIDENT"] = str(resident)
        n = 34
        f0 = sf.feq(np.ones((n, n)), np.zeros((n, n)), np.zeros((n, n)))
        lat = lb.Lattice(n, n, "sf_poiseuille", omega=0.5, u_wall=0.0, rho_in=1.001, rho_out=0.999)
        lat.upload(f0)
        lat.step(5)
        ref = f0.copy()
        for _ in range(5):
            sf.poiseuille_step(ref, 0.5, 1.001, 0.999)
        assert np.array_equal(lat.download(), ref)
        lat.close()


def probe():
    os.environ["LBM_RESIDENT"] = "1"
    f0, uy_k = orc.shear_wave_init(60, 40, a0=0.01)
    lat = lb.Lattice(60, 40, "periodic", omega=1.0)
    lat.upload(f0)
    lat.probe_shear_enable(uy_k, 16)
    lat.step(12)
    a = lat.probe_shear_read(12)
    ref = orc.periodic_run(f0.copy(), 1.0, 12, uy_k)
    assert np.abs(a - ref).max() < 1e-14
    lat.close()


case("single_step", lambda: fused(1, False))
case("single_step_f32", lambda: fused(1, False, dt=np.float32))
case("temporal_tma", lambda: fused(2, False, steps=7))
case("temporal_tma_f32", lambda: fused(2, False, steps=6, dt=np.float32))
case("temporal_blocks_2x2", lambda: fused(2, False, 2, 2, steps=6))
case("single_blocks_3x2", lambda: fused(1, False, 3, 2))
case("resident", lambda: fused(1, True))
case("inplace_aa", lambda: fused(1, False, inplace=True, steps=5))
case("host_step_2x2", host_step)
case("boundary_table", table)
case("poiseuille", poiseuille)
case("probe_resident", probe)
print("all paths ok", flush=True)
