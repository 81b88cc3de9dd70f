"""GPU parity tests proper: the CUDA path (through the C ABI, via ctypes) against the
CPU oracle on identical seeded inputs and against the committed golden fixtures."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from gpu_util import rel_err, require_gpu

pytestmark = pytest.mark.gpu

SHAPES = [(24, 20), (7, 5), (16, 33), (3, 3), (2, 2), (1, 4), (5, 1), (40, 2), (13, 300), (70, 530)]


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("dt", ["float64", "float32"])
def test_stream_and_bounce_back_golden_bitexact(golden_dir, dt):
    """cavity_opt2.py:109-177 through the fused kernel with the collision switched off."""
    lb = require_gpu()
    g = load(golden_dir, "cavity_opt2_bb_%s.npz" % dt)
    for tag in "abcde":
        f = g["in_" + tag]
        lat = lb.Lattice(f.shape[1], f.shape[2], "cavity", omega=1.0, u_wall=float(g["u0_" + tag]), dtype=dt)
        lat.upload(f)
        lat.stream_only(1)
        assert np.array_equal(lat.download(), g["out_" + tag]), tag
        lat.stream_only(2)
        assert np.array_equal(lat.download(), g["out3_" + tag]), tag
        lat.health()
        lat.close()


def test_periodic_stream_golden_bitexact(golden_dir):
    lb = require_gpu()
    g = load(golden_dir, "streaming_roll.npz")            # PyLB/Streaming.py:33-46
    for tag in "abc":
        f = g["in_" + tag]
        lat = lb.Lattice(f.shape[1], f.shape[2], "periodic")
        lat.upload(f)
        lat.stream_only(1)
        assert np.array_equal(lat.download(), g["out_" + tag]), tag
        lat.close()


@pytest.mark.usefixtures("stepping_path")
@pytest.mark.parametrize("dt", ["float64", "float32"])
@pytest.mark.parametrize("boundary", ["periodic", "cavity", "cavity_xperiodic"])
def test_fused_step_bitexact_vs_oracle(dt, boundary):
    """EXACT arithmetic: bit-identical to the no-FMA CPU oracle, ragged and tiny shapes included."""
    lb = require_gpu()
    for nx, ny in SHAPES:
        if boundary != "periodic" and (ny < 2 or (boundary == "cavity" and nx < 2)):
            with pytest.raises(lb.LbmError):     # degenerate walled boxes are rejected, not mis-computed
                lb.Lattice(nx, ny, boundary)
            continue
        f0 = orc.perturbed_state(nx, ny, np.dtype(dt), seed=nx * 1000 + ny)
        # temporal=2: blocks of at least 16 x 16 advance two steps per pass (7 steps = 3 double + 1 single),
        # smaller ones fall back to the single-step kernel
        lat = lb.Lattice(nx, ny, boundary, omega=1.7, u_wall=0.1, dtype=dt, temporal=2)
        lat.upload(f0)
        lat.step(7)
        got = lat.download()
        lat.health()
        lat.close()
        ref = f0.copy()
        if boundary == "periodic":
            orc.periodic_run(ref, 1.7, 7)
        else:
            orc.cavity_run(ref, 1.7, 7, 0.1, walls_lr=(boundary == "cavity"))
        assert np.array_equal(got, ref), (nx, ny, float(np.abs(got - ref).max()))


@pytest.mark.usefixtures("stepping_path")
def test_cavity_1000_steps_exact_and_fast():
    """BASELINE gate: f, rho within 1e-12 relative, u within 1e-12 of max|u| after 1000 fp64 steps.
    EXACT mode is bit-identical; FAST mode (FMA contraction, reciprocal) is the documented deviation."""
    lb = require_gpu()
    nx, ny, omega = 96, 80, 1.7
    ref = orc.init_equilibrium(nx, ny)
    orc.cavity_run(ref, omega, 1000)
    rr, rux, ruy = orc.moments(ref)
    for arith, temporal in (("exact", 2), ("exact", 1), ("fast", 2)):
        lat = lb.Lattice(nx, ny, "cavity", omega=omega, u_wall=0.1, arith=arith, temporal=temporal)
        lat.init_equilibrium()
        lat.step(1000)
        f = lat.download()
        rho, ux, uy = lat.moments()
        lat.health()
        lat.close()
        if arith == "exact":
            assert np.array_equal(f, ref)
        assert rel_err(f, ref) < 1e-12, arith
        assert rel_err(rho, rr) < 1e-12
        umax = max(np.abs(rux).max(), np.abs(ruy).max())
        assert np.abs(ux - rux).max() / umax < 1e-12
        assert np.abs(uy - ruy).max() / umax < 1e-12


def test_cavity_vs_reference_opt1_golden(golden_dir):
    g = load(golden_dir, "cavity_opt1_run.npz")           # cavity_opt1.py numpy path, 50 steps
    lb = require_gpu()
    f0 = g["f0"]
    lat = lb.Lattice(f0.shape[1], f0.shape[2], "cavity", omega=float(g["omega"]), u_wall=float(g["u0"]))
    lat.upload(f0)
    done = 0
    for n in (1, 10, 50):
        lat.step(n - done)
        done = n
        assert rel_err(lat.download(), g["f_%d" % n]) < 1e-12, n
    lat.close()


@pytest.mark.parametrize("temporal", [1, 2])
@pytest.mark.parametrize("boundary", ["periodic", "cavity", "cavity_xperiodic"])
@pytest.mark.parametrize("ndx,ndy", [(2, 1), (1, 2), (2, 2), (3, 2)])
def test_decomposition_bit_exact_larger_blocks(boundary, ndx, ndy, temporal):
    """Blocks large enough (>= 16 x 16) for the temporal-blocking double step: the level-(n+1) frame ghosts
    are exchanged between blocks as well.  Odd step counts mix double and single steps."""
    lb = require_gpu()
    nx, ny = 101, 83
    f0 = orc.perturbed_state(nx, ny, seed=12)
    ref = f0.copy()
    if boundary == "periodic":
        orc.periodic_run(ref, 1.7, 31)
    else:
        orc.cavity_run(ref, 1.7, 31, 0.1, walls_lr=(boundary == "cavity"))
    lat = lb.Lattice(nx, ny, boundary, omega=1.7, ndx=ndx, ndy=ndy, temporal=temporal)
    assert all(b.temporal_active == (temporal == 2) for b in lat.blocks)
    lat.upload(f0)
    lat.step(14)
    lat.step(17)
    got = lat.download()
    lat.health()
    lat.close()
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("boundary", ["periodic", "cavity", "cavity_xperiodic"])
@pytest.mark.parametrize("ndx,ndy", [(2, 1), (1, 2), (2, 2), (3, 2), (4, 1), (8, 1), (2, 4)])
def test_decomposition_bit_exact(boundary, ndx, ndy):
    """Halo indexing: any ndx x ndy block split gives the bit-identical gathered field
    (BASELINE gate "1 GPU == 8 GPUs"), here with all blocks on one device."""
    lb = require_gpu()
    nx, ny = 53, 47
    f0 = orc.perturbed_state(nx, ny, seed=11)
    one = lb.Lattice(nx, ny, boundary, omega=1.7)
    one.upload(f0)
    one.step(25)
    ref = one.download()
    one.close()
    many = lb.Lattice(nx, ny, boundary, omega=1.7, ndx=ndx, ndy=ndy)
    many.upload(f0)
    many.step(25)
    got = many.download()
    many.health()
    many.close()
    assert np.array_equal(got, ref)
    chk = f0.copy()
    if boundary == "periodic":
        orc.periodic_run(chk, 1.7, 25)
    else:
        orc.cavity_run(chk, 1.7, 25, 0.1, walls_lr=(boundary == "cavity"))
    assert np.array_equal(ref, chk)


@pytest.mark.usefixtures("stepping_path")
def test_shear_wave_viscosity_300x200():
    """BASELINE config 1: 300x200 periodic, omega=1, 1000 steps; on-device amplitude probe."""
    lb = require_gpu()
    nx, ny, a0 = 300, 200, 0.01
    f0, uy_k = orc.shear_wave_init(nx, ny, a0=a0)
    lat = lb.Lattice(nx, ny, "periodic", omega=1.0)
    lat.init_equilibrium(uy=np.resize(uy_k, (ny, nx)).T)
    assert np.array_equal(lat.download(), f0)
    lat.probe_shear_enable(uy_k, 1000)
    lat.step(1000)
    ampl = lat.probe_shear_read(1000)
    f = lat.download()
    lat.close()
    ref = f0.copy()
    ref_ampl = orc.periodic_run(ref, 1.0, 1000, uy_k)
    assert np.array_equal(f, ref)
    assert np.max(np.abs(ampl - ref_ampl)) < 1e-14
    a_init = (uy_k * uy_k).sum() * 2 / nx
    assert abs(ampl[-1] / a_init - 0.929500270576) < 1e-9
    kk = (2 * np.pi / nx) ** 2
    nu = -np.polyfit(np.arange(1, 1001), np.log(ampl / a_init), 1)[0] / kk
    assert abs(nu - 1 / 6) / (1 / 6) < 1e-6                   # analytic (1/omega - 1/2)/3


def test_full_size_properties_4096():
    """At BASELINE's 4096^2 the oracle is too slow for a full comparison; check rows against the
    oracle on a strip and size-independent properties (mass conservation in the periodic box)."""
    lb = require_gpu()
    n = 4096
    lat = lb.Lattice(n, n, "periodic", omega=1.0)
    rng = np.random.default_rng(0)
    uy = 0.01 * np.sin(2 * np.pi * np.arange(n) / n)[:, None] * np.ones((1, n))
    lat.init_equilibrium(uy=uy)
    rho0, _, _ = lat.moments()
    lat.step(20)
    rho1, ux1, uy1 = lat.moments()
    lat.health()
    lat.close()
    assert abs(rho1.sum() - rho0.sum()) / rho0.sum() < 1e-13
    # translation invariance along y of the shear wave: every column identical
    assert np.array_equal(uy1[:, 0], uy1[:, n // 2])
    assert np.array_equal(uy1[:, 0], uy1[:, n - 1])
    del rng


def test_exact_constant_division():
    """The EXACT kernel replaces x/9 and x/6 by a 3-operation sequence proven correctly rounded
    (d2q9_math.cuh: rn_div_const); check it against the IEEE division on 2^26 random bit patterns per
    type plus edge cases (zeros, subnormals, binade edges, huge, inf, nan, values around 1)."""
    import ctypes
    lb = require_gpu()
    from latticeboltzmann_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(7)
    n = 1 << 26
    for dt, ut, fn in ((np.float64, np.uint64, lib.lbk_selftest_div_const_f64), (np.float32, np.uint32, lib.lbk_selftest_div_const_f32)):
        info = np.finfo(dt)
        edge = np.array([0.0, -0.0, info.tiny, -info.tiny, info.tiny / 8, info.max, -info.max, np.inf, -np.inf, np.nan,
                         1.0, 9.0, 6.0, 1 - info.epsneg, 1 + info.eps, 2.0, 4 - 2 * info.eps, 0.1, 1 / 3, 4 / 9, 1e-30, 1e30], dtype=dt)
        near = (1 + 0.1 * rng.standard_normal(n // 4)).astype(dt)            # densities
        small = (1e-3 * rng.standard_normal(n // 4)).astype(dt)              # u.u magnitudes
        bits = rng.integers(0, np.iinfo(ut).max, size=n // 2, dtype=ut, endpoint=True)
        allbits = np.ascontiguousarray(np.concatenate([edge.view(ut), near.view(ut), small.view(ut), bits]))
        bad = ctypes.c_int64(-1)
        L.check(fn(L.np_ptr(allbits), allbits.size, ctypes.byref(bad)))
        assert bad.value == 0, (dt, bad.value)


def test_halo_timeout_is_reported_not_hung():
    """A neighbour that never posts its halo flag must surface as LB_ERR_HALO_TIMEOUT, not a hung GPU."""
    lb = require_gpu()
    lat = lb.Lattice(64, 32, "cavity", omega=1.0, ndx=2, ndy=1)
    lat.init_equilibrium()
    a, b = lat.blocks
    a.set_halo_timeout_ms(200)
    a.step(1)                 # step 0 needs flags >= 0: fine
    a.step(1)                 # step 1 needs b's flag >= 1, but b never stepped
    with pytest.raises(lb.LbmError, match="halo flag wait timed out"):
        a.health()
    lat.close()


def test_graph_replay_equals_single_launches(monkeypatch):
    """lb_step replays a CUDA graph of 64 fused steps for long runs; same bits as step-by-step launches."""
    lb = require_gpu()
    monkeypatch.setenv("LBM_RESIDENT", "0")      # the per-step launch path is the subject here
    f0 = orc.perturbed_state(70, 530, seed=5)
    outs = []
    for use_graph, temporal in ((True, 1), (False, 1), (True, 2), (False, 2)):
        lat = lb.Lattice(70, 530, "cavity", omega=1.7, temporal=temporal)
        lat.blocks[0].set_use_graph(use_graph)
        lat.upload(f0)
        lat.step(64 * 3 + 17)
        outs.append(lat.download())
        assert lat.kernel_launches >= 64 * 3 + 17
        lat.health()
        lat.close()
    assert all(np.array_equal(outs[0], o) for o in outs[1:])
    ref = f0.copy()
    orc.cavity_run(ref, 1.7, 64 * 3 + 17)
    assert np.array_equal(outs[0], ref)


def test_decomposition_bit_exact_fp32_and_reupload():
    """fp32 blocks (two-rows-per-iteration interior) and a mid-run download / re-upload cycle."""
    lb = require_gpu()
    nx, ny = 67, 530
    f0 = orc.perturbed_state(nx, ny, np.float32, seed=9)
    ref = f0.copy()
    orc.cavity_run(ref, 1.7, 20)
    for ndx, ndy in ((1, 1), (3, 2)):
        lat = lb.Lattice(nx, ny, "cavity", omega=1.7, dtype=np.float32, ndx=ndx, ndy=ndy, temporal=2)
        lat.upload(f0)
        lat.step(9)
        mid = lat.download()
        lat.upload(mid)                      # state round-trips through the host unchanged
        lat.step(11)
        assert np.array_equal(lat.download(), ref), (ndx, ndy)
        lat.health()
        lat.close()


@pytest.mark.parametrize("omega,nsteps", [(0.5784359093012493, 100), (1.7, 40)])
def test_cavity_4096_bitexact_vs_oracle(omega, nsteps):
    """BASELINE configs[2] at full size: 4096 x 4096 lid-driven cavity (omega = 0.578 for Re = 1000,
    slidingLid.py:28, and the script's hard-coded 1.7, cavity_opt2.py:66), automatic stepping mode (temporal
    blocking is active at this size: 17 x 128 fused tiles), non-uniform exactly representable initial fields
    so that every cell is distinct -- bitwise against the oracle's pull form on all host cores."""
    lb = require_gpu()
    from latticeboltzmann_b200 import selfcheck
    n = 4096
    rho, ux, uy = selfcheck.fields(n, n, "float64")
    lat = lb.Lattice(n, n, "cavity", omega=omega, u_wall=0.1)
    assert lat.blocks[0].temporal_active
    lat.init_equilibrium(rho, ux, uy)
    lat.step(nsteps)
    got = lat.download()
    digest = lat.checksum()
    lat.health()
    lat.close()
    ref = orc.cavity_run_threaded(orc.init_equilibrium(n, n, np.float64, rho, ux, uy), omega, nsteps)
    assert np.array_equal(got, ref)
    # the device-side digest is a function of the field only: the single-step kernel on 2 x 2 blocks agrees
    lat = lb.Lattice(n, n, "cavity", omega=omega, u_wall=0.1, ndx=2, ndy=2, temporal=1)
    lat.init_equilibrium(rho, ux, uy)
    lat.step(nsteps)
    assert lat.checksum() == digest
    lat.health()
    lat.close()


def test_selfcheck_cases_match_committed_oracle_hashes(golden_dir):
    """The cases bench.py runs before timing (forced temporal blocking, ragged sizes), here on one GPU as 1 x 1,
    2 x 2 and 3 x 2 blocks, against tests/golden/bench_parity.json."""
    import json
    lb = require_gpu()
    from latticeboltzmann_b200 import selfcheck
    want = json.load(open(os.path.join(golden_dir, "bench_parity.json")))["sha256"]
    for name, (boundary, nx, ny, dtype, omega, u0, steps) in selfcheck.CASES.items():
        digests = set()
        for ndx, ndy in ((1, 1), (2, 2), (3, 2)):
            lat = lb.Lattice(nx, ny, boundary, omega=omega, u_wall=u0, dtype=np.dtype(dtype), ndx=ndx, ndy=ndy, temporal=2)
            lat.init_equilibrium(*selfcheck.fields(nx, ny, dtype))
            lat.step(steps)
            assert selfcheck.digest(lat.download()) == want[name], (name, ndx, ndy)
            digests.add(lat.checksum())
            lat.health()
            lat.close()
        assert len(digests) == 1          # device digest independent of the decomposition


def test_rows_io_and_checksum_sensitivity():
    lb = require_gpu()
    nx, ny = 70, 300
    f0 = orc.perturbed_state(nx, ny, seed=12)
    lat = lb.Lattice(nx, ny, "cavity", omega=1.1)
    lat.upload(f0)
    blk = lat.blocks[0]
    assert np.array_equal(blk.download_rows(13, 41), f0[:, 13:41])
    d0 = lat.checksum()
    rows = f0[:, 20:22].copy()
    rows[4, 1, 77] = np.nextafter(rows[4, 1, 77], 2.0)        # one ulp in one population of one cell
    blk.upload_rows(20, 22, rows)
    assert lat.checksum() != d0
    blk.upload_rows(20, 22, f0[:, 20:22].copy())
    assert lat.checksum() == d0
    with pytest.raises(lb.LbmError):
        blk.download_rows(60, 71)
    lat.close()


def test_impossible_allocation_fails_cleanly():
    """lb_create of a lattice no GPU can hold returns LB_ERR_CUDA with a message (and the library stays usable)."""
    lb = require_gpu()
    with pytest.raises(lb.LbmError, match="cudaMalloc"):
        lb.Lattice(1 << 20, 1 << 20, "cavity")
    lat = lb.Lattice(32, 32, "cavity")
    lat.init_equilibrium()
    lat.step(3)
    lat.health()
    lat.close()


@pytest.mark.parametrize("boundary", ["cavity", "periodic"])
@pytest.mark.parametrize("ndx,ndy", [(1, 1), (2, 1), (1, 2), (3, 2)])
def test_host_step_decomposed_bitexact(boundary, ndx, ndy):
    """The host-buffer step (state in host memory, slab-pipelined per block, rims exchanged through the neighbours'
    ghosts first) on decomposed lattices, mixed with device-resident steps: bitwise the oracle."""
    lb = require_gpu()
    nx, ny = 67, 301
    f = orc.perturbed_state(nx, ny, seed=17)
    ref = f.copy()
    run = orc.cavity_run if boundary == "cavity" else orc.periodic_run
    lat = lb.Lattice(nx, ny, boundary, omega=1.6, ndx=ndx, ndy=ndy)
    lat.upload(f)
    for _ in range(3):
        lat.step_host(f, nslabs=5)
    run(ref, 1.6, 3)
    assert np.array_equal(f, ref)
    assert np.array_equal(lat.download(), ref)          # the device copy advanced with the host copy
    lat.step(4)                                         # fused steps continue from the host-stepped state
    run(ref, 1.6, 4)
    assert np.array_equal(lat.download(), ref)
    lat.health()
    lat.close()


@pytest.mark.parametrize("dt", ["float64", "float32"])
@pytest.mark.parametrize("boundary", ["periodic", "cavity", "cavity_xperiodic"])
def test_inplace_aa_bitexact_vs_oracle(dt, boundary):
    """LB_CREATE_INPLACE: ONE buffer advanced with the AA pattern.  Odd step counts leave the buffer in the swapped
    layout -- downloads, row downloads, moments and the digest read through it -- and everything is bit-identical
    to the oracle and to the A/B lattice, ragged / tiny / multi-tile shapes included."""
    lb = require_gpu()
    for nx, ny in SHAPES + [(300, 517)]:
        if boundary != "periodic" and (ny < 2 or (boundary == "cavity" and nx < 2)):
            continue
        f0 = orc.perturbed_state(nx, ny, np.dtype(dt), seed=nx * 77 + ny)
        run = (lambda f, n: orc.periodic_run(f, 1.7, n)) if boundary == "periodic" else \
              (lambda f, n: orc.cavity_run(f, 1.7, n, 0.1, walls_lr=(boundary == "cavity")))
        lat = lb.Lattice(nx, ny, boundary, omega=1.7, u_wall=0.1, dtype=dt, inplace=True)
        ab = lb.Lattice(nx, ny, boundary, omega=1.7, u_wall=0.1, dtype=dt, temporal=1)
        lat.upload(f0)
        ab.upload(f0)
        ref = f0.copy()
        done = 0
        for n in (7, 8, 21):                         # swapped, natural, swapped
            lat.step(n - done)
            ab.step(n - done)
            run(ref, n - done)
            done = n
            assert np.array_equal(lat.download(), ref), (nx, ny, n)
            assert lat.checksum() == ab.checksum(), (nx, ny, n)
            for a, b in zip(lat.moments(), ab.moments()):
                assert np.array_equal(a, b), (nx, ny, n)
        if nx >= 5:
            assert np.array_equal(lat.blocks[0].download_rows(1, 4), ref[:, 1:4])
        lat.upload(ref)                              # an upload in the swapped state restarts in the natural layout
        lat.step(2)
        run(ref, 2)
        assert np.array_equal(lat.download(), ref)
        lat.health()
        lat.close()
        ab.close()


def test_inplace_aa_long_run_and_restrictions():
    lb = require_gpu()
    nx, ny, omega = 96, 80, 1.7
    ref = orc.init_equilibrium(nx, ny)
    orc.cavity_run(ref, omega, 1001)
    lat = lb.Lattice(nx, ny, "cavity", omega=omega, inplace=True)
    lat.init_equilibrium()
    lat.step(1001)
    assert np.array_equal(lat.download(), ref)
    lat.health()
    with pytest.raises(lb.LbmError):
        lat.stream_only(1)
    with pytest.raises(lb.LbmError):
        lat.blocks[0].upload_rows(0, 2, ref[:, 0:2].copy())        # swapped layout (1001 steps): partial uploads refused
    lat.close()
    with pytest.raises(lb.LbmError):
        lb.Lattice(64, 64, "sf_couette", inplace=True)
    with pytest.raises(lb.LbmError):
        lb.Lattice(64, 64, "cavity", ndx=2, inplace=True)


def test_probe_switches_the_whole_decomposition_to_single_steps():
    """The stepping mode is collective: enabling the shear probe on a forced two-steps-per-pass block is refused by
    the C ABI (no silent per-block downgrade that would stall its neighbours); Lattice switches all blocks first."""
    lb = require_gpu()
    nx, ny = 64, 48
    f0, uy_k = orc.shear_wave_init(nx, ny, a0=0.01)
    lat = lb.Lattice(nx, ny, "periodic", omega=1.0, ndx=2, ndy=1, temporal=2)
    assert all(b.temporal_active for b in lat.blocks)
    with pytest.raises(lb.LbmError, match="single-step mode"):
        lat.blocks[0].probe_shear_enable(ny // 2, uy_k[:32], 8)
    lat.upload(f0)
    lat.probe_shear_enable(uy_k, 8)
    assert not any(b.temporal_active for b in lat.blocks)
    lat.step(8)
    ampl = lat.probe_shear_read(8)
    ref = orc.periodic_run(f0.copy(), 1.0, 8, uy_k)
    assert np.abs(ampl - ref).max() < 1e-14
    lat.health()
    lat.close()
