"""Pin the CPU oracle (oracle/) against golden vectors produced by executing the
reference's own Python functions (tests/make_golden.py) and against the known
answers recorded in SURVEY.md §8c.  CPU only."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from oracle import simple_flows as sf


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def rel_err(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


def test_stream_is_np_roll_bitexact(golden_dir):
    g = load(golden_dir, "streaming_roll.npz")         # PyLB/Streaming.py:33-46
    for tag in "abc":
        f = g["in_" + tag].copy()
        orc.stream(f)
        assert np.array_equal(f, g["out_" + tag])


@pytest.mark.parametrize("dt", ["float64", "float32"])
def test_cavity_stream_and_bounce_back_bitexact(golden_dir, dt):
    g = load(golden_dir, "cavity_opt2_bb_%s.npz" % dt)  # cavity_opt2.py:109-177
    for tag in "abcde":
        f = g["in_" + tag].copy()
        u0 = float(g["u0_" + tag])
        orc.cavity_stream_and_bounce_back(f, u0)
        assert np.array_equal(f, g["out_" + tag]), tag
        orc.cavity_stream_and_bounce_back(f, u0)
        orc.cavity_stream_and_bounce_back(f, u0)
        assert np.array_equal(f, g["out3_" + tag]), tag


@pytest.mark.parametrize("dt", ["float64", "float32"])
@pytest.mark.parametrize("walls_lr", [True, False])
def test_pull_formulation_equals_literal_sequence(dt, walls_lr):
    """the fused pull rule == roll-then-overwrite, bitwise."""
    for nx, ny in [(24, 20), (7, 5), (16, 33), (3, 3), (2, 2), (40, 2)]:
        f = orc.perturbed_state(nx, ny, np.dtype(dt), seed=nx * ny)
        a = f.copy()
        b = np.empty_like(f)
        for step in range(5):
            orc.cavity_stream_and_bounce_back(a, 0.1, walls_lr)
            orc.cavity_step_pull(f, b, 1.7, 0.1, walls_lr, do_collide=False)
            assert np.array_equal(a, b), (nx, ny, step)
            orc.collide(a, 1.7)
            orc.collide(b, 1.7)
            f, b = b, f


def test_pull_run_equals_literal_run_200_steps():
    f = orc.init_equilibrium(24, 20)
    a = f.copy()
    orc.cavity_run(a, 1.7, 200)
    b = np.empty_like(f)
    for _ in range(200):
        orc.cavity_step_pull(f, b, 1.7)
        f, b = b, f
    assert np.array_equal(a, f)


def test_periodic_pull_equals_roll():
    f = orc.perturbed_state(9, 13, seed=3)
    a = f.copy()
    b = np.empty_like(f)
    orc.stream(a)
    orc.periodic_step_pull(f, b, 1.0, do_collide=False)
    assert np.array_equal(a, b)


def test_equilibrium_and_collide_vs_reference_test_formulas(golden_dir):
    """tests/02-CollideTest.py:94-111 with the reference's tolerance (1e-7, PyLBTest.py:75)."""
    g = load(golden_dir, "collide_test_ref.npz")
    for p in ("eq", "eq2"):
        rho, ux, uy = g[p + "_rho"], g[p + "_ux"], g[p + "_uy"]
        e = np.zeros((9,) + rho.shape)
        orc.equilibrium(rho.reshape(-1).copy(), ux.reshape(-1).copy(), uy.reshape(-1).copy(), e.reshape(9, -1))
        assert np.abs(e - g[p + "_out"]).max() < 1e-7
        assert np.abs(e - g[p + "_out"]).max() < 1e-14      # what is actually achieved
    for omega in (0.5, 1.7):
        c = g["col_in"].copy()
        orc.collide(c.reshape(9, -1), omega)
        assert np.abs(c - g["col_out_%s" % omega]).max() < 1e-7
        assert np.abs(c - g["col_out_%s" % omega]).max() < 1e-13


def test_collide_vs_opt0_matrix_form(golden_dir):
    g = load(golden_dir, "cavity_opt0_collide.npz")    # cavity_opt0.py:93-138
    c = g["f_in"].copy()
    orc.collide(c, float(g["omega"]))
    assert rel_err(c, g["f_out"]) < 1e-12


def test_cavity_run_vs_opt1_numpy(golden_dir):
    """cavity_opt1.py (numpy collide, different algebraic form) for 50 steps: 1e-12 relative."""
    g = load(golden_dir, "cavity_opt1_run.npz")
    f = g["f0"].copy()
    done = 0
    for n in (1, 10, 50):
        orc.cavity_run(f, float(g["omega"]), n - done, float(g["u0"]))
        done = n
        assert rel_err(f, g["f_%d" % n]) < 1e-12, n
    # the lid leaks mass at the corners, 480 -> 480.0409 after 50 steps
    assert abs(f.sum() - 480.0409) < 1e-4
    assert abs(f.sum() - float(g["mass_50"])) < 1e-9


def test_shear_wave_vs_opt1_numpy(golden_dir):
    g = load(golden_dir, "shear_opt1_run.npz")         # shear_wave_opt1.py
    f = g["f0"].copy()
    ampl = orc.periodic_run(f, float(g["omega"]), int(g["nsteps"]), g["uy_k"])
    assert rel_err(f, g["f_end"]) < 1e-12
    assert np.max(np.abs(ampl - g["ampl"])) < 1e-14


def test_shear_wave_known_answers_300x200():
    """SURVEY.md §8c known answers: 300x200, omega=1, 1000 steps."""
    nx, ny, a0 = 300, 200, 0.01
    f, uy_k = orc.shear_wave_init(nx, ny, a0=a0)
    ampl = orc.periodic_run(f, 1.0, 1000, uy_k)
    a_init = (uy_k * uy_k).sum() * 2 / nx
    assert abs(ampl[0] / a_init - 0.999926894492) < 1e-9
    assert abs(ampl[-1] / a_init - 0.929500270576) < 1e-9
    # viscosity from the decay a(t) = a0 exp(-nu k^2 t), analytic nu = (1/omega - 1/2)/3 = 1/6
    kk = (2 * np.pi / nx) ** 2
    t = np.arange(1, 1001)
    nu = -np.polyfit(t, np.log(ampl / a_init), 1)[0] / kk
    assert abs(nu - 1 / 6) / (1 / 6) < 1e-6


def test_scalar_equilibrium_matches_array():
    e = orc.equilibrium1(1.1, 0.05, -0.02)
    f = np.zeros((9, 1))
    orc.equilibrium(np.array([1.1]), np.array([0.05]), np.array([-0.02]), f)
    assert np.array_equal(e, f[:, 0])
    assert abs(e.sum() - 1.1) < 1e-15


def test_simple_flows_couette_bitexact(golden_dir):
    g = load(golden_dir, "simple_couette.npz")          # PoiseuilleFlow.py:93-111
    f = g["f0"].copy()
    assert np.array_equal(f, sf.feq(np.ones(f.shape[1:]), np.zeros(f.shape[1:]), np.zeros(f.shape[1:])))
    for s in range(int(g["nsteps"])):
        ux = sf.couette_step(f, float(g["omega"]), float(g["uw"]))
        if "f_%d" % (s + 1) in g:
            assert np.array_equal(f, g["f_%d" % (s + 1)]), s
    assert np.array_equal(ux, g["ux_last"])


def test_simple_flows_poiseuille_bitexact(golden_dir):
    g = load(golden_dir, "simple_poiseuille.npz")       # PoiseuilleFlow.py:129-148
    f = g["f0"].copy()
    for s in range(int(g["nsteps"])):
        ux = sf.poiseuille_step(f, float(g["omega"]), float(g["rho_in"]), float(g["rho_out"]))
        if "f_%d" % (s + 1) in g:
            assert np.array_equal(f, g["f_%d" % (s + 1)]), s
    assert np.array_equal(ux, g["ux_last"])


def test_simple_flows_sliding_lid_bitexact(golden_dir):
    g = load(golden_dir, "simple_sliding_lid.npz")      # slidingLid.py:68-108
    f = g["f0"].copy()
    for s in range(int(g["nsteps"])):
        sf.sliding_lid_step(f, float(g["omega"]), float(g["uw"]))
        if "f_%d" % (s + 1) in g:
            assert np.array_equal(f, g["f_%d" % (s + 1)]), s


@pytest.mark.parametrize("dt", ["float64", "float32"])
def test_opt2_numpy_structure_bitexact(golden_dir, dt):
    """The numpy-structured step bench.py times as the CPU baseline == the reference's function."""
    from oracle import opt2_numpy
    g = load(golden_dir, "cavity_opt2_bb_%s.npz" % dt)
    for tag in "abcde":
        f = g["in_" + tag].copy()
        opt2_numpy.stream_and_bounce_back(f, float(g["u0_" + tag]))
        assert np.array_equal(f, g["out_" + tag]), tag
    t = opt2_numpy.run_independent_blocks(2, 32, 32, 1.7, 1, 2)
    assert t > 0


@pytest.mark.parametrize("dt", ["float64", "float32"])
@pytest.mark.parametrize("walls_lr", [True, False])
def test_hoisted_pull_equals_literal_pull(dt, walls_lr):
    """The index-hoisted pull loop (what the 4096^2 GPU test and bench.py's strip check run) is the literal
    per-cell rule: bitwise on ragged / degenerate shapes, and the threaded runner equals the literal
    roll-then-overwrite sequence over many steps."""
    for nx, ny in ((37, 29), (5, 2), (2, 2), (64, 50), (3, 7)):
        f = orc.perturbed_state(nx, ny, np.dtype(dt), seed=nx + ny)
        a, b = np.empty_like(f), np.empty_like(f)
        orc.cavity_step_pull(f, a, 1.3, walls_lr=walls_lr)
        orc.cavity_step_pull_hoisted(f, b, 1.3, walls_lr=walls_lr)
        assert np.array_equal(a, b), (nx, ny)
    f = orc.perturbed_state(45, 33, np.dtype(dt), seed=4)
    ref = f.copy()
    orc.cavity_run(ref, 1.7, 25, walls_lr=walls_lr)
    got = orc.cavity_run_threaded(f.copy(), 1.7, 25, walls_lr=walls_lr, threads=3)
    assert np.array_equal(got, ref)


def test_bench_selfcheck_hashes_are_the_oracles(golden_dir):
    """tests/golden/bench_parity.json (what bench.py compares the GPUs' gathered fields with at every N) is
    what the oracle computes for latticeboltzmann_b200/selfcheck.py's cases."""
    import json
    import make_bench_parity
    from latticeboltzmann_b200 import selfcheck
    want = json.load(open(os.path.join(golden_dir, "bench_parity.json")))["sha256"]
    assert set(want) == set(selfcheck.CASES)
    name = "periodic_f32_517x1031_w1.2_s20"
    assert make_bench_parity.expected(name) == want[name]
    rho, ux, uy = selfcheck.fields(64, 48, "float32", 5, 7, 20, 11)
    full = selfcheck.fields(64, 48, "float32")
    assert all(np.array_equal(a, b[5:25, 7:18]) for a, b in zip((rho, ux, uy), full))       # blocks are slices of the global fields
    assert all(np.array_equal(a.astype(np.float64), b) for a, b in zip(full, selfcheck.fields(64, 48, "float64")))   # exactly representable


def _table_cases(golden_dir):
    from latticeboltzmann_b200 import boundary_table as bt
    g = np.load(os.path.join(golden_dir, "table_sliding_lid_mpi.npz"))
    n = g["f0"].shape[1]
    yield ("sliding_lid_mpi", g, bt.sliding_lid_mpi_table(n, n, float(g["uw"])),
           lambda f: sf.sliding_lid_mpi_step(f, float(g["omega"]), float(g["uw"])))
    h = np.load(os.path.join(golden_dir, "table_obstacle_channel.npz"))
    x0, x1, y0, y1 = (int(v) for v in h["obstacle"])
    yield ("obstacle_channel", h, bt.obstacle_channel_table(h["f0"].shape[1], h["f0"].shape[2], x0, x1, y0, y1),
           lambda f: sf.obstacle_channel_step(f, float(h["omega"]), x0, x1, y0, y1))


def test_boundary_tables_bitexact_vs_reference_functions(golden_dir):
    """N4: slidingLidMPI.py's full-range bounce and obstacle_canal.py's rectangular obstacle.  Goldens come from the
    reference's own functions / methods; the literal numpy restatement AND the symbolic per-cell table
    (latticeboltzmann_b200/boundary_table.py, the form the kernel applies) both reproduce them bit for bit."""
    for name, g, (cells, src, add), literal in _table_cases(golden_dir):
        fa, fb = g["f0"].copy(), g["f0"].copy()
        for s in range(1, int(g["nsteps"]) + 1):
            literal(fa)
            sf.table_step(fb, float(g["omega"]), cells, src, add)
            if "f_%d" % s in g.files:
                assert np.array_equal(fa, g["f_%d" % s]), (name, "literal", s)
                assert np.array_equal(fb, g["f_%d" % s]), (name, "table", s)
        assert 0 < len(cells) < fa.shape[1] * fa.shape[2] // 2          # O(perimeter) cells, not the lattice


def test_symbolic_grid_rejects_what_a_table_cannot_hold():
    from latticeboltzmann_b200.boundary_table import SymbolicGrid
    g = SymbolicGrid(6, 5).stream()
    g[7, :, -2] = g[5, :, -1] - 0.01
    with pytest.raises(ValueError):
        g[8, :, 1] = g[7, :, -2] + 0.01        # shifting an already shifted population: two roundings
    with pytest.raises(TypeError):
        g[1, 1, :] = 0.0
    with pytest.raises(RuntimeError):
        SymbolicGrid(4, 4).table()


def test_reference_unit_tests_pin_the_oracle(golden_dir):
    """The reference's OWN unit tests for this path (tests/simpleFlowsTest.py) were executed by
    tests/make_golden.py (all three pass); their helper functions, evaluated on recorded inputs, pin the oracle."""
    g = load(golden_dir, "simpleflows_test_pins.npz")
    assert set(g["reference_tests_passed"]) == {"test_array_allChannel_streaming", "test_faster_principal_calc", "test_equilibrium"}
    # streaming convention, simpleFlowsTest.py:260-297: one line of ones per channel moves by its lattice velocity
    src = g["stream_in"]
    out = src.copy()
    orc.stream(out)
    assert np.array_equal(out, g["stream_out"])
    pulled = np.empty_like(src)
    orc.periodic_step_pull(src, pulled, 1.0, do_collide=False)          # the gather form the kernels use
    assert np.array_equal(pulled, out)
    moves = {1: (1, 0), 2: (0, 1), 3: (-1, 0), 4: (0, -1), 5: (1, 1), 6: (-1, 1), 7: (-1, -1), 8: (1, -1)}     # the test's eight assertion loops
    for ch, (dx, dy) in moves.items():
        for i in range(1, 8):
            assert src[ch, i, 1] == out[ch, i + dx, 1 + dy]
    # equilibrium, :700-720 (array form == scalar form, exactly, on the test's uniform field) and the helpers :933-968
    for tag in ("uniform", "random"):
        rho, ux, uy = g["rho_" + tag], g["ux_" + tag], g["uy_" + tag]
        assert np.array_equal(sf.feq(rho, ux, uy), g["feq_array_" + tag])
        if tag == "uniform":
            assert np.array_equal(g["feq_array_" + tag], g["feq_scalar_" + tag])          # what the reference test asserts
        else:
            # For u != 0 the reference's scalar helper (:933-948) writes channels 6 and 8 with +-3(ux + uy) where the array
            # form (:951-968, the one the simulators use: PoiseuilleFlow.py:25-42) has +-3(ux - uy); the reference only
            # asserts their equality on a field at rest.  The other seven channels agree to rounding.
            same = [0, 1, 2, 3, 4, 5, 7]
            assert rel_err(g["feq_scalar_" + tag][same], g["feq_array_" + tag][same]) < 1e-14
    # moments, :676-698 and the helpers :971-976, :1088-1092
    rho, ux, uy = sf.moments(g["mom_f"])
    assert np.array_equal(rho, g["mom_rho"]) and np.array_equal(ux, g["mom_ux"]) and np.array_equal(uy, g["mom_uy"])
    node = g["mom_node_3_7"][0]
    assert abs(node[0] - rho[3, 7]) <= 2e-16 * abs(rho[3, 7]) and abs(node[1] - ux[3, 7]) < 1e-16 and abs(node[2] - uy[3, 7]) < 1e-16
