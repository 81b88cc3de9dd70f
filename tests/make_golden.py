"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN PYTHON FUNCTIONS.

Run in the build container only (it reads /root/reference, which does not
exist on the GPU box):

    python tests/make_golden.py

The reference's scripts cannot be imported as modules (they import mpi4py /
matplotlib, which are not installed, and run their main loops at import), so
the function definitions and the few module-level tables they need are lifted
out of the unmodified source files with ``ast`` and executed as they are.
No reference source is copied into this repository; only the numeric
input/output vectors are committed.

What each fixture pins (reference file:line):
  streaming_roll.npz      PyLB/Streaming.py:33-46                          (bit-exact)
  cavity_opt2_bb_*.npz    cavity_opt2.py:94-177  stream_and_bounce_back    (bit-exact)
  collide_test_ref.npz    tests/02-CollideTest.py:43-89 numpy formulas     (tol 1e-7 there; ~1e-15 here)
  cavity_opt1_run.npz     cavity_opt1.py:100-160 + :162-245, 50 steps      (tol 1e-12)
  cavity_opt0_collide.npz cavity_opt0.py:93-138 matrix-form collide        (tol 1e-12)
  shear_opt1_run.npz      shear_wave_opt1.py  stream/collide/amplitude     (tol 1e-12)
  simple_couette.npz      PoiseuilleFlow.py:25-74,93-111                   (bit-exact vs numpy oracle)
  simple_poiseuille.npz   PoiseuilleFlow.py:76-88,129-148                  (bit-exact vs numpy oracle)
  simple_sliding_lid.npz  slidingLid.py:33-108                             (bit-exact vs numpy oracle)
  table_sliding_lid_mpi.npz  slidingLidMPI.py:123-204,264-268 (full-range bounce, one rank)      (bit-exact)
  table_obstacle_channel.npz obstacle_canal.py:303-339,413-458 (methods of obstacleWindTunnel)   (bit-exact)
  simpleflows_test_pins.npz  tests/simpleFlowsTest.py: the reference's OWN unit tests for this path, executed
                             (test_array_allChannel_streaming :260-297, test_faster_principal_calc / test_equilibrium
                             :676-720) and their helper functions (:928-990) evaluated on recorded inputs  (bit-exact)
"""
import ast
import os
import sys

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def lift(path, funcs=(), assigns=(), preset=None):
    """Execute selected top-level FunctionDef / Assign nodes of an unmodified
    reference file inside a fresh namespace and return that namespace."""
    with open(os.path.join(REF, path)) as fh:
        tree = ast.parse(fh.read())
    ns = {"np": np, "sys": sys}
    ns.update(preset or {})
    body = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in funcs:
            body.append(node)
        elif isinstance(node, ast.Assign) and any(
                isinstance(t, ast.Name) and t.id in assigns for t in node.targets):
            body.append(node)
    mod = ast.Module(body=body, type_ignores=[])
    exec(compile(mod, path, "exec"), ns)
    for name in list(funcs) + list(assigns):
        assert name in ns, (path, name)
    return ns


def perturbed(nx, ny, dtype, seed):
    rng = np.random.default_rng(seed)
    w = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)
    f = w[:, None, None] * (1.0 + 0.01 * rng.standard_normal((9, nx, ny)))
    return np.ascontiguousarray(f.astype(dtype))


def gen_streaming():
    ns = lift("PyLB/Streaming.py", funcs=("stream",), assigns=("c_ic",))
    rng = np.random.default_rng(1)
    out = {}
    for tag, (nx, ny) in {"a": (5, 7), "b": (8, 3), "c": (1, 4)}.items():
        f = rng.random((9, nx, ny))
        g = f.copy()
        ns["stream"](g)
        out["in_" + tag] = f
        out["out_" + tag] = g
    np.savez_compressed(os.path.join(OUT, "streaming_roll.npz"), **out)


def gen_cavity_opt2_bb():
    from enum import IntEnum
    path = "simulators/parallel_lid_drive_cavity/cavity_opt2.py"
    for dt in (np.float64, np.float32):
        ns = lift(path, funcs=("stream", "stream_and_bounce_back"), assigns=("D", "c_ic", "w_i"),
                  preset={"IntEnum": IntEnum, "dtype": np.dtype(dt)})
        out = {}
        for tag, (nx, ny, u0) in {"a": (24, 20, 0.1), "b": (7, 5, 0.1), "c": (16, 33, 0.05),
                                  "d": (3, 3, 0.1), "e": (2, 2, 0.1)}.items():
            f = perturbed(nx, ny, dt, seed=10 + nx)
            g = f.copy()
            ns["stream_and_bounce_back"](g, u0)
            out["in_" + tag] = f
            out["out_" + tag] = g
            out["u0_" + tag] = np.array(u0)
            # three chained applications (no collide) exercise the corner/wrap quirk repeatedly
            h = f.copy()
            for _ in range(3):
                ns["stream_and_bounce_back"](h, u0)
            out["out3_" + tag] = h
        np.savez_compressed(os.path.join(OUT, "cavity_opt2_bb_%s.npz" % np.dtype(dt).name), **out)


def gen_collide_test_ref():
    ns = lift("tests/02-CollideTest.py", funcs=("d2q9_equilibrium_ref", "d2q9_collide_ref"),
              assigns=("c_ic", "w_i"))
    rng = np.random.default_rng(2)
    out = {}
    # same shapes and value ranges as tests/02-CollideTest.py:94-111 (there unseeded)
    rho = np.abs(rng.random((2, 2)))
    ux = rng.random((2, 2))
    uy = rng.random((2, 2))
    out["eq_rho"], out["eq_ux"], out["eq_uy"] = rho, ux, uy
    out["eq_out"] = ns["d2q9_equilibrium_ref"](rho, np.array([ux, uy]))
    f = np.abs(rng.random((9, 4, 4)))
    out["col_in"] = f
    for omega in (0.5, 1.7):
        c = f.copy()
        ns["d2q9_collide_ref"](c, omega)
        out["col_out_%s" % omega] = c
    # low-Mach sample closer to real use
    rho = 1 + 0.05 * rng.standard_normal((6, 5))
    ux = 0.1 * rng.standard_normal((6, 5))
    uy = 0.1 * rng.standard_normal((6, 5))
    out["eq2_rho"], out["eq2_ux"], out["eq2_uy"] = rho, ux, uy
    out["eq2_out"] = ns["d2q9_equilibrium_ref"](rho, np.array([ux, uy]))
    np.savez_compressed(os.path.join(OUT, "collide_test_ref.npz"), **out)


def gen_cavity_opt1_run():
    from enum import IntEnum
    path = "simulators/parallel_lid_drive_cavity/cavity_opt1.py"
    dt = np.dtype(np.float64)
    ns = lift(path, funcs=("equilibrium", "collide", "stream", "stream_and_bounce_back"),
              assigns=("D", "c_ic", "w_0", "w_1234", "w_5678", "w_i"),
              preset={"IntEnum": IntEnum, "dtype": dt})
    nx, ny, omega, nsteps = 24, 20, 1.7, 50
    f = ns["equilibrium"](np.ones((nx, ny), dt), np.zeros((nx, ny), dt), np.zeros((nx, ny), dt))
    f = np.ascontiguousarray(f)
    out = {"f0": f.copy(), "omega": np.array(omega), "nsteps": np.array(nsteps), "u0": np.array(0.1)}
    for s in range(nsteps):
        ns["stream_and_bounce_back"](f)
        ns["collide"](f, omega)
        if s + 1 in (1, 10, 50):
            out["f_%d" % (s + 1)] = f.copy()
    out["mass_50"] = np.array(f.sum())
    np.savez_compressed(os.path.join(OUT, "cavity_opt1_run.npz"), **out)


def gen_cavity_opt0_collide():
    from enum import IntEnum
    path = "simulators/parallel_lid_drive_cavity/cavity_opt0.py"
    dt = np.dtype(np.float64)
    ns = lift(path, funcs=("equilibrium", "collide"), assigns=("D", "c_ic", "w_i"),
              preset={"IntEnum": IntEnum, "dtype": dt})
    f = perturbed(9, 11, dt, seed=5)
    g = f.copy()
    ns["collide"](g, 1.7)
    np.savez_compressed(os.path.join(OUT, "cavity_opt0_collide.npz"), f_in=f, f_out=g, omega=np.array(1.7))


def gen_shear_opt1_run():
    path = "simulators/serial_shear_wave/Python/shear_wave_opt1.py"
    dt = np.dtype(np.float64)
    ns = lift(path, funcs=("equilibrium", "collide", "stream"),
              assigns=("c_ic", "w_0", "w_1234", "w_5678"), preset={"dtype": np.float64})
    nx, ny, omega, nsteps, a0 = 60, 40, 1.0, 200, 0.05
    x_k = np.arange(nx)
    uy_k = (a0 * np.sin(2 * np.pi / nx * x_k)).astype(dt)
    f = ns["equilibrium"](np.ones((nx, ny), dt), np.zeros((nx, ny), dt), np.resize(uy_k, (ny, nx)).T.copy())
    f = np.ascontiguousarray(f)
    out = {"f0": f.copy(), "uy_k": uy_k, "omega": np.array(omega), "nsteps": np.array(nsteps)}
    c_ic = ns["c_ic"]
    ampl = []
    for s in range(nsteps):
        ns["stream"](f)
        ns["collide"](f, omega)
        # shear_wave_opt2.py:99 (same expression in opt1)
        ampl += [((c_ic[:, 1].dot(f[:, :, ny // 2]) / (f[:, :, ny // 2].sum(axis=0))) * uy_k).sum() * 2 / nx]
    out["ampl"] = np.array(ampl)
    out["f_end"] = f
    np.savez_compressed(os.path.join(OUT, "shear_opt1_run.npz"), **out)


def gen_simple_flows():
    path = "simulators/simple_flows/PoiseuilleFlow.py"
    vs = np.array([[0, 1, 0, -1, 0, 1, -1, -1, 1], [0, 0, 1, 0, -1, 1, 1, -1, -1]]).T
    fns = ("equilibrium_on_array", "collision", "caluculate_real_values", "stream", "bounce_back",
           "periodic_boundary_with_pressure_variations")
    ns = lift(path, funcs=fns, preset={"relaxation": 0.5, "velocity_set": vs})
    # Couette, PoiseuilleFlow.py:93-111
    nx, ny, uw, nsteps = 12, 10, 0.1, 40
    grid = ns["equilibrium_on_array"](np.ones((nx, ny + 2)), np.zeros((nx, ny + 2)), np.zeros((nx, ny + 2)))
    out = {"f0": grid.copy(), "uw": np.array(uw), "omega": np.array(0.5), "nsteps": np.array(nsteps)}
    for s in range(nsteps):
        rho, ux, uy = ns["caluculate_real_values"](grid)
        ns["collision"](grid, rho, ux, uy)
        ns["stream"](grid)
        ns["bounce_back"](grid, uw)
        if s + 1 in (1, 5, nsteps):
            out["f_%d" % (s + 1)] = grid.copy()
    out["ux_last"] = ux
    np.savez_compressed(os.path.join(OUT, "simple_couette.npz"), **out)
    # Poiseuille, PoiseuilleFlow.py:129-148
    nx, ny, nsteps, diff = 12, 10, 40, 0.001
    grid = ns["equilibrium_on_array"](np.ones((nx + 2, ny + 2)), np.zeros((nx + 2, ny + 2)), np.zeros((nx + 2, ny + 2)))
    out = {"f0": grid.copy(), "rho_in": np.array(1 + diff), "rho_out": np.array(1 - diff),
           "omega": np.array(0.5), "nsteps": np.array(nsteps)}
    for s in range(nsteps):
        ns["periodic_boundary_with_pressure_variations"](grid, 1 + diff, 1 - diff)
        ns["stream"](grid)
        ns["bounce_back"](grid, 0.0)
        rho, ux, uy = ns["caluculate_real_values"](grid)
        ns["collision"](grid, rho, ux, uy)
        if s + 1 in (1, 5, nsteps):
            out["f_%d" % (s + 1)] = grid.copy()
    out["ux_last"] = ux
    np.savez_compressed(os.path.join(OUT, "simple_poiseuille.npz"), **out)
    # Sliding lid, slidingLid.py:33-108
    path = "simulators/simple_flows/slidingLid.py"
    L, uw, re, nsteps = 10, 0.1, 1000, 40
    omega = (2 * re) / (6 * L * uw + re)   # slidingLid.py:28
    ns = lift(path, funcs=("stream", "equilibrium", "collision", "caluculate_rho_ux_uy", "bounce_back"),
              preset={"relaxation": omega, "velocity_set": vs})
    n = L + 2
    grid = ns["equilibrium"](np.ones((n, n)), np.zeros((n, n)), np.zeros((n, n)))
    out = {"f0": grid.copy(), "uw": np.array(uw), "omega": np.array(omega), "nsteps": np.array(nsteps)}
    for s in range(nsteps):
        ns["stream"](grid)
        ns["bounce_back"](grid, uw)
        rho, ux, uy = ns["caluculate_rho_ux_uy"](grid)
        ns["collision"](grid, rho, ux, uy)
        if s + 1 in (1, 5, nsteps):
            out["f_%d" % (s + 1)] = grid.copy()
    np.savez_compressed(os.path.join(OUT, "simple_sliding_lid.npz"), **out)


def lift_methods(path, cls, methods, preset=None):
    """The named methods of a class of an unmodified reference file, as plain functions taking `self`."""
    with open(os.path.join(REF, path)) as fh:
        tree = ast.parse(fh.read())
    ns = {"np": np}
    ns.update(preset or {})
    body = []
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            body = [n for n in node.body if isinstance(n, ast.FunctionDef) and n.name in methods]
    assert len(body) == len(methods), (path, cls, methods)
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return ns


def gen_table_flows():
    from types import SimpleNamespace as NS
    vs = np.array([[0, 1, 0, -1, 0, 1, -1, -1, 1], [0, 0, 1, 0, -1, 1, 1, -1, -1]]).T
    # slidingLidMPI.py on one rank: stream -> bounce_back_choosen (all four walls, FULL ranges) -> moments -> collision (:264-268)
    path = "simulators/simple_flows/slidingLidMPI.py"
    ns = lift(path, funcs=("stream", "equilibrium", "collision", "caluculate_rho_ux_uy", "bounce_back_choosen"),
              preset={"velocity_set": vs})
    L, uw, omega, nsteps = 11, 0.1, 1.2, 40
    n = L + 2
    info = NS(boundaries_info=NS(apply_right=True, apply_left=True, apply_bottom=True, apply_top=True))
    grid = perturbed(n, n, np.float64, 41)
    out = {"f0": grid.copy(), "uw": np.array(uw), "omega": np.array(omega), "nsteps": np.array(nsteps)}
    for s in range(nsteps):
        ns["stream"](grid)
        ns["bounce_back_choosen"](grid, uw, info)
        rho, ux, uy = ns["caluculate_rho_ux_uy"](grid)
        ns["collision"](grid, rho, ux, uy, omega)
        if s + 1 in (1, 5, nsteps):
            out["f_%d" % (s + 1)] = grid.copy()
    np.savez_compressed(os.path.join(OUT, "table_sliding_lid_mpi.npz"), **out)
    # obstacle_canal.py: the methods of obstacleWindTunnel, x periodic, bounce-back bottom / top, rectangular obstacle
    path = "simulators/experimantal_flows/obstacle_canal.py"
    import enum
    states = enum.Enum("boundaryStates", "NONE BAUNCE_BACK BAUNCE_BACK_MOVING_WALL PERIODIC_BOUNDARY COMMUNICATE")
    m = lift_methods(path, "obstacleWindTunnel", ("equilibrium", "stream", "bounce_back_choosen", "apply_obstacle",
                                                   "caluculate_rho_ux_uy", "collision"),
                     preset={"velocity_set": vs, "boundaryStates": states})
    nx, ny, omega, nsteps = 24, 16, 0.9, 40
    x0, x1, y0, y1 = 8, 12, 5, 9
    BB, NO = states.BAUNCE_BACK, states.NONE
    self = NS(grid=perturbed(nx, ny, np.float64, 43), relaxation=omega, uw=0.0, rho=None, ux=None, uy=None,
              packed_info=NS(boundaries_info=NS(apply_right=NO, apply_left=NO, apply_bottom=BB, apply_top=BB)),
              local_obstacle=NS(start_x=x0, end_x=x1, start_y=y0, end_y=y1,
                                boundary_state=NS(apply_left=BB, apply_right=BB, apply_top=BB, apply_bottom=BB)))
    self.equilibrium = lambda: m["equilibrium"](self)
    out = {"f0": self.grid.copy(), "omega": np.array(omega), "nsteps": np.array(nsteps), "obstacle": np.array([x0, x1, y0, y1])}
    for s in range(nsteps):
        m["stream"](self)
        m["bounce_back_choosen"](self)
        m["caluculate_rho_ux_uy"](self)       # apply_obstacle zeroes ux / uy inside the obstacle; run() recomputes them right after (:273-274)
        m["apply_obstacle"](self)
        m["caluculate_rho_ux_uy"](self)
        m["collision"](self)
        if s + 1 in (1, 5, nsteps):
            out["f_%d" % (s + 1)] = self.grid.copy()
    np.savez_compressed(os.path.join(OUT, "table_obstacle_channel.npz"), **out)


def gen_reference_test_pins():
    """The unit tests the reference itself holds for this path.  The test classes are lifted
    like the functions above and their methods RUN here (they must pass under this numpy); then the helper functions
    they exercise are evaluated on recorded inputs so that the oracle and the GPU can be held against them."""
    import unittest
    path = "tests/simpleFlowsTest.py"
    with open(os.path.join(REF, path)) as fh:
        tree = ast.parse(fh.read())
    helpers = ("stream", "equlibrium_function", "equilibrium_on_array_test", "calculate_3pincipal_values", "caluculate_real_values")
    classes = ("testsInStreaming", "testsForNewCollision")
    body = [n for n in tree.body
            if (isinstance(n, ast.FunctionDef) and n.name in helpers) or (isinstance(n, ast.ClassDef) and n.name in classes)
            or (isinstance(n, ast.Assign) and any(isinstance(t, ast.Name) and t.id == "c_ic" for t in n.targets))]
    ns = {"np": np, "unittest": unittest}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    ran = {}
    for cls, method in (("testsInStreaming", "test_array_allChannel_streaming"), ("testsForNewCollision", "test_faster_principal_calc"),
                        ("testsForNewCollision", "test_equilibrium")):
        res = unittest.TestResult()
        ns[cls](method).run(res)
        assert res.testsRun == 1 and res.wasSuccessful(), (cls, method, res.errors, res.failures)
        ran[method] = True
    out = {"reference_tests_passed": np.array(sorted(ran))}
    # the streaming test's pattern (simpleFlowsTest.py:262-272): a line of ones in every channel, one periodic step
    grid = np.zeros((9, 9, 9))
    for watch in range(1, 9):
        grid[watch, 1, 1:8] = 1
    out["stream_in"] = grid.copy()
    ns["stream"](grid)                               # simpleFlowsTest.py:928-930 (== PyLB/Streaming.py)
    out["stream_out"] = grid.copy()
    # equilibrium: scalar form (:933-948) and array form (:951-968) on the test's uniform field and on random fields
    rng = np.random.default_rng(11)
    for tag, (rho, ux, uy) in {"uniform": (np.full((5, 4), 9.0), np.zeros((5, 4)), np.zeros((5, 4))),
                               "random": (1 + 0.05 * rng.standard_normal((16, 12)), 0.1 * rng.standard_normal((16, 12)),
                                          0.1 * rng.standard_normal((16, 12)))}.items():
        arr = ns["equilibrium_on_array_test"](rho, ux, uy)
        sca = np.empty_like(arr)
        for k in range(rho.shape[0]):
            for l in range(rho.shape[1]):
                sca[:, k, l] = ns["equlibrium_function"](rho[k, l], ux[k, l], uy[k, l])
        out.update({"rho_" + tag: rho, "ux_" + tag: ux, "uy_" + tag: uy, "feq_array_" + tag: arr, "feq_scalar_" + tag: sca})
    # moments (:1088-1092 and the per-node form :971-976) of a perturbed equilibrium
    f = out["feq_array_random"] * (1 + 0.01 * rng.standard_normal((9, 16, 12)))
    rho, ux, uy = ns["caluculate_real_values"](f)
    out.update({"mom_f": f, "mom_rho": rho, "mom_ux": ux, "mom_uy": uy})
    node = np.array([ns["calculate_3pincipal_values"](f[:, 3, 7])])
    out["mom_node_3_7"] = node
    np.savez_compressed(os.path.join(OUT, "simpleflows_test_pins.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    gen_reference_test_pins()
    gen_streaming()
    gen_cavity_opt2_bb()
    gen_collide_test_ref()
    gen_cavity_opt1_run()
    gen_cavity_opt0_collide()
    gen_shear_opt1_run()
    gen_simple_flows()
    gen_table_flows()
    for n in sorted(os.listdir(OUT)):
        print("%-32s %8d bytes" % (n, os.path.getsize(os.path.join(OUT, n))))


if __name__ == "__main__":
    main()
