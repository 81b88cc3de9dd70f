"""Randomised soak of every stepping path against the CPU checker: random shapes (degenerate ones included), boundary,
dtype, decomposition, mode and step counts; bitwise comparison.  Developer tool (GPU box):  python tools/soak.py [cases] [seed]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latticeboltzmann_b200 as lb
from oracle import oracle as orc

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
MODES = ("per-step", "resident1", "resident2", "temporal", "inplace", "blocks", "blocks-temporal", "host-step")
bad = 0
counts = {m: 0 for m in MODES}
for c in range(cases):
    mode = MODES[rng.integers(len(MODES))]
    boundary = ("periodic", "cavity", "cavity_xperiodic")[rng.integers(3)]
    dt = (np.float64, np.float32)[rng.integers(2)]
    nx = int(rng.choice([1, 2, 3, 5, 16, 17, 31, 33, 48, 70, 129])) if rng.random() < 0.5 else int(rng.integers(1, 140))
    ny = int(rng.choice([1, 2, 3, 31, 32, 33, 64, 255, 256, 257, 300, 511, 530])) if rng.random() < 0.5 else int(rng.integers(1, 600))
    if boundary != "periodic":
        ny = max(ny, 2)
        if boundary == "cavity":
            nx = max(nx, 2)
    steps = [int(s) for s in rng.integers(1, 9, size=int(rng.integers(1, 4)))]
    ndx = ndy = 1
    kw = {}
    os.environ["LBM_RESIDENT"] = "0"
    os.environ["LBM_RESIDENT2"] = "1"
    if mode == "resident1":
        os.environ["LBM_RESIDENT"] = "1"
        os.environ["LBM_RESIDENT2"] = "0"
    elif mode == "resident2":
        os.environ["LBM_RESIDENT"] = "1"
    elif mode == "temporal":
        kw["temporal"] = 2
    elif mode == "inplace":
        kw["inplace"] = True
    elif mode in ("blocks", "blocks-temporal", "host-step"):
        ndx, ndy = int(rng.integers(1, min(4, nx) + 1)), int(rng.integers(1, min(4, ny) + 1))
        kw["temporal"] = 2 if mode == "blocks-temporal" else 1
    f0 = orc.perturbed_state(nx, ny, dt, seed=c)
    ref = f0.copy()
    lat = lb.Lattice(nx, ny, boundary, omega=1.7, u_wall=0.1, dtype=dt, ndx=ndx, ndy=ndy, **kw)
    lat.upload(f0)
    ok = True
    host = f0.copy()
    for n in steps:
        if mode == "host-step":
            for _ in range(n):
                lat.step_host(host, nslabs=int(rng.integers(1, 6)))
        else:
            lat.step(n)
        if boundary == "periodic":
            orc.periodic_run(ref, 1.7, n)
        else:
            orc.cavity_run(ref, 1.7, n, 0.1, walls_lr=(boundary == "cavity"))
        got = host if mode == "host-step" else lat.download()
        ok = ok and bool(np.array_equal(got, ref)) and bool(np.array_equal(lat.download(), ref))
    lat.health()
    lat.close()
    counts[mode] += 1
    if not ok:
        bad += 1
        print("MISMATCH", dict(case=c, mode=mode, boundary=boundary, dtype=np.dtype(dt).name, nx=nx, ny=ny, ndx=ndx, ndy=ndy, steps=steps), flush=True)
print("soak: %d cases, %d mismatches, per mode %s" % (cases, bad, counts), flush=True)
sys.exit(1 if bad else 0)
