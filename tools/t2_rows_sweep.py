"""Fused-tile height sweep of the two-steps-per-pass kernel at several lattice sizes (developer tool, GPU box)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latticeboltzmann_b200 as lb

SIZES = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [2048, 3072, 4096, 8192, 16384]
ROWS = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8, 16, 32, 64]
for n in SIZES:
    res = {"n": n}
    for rows in ROWS:
        os.environ["LBM_T2_ROWS"] = str(rows)
        lat = lb.Lattice(n, n, "cavity", omega=2000.0 / (0.6 * n + 1000.0), temporal=2)
        lat.init_equilibrium()
        lat.step(10)
        lat.sync()
        steps = 200 if n <= 4096 else 40
        best = min(lat.step_timed(steps) for _ in range(3))
        lat.health()
        lat.close()
        res["rows%d" % rows] = round(n * n * steps / (best * 1e-3) / 1e9, 2)
    os.environ.pop("LBM_T2_ROWS")
    lat = lb.Lattice(n, n, "cavity", omega=2000.0 / (0.6 * n + 1000.0), temporal=1)
    lat.init_equilibrium()
    lat.step(10)
    lat.sync()
    best = min(lat.step_timed(steps) for _ in range(3))
    lat.close()
    res["single_step"] = round(n * n * steps / (best * 1e-3) / 1e9, 2)
    print(json.dumps(res), flush=True)
