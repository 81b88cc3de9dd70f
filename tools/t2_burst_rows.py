"""Bench-like (burst) timing of fused-tile heights: for every height a FRESH lattice, 5 warm-up steps, 50 timed steps,
a pause -- the conditions of `bench.py`'s timed region (a cool GPU, 0.15 s of load), as opposed to the sustained,
power-capped regime tools/t2_env_ab.py measures (developer tool, GPU box).

    python tools/t2_burst_rows.py [n] [rows,rows,...] [cycles] [float32|float64]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import latticeboltzmann_b200 as lb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ROWS = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [32, 48, 64, 96]
cycles = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dtype = np.dtype(sys.argv[4] if len(sys.argv) > 4 else "float64")
steps = 50 if n >= 8192 else 400
res = {r: [] for r in ROWS}
for _ in range(cycles):
    for rows in ROWS:
        os.environ["LBM_T2_ROWS"] = str(rows)
        lat = lb.Lattice(n, n, "cavity", omega=2000.0 / (0.6 * n + 1000.0), temporal=2, dtype=dtype)
        lat.init_equilibrium()
        lat.step(5)
        lat.sync()
        ms = lat.step_timed(steps)
        lat.close()
        res[rows].append(round(n * n * steps / (ms * 1e-3) / 1e9, 2))
        time.sleep(1.0)
for rows in ROWS:
    print(json.dumps({"n": n, "dtype": str(dtype), "steps": steps, "rows": rows, "glups_burst": res[rows], "median": sorted(res[rows])[len(res[rows]) // 2]}), flush=True)
