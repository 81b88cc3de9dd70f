"""Developer sweep (not the contract bench): MLUPS of the fused step for a few tunings."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latticeboltzmann_b200 as lb


def run(n, dtype, arith, rows, boundary="cavity", steps=60, warm=10):
    lat = lb.Lattice(n, n, boundary, omega=1.7, dtype=dtype, arith=arith, rows_per_tile=rows)
    lat.init_equilibrium()
    lat.step(warm)
    lat.sync()
    best = 1e30
    for _ in range(3):
        ms = lat.step_timed(steps)
        best = min(best, ms)
    lat.health()
    lat.close()
    mlups = n * n * steps / (best * 1e-3) / 1e6
    bpc = 144 if np.dtype(dtype) == np.float64 else 72
    return mlups, mlups * 1e6 * bpc / 1e9


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [4096]
    for n in sizes:
        for dtype in ("float64", "float32"):
            for arith in ("exact", "fast"):
                for rows in (2, 4, 8, 16, 32):
                    m, gbs = run(n, dtype, arith, rows)
                    print(json.dumps({"n": n, "dtype": dtype, "arith": arith, "rows": rows,
                                      "mlups": round(m, 1), "GBs": round(gbs, 1)}), flush=True)
