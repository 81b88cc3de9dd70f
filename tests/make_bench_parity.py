"""Expected SHA-256 of the self-check cases bench.py runs on the GPUs (latticeboltzmann_b200/selfcheck.py),
computed with the CPU oracle (oracle/d2q9_oracle_impl.h: c/d2q9.h + cavity_opt2.py:109-177 restated).

    python tests/make_bench_parity.py        -> tests/golden/bench_parity.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from latticeboltzmann_b200 import selfcheck          # noqa: E402
from oracle import oracle as orc                      # noqa: E402


def expected(name):
    boundary, nx, ny, dtype, omega, u0, steps = selfcheck.CASES[name]
    rho, ux, uy = selfcheck.fields(nx, ny, dtype)
    f = orc.init_equilibrium(nx, ny, np.dtype(dtype), rho, ux, uy)
    if boundary == "cavity":
        orc.cavity_run(f, omega, steps, u0=u0)
    else:
        orc.periodic_run(f, omega, steps)
    return selfcheck.digest(f)


if __name__ == "__main__":
    out = {name: expected(name) for name in selfcheck.CASES}
    fn = os.path.join(ROOT, "tests", "golden", "bench_parity.json")
    with open(fn, "w") as fh:
        json.dump({"generator": "tests/make_bench_parity.py (CPU oracle)", "sha256": out}, fh, indent=1)
        fh.write("\n")
    print(json.dumps(out, indent=1))
