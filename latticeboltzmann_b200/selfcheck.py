"""Deterministic self-check cases shared by bench.py (which runs them on the GPUs before timing) and
tests/make_bench_parity.py (which computes their expected SHA-256 on the CPU with the test suite's checker
and commits it to tests/golden/bench_parity.json).  Pure numpy: this module is part of the product and
imports nothing from the test infrastructure.

The initial macroscopic fields are dyadic rationals of the cell indices, so they are exactly representable
in fp32 and fp64 and do not depend on any libm; f = feq(rho, ux, uy) is then evaluated by the EXACT
equilibrium on either side (c/d2q9.h:59-81), and the lattices are ragged on purpose (1021 x 1531 does not
divide by 2, 4, 8 blocks, by the 254-column fused tile or by its 32 rows)."""
import hashlib

import numpy as np

CASES = {
    # name: (boundary, nx, ny, dtype, omega, u0, steps) -- odd step counts end with a single step after the double steps
    "cavity_f64_1021x1531_w1.7_s41": ("cavity", 1021, 1531, "float64", 1.7, 0.1, 41),
    "periodic_f32_517x1031_w1.2_s20": ("periodic", 517, 1031, "float32", 1.2, 0.0, 20),
}


def fields(nx, ny, dtype, x0=0, y0=0, lnx=None, lny=None):
    """rho, ux, uy on the block [x0, x0+lnx) x [y0, y0+lny) of the global (nx, ny) lattice."""
    lnx = nx if lnx is None else lnx
    lny = ny if lny is None else lny
    k = np.arange(x0, x0 + lnx, dtype=np.int64)[:, None]
    l = np.arange(y0, y0 + lny, dtype=np.int64)[None, :]
    rho = 1.0 + ((7 * k + 13 * l) % 32 - 16) / 1024.0
    ux = ((5 * k + 3 * l) % 17 - 8) / 512.0
    uy = ((11 * k + 2 * l) % 13 - 6) / 512.0
    return tuple(np.ascontiguousarray(a, dtype=np.dtype(dtype)) for a in (rho, ux, uy))


def digest(f):
    return hashlib.sha256(np.ascontiguousarray(f).tobytes()).hexdigest()
