import numpy as np
import pytest


def rel_err(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def require_gpu():
    import latticeboltzmann_b200 as lb
    lib = lb.load_native()
    if lib.lb_device_count() <= 0:
        pytest.fail("gpu test selected but no CUDA device is visible (no CPU fallback exists)")
    return lb
