import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(params=["resident", "resident-one-step-per-barrier", "per-step"])
def stepping_path(request, monkeypatch):
    """Small single-block lattices can advance through the resident multi-step kernels (one cooperative launch;
    tiny lattices two steps per grid barrier -- the default -- or one step per barrier) or through one fused launch
    per step; tests that use this fixture run once on each path."""
    monkeypatch.setenv("LBM_RESIDENT", "0" if request.param == "per-step" else "1")
    monkeypatch.setenv("LBM_RESIDENT2", "0" if request.param == "resident-one-step-per-barrier" else "1")
    return request.param


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
