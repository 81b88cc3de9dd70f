"""Reference-named modules (``_lbkernels``, ``PyLB``) backed by the CUDA library."""
import os
import sys


def activate():
    """Make ``import _lbkernels`` / ``import PyLB`` resolve to the GPU drop-ins."""
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    return here
