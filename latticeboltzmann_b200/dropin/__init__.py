"""Reference-named modules (``_lbkernels``, ``PyLB``) backed by the CUDA library."""
import importlib.util
import os
import sys


def activate(mpi_shim="auto"):
    """Make ``import _lbkernels`` / ``import PyLB`` resolve to the GPU drop-ins (first on sys.path).

    ``mpi_shim``: "auto" (default) appends the ``mpi4py`` stand-in (``latticeboltzmann_b200/mpi_shim``) to
    sys.path only if no real mpi4py can be imported -- a real installation is never shadowed (under
    ``mpirun`` the stand-in would report one rank per process and every rank would write whole files);
    True forces the stand-in in front, False never adds it."""
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    shim = os.path.join(os.path.dirname(here), "mpi_shim")
    if mpi_shim is True:
        if shim in sys.path:
            sys.path.remove(shim)
        sys.path.insert(0, shim)
    elif mpi_shim == "auto" and shim not in sys.path:
        try:
            real = importlib.util.find_spec("mpi4py")
        except (ImportError, ValueError):
            real = None
        if real is None:
            sys.path.append(shim)
    return here
