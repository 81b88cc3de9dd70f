"""Stand-in for mpi4py (not installed on the B200 image), sufficient for the reference's opt2 cavity
script and PyLB/IO.py to import and run on one rank or on a torchrun world.
It lives in its own directory (``latticeboltzmann_b200/mpi_shim``) which ``dropin.activate()`` puts on
sys.path ONLY when no real mpi4py is importable, so it can never shadow a real MPI installation."""
from . import MPI   # noqa: F401
