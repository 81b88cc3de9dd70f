"""Minimal single-process stand-in for mpi4py (which is not installed on the B200 image), sufficient
for the reference's opt2 cavity script and PyLB/IO.py to import and run on ONE rank
(SURVEY.md H7 / N3).  Only used when the real mpi4py is absent: put
``latticeboltzmann_b200/dropin`` on sys.path AFTER site-packages if you have the real one."""
from . import MPI   # noqa: F401
