"""PyLB/Collision.py:21 re-exports the compiled kernels; so does this module (GPU versions)."""
from _lbkernels import equilibrium, collide   # noqa: F401
