"""latticeboltzmann_b200 -- B200-native D2Q9 BGK lattice-Boltzmann time step.

One hot path (stream + boundaries + collide + halo exchange, fused into one
hand-written sm_100a kernel) behind the reference's PyLB / _lbkernels API.
See DESIGN.md.  There is no CPU fallback: without the CUDA library and a GPU
every compute entry point raises ``LbmError``.
"""
from ._lib import LbmError, load as load_native, library_path   # noqa: F401
from .decomposition import Decomposition                         # noqa: F401
from .lattice import Block, Lattice                              # noqa: F401

__all__ = ["LbmError", "load_native", "library_path", "Decomposition", "Block", "Lattice"]
