"""Drop-in ``PyLB`` package: the reference exports exactly ``collide``, ``equilibrium``
(PyLB/__init__.py:21) and ``stream`` (PyLB/__init__.py:22); all three run on the GPU here."""
from .Collision import collide, equilibrium   # noqa: F401
from .Streaming import stream                 # noqa: F401
