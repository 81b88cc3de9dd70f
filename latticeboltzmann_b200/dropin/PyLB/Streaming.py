"""Drop-in for PyLB/Streaming.py: the velocity table ``c_ic`` (:28-29) and ``stream`` (:33-46)."""
import numpy as np

from latticeboltzmann_b200 import _lib

# "Velocities" of the nine channels, one (cx, cy) row per channel (PyLB/Streaming.py:28-29)
c_ic = np.array([(0, 0), (1, 0), (0, 1), (-1, 0), (0, -1), (1, 1), (-1, 1), (-1, -1), (1, -1)])


def stream(f_ikl):
    """Propagate channel occupations by one cell distance, periodically, IN PLACE:
    ``f[i] <- np.roll(f[i], c_ic[i], axis=(0, 1))`` for i = 1..8 (PyLB/Streaming.py:45-46).

    Pure data movement, so any 4- or 8-byte element type is moved bit for bit on the GPU."""
    if not isinstance(f_ikl, np.ndarray) or f_ikl.ndim != 3 or f_ikl.shape[0] != 9:
        raise TypeError("stream(): expected an ndarray of shape (9, nx, ny)")
    if f_ikl.dtype.itemsize not in (4, 8):
        raise TypeError("stream(): element size must be 4 or 8 bytes, got %s" % f_ikl.dtype)
    lib = _lib.load()
    fn = lib.lbk_stream_f64 if f_ikl.dtype.itemsize == 8 else lib.lbk_stream_f32
    _, nx, ny = f_ikl.shape
    if f_ikl.flags.c_contiguous and f_ikl.flags.writeable:
        _lib.check(fn(_lib.np_ptr(f_ikl), nx, ny))
    else:
        tmp = np.ascontiguousarray(f_ikl)
        _lib.check(fn(_lib.np_ptr(tmp), nx, ny))
        f_ikl[...] = tmp
