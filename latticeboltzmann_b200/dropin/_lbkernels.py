"""Drop-in for the reference's compiled extension ``_lbkernels`` (c/_lbkernels.cpp:33-49).

Put ``latticeboltzmann_b200/dropin`` on ``sys.path`` (or call
``latticeboltzmann_b200.dropin.activate()``) and ``import _lbkernels`` / ``import PyLB``
resolve here; every call runs on the GPU through the C ABI (``lbk_*`` in
include/lbm_b200.h).  Overloads, in the reference's registration order:

    equilibrium(rho: float, ux: float, uy: float) -> ndarray[float32, (9,)]
    equilibrium(rho: f32[N], ux: f32[N], uy: f32[N], f: f32[9, N]) -> None
    collide(f: f32[9, N], omega: float) -> None
    ... and the same three for float64.

pybind11 semantics that callers rely on are kept: array arguments are
``Eigen::Ref`` (non-const) in the reference (c/d2q9.h:99-102,122), so nothing is
ever converted or copied -- dtype must match exactly, arrays must be writeable,
results are written IN PLACE into the caller's buffer (callers pass
``f.reshape(9, -1)`` views); anything else is a ``TypeError``.  Because the
``float`` scalar overload is registered first, ``equilibrium(1.0, 0.1, 0.0)``
returns float32, as in the reference.  ``omega`` may be a Python float or a 0-d
numpy array (cavity_opt2.py:66).
"""
import numpy as np

from latticeboltzmann_b200 import _lib

__doc_module__ = "Lattice Boltzmann kernels"     # m.doc(), c/_lbkernels.cpp:34


def _fail(name, args):
    raise TypeError("%s(): incompatible function arguments. The following argument types are supported:\n%s\n\n"
                    "Invoked with: %s" % (name, _SIGS[name], ", ".join(repr(type(a).__name__) for a in args)))


_SIGS = {
    "equilibrium": "    1. (arg0: float, arg1: float, arg2: float) -> numpy.ndarray[float32[9, 1]]\n"
                   "    2. (arg0: numpy.ndarray[float32[m, 1], flags.writeable], arg1: ..., arg2: ..., "
                   "arg3: numpy.ndarray[float32[9, n], flags.writeable, flags.c_contiguous]) -> None\n"
                   "    3. (arg0: float, arg1: float, arg2: float) -> numpy.ndarray[float64[9, 1]]\n"
                   "    4. (arg0: numpy.ndarray[float64[m, 1], flags.writeable], arg1: ..., arg2: ..., "
                   "arg3: numpy.ndarray[float64[9, n], flags.writeable, flags.c_contiguous]) -> None",
    "collide": "    1. (arg0: numpy.ndarray[float32[9, n], flags.writeable, flags.c_contiguous], arg1: float) -> None\n"
               "    2. (arg0: numpy.ndarray[float64[9, n], flags.writeable, flags.c_contiguous], arg1: float) -> None",
}


def _as_scalar(x):
    """pybind11's float caster with implicit conversion: Python numbers, numpy scalars, 0-d arrays."""
    if isinstance(x, (bool, np.bool_)):
        return float(x)
    if isinstance(x, (int, float, np.integer, np.floating)):
        return float(x)
    if isinstance(x, np.ndarray) and x.ndim == 0:
        return float(x)
    return None


def _vec(a, dtype):
    """Eigen::Ref<Array<T, Dynamic, 1>>: exact dtype, writeable, 1-D (or N x 1) with unit inner stride."""
    if not isinstance(a, np.ndarray) or a.dtype != dtype or not a.flags.writeable:
        return None
    if a.ndim == 2 and a.shape[1] == 1:
        a = a[:, 0]
    if a.ndim != 1 or (a.size > 1 and a.strides[0] != a.itemsize):
        return None
    return a


def _field(a, dtype):
    """Eigen::Ref<Array<T, 9, Dynamic, RowMajor>>: exact dtype, writeable, (9, N), rows with unit inner stride."""
    if not isinstance(a, np.ndarray) or a.dtype != dtype or not a.flags.writeable:
        return None
    if a.ndim != 2 or a.shape[0] != 9 or (a.shape[1] > 1 and a.strides[1] != a.itemsize):
        return None
    return a


def _run_in_place(f, fn):
    """Call fn on a C-contiguous (9, N) buffer and write the result through to `f` (which may
    have a non-trivial outer stride, e.g. a column slice)."""
    if f.flags.c_contiguous:
        fn(f)
    else:
        tmp = np.ascontiguousarray(f)
        fn(tmp)
        f[...] = tmp


def equilibrium(*args):
    """Return the equilibrium distribution function (c/d2q9.h:59-81, 98-108)."""
    lib = _lib.load()
    if len(args) == 3:
        s = [_as_scalar(a) for a in args]
        if None in s:
            _fail("equilibrium", args)
        out = np.empty(9, np.float32)            # the float overload is registered first (c/_lbkernels.cpp:36)
        _lib.check(lib.lbk_equilibrium1_f32(s[0], s[1], s[2], _lib.np_ptr(out)))
        return out
    if len(args) == 4:
        for dtype, suf in ((np.float32, "f32"), (np.float64, "f64")):
            rho, ux, uy = (_vec(a, dtype) for a in args[:3])
            f = _field(args[3], dtype)
            if rho is None or ux is None or uy is None or f is None:
                continue
            n = f.shape[1]
            if min(rho.size, ux.size, uy.size) < n:
                # the reference reads rho(kl) for kl < f.cols() unchecked (c/d2q9.h:104-107): undefined
                # behaviour there, an error here
                raise TypeError("equilibrium(): rho/ux/uy are shorter than f.shape[1]")
            fn = getattr(lib, "lbk_equilibriumn_" + suf)
            _run_in_place(f, lambda buf: _lib.check(fn(_lib.np_ptr(rho), _lib.np_ptr(ux), _lib.np_ptr(uy),
                                                       _lib.np_ptr(buf), n)))
            return None
    _fail("equilibrium", args)


def collide(*args):
    """Carry out collision operation for an array of values (c/d2q9.h:121-131), in place."""
    lib = _lib.load()
    if len(args) == 2:
        omega = _as_scalar(args[1])
        if omega is not None:
            for dtype, suf in ((np.float32, "f32"), (np.float64, "f64")):
                f = _field(args[0], dtype)
                if f is None:
                    continue
                fn = getattr(lib, "lbk_collide_" + suf)
                _run_in_place(f, lambda buf: _lib.check(fn(_lib.np_ptr(buf), f.shape[1], omega)))
                return None
    _fail("collide", args)
