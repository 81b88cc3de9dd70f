"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle.

ctypes bindings of ``oracle/_build/liboracle.so`` (the plain-C restatement of
the reference's opt2 arithmetic, ``d2q9_oracle_impl.h``) plus thin numpy
helpers.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module; the
product package ``latticeboltzmann_b200`` never does.

Parity status: pinned against golden vectors generated
by executing the reference's own Python functions (``tests/make_golden.py``);
the association order of Eigen's ``.sum()`` (c/d2q9.h:126) is restated from
Eigen 3.4.0's published algorithm because Eigen is not available here.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

# PyLB/Streaming.py:28-29
C_IC = np.array([[0, 1, 0, -1, 0, 1, -1, -1, 1],
                 [0, 0, 1, 0, -1, 1, 1, -1, -1]]).T
OPPOSITE = np.array([0, 3, 4, 1, 2, 7, 8, 5, 6])


def build(force=False):
    """Compile the oracle with the Makefile next to this file."""
    src = [os.path.join(_HERE, n) for n in ("d2q9_oracle.c", "d2q9_oracle_impl.h", "Makefile")]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in src)):
        return _SO
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _declare(_lib)
    return _lib


_i64 = ctypes.c_int64
_int = ctypes.c_int
_vp = ctypes.c_void_p


def _declare(L):
    for suf, ct in (("f32", ctypes.c_float), ("f64", ctypes.c_double)):
        def fn(name, *args):
            f = getattr(L, "%s_%s" % (name, suf))
            f.restype = None
            f.argtypes = list(args)
        fn("orc_equilibrium1", ct, ct, ct, _vp)
        fn("orc_equilibriumn", _vp, _vp, _vp, _vp, _i64)
        fn("orc_colliden", _vp, _i64, ct)
        fn("orc_colliden_range", _vp, _i64, ct, _i64, _i64)
        fn("orc_stream", _vp, _i64, _i64, _vp)
        fn("orc_cavity_stream_and_bounce_back", _vp, _i64, _i64, ct, _int, _vp)
        fn("orc_cavity_run", _vp, _i64, _i64, ct, ct, _int, _i64, _vp)
        fn("orc_periodic_run", _vp, _i64, _i64, ct, _i64, _vp, _vp, _vp)
        fn("orc_cavity_step_pull_rows", _vp, _vp, _i64, _i64, ct, ct, _int, _int, _i64, _i64)
        fn("orc_cavity_step_pull_rows_hoisted", _vp, _vp, _i64, _i64, ct, ct, _int, _int, _i64, _i64)
        fn("orc_periodic_step_pull_rows", _vp, _vp, _i64, _i64, ct, _int, _i64, _i64)
        fn("orc_moments", _vp, _i64, _vp, _vp, _vp)


def _suf(a):
    if a.dtype == np.float64:
        return "f64"
    if a.dtype == np.float32:
        return "f32"
    raise TypeError("oracle supports float32/float64, got %s" % a.dtype)


def _p(a):
    assert a.flags.c_contiguous
    return a.ctypes.data_as(_vp)


def _call(name, ref, *args):
    getattr(lib(), "%s_%s" % (name, _suf(ref)))(*args)


def equilibrium1(rho, ux, uy, dtype=np.float64):
    """c/d2q9.h:59-81"""
    out = np.empty(9, dtype=dtype)
    _call("orc_equilibrium1", out, dtype(rho), dtype(ux), dtype(uy), _p(out))
    return out


def equilibrium(rho, ux, uy, f):
    """c/d2q9.h:98-108 -- fills f (9, N) in place."""
    assert f.shape[0] == 9 and f.ndim == 2
    _call("orc_equilibriumn", f, _p(rho), _p(ux), _p(uy), _p(f), f.shape[1])


def collide(f, omega):
    """c/d2q9.h:121-131 -- in place on f (9, N) (any trailing shape)."""
    assert f.shape[0] == 9
    _call("orc_colliden", f, _p(f), f.size // 9, f.dtype.type(omega))


def stream(f):
    """PyLB/Streaming.py:33-46 -- in place periodic roll on f (9, nx, ny)."""
    _, nx, ny = f.shape
    tmp = np.empty(nx * ny, dtype=f.dtype)
    _call("orc_stream", f, _p(f), nx, ny, _p(tmp))


def _scratch(f):
    _, nx, ny = f.shape
    return np.empty(nx * ny + 18 * nx + 18 * ny, dtype=f.dtype)


def cavity_stream_and_bounce_back(f, u0=0.1, walls_lr=True):
    """cavity_opt2.py:109-177 -- in place."""
    _, nx, ny = f.shape
    s = _scratch(f)
    _call("orc_cavity_stream_and_bounce_back", f, _p(f), nx, ny, f.dtype.type(u0), int(walls_lr), _p(s))


def cavity_run(f, omega, nsteps, u0=0.1, walls_lr=True):
    """cavity_opt2.py:272-277 on one rank -- in place."""
    _, nx, ny = f.shape
    s = _scratch(f)
    _call("orc_cavity_run", f, _p(f), nx, ny, f.dtype.type(omega), f.dtype.type(u0), int(walls_lr),
          int(nsteps), _p(s))


def periodic_run(f, omega, nsteps, uy_k=None):
    """shear_wave_opt2.py:95-99 -- in place; returns the amplitude series if uy_k is given."""
    _, nx, ny = f.shape
    s = _scratch(f)
    ampl = None
    if uy_k is not None:
        uy_k = np.ascontiguousarray(uy_k, dtype=f.dtype)
        ampl = np.zeros(nsteps, dtype=f.dtype)
    _call("orc_periodic_run", f, _p(f), nx, ny, f.dtype.type(omega), int(nsteps),
          _p(uy_k) if uy_k is not None else None, _p(ampl) if ampl is not None else None, _p(s))
    return ampl


def cavity_step_pull(src, dst, omega, u0=0.1, walls_lr=True, do_collide=True, k_lo=0, k_hi=None):
    """Fused pull formulation; rows [k_lo, k_hi) of dst."""
    _, nx, ny = src.shape
    k_hi = nx if k_hi is None else k_hi
    _call("orc_cavity_step_pull_rows", src, _p(src), _p(dst), nx, ny, src.dtype.type(omega),
          src.dtype.type(u0), int(walls_lr), int(do_collide), int(k_lo), int(k_hi))


def cavity_step_pull_hoisted(src, dst, omega, u0=0.1, walls_lr=True, do_collide=True, k_lo=0, k_hi=None):
    """cavity_step_pull with the index arithmetic hoisted out of the column loop (same results, ~3x faster)."""
    _, nx, ny = src.shape
    k_hi = nx if k_hi is None else k_hi
    _call("orc_cavity_step_pull_rows_hoisted", src, _p(src), _p(dst), nx, ny, src.dtype.type(omega),
          src.dtype.type(u0), int(walls_lr), int(do_collide), int(k_lo), int(k_hi))


def cavity_run_threaded(f, omega, nsteps, u0=0.1, walls_lr=True, threads=None):
    """nsteps cavity steps with the pull form, rows split over `threads` host threads (ctypes releases the GIL;
    every cell of a step reads only the previous step's array, so the split cannot change a bit).  Returns the
    final array (f or the scratch copy, whichever holds it)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    _, nx, ny = f.shape
    threads = threads or max(1, min(len(os.sched_getaffinity(0)), 64, nx))
    edges = [nx * i // threads for i in range(threads + 1)]
    a, b = f, np.empty_like(f)
    with ThreadPoolExecutor(threads) as ex:
        for _ in range(nsteps):
            list(ex.map(lambda i: cavity_step_pull_hoisted(a, b, omega, u0, walls_lr, True, edges[i], edges[i + 1]), range(threads)))
            a, b = b, a
    return a


def periodic_step_pull(src, dst, omega, do_collide=True, k_lo=0, k_hi=None):
    _, nx, ny = src.shape
    k_hi = nx if k_hi is None else k_hi
    _call("orc_periodic_step_pull_rows", src, _p(src), _p(dst), nx, ny, src.dtype.type(omega),
          int(do_collide), int(k_lo), int(k_hi))


def moments(f):
    """cavity_opt2.py:280-281 -- returns rho, ux, uy with f's trailing shape."""
    shape = f.shape[1:]
    n = f.size // 9
    rho = np.empty(n, dtype=f.dtype)
    ux = np.empty(n, dtype=f.dtype)
    uy = np.empty(n, dtype=f.dtype)
    _call("orc_moments", f, _p(f), n, _p(rho), _p(ux), _p(uy))
    return rho.reshape(shape), ux.reshape(shape), uy.reshape(shape)


def init_equilibrium(nx, ny, dtype=np.float64, rho=None, ux=None, uy=None):
    """cavity_opt2.py:265-269 / shear_wave_opt2.py:84-88: f = feq(rho, ux, uy)."""
    dtype = np.dtype(dtype)
    rho = np.ones((nx, ny), dtype) if rho is None else np.ascontiguousarray(np.broadcast_to(rho, (nx, ny)), dtype)
    ux = np.zeros((nx, ny), dtype) if ux is None else np.ascontiguousarray(np.broadcast_to(ux, (nx, ny)), dtype)
    uy = np.zeros((nx, ny), dtype) if uy is None else np.ascontiguousarray(np.broadcast_to(uy, (nx, ny)), dtype)
    f = np.zeros((9, nx, ny), dtype)
    equilibrium(rho.reshape(-1), ux.reshape(-1), uy.reshape(-1), f.reshape(9, -1))
    return f


def shear_wave_init(nx, ny, dtype=np.float64, a0=1.0):
    """shear_wave_opt2.py:79-88: rho=1, ux=0, uy(k,l) = a0*sin(2*pi*k/nx)."""
    dtype = np.dtype(dtype)
    x_k = np.arange(nx)
    uy_k = (a0 * np.sin(2 * np.pi / nx * x_k)).astype(dtype)
    f = init_equilibrium(nx, ny, dtype, uy=np.resize(uy_k, (ny, nx)).T)
    return f, uy_k


def perturbed_state(nx, ny, dtype=np.float64, seed=0, amp=0.01):
    """SURVEY.md §8d value distribution for kernel unit tests:
    f = feq(1,0,0) * (1 + amp*N(0,1)), numpy default_rng(seed)."""
    rng = np.random.default_rng(seed)
    f = init_equilibrium(nx, ny, np.float64)
    f *= 1.0 + amp * rng.standard_normal(f.shape)
    return np.ascontiguousarray(f.astype(dtype))
