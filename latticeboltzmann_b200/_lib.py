"""ctypes binding of liblbm_b200.so (include/lbm_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device
is visible, every compute entry point raises.  The library is built in-tree
(latticeboltzmann_b200/build.py) so the driver sees which .so is loaded.
"""
import ctypes
import os

import numpy as np

from . import build as _build

c_i64 = ctypes.c_int64
c_vp = ctypes.c_void_p

LB_F32, LB_F64 = 0, 1
BOUNDARY = {"periodic": 0, "cavity": 1, "cavity_xperiodic": 2,
            "sf_couette": 3, "sf_poiseuille": 4, "sf_sliding_lid": 5, "sf_table": 6}
ARITH = {"exact": 0, "fast": 1}
NUM_DIRS = 8
# (dx, dy) of direction slot d -- include/lbm_b200.h
DIRS = [(-1, 0), (1, 0), (0, -1), (0, 1), (-1, -1), (-1, 1), (1, -1), (1, 1)]


class LbConfig(ctypes.Structure):
    _fields_ = [("device", ctypes.c_int32), ("dtype", ctypes.c_int32), ("boundary", ctypes.c_int32),
                ("arith", ctypes.c_int32), ("gnx", c_i64), ("gny", c_i64), ("x0", c_i64), ("y0", c_i64),
                ("lnx", c_i64), ("lny", c_i64), ("omega", ctypes.c_double), ("u_wall", ctypes.c_double),
                ("rho_in", ctypes.c_double), ("rho_out", ctypes.c_double)]


class LbExport(ctypes.Structure):
    _fields_ = [("ipc_mem_handle", ctypes.c_uint8 * 64), ("local_base", ctypes.c_uint64), ("pid", c_i64),
                ("device", ctypes.c_int32), ("dtype", ctypes.c_int32), ("lnx", c_i64), ("lny", c_i64),
                ("pitch", c_i64), ("pop_stride", c_i64), ("buf_bytes", c_i64), ("ycol_offset", c_i64),
                ("ycol_bytes", c_i64), ("frame_offset", c_i64), ("state_offset", c_i64),
                ("total_bytes", c_i64)]


class LbmError(RuntimeError):
    pass


# Every symbol include/lbm_b200.h declares: name -> (restype, argtypes)
_P = ctypes.POINTER
SYMBOLS = {
    "lb_last_error": (ctypes.c_char_p, []),
    "lb_abi_version": (ctypes.c_int, []),
    "lb_sizeof_config": (c_i64, []),
    "lb_sizeof_export": (c_i64, []),
    "lb_device_count": (ctypes.c_int, []),
    "lb_create": (ctypes.c_int, [_P(LbConfig), _P(c_vp)]),
    "lb_create_ex": (ctypes.c_int, [_P(LbConfig), ctypes.c_int, _P(c_vp)]),
    "lb_destroy": (ctypes.c_int, [c_vp]),
    "lb_set_stream": (ctypes.c_int, [c_vp, c_vp]),
    "lb_get_stream": (c_vp, [c_vp]),
    "lb_sync": (ctypes.c_int, [c_vp]),
    "lb_get_export": (ctypes.c_int, [c_vp, _P(LbExport)]),
    "lb_connect": (ctypes.c_int, [c_vp, ctypes.c_int, _P(LbExport)]),
    "lb_halo_refresh": (ctypes.c_int, [c_vp]),
    "lb_upload_f": (ctypes.c_int, [c_vp, c_vp]),
    "lb_download_f": (ctypes.c_int, [c_vp, c_vp]),
    "lb_download_rows": (ctypes.c_int, [c_vp, c_i64, c_i64, c_vp]),
    "lb_upload_rows": (ctypes.c_int, [c_vp, c_i64, c_i64, c_vp]),
    "lb_checksum": (ctypes.c_int, [c_vp, _P(ctypes.c_uint64)]),
    "lb_init_equilibrium": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "lb_step": (ctypes.c_int, [c_vp, c_i64]),
    "lb_stream_only": (ctypes.c_int, [c_vp, c_i64]),
    "lb_step_host": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_int]),
    "lb_step_host_begin": (ctypes.c_int, [c_vp, c_vp]),
    "lb_step_timed": (ctypes.c_int, [c_vp, c_i64, _P(ctypes.c_float)]),
    "lb_steps_done": (c_i64, [c_vp]),
    "lb_health": (ctypes.c_int, [c_vp]),
    "lb_moments": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "lb_probe_shear_enable": (ctypes.c_int, [c_vp, c_i64, c_vp, c_i64]),
    "lb_probe_shear_read": (ctypes.c_int, [c_vp, c_vp, c_i64]),
    "lb_set_rows_per_tile": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "lb_set_halo_timeout_ms": (ctypes.c_int, [c_vp, c_i64]),
    "lb_set_boundary_table": (ctypes.c_int, [c_vp, c_i64, c_vp, c_vp, c_vp]),
    "lb_set_resident": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "lb_set_use_graph": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "lb_set_temporal": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int]),
    "lb_double_step_phase": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "lb_temporal_active": (ctypes.c_int, [c_vp]),
    "lb_temporal_rows": (ctypes.c_int, [c_vp]),
    "lb_pitch": (c_i64, [c_vp]),
    "lb_pop_stride": (c_i64, [c_vp]),
    "lb_kernel_launches": (ctypes.c_int, [c_vp, _P(c_i64)]),
    "lbk_equilibrium1_f32": (ctypes.c_int, [ctypes.c_float] * 3 + [c_vp]),
    "lbk_equilibrium1_f64": (ctypes.c_int, [ctypes.c_double] * 3 + [c_vp]),
    "lbk_equilibriumn_f32": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64]),
    "lbk_equilibriumn_f64": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64]),
    "lbk_collide_f32": (ctypes.c_int, [c_vp, c_i64, ctypes.c_float]),
    "lbk_collide_f64": (ctypes.c_int, [c_vp, c_i64, ctypes.c_double]),
    "lbk_stream_f32": (ctypes.c_int, [c_vp, c_i64, c_i64]),
    "lbk_stream_f64": (ctypes.c_int, [c_vp, c_i64, c_i64]),
    "lbk_selftest_div_const_f64": (ctypes.c_int, [c_vp, c_i64, _P(c_i64)]),
    "lbk_selftest_div_const_f32": (ctypes.c_int, [c_vp, c_i64, _P(c_i64)]),
    "lbk_step_host_f32": (ctypes.c_int, [c_vp, c_i64, c_i64, ctypes.c_int, ctypes.c_float, ctypes.c_float, c_i64]),
    "lbk_step_host_f64": (ctypes.c_int, [c_vp, c_i64, c_i64, ctypes.c_int, ctypes.c_double, ctypes.c_double, c_i64]),
}

_lib = None


def library_path():
    return _build.SO


def load(build_if_missing=True):
    """Load liblbm_b200.so; raises LbmError (never falls back) if it cannot be loaded."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("LBM_NATIVE_LIB") or _build.SO      # developer override (A/B kernel experiments)
    if not os.path.exists(path):
        if not build_if_missing:
            raise LbmError("native library %s is missing (run `python -m latticeboltzmann_b200.build`)" % path)
        _build.build_native()
    try:
        lib = ctypes.CDLL(path)
    except OSError as e:
        raise LbmError("cannot load native library %s: %s" % (path, e))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    if lib.lb_sizeof_config() != ctypes.sizeof(LbConfig) or lib.lb_sizeof_export() != ctypes.sizeof(LbExport):
        raise LbmError("ctypes declarations of lb_config / lb_export do not match %s (rebuild the library)" % path)
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().lb_last_error()
        raise LbmError("liblbm_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def require_device():
    n = load().lb_device_count()
    if n <= 0:
        raise LbmError("no CUDA device visible: latticeboltzmann_b200 has no CPU fallback")
    return n


def np_ptr(a):
    return a.ctypes.data_as(c_vp)


def dtype_code(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return LB_F64
    if dtype == np.float32:
        return LB_F32
    raise TypeError("dtype must be float32 or float64, got %s" % dtype)
