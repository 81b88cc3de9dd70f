"""CPU: the C-ABI library loads and exports every symbol include/lbm_b200.h declares;
no compute without a GPU (and no silent fallback)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "lbm_b200.h")).read()
    return sorted(set(re.findall(r"LB_API[^;(]*?\b(lbk?_\w+)\s*\(", src)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    import ctypes
    from latticeboltzmann_b200 import _lib
    from latticeboltzmann_b200.build import build_native
    so = build_native()
    lib = ctypes.CDLL(so)
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), n
    # the ctypes table binds exactly the declared ABI
    assert sorted(_lib.SYMBOLS) == names
    lib2 = _lib.load()
    assert lib2.lb_abi_version() == 1
    # the ctypes mirrors of the ABI structs have the C layout (the export struct carries the halo wiring)
    assert lib2.lb_sizeof_config() == ctypes.sizeof(_lib.LbConfig) == 96
    assert lib2.lb_sizeof_export() == ctypes.sizeof(_lib.LbExport)


def test_no_cpu_fallback_without_device():
    import latticeboltzmann_b200 as lb
    from latticeboltzmann_b200 import _lib
    lib = _lib.load()
    if lib.lb_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(lb.LbmError):
        lb.Lattice(8, 8)
    f = np.zeros((9, 4))
    assert lib.lbk_collide_f64(_lib.np_ptr(f), 4, 1.0) == -3      # LB_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.lb_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "latticeboltzmann_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                assert "oracle" not in txt.lower() or fn == "never", os.path.join(dp, fn)


def test_plain_c_client_compiles_links_and_fails_loudly_without_gpu(tmp_path):
    """examples/c_abi_cavity.c: the ABI is usable from plain C (no C++ / torch types in the header)."""
    import subprocess
    from latticeboltzmann_b200 import _lib
    from latticeboltzmann_b200.build import build_native
    so_dir = os.path.dirname(build_native())
    exe = str(tmp_path / "c_abi_cavity")
    subprocess.run(["gcc", "-std=c11", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "c_abi_cavity.c"), "-o", exe, "-L", so_dir, "-llbm_b200",
                    "-Wl,-rpath," + so_dir], check=True)
    r = subprocess.run([exe, "64", "48", "20"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if _lib.load().lb_device_count() > 0:
        assert r.returncode == 0 and "MLUPS" in r.stdout, r.stderr
    else:
        assert r.returncode == 3 and "no CPU fallback" in r.stderr
