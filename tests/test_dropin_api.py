"""The reference-named drop-in modules (_lbkernels, PyLB).  The GPU tests restate the
reference's own tests (tests/01-ImportTest.py:44-45, tests/02-CollideTest.py:94-111):
same shapes, same unseeded random inputs, same tolerance (1e-7, tests/PyLBTest.py:75),
numpy formulas as the expected values."""
import numpy as np
import pytest

from latticeboltzmann_b200 import dropin

dropin.activate()

C = np.array([(0, 0), (1, 0), (0, 1), (-1, 0), (0, -1), (1, 1), (-1, 1), (-1, -1), (1, -1)])
W = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)


def feq_numpy(rho, u):
    """w_i rho (1 + 3 c.u + 9/2 (c.u)^2 - 3/2 u.u), the textbook form 02-CollideTest.py checks against."""
    cu = np.einsum("ic,c...->i...", C, u)
    uu = np.einsum("c...,c...->...", u, u)
    return W.reshape((9,) + (1,) * rho.ndim) * rho * (1 + 3 * cu + 4.5 * cu ** 2 - 1.5 * uu)


def collide_numpy(f, omega):
    rho = f.sum(axis=0)
    u = np.einsum("ic,i...->c...", C, f) / rho
    f += omega * (feq_numpy(rho, u) - f)


def test_import_and_exports():
    """tests/01-ImportTest.py: `import PyLB` works and exposes the three compute symbols."""
    import PyLB
    import _lbkernels
    assert callable(PyLB.collide) and callable(PyLB.equilibrium) and callable(PyLB.stream)
    assert PyLB.collide is _lbkernels.collide and PyLB.equilibrium is _lbkernels.equilibrium
    from PyLB.Streaming import c_ic
    assert np.array_equal(c_ic, C)


def test_argument_contract_raises_typeerror_like_pybind11():
    """Non-const Eigen::Ref never converts: wrong dtype / read-only / wrong shape -> TypeError (no GPU needed)."""
    import PyLB
    f = np.zeros((9, 4))
    ro = np.zeros((9, 4))
    ro.flags.writeable = False
    bad = [(np.zeros((9, 4), np.int32), 1.0), (np.zeros((8, 4)), 1.0), (np.zeros((9, 2, 2)), 1.0),
           (ro, 1.0), (np.zeros((4, 9)).T, 1.0), (f, "x"), ([[0.0] * 4] * 9, 1.0)]
    for args in bad:
        with pytest.raises(TypeError):
            PyLB.collide(*args)
    with pytest.raises(TypeError):
        PyLB.collide(f)
    r = np.ones(4)
    with pytest.raises(TypeError):       # mixed precision: no overload matches
        PyLB.equilibrium(r.astype(np.float32), r, r, f)
    with pytest.raises(TypeError):
        PyLB.equilibrium(r, r, r, np.zeros((9, 4), np.float32))
    with pytest.raises(TypeError):
        PyLB.equilibrium(r[::2], r[::2], r[::2], np.zeros((9, 2)))   # strided vectors are not Ref-compatible
    with pytest.raises(TypeError):
        PyLB.equilibrium(r, r)
    with pytest.raises(TypeError):
        PyLB.stream(np.zeros((8, 3, 3)))
    with pytest.raises(TypeError):
        PyLB.stream(np.zeros((9, 3, 3), np.int16))


@pytest.mark.gpu
def test_D2Q9_equilibrium():
    """tests/02-CollideTest.py:94-102."""
    from PyLB import equilibrium
    rho_kl = np.abs(np.random.random([2, 2]))
    ux_kl = np.random.random(rho_kl.shape)
    uy_kl = np.random.random(rho_kl.shape)
    e1 = np.zeros([9] + list(rho_kl.shape))
    assert equilibrium(rho_kl.reshape(-1), ux_kl.reshape(-1), uy_kl.reshape(-1), e1.reshape(9, -1)) is None
    e2 = feq_numpy(rho_kl, np.array([ux_kl, uy_kl]))
    assert not np.isnan(e1).any()
    assert np.abs(e1 - e2).max() < 1e-7


@pytest.mark.gpu
def test_D2Q9_collide():
    """tests/02-CollideTest.py:104-111: in place on a reshape(9, -1) VIEW of the caller's array."""
    from PyLB import collide
    f_ikl = np.abs(np.random.random([9, 4, 4]))
    for omega in [0.5, 1.7]:
        c1 = f_ikl.copy()
        c2 = f_ikl.copy()
        assert collide(c1.reshape(9, -1), omega) is None
        collide_numpy(c2, omega)
        assert np.abs(c1 - c2).max() < 1e-7
        assert np.abs(c1 - f_ikl).max() > 0         # written through the view


@pytest.mark.gpu
def test_overload_quirks_and_omega_forms():
    from PyLB import collide, equilibrium
    e = equilibrium(1.0, 0.1, 0.0)
    assert e.dtype == np.float32 and e.shape == (9,)          # float overload registered first
    assert abs(float(e.sum()) - 1.0) < 1e-6
    f32 = np.abs(np.random.random([9, 6])).astype(np.float32)
    ref = f32.astype(np.float64)
    collide(f32, np.array(1.7, dtype=np.float32))             # 0-d array omega, cavity_opt2.py:66
    collide_numpy(ref, float(np.float32(1.7)))
    assert np.abs(f32 - ref).max() < 1e-5
    wide = np.abs(np.random.random([9, 10]))
    view = wide[:, :5]                                        # outer stride != N: still a valid Ref
    expect = view.copy()
    collide(view, 0.5)
    collide_numpy(expect, 0.5)
    assert np.abs(view - expect).max() < 1e-12
    assert np.array_equal(wide[:, 5:], wide[:, 5:])


@pytest.mark.gpu
def test_stream_is_np_roll_for_any_8_byte_type():
    """shear_wave_opt2.py:80 builds f with np.arange (then overwrites it); stream must move any values."""
    from PyLB import stream
    from PyLB.Streaming import c_ic
    for dt in (np.float64, np.int64, np.float32, np.int32):
        f = np.arange(9 * 5 * 7, dtype=dt).reshape(9, 5, 7)
        ref = f.copy()
        for i in range(1, 9):
            ref[i] = np.roll(ref[i], c_ic[i], axis=(0, 1))
        stream(f)
        assert np.array_equal(f, ref), dt


@pytest.mark.gpu
def test_shear_wave_opt2_script_body():
    """The body of simulators/serial_shear_wave/Python/shear_wave_opt2.py:79-99 (60 steps, 30x20),
    run against the drop-in PyLB exactly as the script calls it, vs the oracle."""
    import PyLB as D2Q9
    from oracle import oracle as orc
    nx, ny, nsteps, omega, dtype = 30, 20, 60, 0.3, np.float64
    c_ic = D2Q9.Streaming.c_ic
    x_k = np.arange(nx)
    uy_k = np.sin(2 * np.pi / nx * x_k, dtype=dtype)
    f_ikl = np.arange(9 * nx * ny, dtype=dtype).reshape(9, nx, ny)
    D2Q9.equilibrium(np.ones((nx, ny), dtype=dtype).reshape(-1), np.zeros((nx, ny), dtype=dtype).reshape(-1),
                     np.resize(uy_k, (ny, nx)).T.reshape(-1),     # == np.resize(uy_k, (nx, ny)).T for the script's nx == ny
                     f_ikl.reshape(9, -1))
    ref, _ = orc.shear_wave_init(nx, ny)
    assert np.array_equal(f_ikl, ref)
    ampl = []
    for _ in range(nsteps):
        D2Q9.stream(f_ikl)
        D2Q9.collide(f_ikl.reshape(9, -1), omega)
        ampl += [((c_ic[:, 1].dot(f_ikl[:, :, ny // 2]) / (f_ikl[:, :, ny // 2].sum(axis=0))) * uy_k).sum() * 2 / nx]
    ref_ampl = orc.periodic_run(ref, omega, nsteps, uy_k)
    assert np.array_equal(f_ikl, ref)
    assert np.abs(np.array(ampl) - ref_ampl).max() < 1e-13
