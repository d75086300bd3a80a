"""The reference's system constructors, run on the host by the product library (csrc/host_ctor.hpp through
sadmc_reference_system), against the oracle's restatement of the same constructors.  No GPU needed: these are
the images SADMC_INIT_REFERENCE replicates to every walker."""
import ctypes as C

import numpy as np
import pytest

from sad_monte_carlo_b200 import _abi
from sad_monte_carlo_b200._abi import make_config
from tests.oracle_lib import OracleMC

f64p = C.POINTER(C.c_double)


def _image(gpu_lib, cfg):
    need = C.c_size_t()
    assert gpu_lib.sadmc_reference_system(C.byref(cfg), None, 0, C.byref(need)) == 0, gpu_lib.sadmc_last_error()
    buf = np.zeros(need.value)
    rc = gpu_lib.sadmc_reference_system(C.byref(cfg), buf.ctypes.data_as(f64p), buf.size, None)
    assert rc == 0, gpu_lib.sadmc_last_error()
    return buf


@pytest.mark.parametrize("N,rho", [(20, 0.3), (40, 0.8), (64, 1.0)])
def test_wca_n_squared_attempts_constructor(gpu_lib, N, rho):
    # wca.rs:448-496: N*N random fillings, the lowest running energy wins; positions are raw RNG values -> identical
    cfg = make_config("wca", N=N, reduced_density=rho, seed=1)
    img = _image(gpu_lib, cfg)
    o = OracleMC(cfg).system()
    assert np.array_equal(img[:-2], o[:-2])
    assert abs(img[-2] - o[-2]) <= 1e-13 * max(1.0, abs(o[-2]))  # compute_energy: same list order on both sides
    assert img[-1] == pytest.approx(o[-1], rel=1e-9, abs=1e-300)  # the error budget left by the last N adds
    L = (N / rho) ** (1.0 / 3.0)
    assert (img[:-2] >= 0).all() and (img[:-2] < L).all()


def test_wca_constructor_does_not_depend_on_the_thread_count(gpu_lib, monkeypatch):
    cfg = make_config("wca", N=30, reduced_density=0.6, seed=1)
    a = _image(gpu_lib, cfg)
    monkeypatch.setenv("SADMC_HOST_THREADS", "1")
    b = _image(gpu_lib, cfg)
    monkeypatch.setenv("SADMC_HOST_THREADS", "7")
    c = _image(gpu_lib, cfg)
    assert np.array_equal(a, b) and np.array_equal(a, c)


@pytest.mark.parametrize("N,R", [(7, 2.0), (13, 2.5), (31, 2.5)])
def test_lj_constructor(gpu_lib, N, R):
    cfg = make_config("lj", N=N, lj_radius=R, seed=1)
    assert np.array_equal(_image(gpu_lib, cfg), OracleMC(cfg).system())


def test_ising_and_square_well_constructors(gpu_lib):
    cfg = make_config("ising", N=32, seed=1)
    assert np.array_equal(_image(gpu_lib, cfg), OracleMC(cfg).system())
    cfg = make_config("sw", N=50, filling_fraction=0.3, sw_well_width=1.3, seed=1)
    img, o = _image(gpu_lib, cfg), OracleMC(cfg).system()
    assert np.array_equal(img[:-2], o[:-2]) and np.isnan(img[-2])  # the device counts the wells when it loads the image


def test_box_smaller_than_the_cutoff_is_an_error_code(gpu_lib):
    cfg = make_config("wca", N=1, reduced_density=2.0, seed=1)  # L = 0.79 < 2^(1/6): wca.rs:186-191 panics
    assert gpu_lib.sadmc_reference_system(C.byref(cfg), None, 0, None) == _abi.ERR_INVALID
    assert b"not large enough" in gpu_lib.sadmc_last_error()


def test_analytic_systems_start_where_the_reference_puts_them(gpu_lib):
    # fake.rs:85-93 (the origin), erfinv.rs:50-58 (0.5 everywhere), two_wells.rs:248-250 (x0 = -0.99, d^2 cached)
    img = _image(gpu_lib, make_config("fake", fake_function=_abi.FAKE_QUADRATIC, N=4, seed=1))
    assert img.tolist() == [0.0] * 4
    img = _image(gpu_lib, make_config("fake-erfinv", N=3, erfinv_mean_energy=0.0, seed=1))
    assert img.tolist() == [0.5] * 3
    cfg = make_config("two-wells", N=12, tw_h2_to_h1=1.1, tw_barrier_over_h1=0.1, tw_r2=0.5, seed=1)
    img = _image(gpu_lib, cfg)
    assert img[0] == -0.99 and not img[1:12].any() and img[12] == (-0.99) ** 2
    assert np.array_equal(img, OracleMC(cfg).system())


@pytest.mark.parametrize("system,kw,msg", [
    ("two-wells", dict(N=10, tw_h2_to_h1=1.1, tw_barrier_over_h1=0.1, tw_r2=0.5), b"not divisible by three"),   # two_wells.rs:240-245
    ("ising", dict(N=1), b"N must be > 1"),                                                                         # ising.rs:39
    ("lj", dict(N=31, lj_radius=0.0), b"radius must be > 0"),
])
def test_constructor_panics_become_error_codes(gpu_lib, system, kw, msg):
    cfg = make_config(system, seed=1, **kw)
    assert gpu_lib.sadmc_reference_system(C.byref(cfg), None, 0, None) == _abi.ERR_INVALID
    assert msg in gpu_lib.sadmc_last_error()


def test_buffer_too_small_is_reported_not_overrun(gpu_lib):
    cfg = make_config("ising", N=8, seed=1)
    buf = np.full(10, 7.0)
    assert gpu_lib.sadmc_reference_system(C.byref(cfg), buf.ctypes.data_as(f64p), buf.size, None) == _abi.ERR_INVALID
    assert (buf == 7.0).all() and b"needs 65 doubles" in gpu_lib.sadmc_last_error()
