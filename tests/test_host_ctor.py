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
