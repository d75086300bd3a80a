"""Resume == continuous run (the reference's tests/resume-sad.rs), through the C ABI: every walker of an engine is
read back (sadmc_get_walker / get_bins / get_system / get_rngs), restored into a FRESH engine created with
SADMC_INIT_EXTERNAL (sadmc_set_systems / set_rngs / set_walker_bins / sadmc_resume), and the continuation must be
bit-identical to an uninterrupted run -- scalars, every per-bin vector, RNG state and configuration."""
import numpy as np
import pytest

from sad_monte_carlo_b200 import WalkerEngine, make_config, _abi
from tests.gpu_common import BINS_EXACT, clone_config, method_fields

pytestmark = pytest.mark.gpu


def snapshot(eng):
    return {"moves": eng.num_moves(), "systems": eng.systems(), "rngs": eng.rngs(),
            "walkers": [eng.walker(w) for w in range(eng.n_walkers)], "bins": [eng.bins(w) for w in range(eng.n_walkers)]}


def restore(cfg, snap):
    eng = WalkerEngine(clone_config(cfg, init_mode=_abi.INIT_EXTERNAL))
    eng.set_systems(snap["systems"])
    eng.set_rngs(snap["rngs"])
    for w in range(eng.n_walkers):
        eng.set_walker_bins(w, snap["walkers"][w], snap["bins"][w])
    eng.resume(snap["moves"])
    return eng


def assert_engines_identical(a, b, context):
    assert a.num_moves() == b.num_moves()
    assert np.array_equal(a.rngs(), b.rngs()), context
    assert np.array_equal(a.systems(), b.systems()), context
    for w in range(a.n_walkers):
        wa, wb = a.walker(w), b.walker(w)
        assert wa.status == 0 and wb.status == 0
        for f in method_fields(wa.method) + ["energy"]:
            assert getattr(wa, f) == getattr(wb, f), "%s walker %d: %s %r vs %r" % (context, w, f, getattr(wa, f), getattr(wb, f))
        ba, bb = a.bins(w), b.bins(w)
        for k in BINS_EXACT:
            assert np.array_equal(ba[k], bb[k]), "%s walker %d: bins.%s" % (context, w, k)


CASES = {
    "ising_sad": (dict(system="ising", method="sad", N=16, sad_min_T=1.0, n_walkers=9, seed=4), 7000, 9000),
    "ising_wl": (dict(system="ising", method="wl", N=8, wl_min_gamma=1e-3, min_allowed_energy=-128.0, max_allowed_energy=50.0,
                      n_walkers=5, seed=3), 30000, 30000),
    "ising_inv_t_wl": (dict(system="ising", method="inv-t-wl", N=8, min_allowed_energy=-128.0, max_allowed_energy=50.0,
                            n_walkers=5, seed=2), 40000, 40000),
    "sw_sad": (dict(system="sw", method="sad", N=64, filling_fraction=0.25, sad_min_T=0.5, n_walkers=4, seed=1), 6000, 6000),  # resume-sad.rs
    "lj31_sad_exact": (dict(system="lj", method="sad", N=31, lj_radius=2.5, max_allowed_energy=0.0, sad_min_T=0.01, energy_bin=0.01,
                            n_walkers=40, init_mode=_abi.INIT_RANDOMIZE, lanes_per_walker=1, bin_window_lo=-133.62,
                            bin_window_hi=0.02), 20000, 25000),
    "lj31_sad_fast": (dict(system="lj", method="sad", N=31, lj_radius=2.5, max_allowed_energy=0.0, sad_min_T=0.01, energy_bin=0.01,
                           n_walkers=40, init_mode=_abi.INIT_RANDOMIZE, lanes_per_walker=1, flags=_abi.FLAG_FAST_MATH,
                           bin_window_lo=-133.62, bin_window_hi=0.02), 20000, 25000),
    "two_wells_sad": (dict(system="two-wells", method="sad", N=12, tw_h2_to_h1=1.1, tw_barrier_over_h1=0.1, tw_r2=0.5, sad_min_T=0.001,
                           energy_bin=1e-3, move_value=1e-2, n_walkers=6, seed=1), 20000, 20000),
    "fake_samc": (dict(system="fake", method="samc", fake_function=_abi.FAKE_QUADRATIC, N=3, samc_t0=1e3, energy_bin=0.01,
                       n_walkers=6, seed=3, bin_window_lo=-2.5, bin_window_hi=4.0), 20000, 20000),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_resume_equals_continuous(name):
    kw, n1, n2 = CASES[name]
    kw = dict(kw)
    cfg = make_config(kw.pop("system"), kw.pop("method"), **kw)
    full = WalkerEngine(cfg)
    full.run(n1)
    full.run(n2)
    first = WalkerEngine(cfg)
    first.run(n1)
    snap = snapshot(first)
    first.close()
    second = restore(cfg, snap)
    # a restored engine reads back exactly what was put in
    for w in (0, second.n_walkers - 1):
        for k in BINS_EXACT:
            assert np.array_equal(second.bins(w)[k], snap["bins"][w][k]), (name, k)
    second.run(n2)
    assert_engines_identical(full, second, name)


def test_resume_refuses_bins_outside_the_window_and_wrong_width():
    cfg = make_config("ising", "sad", N=8, sad_min_T=1.0, n_walkers=2)
    a = WalkerEngine(cfg)
    a.run(5000)
    snap = snapshot(a)
    small = WalkerEngine(clone_config(cfg, init_mode=_abi.INIT_EXTERNAL, bin_window_lo=-20.0, bin_window_hi=20.0))
    small.set_systems(snap["systems"])
    with pytest.raises(Exception):
        small.set_walker_bins(0, snap["walkers"][0], snap["bins"][0])
    started = WalkerEngine(cfg)
    with pytest.raises(Exception):
        started.resume(10)
