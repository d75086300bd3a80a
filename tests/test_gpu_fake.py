"""GPU parity for the analytic test systems (BASELINE.json config 2): fake linear / quadratic / pieces / gaussian and
two-wells are bit-exact against the CPU oracle; erfinv is a tolerance-tier system (platform erf/exp/log)."""
import numpy as np
import pytest

from sad_monte_carlo_b200 import WalkerEngine, make_config, _abi
from tests.gpu_common import assert_walker_equal
from tests.oracle_lib import OracleMC

pytestmark = pytest.mark.gpu


def _check(cfg, moves, walkers, exact=True):
    eng = WalkerEngine(cfg)
    oracles = {w: OracleMC(cfg, walker=cfg.walker_offset + w) for w in walkers}
    for w, o in oracles.items():
        assert_walker_equal(eng, w, o, exact=exact, context="init")
    for n in moves:
        eng.run(n)
        for w, o in oracles.items():
            o.run(n)
            assert_walker_equal(eng, w, o, exact=exact, context="after %d" % eng.num_moves())
    return eng


@pytest.mark.parametrize("fn,kw", [
    (_abi.FAKE_LINEAR, {}),
    (_abi.FAKE_QUADRATIC, dict(N=3)),
    (_abi.FAKE_QUADRATIC, dict(N=7)),
    (_abi.FAKE_PIECES, dict(fake_a=0.1, fake_b=0.5, fake_e1=2.0, fake_e2=1.0)),
    (_abi.FAKE_GAUSSIAN, dict(fake_sigma=0.3)),
])
@pytest.mark.parametrize("method,mkw", [("sad", dict(sad_min_T=0.001)), ("samc", dict(samc_t0=1e3)), ("wl", {})])
def test_fake_systems_bit_exact(fn, kw, method, mkw):
    # fake/run-fake.py:28-36,76-79: min_T 0.001, translation scale 0.05, energy bins 0.001..0.1
    cfg = make_config("fake", method, fake_function=fn, energy_bin=0.01, move_value=0.05, n_walkers=70, seed=3,
                      bin_window_lo=-2.5, bin_window_hi=4.0, **kw, **mkw)
    _check(cfg, [3000, 60000], walkers=(0, 33, 69))


def test_fake_randomized_starts():
    cfg = make_config("fake", "sad", fake_function=_abi.FAKE_QUADRATIC, N=3, sad_min_T=0.001, energy_bin=0.01,
                      n_walkers=64, seed=9, init_mode=_abi.INIT_RANDOMIZE)
    _check(cfg, [20000], walkers=(0, 63))


@pytest.mark.parametrize("barrier", [0.0, 0.1, 0.2])
def test_two_wells_bit_exact_including_which_well_counts(barrier):
    # two-wells/run-two-wells.py:144-148 "T-trans-1": N = 12, h2/h1 = 1.1, r2 = 0.5
    cfg = make_config("two-wells", "sad", N=12, tw_h2_to_h1=1.1, tw_barrier_over_h1=barrier, tw_r2=0.5,
                      sad_min_T=0.001, energy_bin=1e-3, move_value=1e-2, n_walkers=40, seed=1)
    eng = _check(cfg, [2000, 80000], walkers=(0, 39))
    b = eng.bins(5)
    assert b["extra_count"].sum() == 82000 and np.all(b["extra_total"] <= b["extra_count"])


def test_two_wells_rejects_bad_dimension():
    with pytest.raises(Exception) as ei:
        WalkerEngine(make_config("two-wells", "sad", N=10, tw_h2_to_h1=1.1, tw_barrier_over_h1=0.0, tw_r2=0.5))
    assert "divisible by three" in str(ei.value)


def test_erfinv_per_move_energy_within_tolerance():
    cfg = make_config("fake-erfinv", "sad", N=8, erfinv_mean_energy=0.0, sad_min_T=0.05, energy_bin=0.05,
                      move_value=0.05, n_walkers=4, seed=2, bin_window_lo=-20.0, bin_window_hi=20.0)
    eng = WalkerEngine(cfg)
    o = OracleMC(cfg, walker=1)
    for step in range(1500):
        eng.set_system(1, o.system())
        st = o.walker()
        r = eng.rngs()
        r[1] = (st.rng_s0, st.rng_s1)
        eng.set_rngs(r)
        eg, eo = eng.plan_move(1, 0.1), o.plan_move(0.1)
        assert (eg is None) == (eo is None)
        if eo is not None:
            assert abs(eg - eo) <= 1e-12 * max(1.0, abs(eo))
            o.confirm()


def test_erfinv_short_trajectory_tracks_oracle():
    # |erfinv(x)| < 5.9 for every double |x| < 1: eight coordinates stay inside [-48, 48]
    cfg = make_config("fake-erfinv", "sad", N=8, erfinv_mean_energy=0.0, sad_min_T=0.05, energy_bin=0.05,
                      move_value=0.05, n_walkers=33, seed=2, bin_window_lo=-50.0, bin_window_hi=50.0)
    eng = WalkerEngine(cfg)
    o = OracleMC(cfg, walker=32)
    eng.run(20000)
    o.run(20000)
    g, s = eng.walker(32), o.walker()
    assert (g.rng_s0, g.rng_s1) == (s.rng_s0, s.rng_s1)
    assert abs(g.energy - s.energy) <= 1e-12 * max(1.0, abs(s.energy))
    assert np.array_equal(eng.bins(32)["histogram"], o.bins()["histogram"])
