"""GPU parity for SADMC_FLAG_BINNING: the bookkeeping of the reference's `binning` binary (src/mc/energy_binning.rs over
src/mc/binning/histogram.rs) on the device against its CPU restatement (oracle/oracle_binning.hpp).  Bit-exact tier:
every scalar the sampler keeps, every per-bin vector (ln w, counts, the "energy" / "t_found" / "hist" accumulators, the
system's own data_to_collect accumulator), generator state and configuration."""
import numpy as np
import pytest

from sad_monte_carlo_b200 import WalkerEngine, _abi, make_config
from tests.oracle_lib import OracleBinningMC

pytestmark = pytest.mark.gpu

SCALARS = ["moves", "accepted_moves", "acceptance_rate", "translation_scale", "rng_s0", "rng_s1", "energy", "bins_min",
           "bins_width", "bins_min_e", "bins_max_e", "bins_len", "method", "too_lo", "too_hi", "latest_parameter", "tF", "tL",
           "num_states", "samc_t0", "wl_gamma", "lnw_max_count", "lnw_total_count", "t_found_max_total", "hist_min_count",
           "hist_total_count"]
BINS = ["lnw_total", "lnw_count", "energy_total", "energy_count", "t_found_total", "t_found_count", "hist_count",
        "extra_total", "extra_count"]


def assert_binning_equal(eng, w, o, context="", exact=True, rtol=1e-12):
    g, s = eng.binning_walker(w), o.walker()
    assert g.status == 0, "%s walker %d status %d" % (context, w, g.status)
    sad_only = ("too_lo", "too_hi", "latest_parameter", "tF", "tL", "num_states", "t_found_max_total")
    wl_only = ("wl_gamma", "hist_min_count", "hist_total_count")
    for f in SCALARS:
        a, b = getattr(g, f), getattr(s, f)
        if (f == "samc_t0" and s.method != _abi.METHOD_SAMC) or (f in sad_only and s.method != _abi.METHOD_SAD) or (
                f in wl_only and s.method not in (_abi.METHOD_WL, _abi.METHOD_INV_T_WL)):
            continue
        if exact or isinstance(a, int):
            assert a == b, "%s walker %d: %s gpu=%r oracle=%r" % (context, w, f, a, b)
        else:
            assert abs(a - b) <= rtol * max(1.0, abs(b)), "%s walker %d: %s gpu=%r oracle=%r" % (context, w, f, a, b)
    gb, ob = eng.binning_bins(w), o.bins()
    for k in BINS:
        if exact or gb[k].dtype != np.float64:
            assert np.array_equal(gb[k], ob[k]), "%s walker %d: bins.%s differ at %s" % (context, w, k, np.nonzero(gb[k] != ob[k])[0][:5])
        else:
            assert np.allclose(gb[k], ob[k], rtol=rtol, atol=rtol), "%s walker %d: bins.%s" % (context, w, k)
    if exact:
        assert np.array_equal(eng.system(w), o.system()), "%s walker %d: system differs" % (context, w)


def _check(cfg, moves, walkers, exact=True):
    cfg.flags |= _abi.FLAG_BINNING
    eng = WalkerEngine(cfg)
    oracles = {w: OracleBinningMC(cfg, walker=cfg.walker_offset + w) for w in walkers}
    for w, o in oracles.items():
        assert_binning_equal(eng, w, o, "init", exact)
    for n in moves:
        eng.run(n)
        for w, o in oracles.items():
            o.run(n)
            assert_binning_equal(eng, w, o, "after %d" % eng.num_moves(), exact)
    return eng


METHODS = [("sad", dict(sad_min_T=0.001)), ("samc", dict(samc_t0=1e3)), ("wl", {}), ("wl", dict(wl_min_gamma=1e-2)), ("inv-t-wl", {})]


@pytest.mark.parametrize("fn,kw,bounds", [
    (_abi.FAKE_LINEAR, {}, (0.0, 0.995)),
    (_abi.FAKE_QUADRATIC, dict(N=3), (0.0, 0.995)),
    (_abi.FAKE_PIECES, dict(fake_a=0.1, fake_b=0.5, fake_e1=2.0, fake_e2=1.0), (-1.9, 0.5)),
    (_abi.FAKE_GAUSSIAN, dict(fake_sigma=0.3), (-0.95, -0.05)),
])
@pytest.mark.parametrize("method,mkw", METHODS)
@pytest.mark.parametrize("de", [0.01, 0.0078125])
def test_fake_systems_bit_exact(fn, kw, bounds, method, mkw, de):
    # fake/run-fake.py:25-48: `binning --histogram-bin de --translation-scale 0.05 --sad-min-T 0.001`, WL runs bounded
    b = dict(min_allowed_energy=bounds[0], max_allowed_energy=bounds[1]) if "wl" in method else {}
    cfg = make_config("fake", method, fake_function=fn, energy_bin=de, move_value=0.05, n_walkers=70, seed=3,
                      bin_window_lo=-2.5, bin_window_hi=4.0, **kw, **mkw, **b)
    _check(cfg, [1, 2999, 60000], walkers=(0, 33, 69))


def test_fake_linear_long_run_with_range_extensions_and_refound_bins():
    cfg = make_config("fake", "sad", fake_function=_abi.FAKE_LINEAR, sad_min_T=0.001, energy_bin=0.001, move_value=0.05,
                      n_walkers=33, seed=11)
    eng = _check(cfg, [400000], walkers=(0, 32))
    b = eng.binning_bins(7)
    assert b["t_found_count"].max() > 1 and b["lnw_count"].sum() < 400000 == b["energy_count"].sum()


def test_acceptance_rate_move_plan_rescales_on_new_tF():
    cfg = make_config("fake", "sad", fake_function=_abi.FAKE_QUADRATIC, N=3, sad_min_T=0.01, energy_bin=0.01,
                      move_plan=_abi.MOVE_ACCEPTANCE_RATE, move_value=0.5, n_walkers=40, seed=5)
    eng = _check(cfg, [50000], walkers=(0, 39))
    assert eng.binning_walker(3).translation_scale != 0.05


def test_randomized_starts():
    cfg = make_config("fake", "sad", fake_function=_abi.FAKE_QUADRATIC, N=3, sad_min_T=0.001, energy_bin=0.01,
                      n_walkers=64, seed=9, init_mode=_abi.INIT_RANDOMIZE)
    _check(cfg, [20000], walkers=(0, 63))


@pytest.mark.parametrize("method,mkw", [("sad", dict(sad_min_T=0.001)), ("samc", dict(samc_t0=1e4))])
def test_two_wells_with_the_which_accumulator(method, mkw):
    cfg = make_config("two-wells", method, N=12, tw_h2_to_h1=1.1, tw_barrier_over_h1=0.1, tw_r2=0.5, energy_bin=1e-3,
                      move_value=1e-2, n_walkers=40, seed=1, **mkw)
    eng = _check(cfg, [2000, 80000], walkers=(0, 39))
    b = eng.binning_bins(5)
    assert b["extra_count"].sum() == 82000 and np.all(b["extra_total"] <= b["extra_count"])


@pytest.mark.parametrize("method,mkw", [("sad", dict(sad_min_T=1.0)), ("wl", dict(min_allowed_energy=-1200.0, max_allowed_energy=0.0)),
                                        ("inv-t-wl", dict(min_allowed_energy=-1200.0, max_allowed_energy=0.0))])
@pytest.mark.parametrize("de", [4.0, 1.0])
def test_ising_bit_exact(method, mkw, de):
    cfg = make_config("ising", method, N=32, energy_bin=de, n_walkers=96, seed=2, **mkw)
    _check(cfg, [500, 40000], walkers=(0, 95))


@pytest.mark.parametrize("method,mkw", [("sad", dict(sad_min_T=0.05)), ("samc", dict(samc_t0=1e3))])
def test_lj13_reference_order_arithmetic_bit_exact(method, mkw):
    cfg = make_config("lj", method, N=13, lj_radius=2.0, max_allowed_energy=0.0, energy_bin=0.05, move_value=0.05,
                      n_walkers=40, seed=7, lanes_per_walker=1, init_mode=_abi.INIT_RANDOMIZE, bin_window_lo=-46.0,
                      bin_window_hi=0.5, **mkw)
    _check(cfg, [300, 20000], walkers=(0, 39))


def test_lj31_fast_math_tier():
    # same tolerance tier as the headline kernel: per-move energies within 1e-12; decisions can only differ for an
    # energy within that distance of a bin edge, which these 6 000 moves do not meet
    cfg = make_config("lj", "sad", N=31, lj_radius=2.5, max_allowed_energy=0.0, sad_min_T=0.01, energy_bin=0.01,
                      move_value=0.05, n_walkers=64, seed=1, lanes_per_walker=1, init_mode=_abi.INIT_RANDOMIZE,
                      bin_window_lo=-133.62, bin_window_hi=0.02, flags=_abi.FLAG_FAST_MATH)
    cfg.flags |= _abi.FLAG_BINNING
    eng = WalkerEngine(cfg)
    eng.run(6000)
    ocfg = make_config("lj", "sad", N=31, lj_radius=2.5, max_allowed_energy=0.0, sad_min_T=0.01, energy_bin=0.01,
                       move_value=0.05, n_walkers=64, seed=1, lanes_per_walker=1, init_mode=_abi.INIT_RANDOMIZE)
    for w in (0, 63):
        o = OracleBinningMC(ocfg, walker=w)
        o.run(6000)
        g, s = eng.binning_walker(w), o.walker()
        assert (g.rng_s0, g.rng_s1, g.accepted_moves, g.bins_len) == (s.rng_s0, s.rng_s1, s.accepted_moves, s.bins_len)
        assert abs(g.energy - s.energy) <= 1e-12 * abs(s.energy)
        gb, ob = eng.binning_bins(w), o.bins()
        assert np.array_equal(gb["energy_count"], ob["energy_count"]) and np.array_equal(gb["lnw_count"], ob["lnw_count"])
        assert np.allclose(gb["lnw_total"], ob["lnw_total"], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("method,mkw", [("sad", dict(sad_min_T=0.5)), ("wl", dict(min_allowed_energy=-400.0, max_allowed_energy=0.0))])
def test_square_well_bit_exact(method, mkw):
    cfg = make_config("sw", method, N=50, filling_fraction=0.3, sw_well_width=1.3, move_value=0.05, n_walkers=24, seed=3, **mkw)
    _check(cfg, [200, 8000], walkers=(0, 23))


def test_wca_with_the_pressure_accumulator():
    # wca-transposed/run.py:52: `binning ... --sad-min-T .. --max-allowed-energy .. --histogram-bin dE`; tolerance tier
    # like the energy.rs WCA kernels (different summation order): same decisions, energies within 1e-12
    from tests.test_gpu_fluids import _relaxed_wca_state
    N, rho = 27, 0.3
    state = _relaxed_wca_state(N, rho)
    cfg = make_config("wca", "sad", N=N, reduced_density=rho, sad_min_T=0.5, energy_bin=1.0, n_walkers=3, seed=5,
                      max_allowed_energy=10.0 * N, lanes_per_walker=32, init_mode=_abi.INIT_EXTERNAL, flags=_abi.FLAG_BINNING)
    eng = WalkerEngine(cfg)
    eng.set_systems(np.tile(state, (3, 1)))
    eng.start()
    o = OracleBinningMC(cfg, walker=2, system_state=state)
    n = N * N * 20 + 5
    eng.run(n)
    o.run(n)
    g, s = eng.binning_walker(2), o.walker()
    assert g.status == 0 and (g.rng_s0, g.rng_s1, g.accepted_moves, g.bins_len, g.tL, g.num_states) == (
        s.rng_s0, s.rng_s1, s.accepted_moves, s.bins_len, s.tL, s.num_states)
    assert abs(g.energy - s.energy) <= 1e-12 * max(1.0, abs(s.energy))
    gb, ob = eng.binning_bins(2), o.bins()
    for k in ("lnw_count", "energy_count", "t_found_count", "extra_count"):
        assert np.array_equal(gb[k], ob[k]), k
    assert gb["extra_count"].sum() == 20  # pressure every N^2 moves (wca.rs:203)
    assert np.allclose(gb["extra_total"], ob["extra_total"], rtol=1e-11, atol=1e-12)
    assert np.allclose(gb["lnw_total"], ob["lnw_total"], rtol=1e-12, atol=1e-12)
    assert np.allclose(gb["energy_total"], ob["energy_total"], rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("system,kw", [
    ("fake", dict(fake_function=_abi.FAKE_QUADRATIC, N=3, sad_min_T=0.001, energy_bin=0.01, high_resolution_de=0.0013)),
    ("ising", dict(N=16, sad_min_T=1.0, energy_bin=8.0, high_resolution_de=4.0)),
])
def test_high_resolution_histogram_rides_along(system, kw):
    # energy_binning.rs:124-125, 328-330: a second histogram::Bins with its own width, counts only
    cfg = make_config(system, "sad", n_walkers=40, seed=6, **kw)
    eng = _check(cfg, [1, 30000], walkers=(0, 39))
    for w in (0, 39):
        o = OracleBinningMC(cfg, walker=w)
        o.run(30001)
        (gm, gc), (om, oc) = eng.high_resolution(w), o.high_resolution()
        assert gm == om and np.array_equal(gc, oc) and int(gc.sum()) == 30001


# ---- binning::linear (SADMC_FLAG_BINNING_LINEAR, csrc/book_linear.cuh) ------------------------------------------------
LSCALARS = [f for f in SCALARS if f not in ("lnw_max_count", "hist_min_count")] + ["lnw_max_count_f64", "hist_min_count_f64"]


def assert_linear_equal(eng, w, o, context=""):
    g, s = eng.binning_walker(w), o.walker()
    assert g.status == 0, "%s walker %d status %d" % (context, w, g.status)
    sad_only = ("too_lo", "too_hi", "latest_parameter", "tF", "tL", "num_states", "t_found_max_total")
    wl_only = ("wl_gamma", "hist_min_count_f64", "hist_total_count")
    for f in LSCALARS:
        if (f == "samc_t0" and s.method != _abi.METHOD_SAMC) or (f in sad_only and s.method != _abi.METHOD_SAD) or (
                f in wl_only and s.method not in (_abi.METHOD_WL, _abi.METHOD_INV_T_WL)):
            continue
        assert getattr(g, f) == getattr(s, f), "%s walker %d: %s gpu=%r oracle=%r" % (context, w, f, getattr(g, f), getattr(s, f))
    gb, ob = eng.binning_bins_f64(w), o.bins_f64()
    for k in BINS:
        assert np.array_equal(gb[k], ob[k]), "%s walker %d: bins.%s differ at %s" % (context, w, k, np.nonzero(gb[k] != ob[k])[0][:5])
    assert np.array_equal(eng.system(w), o.system())


def _check_linear(cfg, moves, walkers):
    cfg.flags |= _abi.FLAG_BINNING | _abi.FLAG_BINNING_LINEAR
    eng = WalkerEngine(cfg)
    oracles = {w: OracleBinningMC(cfg, walker=cfg.walker_offset + w) for w in walkers}
    for w, o in oracles.items():
        assert_linear_equal(eng, w, o, "init")
    for n in moves:
        eng.run(n)
        for w, o in oracles.items():
            o.run(n)
            assert_linear_equal(eng, w, o, "after %d" % eng.num_moves())
    return eng


@pytest.mark.parametrize("fn,kw,bounds", [
    (_abi.FAKE_LINEAR, {}, (0.0, 0.99)),
    (_abi.FAKE_QUADRATIC, dict(N=3), (0.0, 0.99)),
    (_abi.FAKE_GAUSSIAN, dict(fake_sigma=0.3), (-0.95, -0.05)),
])
@pytest.mark.parametrize("method,mkw", METHODS)
def test_linear_bins_fake_systems_bit_exact(fn, kw, bounds, method, mkw):
    b = dict(min_allowed_energy=bounds[0], max_allowed_energy=bounds[1]) if "wl" in method else {}
    cfg = make_config("fake", method, fake_function=fn, energy_bin=0.01, move_value=0.05, n_walkers=40, seed=3,
                      bin_window_lo=-2.5, bin_window_hi=4.0, **kw, **mkw, **b)
    _check_linear(cfg, [1, 2999, 40000], walkers=(0, 39))


def test_linear_bins_two_wells_with_which_and_high_resolution():
    cfg = make_config("two-wells", "sad", N=12, tw_h2_to_h1=1.1, tw_barrier_over_h1=0.1, tw_r2=0.5, energy_bin=1e-3, sad_min_T=0.001,
                      move_value=1e-2, n_walkers=20, seed=1, high_resolution_de=2.5e-4)
    eng = _check_linear(cfg, [2000, 30000], walkers=(0, 19))
    o = OracleBinningMC(cfg, walker=7)
    o.run(32000)
    (gm, gc), (om, oc) = eng.high_resolution(7), o.high_resolution()
    assert gm == om and np.array_equal(gc, oc)
    assert abs(eng.binning_bins_f64(7)["extra_count"].sum() - 32000) < 1e-6


@pytest.mark.parametrize("method,mkw", [("sad", dict(sad_min_T=1.0)), ("inv-t-wl", dict(min_allowed_energy=-300.0, max_allowed_energy=0.0))])
def test_linear_bins_ising(method, mkw):
    cfg = make_config("ising", method, N=16, energy_bin=4.0, n_walkers=33, seed=2, **mkw)
    _check_linear(cfg, [500, 20000], walkers=(0, 32))


def test_linear_bins_lj13_reference_order():
    cfg = make_config("lj", "sad", N=13, lj_radius=2.0, max_allowed_energy=0.0, sad_min_T=0.05, energy_bin=0.05, move_value=0.05,
                      n_walkers=20, seed=7, lanes_per_walker=1, init_mode=_abi.INIT_RANDOMIZE, bin_window_lo=-46.0, bin_window_hi=0.5)
    _check_linear(cfg, [300, 10000], walkers=(0, 19))


def test_linear_needs_the_binning_flag_and_a_thread_per_walker_system():
    with pytest.raises(Exception):
        WalkerEngine(make_config("fake", "sad", fake_function=_abi.FAKE_LINEAR, energy_bin=0.01, flags=_abi.FLAG_BINNING_LINEAR))
    with pytest.raises(Exception):
        WalkerEngine(make_config("sw", "sad", N=50, filling_fraction=0.3, sw_well_width=1.3, flags=_abi.FLAG_BINNING | _abi.FLAG_BINNING_LINEAR))


def test_energy_layout_calls_refuse_a_binning_engine_and_canonical_is_rejected():
    cfg = make_config("fake", "sad", fake_function=_abi.FAKE_LINEAR, sad_min_T=0.001, energy_bin=0.01, n_walkers=4,
                      flags=_abi.FLAG_BINNING)
    eng = WalkerEngine(cfg)
    with pytest.raises(Exception) as ei:
        eng.bins(0)
    assert "SADMC_FLAG_BINNING" in str(ei.value)
    with pytest.raises(Exception):
        WalkerEngine(make_config("fake", "canonical", fake_function=_abi.FAKE_LINEAR, canonical_T=1.0, energy_bin=0.01,
                                 flags=_abi.FLAG_BINNING))
    with pytest.raises(Exception):  # group kernels carry no energy_binning.rs bookkeeping
        WalkerEngine(make_config("lj", "sad", N=13, lj_radius=2.0, max_allowed_energy=0.0, lanes_per_walker=8,
                                 bin_window_lo=-46.0, bin_window_hi=0.5, flags=_abi.FLAG_BINNING))


def test_fold_of_a_binning_engine_merges_visits_and_aligned_lnw():
    cfg = make_config("fake", "sad", fake_function=_abi.FAKE_LINEAR, sad_min_T=0.001, energy_bin=0.01, move_value=0.05,
                      n_walkers=256, seed=1, flags=_abi.FLAG_BINNING)
    eng = WalkerEngine(cfg)
    eng.run(300000)
    f = eng.fold()
    assert int(f["histogram"].sum()) == 256 * 300000
    assert np.all(f["energy_squared_total"] == 0.0)
    lo, width, nb = eng.window()
    j0 = int(round((0.0 - lo) / width))
    mean = f["lnw_sum"][j0 + 5:j0 + 85] / f["lnw_count"][j0 + 5:j0 + 85]
    assert np.all(f["lnw_count"][j0 + 5:j0 + 85] == 256)
    assert mean.std() < 0.1, mean  # flat density of states of fake-linear
