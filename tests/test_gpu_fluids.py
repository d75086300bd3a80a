"""GPU parity for the periodic fluids on cell lists (BASELINE.json config 5): square well / hard spheres are in the
bit-exact tier (integer energies), WCA in the floating-point tier (<= 1e-12 relative per move)."""
import numpy as np
import pytest

from sad_monte_carlo_b200 import WalkerEngine, make_config, _abi
from tests.gpu_common import assert_walker_equal, clone_config
from tests.oracle_lib import OracleMC

pytestmark = pytest.mark.gpu
RTOL = 1e-12


# ---- square well -----------------------------------------------------------------------------------------------

@pytest.mark.parametrize("kw", [dict(N=100, filling_fraction=0.3), dict(N=50, filling_fraction=0.3),
                                dict(N=100, cell_width=(6.0, 6.0, 6.0)), dict(N=200, filling_fraction=0.2)])
def test_sw_sad_trajectory_bit_exact(kw):
    # tests/square-sad-test.rs drives 10^4 SAD moves with the defaults; resume-sad.rs uses --sad-min-T 0.5
    cfg = make_config("sw", "sad", sad_min_T=0.5, n_walkers=6, seed=0, **kw)
    eng = WalkerEngine(cfg)
    oracles = {w: OracleMC(cfg, walker=w) for w in (0, 5)}
    for w, o in oracles.items():
        assert_walker_equal(eng, w, o, exact=True, context="init")
    for n in (500, 20000):
        eng.run(n)
        for w, o in oracles.items():
            o.run(n)
            assert_walker_equal(eng, w, o, exact=True, context="after %d" % eng.num_moves())
    assert eng.verify_energy(0)  # == compute_energy_slowly, optsquare.rs:199-201


@pytest.mark.parametrize("method,kw", [("samc", dict(samc_t0=1e3)), ("wl", {}), ("inv-t-wl", dict(min_allowed_energy=-300.0, max_allowed_energy=-50.0))])
def test_sw_other_methods_bit_exact(method, kw):
    cfg = make_config("sw", method, N=64, filling_fraction=0.25, n_walkers=5, seed=3, **kw)
    eng = WalkerEngine(cfg)
    o = OracleMC(cfg, walker=4)
    eng.run(30000)
    o.run(30000)
    assert_walker_equal(eng, 4, o, exact=True, context=method)


def test_sw_shims_like_the_reference_unit_test():
    # src/system/optsquare.rs:602-618: plan_move + confirm 1000 times at scale 1.0 (confirm after None is a no-op)
    cfg = make_config("sw", "sad", N=50, filling_fraction=0.3, n_walkers=2, seed=1)
    eng = WalkerEngine(cfg)
    o = OracleMC(cfg, walker=1)
    assert eng.energy(1) == o.energy() == eng.compute_energy(1)
    for _ in range(400):
        eg, eo = eng.plan_move(1, 1.0), o.plan_move(1.0)
        assert eg == eo
        eng.confirm(1)
        o.confirm()
        assert eng.energy(1) == o.energy()
    assert eng.energy(1) == eng.compute_energy(1)
    assert np.array_equal(eng.system(1), o.system())


def test_sw_rejects_boxes_the_reference_rejects():
    with pytest.raises(Exception) as ei:
        WalkerEngine(make_config("sw", "sad", N=10, cell_width=(1.2, 5.0, 5.0)))
    assert "not large enough" in str(ei.value)


# ---- WCA ----------------------------------------------------------------------------------------------------------

_RELAXED = {}


def _relaxed_wca_state(N, rho):
    """A thermalised WCA configuration: the reference constructor (reduced attempt count) followed by a short
    canonical run at T = 1 in the CPU oracle.  Cached per (N, rho)."""
    if (N, rho) not in _RELAXED:
        o = OracleMC(make_config("wca", "canonical", N=N, reduced_density=rho, canonical_T=1.0, energy_bin=1e9,
                                 move_value=0.1, seed=77), attempts_override=5)
        o.run(300 * N)
        assert o.energy() < 8.0 * N
        _RELAXED[(N, rho)] = o.system()
    return _RELAXED[(N, rho)]


def _wca_pair(N, rho, n_walkers, method="samc", **kw):
    """GPU engine whose walkers all start from one relaxed configuration, handed over through SADMC_INIT_EXTERNAL
    (the resume path); the matching oracle walkers are built from the same image."""
    # max_allowed_energy as in the reference's WCA runs (wca/run-wca.py run_sad(max_E=...)): without it one
    # overlapping proposal (E ~ 1e6) makes the reference grow millions of bins
    base = dict(N=N, reduced_density=rho, energy_bin=1.0, n_walkers=n_walkers, seed=5, max_allowed_energy=10.0 * N,
                samc_t0=1e3)
    base.update(kw)
    state = _relaxed_wca_state(N, rho)
    cfg = make_config("wca", method, init_mode=_abi.INIT_EXTERNAL, **base)
    eng = WalkerEngine(cfg)
    eng.set_systems(np.tile(state, (n_walkers, 1)))
    eng.start()
    return cfg, eng, state


# kernels: lanes_per_walker 32 = one warp per walker (sys_cell_fluid.cuh); 4 / 8 / 16 = lane groups (sys_wca_group.cuh),
# with SADMC_FLAG_FAST_MATH their fast arithmetic and relaxed re-summation cadence
KERNELS = [dict(lanes_per_walker=32), dict(lanes_per_walker=8), dict(lanes_per_walker=8, flags=_abi.FLAG_FAST_MATH),
           dict(lanes_per_walker=4, flags=_abi.FLAG_FAST_MATH), dict(lanes_per_walker=16), dict(lanes_per_walker=16, flags=_abi.FLAG_FAST_MATH)]
KERNEL_IDS = ["warp", "g8", "g8fast", "g4fast", "g16", "g16fast"]


@pytest.mark.parametrize("kern", KERNELS, ids=KERNEL_IDS)
@pytest.mark.parametrize("N,rho", [(50, 1.0), (100, 0.3), (256, 0.8)])
def test_wca_per_move_energy_within_1e12(N, rho, kern):
    cfg, eng, state = _wca_pair(N, rho, 2, **kern)
    o = OracleMC(cfg, walker=1, system_state=state)
    rng = np.random.default_rng(0)
    for step in range(1200):
        eng.set_system(1, o.system())
        st = o.walker()
        r = eng.rngs()
        r[1] = (st.rng_s0, st.rng_s1)
        eng.set_rngs(r)
        eg, eo = eng.plan_move(1, 0.3), o.plan_move(0.3)
        assert eg is not None and eo is not None  # WCA never returns None (wca.rs:119-140)
        scale_e = max(1.0, abs(eo), abs(o.energy()))
        assert abs(eg - eo) <= RTOL * scale_e, (step, eg, eo)
        if eo < o.energy() or (eo < 9.0 * N and rng.random() < 0.5):
            o.confirm()
    eng.set_system(1, o.system())
    assert abs(eng.compute_energy(1) - o.compute_energy()) <= RTOL * max(1.0, abs(o.energy()))


@pytest.mark.parametrize("kern", KERNELS, ids=KERNEL_IDS)
@pytest.mark.parametrize("method,kw", [("samc", {}), ("sad", dict(sad_min_T=0.5)), ("inv-t-wl", dict(min_allowed_energy=0.0, max_allowed_energy=600.0))])
def test_wca_trajectory_tracks_oracle(method, kw, kern):
    cfg, eng, state = _wca_pair(64, 0.7, 4, method=method, **kw, **kern)
    oracles = {w: OracleMC(cfg, walker=w, system_state=state) for w in (0, 3)}
    eng.run(20000)
    for w, o in oracles.items():
        o.run(20000)
        g, s = eng.walker(w), o.walker()
        assert g.status == 0
        assert (g.rng_s0, g.rng_s1) == (s.rng_s0, s.rng_s1)
        assert g.accepted_moves == s.accepted_moves
        assert abs(g.energy - s.energy) <= RTOL * max(1.0, abs(s.energy))
        gb, ob = eng.bins(w), o.bins()
        assert np.array_equal(gb["histogram"], ob["histogram"])
        assert np.allclose(gb["lnw"], ob["lnw"], rtol=1e-12, atol=1e-12)
        assert np.array_equal(eng.system(w)[:-2], o.system()[:-2])  # positions: identical arithmetic
        assert eng.verify_energy(w)


def test_wca_fast_tier_energy_drift_between_resummations_is_far_below_1e12():
    """The fast tier re-sums the whole energy every 65 536 accepted moves instead of every ~10 (wca.rs:164-177): in
    between the cached energy only collects the rounding of the per-move differences."""
    cfg, eng, state = _wca_pair(256, 0.8, 64, lanes_per_walker=8, flags=_abi.FLAG_FAST_MATH)
    eng.run(60000)  # fewer accepted moves than one re-summation period: pure accumulation
    worst = 0.0
    for w in range(0, 64, 7):
        assert eng.walker(w).accepted_moves < 65536
        e, good = eng.energy(w), eng.compute_energy(w)
        worst = max(worst, abs(e - good) / max(1.0, abs(good)))
        assert eng.verify_energy(w)  # the reference's own tolerance (wca.rs:237-251)
    assert worst <= 1e-13, worst


@pytest.mark.parametrize("kern", KERNELS[:3], ids=KERNEL_IDS[:3])
def test_wca_pressure_extra_every_n_squared_moves(kern):
    cfg, eng, state = _wca_pair(27, 0.3, 3, **kern)
    o = OracleMC(cfg, walker=2, system_state=state)
    eng.run(27 * 27 * 20 + 5)
    o.run(27 * 27 * 20 + 5)
    gb, ob = eng.bins(2), o.bins()
    assert np.array_equal(gb["extra_count"], ob["extra_count"]) and gb["extra_count"].sum() == 20
    assert np.allclose(gb["extra_total"], ob["extra_total"], rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("kern", KERNELS[:3], ids=KERNEL_IDS[:3])
def test_wca_randomize_matches_reference_randomize(kern):
    cfg = make_config("wca", "samc", N=30, reduced_density=0.05, samc_t0=10.0, energy_bin=1e9, n_walkers=3, seed=11,
                      init_mode=_abi.INIT_RANDOMIZE, bin_window_lo=0.0, bin_window_hi=1e12, **kern)
    eng = WalkerEngine(cfg)
    for w in (0, 2):
        o = OracleMC(cfg, walker=w)
        gs, osys = eng.system(w), o.system()
        assert np.array_equal(gs[:-2], osys[:-2])
        assert abs(gs[-2] - osys[-2]) <= RTOL * max(1.0, abs(osys[-2]))
        assert eng.walker(w).status == 0


@pytest.mark.parametrize("kern", KERNELS[:3], ids=KERNEL_IDS[:3])
def test_wca_reference_constructor_start_tracks_oracle(kern):
    # SADMC_INIT_REFERENCE: every walker starts from the N*N-attempt configuration of wca.rs:448-496 (built on the
    # host by the library), then from_params relaxes it below max_allowed_energy (energy.rs:840-851) on the device
    cfg = make_config("wca", "samc", N=40, reduced_density=0.5, energy_bin=1.0, n_walkers=4, seed=9, samc_t0=1e3,
                      max_allowed_energy=400.0, **kern)
    eng = WalkerEngine(cfg)
    for w in (0, 3):
        o = OracleMC(cfg, walker=w)
        g, s = eng.walker(w), o.walker()
        assert g.status == 0 and g.energy < 400.0
        assert (g.rng_s0, g.rng_s1) == (s.rng_s0, s.rng_s1)  # the relaxation used the same number of draws
        assert np.array_equal(eng.system(w)[:-2], o.system()[:-2])
        assert abs(g.energy - s.energy) <= RTOL * max(1.0, abs(s.energy))
    eng.run(5000)
    o.run(5000)
    g, s = eng.walker(3), o.walker()
    assert (g.rng_s0, g.rng_s1, g.accepted_moves) == (s.rng_s0, s.rng_s1, s.accepted_moves)
    assert np.array_equal(eng.bins(3)["histogram"], o.bins()["histogram"])
    assert eng.verify_energy(3)
