"""GPU parity, bit-exact tier: 2-D Ising (BASELINE.json config 1) against the CPU oracle.

Every comparison goes through the C ABI (sad_monte_carlo_b200.WalkerEngine -> libsadmc_gpu.so)."""
import numpy as np
import pytest

from sad_monte_carlo_b200 import WalkerEngine, make_config, _abi
from tests.gpu_common import assert_walker_equal, clone_config
from tests.oracle_lib import OracleMC

from sad_monte_carlo_b200.engine import SadmcError

pytestmark = pytest.mark.gpu


def _check(cfg, n_moves_list, walkers=(0, 1, 7), exact=True):
    eng = WalkerEngine(cfg)
    oracles = {w: OracleMC(cfg, walker=cfg.walker_offset + w) for w in walkers}
    for w, o in oracles.items():
        assert_walker_equal(eng, w, o, exact=exact, context="initial")
    for n in n_moves_list:
        eng.run(n)
        for w, o in oracles.items():
            o.run(n)
            assert_walker_equal(eng, w, o, exact=exact, context="after %d" % eng.num_moves())
    return eng


def test_ising32_sad_bit_exact_histograms():
    # config 1: 32x32, SAD, reference constructor (all walkers share the seed-10137 lattice), walker w <-> --seed w
    cfg = make_config("ising", "sad", N=32, sad_min_T=1.0, n_walkers=64)
    _check(cfg, [1, 999, 20000, 180000], walkers=(0, 1, 2, 31, 32, 63))


def test_ising_small_lattices_and_odd_sizes():
    for N in (2, 3, 10, 15):
        cfg = make_config("ising", "sad", N=N, sad_min_T=0.5, n_walkers=40, seed=100)
        _check(cfg, [5000, 45000], walkers=(0, 13, 39))


def test_ising_samc_and_canonical():
    cfg = make_config("ising", "samc", N=16, samc_t0=1000.0, n_walkers=33, seed=7)
    _check(cfg, [30000, 30000], walkers=(0, 32))
    cfg = make_config("ising", "canonical", N=16, canonical_T=2.5, n_walkers=33, seed=9)
    _check(cfg, [30000, 30000], walkers=(0, 32))


def test_ising_wl_and_inv_t_wl():
    # the reference's own Ising script (ising-wl-min-gamma.sh): bounded WL with min_gamma
    cfg = make_config("ising", "wl", N=8, wl_min_gamma=1e-3, min_allowed_energy=-128.0, max_allowed_energy=50.0,
                      n_walkers=34, seed=3)
    _check(cfg, [20000, 200000], walkers=(0, 1, 33))
    cfg = make_config("ising", "wl", N=8, n_walkers=34, seed=4)  # unbounded: num_states counts first visits
    _check(cfg, [20000, 100000], walkers=(0, 33))
    cfg = make_config("ising", "inv-t-wl", N=8, min_allowed_energy=-128.0, max_allowed_energy=50.0, n_walkers=34,
                      seed=5)
    eng = _check(cfg, [20000, 300000], walkers=(0, 1, 33))
    # 1/t-WL must have switched to SAMC on at least one walker by now (energy.rs:747-756)
    assert any(eng.walker(w).method == _abi.METHOD_SAMC for w in range(34))


def test_ising_acceptance_rate_move_plan_and_energy_bounds():
    cfg = make_config("ising", "sad", N=12, sad_min_T=0.7, move_plan=_abi.MOVE_ACCEPTANCE_RATE, move_value=0.5,
                      min_allowed_energy=-200.0, max_allowed_energy=100.0, n_walkers=32, seed=11)
    _check(cfg, [40000, 40000], walkers=(0, 31))


def test_ising_randomized_starts():
    cfg = make_config("ising", "sad", N=32, sad_min_T=1.0, n_walkers=48, seed=21, init_mode=_abi.INIT_RANDOMIZE)
    eng = _check(cfg, [50000], walkers=(0, 5, 47))
    assert len(set(eng.energies())) > 5


def test_launch_splitting_is_invisible():
    cfg = make_config("ising", "sad", N=16, sad_min_T=1.0, n_walkers=64, seed=2)
    a, b = WalkerEngine(cfg), WalkerEngine(cfg)
    a.run(30000)
    for n in (1, 2, 7, 990, 9000, 20000):
        b.run(n)
    assert a.num_moves() == b.num_moves() == 30000
    for w in (0, 17, 63):
        ga, gb = a.bins(w), b.bins(w)
        for k in ga:
            assert np.array_equal(ga[k], gb[k])
        assert a.walker(w).as_dict() == b.walker(w).as_dict()
    assert np.array_equal(a.systems(), b.systems())


def test_walker_offset_shards_are_the_same_walkers():
    # walkers [32, 64) of a 64-walker job == a 32-walker engine with walker_offset 32 (multi-GPU sharding rule)
    cfg = make_config("ising", "sad", N=16, sad_min_T=1.0, n_walkers=64, seed=5)
    full = WalkerEngine(cfg)
    half = WalkerEngine(clone_config(cfg, n_walkers=32, walker_offset=32))
    full.run(20000)
    half.run(20000)
    for w in (0, 9, 31):
        fa, ha = full.bins(32 + w), half.bins(w)
        for k in fa:
            assert np.array_equal(fa[k], ha[k])


def test_fold_equals_sum_over_walkers():
    cfg = make_config("ising", "sad", N=8, sad_min_T=1.0, n_walkers=50, seed=1)
    eng = WalkerEngine(cfg)
    eng.run(20000)
    lo, width, n = eng.window()
    f = eng.fold()
    hist = np.zeros(n, np.uint64)
    etot = np.zeros(n)
    cnt = np.zeros(n, np.uint64)
    lsum = np.zeros(n)
    for w in range(50):
        s, b = eng.walker(w), eng.bins(w)
        sl = slice(s.window_first, s.window_first + s.bins_len)
        assert abs((lo + s.window_first * width) - s.bins_min) < 1e-9
        hist[sl] += b["histogram"]
        etot[sl] += b["energy_total"]
        vis = b["histogram"] != 0
        cnt[sl] += vis.astype(np.uint64)
        lsum[sl] += np.where(vis, b["lnw"] - b["lnw"][vis].max(), 0.0)
    assert np.array_equal(f["histogram"], hist)
    assert np.array_equal(f["lnw_count"], cnt)
    assert np.allclose(f["energy_total"], etot)
    assert np.allclose(f["lnw_sum"], lsum)
    assert int(hist.sum()) == 50 * (20000 + 1)


def test_window_overflow_is_reported_not_clamped():
    cfg = make_config("ising", "sad", N=16, sad_min_T=1.0, n_walkers=8, seed=1, bin_window_lo=-40.0, bin_window_hi=40.0)
    eng = WalkerEngine(cfg)
    with pytest.raises(SadmcError) as ei:  # the run says so at once (the reference would have grown its vectors)
        eng.run(20000)
    assert ei.value.code == _abi.ERR_WINDOW and "8 walker(s) left the device bin window" in str(ei.value)
    assert all(eng.walker(w).status == _abi.ERR_WINDOW for w in range(8))
    assert eng.num_halted() == (8, 0)
    eng.run(100)  # every halted walker is reported once; the engine stays usable
    assert eng.num_halted() == (8, 0)


def test_trait_shims_match_oracle_move_by_move():
    cfg = make_config("ising", "sad", N=10, n_walkers=3, seed=10137)
    eng = WalkerEngine(cfg)
    o = OracleMC(cfg, walker=2)
    assert eng.energy(2) == o.energy() == eng.compute_energy(2)
    for _ in range(300):
        eg, eo = eng.plan_move(2, 0.0), o.plan_move(0.0)
        assert eg == eo
        eng.confirm(2)
        o.confirm()
        assert eng.energy(2) == o.energy()
    assert eng.energy(2) == eng.compute_energy(2)
    assert np.array_equal(eng.system(2), o.system())
