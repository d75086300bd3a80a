"""GPU parity, floating-point tier: Lennard-Jones clusters (BASELINE.json configs 3 and 4).

Tolerance (stated by BASELINE.json's north_star): per-move energies within 1e-12 relative in f64.  The kernel sums
the O(N) pair terms in a lane-tree order with FMA-contracted r^2, the reference sums them sequentially; everything
else (RNG stream, positions, accept decisions, bookkeeping) is arithmetic-identical, which the SUM_TREE tests pin
bit for bit by letting the oracle add the pair terms in the kernel's order."""
import numpy as np
import pytest

from sad_monte_carlo_b200 import WalkerEngine, make_config, _abi
from tests.gpu_common import assert_walker_equal, clone_config
from tests.oracle_lib import OracleMC

pytestmark = pytest.mark.gpu
RTOL = 1e-12
FLAG_SUM_TREE = 2


def lj_cfg(N=31, R=2.5, method="sad", lanes=8, **kw):
    base = dict(N=N, lj_radius=R, sad_min_T=0.01, energy_bin=0.01, max_allowed_energy=0.0, move_value=0.05,
                n_walkers=16, seed=0, init_mode=_abi.INIT_RANDOMIZE, lanes_per_walker=lanes,
                bin_window_lo=-8.0 * N, bin_window_hi=0.1)
    base.update(kw)
    return make_config("lj", method, **base)


@pytest.mark.parametrize("N,R,lanes", [(31, 2.5, 32), (31, 2.5, 8), (31, 2.5, 16), (31, 2.5, 4), (38, 3.0, 8),
                                       (38, 3.0, 32), (13, 2.0, 4), (7, 2.0, 8)])
def test_lj_sad_trajectory_bit_exact_with_tree_ordered_oracle(N, R, lanes):
    cfg = lj_cfg(N=N, R=R, lanes=lanes, flags=FLAG_SUM_TREE)
    eng = WalkerEngine(cfg)
    walkers = (0, 5, 15)
    oracles = {w: OracleMC(cfg, walker=w) for w in walkers}
    for w, o in oracles.items():
        assert_walker_equal(eng, w, o, exact=True, context="init")
    for n in (1000, 30000):
        eng.run(n)
        for w, o in oracles.items():
            o.run(n)
            assert_walker_equal(eng, w, o, exact=True, context="N=%d G=%d after %d" % (N, lanes, eng.num_moves()))


@pytest.mark.parametrize("method,kw", [("samc", dict(samc_t0=1e4)), ("wl", dict(min_allowed_energy=-110.0)),
                                       ("inv-t-wl", dict(min_allowed_energy=-110.0)),
                                       ("canonical", dict(canonical_T=0.3))])
def test_lj31_other_methods_bit_exact_with_tree_ordered_oracle(method, kw):
    cfg = lj_cfg(method=method, flags=FLAG_SUM_TREE, energy_bin=0.5, **kw)
    eng = WalkerEngine(cfg)
    oracles = {w: OracleMC(cfg, walker=w) for w in (0, 9)}
    for n in (2000, 40000):
        eng.run(n)
        for w, o in oracles.items():
            o.run(n)
            assert_walker_equal(eng, w, o, exact=True, context="%s after %d" % (method, eng.num_moves()))


def test_lj31_per_move_energy_within_1e12_of_reference_order():
    """The parity claim proper: same configuration + same RNG state -> proposed energy within 1e-12 relative of
    the reference-order (sequential, no FMA) sum, for 3000 consecutive proposals along a SAD trajectory."""
    cfg = lj_cfg(n_walkers=4, lanes=8)
    eng = WalkerEngine(cfg)
    o = OracleMC(cfg, walker=1)
    worst = 0.0
    rng = np.random.default_rng(0)
    for step in range(3000):
        # put the GPU walker exactly where the oracle is
        eng.set_system(1, o.system())
        st = o.walker()
        rngs = eng.rngs()
        rngs[1] = (st.rng_s0, st.rng_s1)
        eng.set_rngs(rngs)
        scale = 0.05 if step % 3 else 0.3
        eg, eo = eng.plan_move(1, scale), o.plan_move(scale)
        assert (eg is None) == (eo is None)
        if eo is not None:
            # relative to the larger of the two energies the move connects: e2 = E + sum(terms), so a move that
            # leaves a high-energy state cancels digits in ANY summation order, the reference's included
            scale_e = max(1.0, abs(eo), abs(o.energy()))
            worst = max(worst, abs(eg - eo) / scale_e)
            assert abs(eg - eo) <= RTOL * scale_e, (step, eg, eo, o.energy())
            if eo < o.energy() or (eo < 0.0 and rng.random() < 0.5):
                o.confirm()
    assert worst < RTOL
    print("worst relative per-move energy error: %.3g" % worst)


def test_lj31_sad_trajectory_tracks_reference_order_oracle():
    """Reference-order oracle vs kernel over a whole SAD run: energies stay within 1e-12 and, because a 1e-13
    perturbation almost never flips a bin index or an accept test, the histograms are identical."""
    cfg = lj_cfg(n_walkers=16, lanes=8)
    eng = WalkerEngine(cfg)
    oracles = {w: OracleMC(cfg, walker=w) for w in (0, 3)}
    eng.run(50000)
    for w, o in oracles.items():
        o.run(50000)
        g, s = eng.walker(w), o.walker()
        assert abs(g.energy - s.energy) <= RTOL * abs(s.energy)
        assert (g.rng_s0, g.rng_s1) == (s.rng_s0, s.rng_s1)
        assert g.accepted_moves == s.accepted_moves
        gb, ob = eng.bins(w), o.bins()
        assert np.array_equal(gb["histogram"], ob["histogram"])
        assert np.allclose(gb["lnw"], ob["lnw"], rtol=1e-12, atol=1e-12)
        assert np.allclose(gb["energy_total"], ob["energy_total"], rtol=1e-12)
        assert np.allclose(eng.system(w)[:-2], o.system()[:-2], rtol=0, atol=0)  # positions: identical arithmetic


@pytest.mark.parametrize("N", [3, 50])
def test_lj_verify_energy_like_the_reference_test(N):
    # src/system/lj.rs:380-434, through the trait shims
    radius = 10.0 * N ** (1.0 / 3.0)
    cfg = make_config("lj", "sad", N=N, lj_radius=radius, n_walkers=2, seed=1, init_mode=_abi.INIT_RANDOMIZE,
                      lanes_per_walker=32, bin_window_lo=-8.0 * N, bin_window_hi=1e7, energy_bin=1e3)
    eng = WalkerEngine(cfg)
    assert abs(eng.energy(0) - eng.compute_energy(0)) <= abs(eng.energy(0)) * 1e-14 * N * N + 1e-300
    old = eng.energy(0)
    maxe = N * 16.0
    done = 0
    tries = 0
    while done < 150 and tries < 3000:
        tries += 1
        e = eng.plan_move(0, 1.0)
        if e is not None and (e < maxe or e < old):
            eng.confirm(0)
            assert eng.verify_energy(0)
            old = e
            done += 1
    assert done > 20


def test_lj_hard_wall_returns_none():
    cfg = lj_cfg(N=7, R=2.0, n_walkers=2, lanes=8)
    eng = WalkerEngine(cfg)
    nones = sum(eng.plan_move(0, 3.0) is None for _ in range(200))
    assert nones > 50


# ---- one thread per walker (lanes_per_walker = 1), cluster resident in shared memory ----------------------------

@pytest.mark.parametrize("N,R,walkers", [(31, 2.5, 70), (38, 3.0, 33), (7, 2.0, 64)])
def test_lj_thread_per_walker_exact_mode_is_bit_exact_with_the_reference_order_oracle(N, R, walkers):
    """EXACT arithmetic: sequential pair sum, IEEE divide, no FMA == the reference's own operation order, so the
    plain (reference-order) oracle must agree bit for bit -- positions, energies, lnw, histograms, RNG state."""
    cfg = lj_cfg(N=N, R=R, lanes=1, n_walkers=walkers)
    eng = WalkerEngine(cfg)
    ws = (0, 31, walkers - 1)
    oracles = {w: OracleMC(cfg, walker=w) for w in ws}
    for w, o in oracles.items():
        assert_walker_equal(eng, w, o, exact=True, context="init")
    for n in (1500, 40000):
        eng.run(n)
        for w, o in oracles.items():
            o.run(n)
            assert_walker_equal(eng, w, o, exact=True, context="N=%d thread/walker after %d" % (N, eng.num_moves()))


@pytest.mark.parametrize("method,kw", [("samc", dict(samc_t0=1e4)), ("wl", dict(min_allowed_energy=-110.0)),
                                       ("canonical", dict(canonical_T=0.3))])
def test_lj_thread_per_walker_exact_other_methods(method, kw):
    cfg = lj_cfg(method=method, lanes=1, n_walkers=40, energy_bin=0.5, **kw)
    eng = WalkerEngine(cfg)
    oracles = {w: OracleMC(cfg, walker=w) for w in (0, 39)}
    eng.run(30000)
    for w, o in oracles.items():
        o.run(30000)
        assert_walker_equal(eng, w, o, exact=True, context=method)


@pytest.mark.parametrize("lanes", [1, 2, 4])
def test_lj_thread_per_walker_fast_math_per_move_energy_within_1e12(lanes):
    cfg = lj_cfg(n_walkers=4, lanes=lanes, flags=_abi.FLAG_FAST_MATH)
    eng = WalkerEngine(cfg)
    o = OracleMC(lj_cfg(n_walkers=4, lanes=1), walker=1)
    worst = 0.0
    rng = np.random.default_rng(1)
    for step in range(2500):
        eng.set_system(1, o.system())
        st = o.walker()
        rngs = eng.rngs()
        rngs[1] = (st.rng_s0, st.rng_s1)
        eng.set_rngs(rngs)
        scale = 0.05 if step % 3 else 0.3
        eg, eo = eng.plan_move(1, scale), o.plan_move(scale)
        assert (eg is None) == (eo is None)
        if eo is not None:
            # relative to the larger of the two energies the move connects: e2 = E + sum(terms), so a move that
            # leaves a high-energy state cancels digits in ANY summation order, the reference's included
            scale_e = max(1.0, abs(eo), abs(o.energy()))
            worst = max(worst, abs(eg - eo) / scale_e)
            assert abs(eg - eo) <= RTOL * scale_e, (step, eg, eo, o.energy())
            if eo < o.energy() or (eo < 0.0 and rng.random() < 0.5):
                o.confirm()
    print("fast-math worst relative per-move energy error: %.3g" % worst)


@pytest.mark.parametrize("lanes", [1, 2, 4])
def test_lj_thread_per_walker_fast_math_tracks_reference_trajectory(lanes):
    cfg = lj_cfg(n_walkers=100, lanes=lanes, flags=_abi.FLAG_FAST_MATH)
    eng = WalkerEngine(cfg)
    oracles = {w: OracleMC(lj_cfg(n_walkers=100, lanes=1), walker=w) for w in (0, 50, 99)}
    eng.run(60000)  # long enough for several warp-cooperative energy recomputations per walker
    for w, o in oracles.items():
        o.run(60000)
        g, s = eng.walker(w), o.walker()
        assert abs(g.energy - s.energy) <= RTOL * abs(s.energy)
        assert (g.rng_s0, g.rng_s1) == (s.rng_s0, s.rng_s1)
        assert np.array_equal(eng.bins(w)["histogram"], o.bins()["histogram"])
        assert np.allclose(eng.system(w)[:-2], o.system()[:-2], rtol=0, atol=0)
        assert abs(eng.compute_energy(w) - g.energy) <= 1e-14 * 31 * 31 * abs(g.energy)


# ---- converged physics: heat capacity of LJ31 against the literature curves the reference ships -------------------

@pytest.mark.timeout(600)
@pytest.mark.parametrize("lanes", [1, 2])  # 1: the kernel bench.py times
def test_lj31_heat_capacity_short_run_approaches_the_reference_curve(lanes):
    """A 15-second version of tools/lj31_cv_run.py (the full 1e8-move run is pinned on the CPU side by
    tests/test_analysis.py from its committed folds): 37 888 SAD walkers x 3e6 moves, min_T 0.15, energy bin 0.1.
    Stated tolerance for this SHORT run: within 12 % of tRem_Ref.csv (the curve of the reference's own error metric,
    plotting/final_heat_capacity.py:185) for T in [0.28, 0.40]; walker-group standard error below 2 %."""
    import os
    from sad_monte_carlo_b200 import analysis
    W, G = 37888, 8
    cfg = make_config("lj", "sad", N=31, lj_radius=2.5, max_allowed_energy=0.0, sad_min_T=0.15, energy_bin=0.1,
                      move_value=0.05, n_walkers=W, init_mode=_abi.INIT_RANDOMIZE, lanes_per_walker=lanes, seed=0,
                      flags=_abi.FLAG_FAST_MATH, bin_window_lo=-133.7, bin_window_hi=0.2)
    eng = WalkerEngine(cfg)
    for _ in range(3):
        eng.run(1_000_000)
    lo, width, nb = eng.window()
    folds = {"window_lo": lo, "width": width, "walkers": W, "groups": G}
    for g in range(G):
        eng.fold_select(g, G, True)
        f = eng.fold()
        folds["lnw_sum_%d" % g], folds["lnw_count_%d" % g] = f["lnw_sum"], f["lnw_count"]
    assert all(eng.walker(w).status == 0 for w in range(0, W, 997))
    ref = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lj31_literature", "tRem_Ref.csv")
    T, cv, sem, cref, err = analysis.cv_error_vs_reference(folds, ref, 0.28, 0.40)
    print("T", T, "Cv", cv, "ref", cref, "err", err)
    assert len(T) >= 3
    assert np.abs(err).max() < 0.12
    assert (sem / cv).max() < 0.02


@pytest.mark.parametrize("flags", [0, _abi.FLAG_FAST_MATH])
def test_lj38_one_224_thread_cta_per_sm_tracks_the_oracle_across_cta_boundaries(flags):
    """LJ38 runs one 224-thread CTA per SM (7 warps; two 128-thread CTAs do not fit shared memory): walkers at CTA and warp
    boundaries, a partial last CTA, exact tier bit for bit and tolerance tier within 1e-12."""
    walkers = 2 * 224 + 37
    cfg = lj_cfg(N=38, R=3.0, lanes=1, n_walkers=walkers, flags=flags)
    eng = WalkerEngine(cfg)
    ocfg = lj_cfg(N=38, R=3.0, lanes=1, n_walkers=walkers)
    ws = (0, 223, 224, 447, 448, walkers - 1)
    eng.run(4000)
    for w in ws:
        o = OracleMC(ocfg, walker=w)
        o.run(4000)
        if flags == 0:
            assert_walker_equal(eng, w, o, exact=True, context="LJ38 walker %d" % w)
        else:
            g, s = eng.walker(w), o.walker()
            assert g.status == 0 and (g.rng_s0, g.rng_s1, g.accepted_moves) == (s.rng_s0, s.rng_s1, s.accepted_moves), w
            assert abs(g.energy - s.energy) <= 1e-12 * abs(s.energy)
            assert np.array_equal(eng.bins(w)["histogram"], o.bins()["histogram"])
            assert abs(eng.compute_energy(w) - g.energy) <= 1e-11 * abs(g.energy)


@pytest.mark.parametrize("N,R,method,kw", [(31, 2.5, "sad", {}), (31, 2.5, "wl", dict(min_allowed_energy=-110.0)), (31, 2.5, "samc", dict(samc_t0=1e4)),
                                           (31, 2.5, "canonical", dict(canonical_T=0.3)), (38, 3.0, "sad", {}),
                                           (38, 3.0, "inv-t-wl", dict(min_allowed_energy=-150.0))])
def test_lj_z_streamed_from_l2_is_bit_identical_to_the_shared_memory_layout(N, R, method, kw):
    """The LJ31 / LJ38 tolerance-tier move kernels keep x and y in shared memory and stream z from an L2-resident array through a
    cp.async ring (LJ31: three 128-thread CTAs per SM, LJ38: one of 320 threads; SADMC_FLAG_LJ_STREAM_Z) or keep all three in shared
    memory (SADMC_FLAG_LJ_SMEM_Z).  Same operations in the same order: every configuration, energy, generator state and bin vector
    must agree bit for bit -- across CTA and warp boundaries, a partial last CTA, several launches (the stream is rebuilt from the
    stored configurations at every launch), and through the warp-cooperative energy re-summation (every ~10 N accepted moves of a
    walker)."""
    block = 128 if N == 31 else 320
    walkers = 3 * block + 45
    a = WalkerEngine(lj_cfg(N=N, R=R, lanes=1, n_walkers=walkers, method=method, flags=_abi.FLAG_FAST_MATH | _abi.FLAG_LJ_STREAM_Z, **kw))
    b = WalkerEngine(lj_cfg(N=N, R=R, lanes=1, n_walkers=walkers, method=method, flags=_abi.FLAG_FAST_MATH | _abi.FLAG_LJ_SMEM_Z, **kw))
    groups = (N + 3) // 4
    # x, y + a two-stage ring per warp + the ziggurat tables; whole groups of four z values streamed
    assert a.move_launch_shape() == (block, 1, 4128 + block * N * 16 + (block // 32) * 2048, groups * 32)
    assert b.move_launch_shape() == (128 if N == 31 else 224, 1, 4128 + (128 if N == 31 else 224) * N * 24, 0)
    for n in (1, 999, 20000, 30000):
        a.run(n)
        b.run(n)
        assert a.num_halted() == (0, 0) and b.num_halted() == (0, 0)
        assert np.array_equal(a.systems(), b.systems()), "configurations differ after %d moves" % a.num_moves()
        assert np.array_equal(a.rngs(), b.rngs())
        assert np.array_equal(a.energies(), b.energies())
        for w in (0, 31, 32, block - 1, block, 3 * block - 1, 3 * block, walkers - 1):
            ga, gb = a.walker(w), b.walker(w)
            assert (ga.energy, ga.accepted_moves, ga.rng_s0, ga.rng_s1) == (gb.energy, gb.accepted_moves, gb.rng_s0, gb.rng_s1), w
            ba, bb = a.bins(w), b.bins(w)
            for k in ba:
                assert np.array_equal(ba[k], bb[k]), (w, k)
    # and the streamed layout against the reference-order oracle, in the tolerance tier
    o = OracleMC(lj_cfg(N=N, R=R, lanes=1, n_walkers=walkers, method=method, **kw), walker=block)
    o.run(a.num_moves())
    g, s = a.walker(block), o.walker()
    if (g.rng_s0, g.rng_s1, g.accepted_moves) == (s.rng_s0, s.rng_s1, s.accepted_moves):  # same accept decisions over 5e4 moves
        assert abs(g.energy - s.energy) <= 1e-11 * abs(s.energy)


def test_lj_layout_follows_the_walker_count_and_method():
    """Without a forcing flag the engine takes the layout that needs less time for the walker count: waves of 384 (LJ31) / 320
    (LJ38) walkers per SM with z streamed from L2 against waves of 256 / 224 with everything in shared memory, weighted by the
    measured time of a wave (kernels_lj_thread_fast.cu).  Narrow bin windows: the engines are created, not run."""
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count

    def picks_stream(N, R, method, walkers, **kw):
        kw.setdefault("bin_window_lo", -2.0)
        eng = WalkerEngine(lj_cfg(N=N, R=R, lanes=1, method=method, n_walkers=walkers, flags=_abi.FLAG_FAST_MATH,
                                  init_mode=_abi.INIT_EXTERNAL, bin_window_hi=0.6, **kw))
        z = eng.streams_z()
        eng.close()
        return z

    assert picks_stream(31, 2.5, "sad", 384 * sms)            # bench.py's count: one wave of three CTAs
    assert picks_stream(31, 2.5, "sad", 2 * 384 * sms)        # 2 x 1.44 < 3
    assert not picks_stream(31, 2.5, "sad", 2 * 256 * sms)    # two waves either way: shared memory
    assert not picks_stream(31, 2.5, "sad", 4096)             # a partial wave either way
    assert picks_stream(38, 3.0, "inv-t-wl", 320 * sms, min_allowed_energy=-150.0, energy_bin=0.5, bin_window_lo=-151.0)  # 1.16 < 2
    assert not picks_stream(38, 3.0, "sad", 224 * sms)        # 1.40 > 1
    assert picks_stream(38, 3.0, "sad", 2 * 320 * sms)        # 2.8 < 3
    assert not picks_stream(13, 2.0, "sad", 384 * sms)        # other sizes have one layout
