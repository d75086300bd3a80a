"""GPU parity for replica exchange (`sadmc_tempering_*`, csrc/tempering.cuh) against the CPU restatement of the
reference's `tempering` binary (src/mc/tempering.rs; oracle/oracle_tempering.hpp): every replica's counters, energy
moments, generator state and configuration, and the simulation's own generator, after canonical sweeps and swaps."""
import numpy as np
import pytest

from sad_monte_carlo_b200 import _abi, make_config
from sad_monte_carlo_b200.tempering import TemperingMC, geometric_spacing
from tests.oracle_lib import OracleMC, OracleTempering

pytestmark = pytest.mark.gpu

FIELDS = ["T", "rejected_count", "accepted_count", "rejected_swap_count", "accepted_swap_count", "ignored_count",
          "total_energy", "total_energy_squared", "translation_scale", "rng_s0", "rng_s1", "energy"]


def assert_sim_equal(mc, sim, o, context="", exact=True, rtol=1e-12):
    assert mc.rng(sim) == o.rng(), context
    for r, (g, s) in enumerate(zip(mc.replicas(sim), o.replicas())):
        for f in FIELDS:
            a, b = getattr(g, f), getattr(s, f)
            if exact or isinstance(a, int):
                assert a == b, "%s sim %d replica %d: %s gpu=%r oracle=%r" % (context, sim, r, f, a, b)
            else:
                assert abs(a - b) <= rtol * max(1.0, abs(b)), "%s sim %d replica %d: %s gpu=%r oracle=%r" % (context, sim, r, f, a, b)
        if exact:
            assert np.array_equal(mc.system(sim, r), o.system(r)), "%s sim %d replica %d: system differs" % (context, sim, r)


def _check(cfg, T, can_steps, rounds, sims, state=None, exact=True):
    mc = TemperingMC(cfg, T, can_steps)
    if state is not None:
        for k in range(mc.n_sim):
            for r in range(mc.n_T):
                mc.set_system(k, r, state)
    oracles = {k: OracleTempering(cfg, T, can_steps, sim=cfg.walker_offset + k, system_state=state) for k in sims}
    for k, o in oracles.items():
        assert_sim_equal(mc, k, o, "init", exact)
    for n in rounds:
        mc.run_once(n)
        for k, o in oracles.items():
            o.run_once(n)
            assert mc.moves == o.moves
            assert_sim_equal(mc, k, o, "after %d moves" % mc.moves, exact)
    return mc


def test_two_wells_ladder_as_the_job_script_runs_it():
    # two-wells/run-two-wells.py:45-61, 204: geometric ladder, --canonical-steps 10, system T-trans-1
    cfg = make_config("two-wells", N=12, tw_h2_to_h1=1.1, tw_barrier_over_h1=0.1, tw_r2=0.5, n_walkers=33, seed=2)
    mc = _check(cfg, geometric_spacing(0.001, 1.0, 10), 10, [1, 3, 400], sims=(0, 17, 32))
    reps = mc.replicas(5)
    assert sum(r.accepted_swap_count + r.rejected_swap_count for r in reps) > 0
    assert all(r.translation_scale == 1.0 for r in reps)
    assert mc.steps_per_round == 12 * 10 and mc.moves == 404 * 120 * 10


@pytest.mark.parametrize("n_T", [1, 2, 3, 7])
def test_odd_and_even_ladders_pair_up_like_chunks_exact_mut(n_T):
    cfg = make_config("fake", fake_function=_abi.FAKE_QUADRATIC, N=3, n_walkers=16, seed=11)
    _check(cfg, [0.05 * 2 ** i for i in range(n_T)], 2, [500], sims=(0, 15))


@pytest.mark.parametrize("fn,kw", [(_abi.FAKE_LINEAR, {}), (_abi.FAKE_GAUSSIAN, dict(fake_sigma=0.3)),
                                   (_abi.FAKE_PIECES, dict(fake_a=0.1, fake_b=0.5, fake_e1=2.0, fake_e2=1.0))])
def test_fake_systems(fn, kw):
    cfg = make_config("fake", fake_function=fn, n_walkers=20, seed=4, **kw)
    _check(cfg, [0.01, 0.03, 0.1, 0.3, 1.0], 5, [2000], sims=(0, 19))


def test_ising():
    cfg = make_config("ising", N=8, n_walkers=12, seed=1)
    mc = _check(cfg, [1.0, 1.5, 2.0, 2.5, 3.0, 4.0], 1, [60], sims=(0, 11))
    assert mc.steps_per_round == 64  # min_moves_to_randomize = N^2 (ising.rs:86-88)


def _lj_state(N, R):
    o = OracleMC(make_config("lj", "canonical", N=N, lj_radius=R, canonical_T=0.3, energy_bin=1e9, move_value=0.05, seed=5,
                             init_mode=_abi.INIT_RANDOMIZE, max_allowed_energy=0.0))
    o.run(3000)
    return o.system()


def test_lj13_reference_order_arithmetic():
    state = _lj_state(13, 2.0)
    cfg = make_config("lj", N=13, lj_radius=2.0, n_walkers=8, seed=9, lanes_per_walker=1, init_mode=_abi.INIT_EXTERNAL)
    _check(cfg, geometric_spacing(0.05, 0.5, 6), 3, [1, 150], sims=(0, 7), state=state)


def test_lj13_with_per_temperature_translation_scales():
    # Replica::translation_scale is serialised state (tempering.rs:71-72); the constructor's 1.0 accepts next to nothing
    # for a cluster, a step that shrinks with temperature does
    state = _lj_state(13, 2.0)
    cfg = make_config("lj", N=13, lj_radius=2.0, n_walkers=8, seed=9, lanes_per_walker=1, init_mode=_abi.INIT_EXTERNAL)
    T = geometric_spacing(0.05, 0.5, 6)
    scales = [0.1 * np.sqrt(t) for t in T]
    mc = TemperingMC(cfg, T, 3)
    mc.set_translation_scales(scales)
    for k in range(mc.n_sim):
        for r in range(mc.n_T):
            mc.set_system(k, r, state)
    o = OracleTempering(cfg, T, 3, sim=5, system_state=state)
    o.set_translation_scales(scales)
    mc.run_once(200)
    o.run_once(200)
    assert_sim_equal(mc, 5, o, "scaled")
    reps = mc.replicas(5)
    assert all(r.accepted_count > 0.1 * (r.accepted_count + r.rejected_count) for r in reps)
    assert [r.translation_scale for r in reps] == scales


def test_lj31_fast_math_tier():
    state = _lj_state(31, 2.5)
    cfg = make_config("lj", N=31, lj_radius=2.5, n_walkers=40, seed=9, lanes_per_walker=1, init_mode=_abi.INIT_EXTERNAL,
                      flags=_abi.FLAG_FAST_MATH)
    ocfg = make_config("lj", N=31, lj_radius=2.5, n_walkers=40, seed=9, lanes_per_walker=1, init_mode=_abi.INIT_EXTERNAL)
    T = geometric_spacing(0.05, 0.5, 5)
    mc = TemperingMC(cfg, T, 1)
    for k in range(mc.n_sim):
        for r in range(mc.n_T):
            mc.set_system(k, r, state)
    mc.run_once(60)
    for k in (0, 39):
        o = OracleTempering(ocfg, T, 1, sim=k, system_state=state)
        o.run_once(60)
        assert mc.rng(k) == o.rng()
        for g, s in zip(mc.replicas(k), o.replicas()):
            assert (g.rng_s0, g.rng_s1, g.accepted_count, g.rejected_count, g.accepted_swap_count) == (
                s.rng_s0, s.rng_s1, s.accepted_count, s.rejected_count, s.accepted_swap_count)
            assert abs(g.energy - s.energy) <= 1e-12 * abs(s.energy)
            assert abs(g.total_energy - s.total_energy) <= 1e-11 * abs(s.total_energy)


def test_square_well():
    cfg = make_config("sw", N=50, filling_fraction=0.3, sw_well_width=1.3, n_walkers=6, seed=3)
    _check(cfg, [0.5, 1.0, 2.0, 4.0], 1, [40], sims=(0, 5))


def test_canonical_energies_order_with_temperature():
    # physics sanity on the GPU alone: <E> rises with T for the two-wells ladder, averaged over 256 simulations
    cfg = make_config("two-wells", N=12, tw_h2_to_h1=1.1, tw_barrier_over_h1=0.0, tw_r2=0.5, n_walkers=256, seed=1)
    T = geometric_spacing(0.01, 1.0, 8)
    mc = TemperingMC(cfg, T, 10)
    mc.run_once(3000)
    e = np.mean([mc.mean_energy(k)[0] for k in range(0, 256, 8)], axis=0)
    assert np.all(np.diff(e) > 0), e


def test_bad_arguments():
    cfg = make_config("two-wells", N=12, tw_h2_to_h1=1.1, tw_barrier_over_h1=0.1, tw_r2=0.5, n_walkers=2)
    with pytest.raises(Exception):
        TemperingMC(cfg, [0.1, -1.0], 1)
    with pytest.raises(Exception):  # lane-group kernels carry no tempering kernel
        TemperingMC(make_config("lj", N=13, lj_radius=2.0, n_walkers=2, lanes_per_walker=8), [0.1, 0.2], 1)
