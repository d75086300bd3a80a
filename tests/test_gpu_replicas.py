"""GPU parity for the energy-ceiling replicas (`sadmc_replicas_*`, csrc/replicas.cuh + replicas_round.cuh) against the CPU
restatement of the reference's `replicas` binary (src/mc/energy_replicas.rs; oracle/oracle_replicas.hpp): after the set-up
sweep and after rounds of moves, swaps, median updates and replica splits -- every replica's ceilings, counters, moments,
step size, generator and configuration, the simulation's generator and its median estimator."""
import numpy as np
import pytest

from sad_monte_carlo_b200 import _abi, make_config
from sad_monte_carlo_b200.replicas import ReplicasMC
from tests.oracle_lib import OracleReplicas

pytestmark = pytest.mark.gpu

FIELDS = ["max_energy", "cutoff_energy", "lowest_max_energy", "translation_scale", "rejected_count", "accepted_count", "above_count",
          "below_count", "upwelling_count", "unique_visitors", "above_total", "below_total", "above_total_squared", "below_total_squared",
          "above_extra_total", "above_extra_count", "collecting_data", "rng_s0", "rng_s1", "energy"]


def assert_sim_equal(mc, sim, o, context="", exact=True, rtol=1e-12):
    assert mc.num_replicas(sim) == o.num_replicas(), "%s: %d replicas on the GPU, %d in the oracle" % (context, mc.num_replicas(sim), o.num_replicas())
    assert mc.moves(sim) == o.moves(), context
    assert mc.rng(sim) == o.rng(), context
    gm, om = mc.median(sim), o.median()
    assert len(gm) == len(om) and (np.array_equal(gm, om) if exact else np.allclose(gm, om, rtol=rtol, atol=0)), context
    for r, (g, s) in enumerate(zip(mc.replicas(sim), o.replicas())):
        for f in FIELDS:
            a, b = getattr(g, f), getattr(s, f)
            if exact or isinstance(a, int):
                assert a == b, "%s sim %d replica %d: %s gpu=%r oracle=%r" % (context, sim, r, f, a, b)
            else:
                assert a == b or abs(a - b) <= rtol * max(1.0, abs(b)), "%s sim %d replica %d: %s gpu=%r oracle=%r" % (context, sim, r, f, a, b)
        if exact:
            assert np.array_equal(mc.system(sim, r), o.system(r)), "%s sim %d replica %d: system differs" % (context, sim, r)


def _check(cfg, rounds, sims, min_T=0.001, indep=8, max_replicas=48, max_init=512, exact=True, rtol=1e-12):
    mc = ReplicasMC(cfg, min_T, indep, max_replicas, max_init)
    oracles = {k: OracleReplicas(cfg, min_T, indep, sim=cfg.walker_offset + k, max_init=max_init) for k in sims}
    for k, o in oracles.items():
        assert_sim_equal(mc, k, o, "init", exact, rtol)
    for n in rounds:
        mc.run_once(n)
        for k, o in oracles.items():
            o.run_once(n)
            assert_sim_equal(mc, k, o, "after %d more rounds" % n, exact, rtol)
    return mc


@pytest.mark.parametrize("fn,kw", [(_abi.FAKE_LINEAR, {}), (_abi.FAKE_QUADRATIC, dict(N=3)), (_abi.FAKE_GAUSSIAN, dict(fake_sigma=0.3)),
                                   (_abi.FAKE_PIECES, dict(fake_a=0.1, fake_b=0.2, fake_e1=1.0, fake_e2=0.5))])  # fake/run-fake.py:60-64
def test_fake_systems_split_off_replicas_like_the_reference(fn, kw):
    cfg = make_config("fake", fake_function=fn, n_walkers=20, seed=3, **kw)
    mc = _check(cfg, [1, 50, 30000], sims=(0, 19))
    assert mc.num_replicas(7) > 4  # the ladder has grown


def test_reference_max_init_and_default_threshold():
    # MAX_INIT = 1 << 15 randomized energies and 64 independent systems before a new bin (energy_replicas.rs:350, 390)
    cfg = make_config("fake", fake_function=_abi.FAKE_QUADRATIC, N=3, n_walkers=3, seed=1)
    _check(cfg, [20000], sims=(1,), indep=64, max_init=0)


def test_ising():
    cfg = make_config("ising", N=8, n_walkers=6, seed=2)
    _check(cfg, [1, 600], sims=(0, 5), min_T=2.0, indep=8, max_replicas=128)


def test_erfinv():
    cfg = make_config("fake-erfinv", N=3, erfinv_mean_energy=0.0, n_walkers=8, seed=5)
    _check(cfg, [1, 5000], sims=(0, 7), min_T=0.1, indep=8, exact=False, rtol=1e-11)


def test_lj13_reference_order():
    cfg = make_config("lj", N=13, lj_radius=2.0, n_walkers=4, seed=7, lanes_per_walker=1)
    _check(cfg, [1, 800], sims=(0, 3), min_T=0.1, indep=4, max_init=256)


def test_wca_with_the_pressure_extra():
    cfg = make_config("wca", N=20, reduced_density=0.3, n_walkers=2, seed=4, lanes_per_walker=32)
    mc = ReplicasMC(cfg, 0.5, 4, 16, 128)
    o = OracleReplicas(cfg, 0.5, 4, sim=1, max_init=128)
    mc.run_once(450)  # 450 x 20 moves: pressure sampled every N^2 = 400 moves above the cutoff
    o.run_once(450)
    assert mc.num_replicas(1) == o.num_replicas() and mc.rng(1) == o.rng() and mc.moves(1) == o.moves()
    for g, s in zip(mc.replicas(1), o.replicas()):
        assert (g.accepted_count, g.rejected_count, g.above_count, g.below_count, g.unique_visitors, g.above_extra_count, g.rng_s0) == (
            s.accepted_count, s.rejected_count, s.above_count, s.below_count, s.unique_visitors, s.above_extra_count, s.rng_s0)
        assert abs(g.energy - s.energy) <= 1e-11 * max(1.0, abs(s.energy)) and abs(g.above_total - s.above_total) <= 1e-10 * max(1.0, abs(s.above_total))


def test_no_free_slot_is_reported_and_unsupported_systems_are_refused():
    cfg = make_config("fake", fake_function=_abi.FAKE_QUADRATIC, N=3, n_walkers=4, seed=3)
    mc = ReplicasMC(cfg, 0.001, 4, 3, 256)
    with pytest.raises(Exception) as ei:
        mc.run_once(20000)
    assert "max_replicas" in str(ei.value)
    with pytest.raises(Exception):
        ReplicasMC(make_config("two-wells", N=12, tw_h2_to_h1=1.1, tw_barrier_over_h1=0.1, tw_r2=0.5, n_walkers=2), 0.001)
