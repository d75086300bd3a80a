"""The reference's own self-consistency tests, restated against the CPU oracle.

The reference ships no golden vectors for this path; these are the tests it does have."""
import math

import numpy as np
import pytest

from sad_monte_carlo_b200._abi import make_config, INIT_REFERENCE
from tests.oracle_lib import OracleMC


@pytest.mark.parametrize("N", [2, 3, 10, 15, 137, 150])
def test_ising_cached_energy_equals_recomputed(N):
    # src/system/ising.rs:127-147 -- the MC rng there is seed 10137, 10^4 confirmed flips, exact equality
    mc = OracleMC(make_config("ising", N=N, seed=10137))
    assert mc.energy() == mc.compute_energy()
    for _ in range(10000 if N < 100 else 2000):
        mc.plan_move(0.0)
        mc.confirm()
        assert mc.energy() == mc.compute_energy()


def test_ising_initial_spins_come_from_seed_10137():
    mc = OracleMC(make_config("ising", N=32))
    s = mc.system()
    spins, E = s[:-1], s[-1]
    assert set(np.unique(spins)) == {-1.0, 1.0}
    # bit 0 of successive next_u64 of seed_from_u64(10137) (ising.rs:45-49)
    from tests.oracle_lib import rng_stream
    st = np.array([0x02edf8daf2766d59, 0x5d8be32458db11f2], np.uint64)
    bits = rng_stream(st, 0, 1024) & np.uint64(1)
    assert np.array_equal(spins, bits.astype(np.float64) * 2 - 1)
    # energy convention E = sum over right/down bonds of S*S' (ising.rs:59-75)
    g = spins.reshape(32, 32)  # index i + j*N  -> g[j, i]
    e = np.sum(g * np.roll(g, -1, axis=0)) + np.sum(g * np.roll(g, -1, axis=1))
    assert E == e


@pytest.mark.parametrize("N", [3, 50, 100, 200])
def test_lj_verify_energy_after_accepted_moves(N):
    # src/system/lj.rs:380-434: R = 10 N^(1/3), seed 1, scale 1.0, 1000 accepted moves
    radius = 5.0 * (2.0 * N ** (1.0 / 3.0))
    mc = OracleMC(make_config("lj", N=N, lj_radius=radius, seed=1))
    assert mc.energy() == mc.compute_energy()
    old_energy = mc.energy()
    maxe = N * 16.0
    i = 0.0
    while i < (1000.0 if N <= 100 else 300.0):
        newe = mc.plan_move(1.0)
        if newe is not None:
            if newe < maxe or newe < old_energy:
                mc.confirm()
                assert mc.verify_energy()
                assert abs(mc.energy() - mc.compute_energy()) <= abs(mc.energy()) * 1e-14 * N * N
                old_energy = newe
                i += 1.0
            else:
                i += 1e-6


@pytest.mark.parametrize("N,rho", [(50, 1.0), (50, 0.3), (100, 0.3)])
def test_wca_verify_energy_after_accepted_moves(N, rho):
    # src/system/wca.rs:647-694
    mc = OracleMC(make_config("wca", N=N, reduced_density=rho, seed=1))
    assert mc.verify_energy()
    old_energy = mc.energy()
    maxe = N * 16.0
    i = 0.0
    while i < 500.0:
        newe = mc.plan_move(1.0)
        if newe is not None:
            if newe < maxe or newe < old_energy:
                mc.confirm()
                assert mc.verify_energy()
                old_energy = newe
                i += 1.0
            else:
                i += 1e-6


def _min_image_energy_wca(state, L):
    pos = state[:-2].reshape(-1, 3)
    d = pos[:, None, :] - pos[None, :, :]
    d -= L * np.round(d / L)
    r2 = (d ** 2).sum(-1)
    iu = np.triu_indices(len(pos), 1)
    r2 = r2[iu]
    rc2 = 2.0 ** (1.0 / 3.0)
    s = 1.0 / r2[r2 < rc2]
    return float(np.sum(4.0 * (s ** 6 - s ** 3) + 1.0))


def test_wca_cell_list_finds_every_interacting_pair():
    # wca.rs:500-645 check that maybe_interacting_atoms misses nothing: compare with a numpy all-pairs minimum-image sum
    for N, rho in [(50, 1.0), (100, 0.3), (200, 0.8)]:
        mc = OracleMC(make_config("wca", N=N, reduced_density=rho, seed=3), attempts_override=50)
        L = (N / rho) ** (1.0 / 3.0)
        for _ in range(300):
            if mc.plan_move(0.5) is not None:
                mc.confirm()
        s = mc.system()
        e = _min_image_energy_wca(s, L)
        assert abs(e - mc.compute_energy()) <= 1e-11 * max(1.0, abs(e))
        assert abs(e - mc.energy()) <= 1e-10 * max(1.0, abs(e))


@pytest.mark.parametrize("N,ff", [(3, 0.1), (50, 0.3), (100, 0.3), (200, 0.3)])
def test_sw_cached_energy_equals_recomputed(N, ff):
    # src/system/optsquare.rs:585-618: plan_move + confirm 1000 times (confirm after None is a no-op), exact equality
    mc = OracleMC(make_config("sw", N=N, filling_fraction=ff, seed=1))
    assert mc.energy() == mc.compute_energy()
    for _ in range(1000):
        mc.plan_move(1.0)
        mc.confirm()
        assert mc.energy() == mc.compute_energy()
    # verify_energy == the slow all-pairs/all-images recount (optsquare.rs:108-152, 199-201)
    assert mc.verify_energy()


def test_sw_sad_trajectory_matches_all_pairs_recount():
    # tests/square-sad-test.rs: optsquare (cell list) vs square (all pairs) under 10^4 SAD moves, box 6 sigma,
    # defaults (N = 100, well 1.3, SAD min_T 0.2).  The all-pairs twin here is compute_energy_slowly.
    cfg = make_config("sw", cell_width=(6.0, 6.0, 6.0))
    mc = OracleMC(cfg)
    for k in range(50):
        mc.run(200)
        assert mc.verify_energy()
        e = mc.energy()
        assert e == round(e)
    w = mc.walker()
    assert w.moves == 10000 and w.accepted_moves > 100


def test_fake_linear_energy_is_radius_and_none_outside():
    mc = OracleMC(make_config("fake", fake_function=0, seed=5))
    assert mc.energy() == 0.0
    nones = 0
    for _ in range(5000):
        e = mc.plan_move(0.3)
        if e is None:
            nones += 1
        else:
            assert 0.0 <= e <= 1.0
            mc.confirm()
            assert mc.energy() == e == abs(mc.system()[0])
    assert nones > 100


def test_two_wells_regions_and_incremental_d_squared():
    cfg = make_config("two-wells", N=12, tw_h2_to_h1=1.1, tw_barrier_over_h1=0.1, tw_r2=0.5, seed=2)
    mc = OracleMC(cfg)
    s = mc.system()
    assert s[0] == -0.99 and np.all(s[1:12] == 0) and abs(s[12] - 0.99 ** 2) < 1e-16
    assert abs(mc.energy() - (0.99 ** 2 - 1.0)) < 1e-15
    for _ in range(20000):
        e = mc.plan_move(0.05)
        if e is not None:
            mc.confirm()
            assert mc.energy() == e
            s = mc.system()
            assert abs(s[12] - np.sum(s[:12] ** 2)) < 1e-12
            assert -1.1 - 1e-12 <= e <= 0.1 + 1e-12


def test_two_wells_rejects_bad_dimension():
    with pytest.raises(RuntimeError):
        OracleMC(make_config("two-wells", N=10, tw_h2_to_h1=1.1, tw_barrier_over_h1=0.0, tw_r2=0.5))


def test_ising_rejects_n1():
    with pytest.raises(RuntimeError):
        OracleMC(make_config("ising", N=1))
