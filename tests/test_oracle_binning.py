"""CPU checks of the `binning` oracle (oracle/oracle_binning.hpp): the reference's own trait test, the invariants
histogram.rs keeps, and convergence of SAD / WL / 1-t-WL / SAMC to the exact density of states of the fake systems."""
import numpy as np
import pytest

from sad_monte_carlo_b200 import _abi, make_config
from tests.oracle_lib import OracleBinningMC, OracleMC, load_oracle


def test_reference_trait_test_passes():
    # binning.rs:336-364 `test_binning`, instantiated for histogram::Bins at histogram.rs:113-116
    assert load_oracle().oracle_binning_reference_test() == 0


def test_reference_linear_tests_pass():
    # `test_linear` (linear.rs:167-186) including test_binning::<linear::Bins>
    assert load_oracle().oracle_binning_reference_test_linear() == 0


def test_linear_bins_interpolate_and_converge_to_the_flat_density():
    cfg = make_config("fake", "sad", fake_function=_abi.FAKE_LINEAR, sad_min_T=0.001, energy_bin=0.01, move_value=0.05,
                      flags=_abi.FLAG_BINNING | _abi.FLAG_BINNING_LINEAR)
    o = OracleBinningMC(cfg, walker=3)
    o2 = OracleBinningMC(cfg, walker=3)
    assert o2.walker().bins_len == 0
    o.run(3_000_001)
    s, b = o.walker(), o.bins_f64()
    assert abs(b["energy_count"].sum() - 3_000_001) < 1e-3  # each visit adds (1 - offset) + offset
    assert b["lnw_count"].min() >= 0.0 and s.lnw_max_count_f64 >= b["lnw_count"].max()
    lnw = b["lnw_total"][2:-2] * 0.01  # the totals are per unit energy (rescaled_gamma = gamma / width, linear.rs:250)
    assert len(lnw) >= 95 and lnw.std() < 0.05  # the flat density of states of fake-linear


def _linear(method, **kw):
    return make_config("fake", method, fake_function=_abi.FAKE_LINEAR, energy_bin=0.01, move_value=0.05, **kw)


def test_bins_are_empty_before_the_first_move_and_edges_sit_on_multiples_of_the_width():
    o = OracleBinningMC(_linear("sad", sad_min_T=0.001), walker=0)
    s = o.walker()
    assert s.bins_len == 0 and s.bins_min == -0.005  # Bins::new: (round(e / w) - 0.5) w, histogram.rs:171
    o.run(1)
    s = o.walker()
    assert s.bins_len >= 1 and s.bins_min == 0.0  # prep_for_e on empty vectors: floor(e / w) w, histogram.rs:149-151
    o.run(5000)
    s, b = o.walker(), o.bins()
    assert b["energy_count"].sum() == 5001 and s.lnw_total_count == 5001
    assert s.bins_min_e <= s.too_lo <= s.too_hi <= s.bins_max_e


def test_counts_are_zeroed_by_range_extensions_and_t_found_adds_up():
    # set_lnw zeroes the count of every rewritten bin (histogram.rs:196-197): lnw.count sums to less than the
    # moves, the "energy" accumulator keeps every visit, and a re-found bin adds its move number again
    o = OracleBinningMC(_linear("sad", sad_min_T=0.001), walker=2)
    o.run(200000)
    b, s = o.bins(), o.walker()
    assert b["energy_count"].sum() == 200000
    assert b["lnw_count"].sum() < 200000
    assert b["t_found_count"].max() > 1
    assert s.tF == s.t_found_max_total == b["t_found_total"].max()
    assert s.bins_len - 3 <= s.num_states <= s.bins_len  # (nearly) every bin centre of [0, 1) lies inside [too_lo, too_hi]


@pytest.mark.parametrize("method,kw,tol", [
    ("sad", dict(sad_min_T=0.001), 0.15),
    ("samc", dict(samc_t0=1e4), 0.15),
    ("wl", dict(min_allowed_energy=0.0, max_allowed_energy=0.999), 0.15),
    ("inv-t-wl", dict(min_allowed_energy=0.0, max_allowed_energy=0.999), 0.15),
])
def test_entropy_of_fake_linear_converges_to_the_exact_flat_density(method, kw, tol):
    # D(E) = 1 on [0, 1] (plotting/analyze-boundaries.py:22-23): ln w must become flat
    o = OracleBinningMC(_linear(method, **kw), walker=5)
    o.run(3_000_000)
    b = o.bins()
    lnw = b["lnw_total"][1:-1]
    assert len(lnw) >= 97
    assert lnw.std() < tol, (method, lnw.std())


def test_wl_gamma_halves_on_flatness_and_inv_t_wl_switches_to_samc():
    o = OracleBinningMC(_linear("wl", min_allowed_energy=0.0, max_allowed_energy=0.999, wl_min_gamma=1e-3), walker=1)
    o.run(2_000_000)
    s = o.walker()
    assert s.wl_gamma == 0.0 and s.method == _abi.METHOD_WL  # production run, energy_binning.rs:477-482
    o = OracleBinningMC(_linear("inv-t-wl", min_allowed_energy=0.0, max_allowed_energy=0.999), walker=1)
    o.run(2_000_000)
    s = o.walker()
    assert s.method == _abi.METHOD_SAMC and s.samc_t0 == s.bins_len  # energy_binning.rs:489-498


def test_same_proposals_as_the_histogram_oracle_until_the_first_difference_in_weights():
    # both Monte Carlos draw from the same generator in the same order; with SAMC and a huge t0 (gamma = 1 for all
    # moves, weights grow identically bin for bin when the bin edges coincide) the trajectories agree move for move
    kw = dict(fake_function=_abi.FAKE_QUADRATIC, N=3, energy_bin=0.25, move_value=0.05, samc_t0=1e12, seed=4)
    a = OracleBinningMC(make_config("fake", "samc", **kw), walker=0)
    a.run(20000)
    s = a.walker()
    assert s.accepted_moves > 0 and s.lnw_total_count == 20000
