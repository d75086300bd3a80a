"""Post-processing restated from the reference's plotting scripts: SAD entropy reconstruction and heat capacity."""
import os

import numpy as np
import pytest

from sad_monte_carlo_b200 import analysis


def test_heat_capacity_of_a_gaussian_density_of_states_is_its_variance():
    # S(E) = -(E - E0)^2 / (2 s^2)  ->  canonical P(E) is Gaussian with the same variance: C = s^2 / T^2
    E = np.linspace(-50, 50, 20001)
    s2 = 4.0
    S = -(E - 3.0) ** 2 / (2 * s2)
    T = np.array([0.5, 1.0, 2.0])
    C = analysis.heat_capacity(T, E, S)
    assert np.allclose(C, s2 / T ** 2, rtol=1e-6)


def test_sad_entropy_reconstruction_outside_the_range():
    # plotting/parse-binning.py:150-169: below too_lo the entropy continues with slope 1/min_T plus ln(hist/mean)
    E = analysis.bin_centres(-1.0, 0.1, 30)
    lnw = np.linspace(0, 5, 30)
    hist = np.full(30, 100.0)
    too_lo, too_hi, min_T = E[5], E[20], 0.5
    S = analysis.sad_excess_entropy(lnw, hist, E, too_lo, too_hi, min_T)
    inside = (E >= too_lo) & (E <= too_hi)
    ref = lnw.copy()
    ref[E < too_lo] = lnw[5] + (E[E < too_lo] - too_lo) / min_T
    ref[E > too_hi] = lnw[20]
    ref -= ref.max()
    assert np.allclose(S, ref)
    assert np.allclose(np.diff(S[inside]), np.diff(lnw[inside]))


def test_exact_dos_and_rms_error():
    E = analysis.bin_centres(0.0, 0.01, 100)
    D = analysis.fake_exact_dos("quadratic", E, dimensions=3)
    assert np.allclose(D, 1.5 * np.sqrt(E))
    S = np.log(D) + 7.0 + 1e-3 * np.sin(40 * E)
    rms, n = analysis.entropy_rms_error(S, E, D, np.ones_like(E, bool))
    assert n == 100 and rms < 1e-3
    assert np.all(analysis.fake_exact_dos("linear", E) == 1.0)


def test_merged_entropy_mean_and_standard_error():
    fold = {"lnw_count": np.array([4, 1, 0]), "lnw_sum": np.array([-4.0, -2.0, 0.0]),
            "lnw_sq_sum": np.array([4.0 + 4 * 0.25, 4.0, 0.0])}
    mean, err, ok = analysis.merged_entropy(fold)
    assert list(ok) == [True, True, False]
    assert np.allclose(mean[:2], [-1.0, -2.0])
    assert np.isclose(err[0], np.sqrt(0.25 / 3)) and err[1] == 0.0


# ---- LJ31 heat capacity from the committed GPU production run (tests/golden/lj31_cv_run, tools/lj31_cv_run.py) ----

def _cv_run(tag):
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lj31_cv_run", "lj31_cv_%s.npz" % tag))


def _lit(name):
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lj31_literature", name)


def test_lj31_heat_capacity_of_the_gpu_run_matches_the_reference_curves():
    """37 888 walkers x 1e8 SAD moves (min_T 0.1, energy bin 0.1) on one B200.  Stated tolerance: the reference's own
    error metric (plotting/final_heat_capacity.py:185-194, against tRem_Ref.csv) stays below 2 % for every
    literature temperature in [0.165, 0.40]; ensemble standard error below 0.3 %."""
    d = _cv_run("1e+08")
    T, cv, sem, ref, err = analysis.cv_error_vs_reference(d, _lit("tRem_Ref.csv"), 0.165, 0.40)
    assert len(T) >= 7
    assert np.abs(err).max() < 0.02, (T, err)
    assert (sem / cv).max() < 0.003
    # the RESTMC curve (same constraining radius 2.5 sigma, plotting/final_heat_capacity.py:66-67) agrees as well
    T, cv, sem, ref, err = analysis.cv_error_vs_reference(d, _lit("LJ31_Cv_Reference_alt.csv"), 0.165, 0.40)
    assert np.abs(err).max() < 0.025
    # LJ31_Cv_Reference.csv (REM) lies 10-15 % above t-REM / RESTMC around the melting peak -- the literature curves
    # disagree among themselves by that much; this run sits with t-REM / RESTMC
    T, cv, sem, ref, err = analysis.cv_error_vs_reference(d, _lit("LJ31_Cv_Reference.csv"), 0.165, 0.40)
    Tt, ct = analysis.load_lj31_reference(_lit("tRem_Ref.csv"))
    gap = (np.interp(T, Tt, ct) - ref) / ref
    assert np.abs(err - gap).max() < 0.03
    assert 0.08 < np.abs(err).max() < 0.17


def test_lj31_heat_capacity_converges_with_moves():
    errs = []
    for tag in ("2e+07", "5e+07", "1e+08"):
        T, cv, sem, ref, err = analysis.cv_error_vs_reference(_cv_run(tag), _lit("tRem_Ref.csv"), 0.165, 0.40)
        errs.append(np.abs(err).mean())
    assert errs[0] > errs[1] > errs[2]
    assert errs[2] < 0.01


# ---- exact bin weights of the analytic systems and the DOS gate (sad_monte_carlo_b200.analysis) ---------------------------

def test_fake_bin_weights_integrate_the_exact_dos_over_each_bin():
    lo, w, n = -0.025, 0.01, 106  # bins centred on multiples of the width: half bins at E = 0 and E = 1
    lin = analysis.fake_bin_weights("linear", lo, w, n)
    assert np.isclose(lin.sum(), 1.0) and np.isclose(lin[2], 0.005) and np.isclose(lin[50], 0.01) and lin[1] == 0.0
    quad = analysis.fake_bin_weights("quadratic", lo, w, n, 3)
    E = lo + (np.arange(n) + 0.5) * w
    inside = (E > 0.05) & (E < 0.95)
    assert np.allclose(quad[inside], 1.5 * np.sqrt(E[inside]) * w, rtol=2e-3)  # analyze-boundaries.py:24-25 at the centre
    assert np.isclose(quad.sum(), 1.0)


@pytest.mark.parametrize("barrier", [0.0, 0.1, 0.2])
def test_two_wells_closed_form_equals_quadrature_of_find_energy(barrier):
    """two-wells/system.py:86-90 (sum of two hypersphere wells) is EXACT for the geometry of two_wells.rs:266-315 at
    every barrier height: the lens the wells share inside the big sphere reappears, mirrored, in the small sphere."""
    N, h, r2 = 12, 1.1, 0.5
    lo, w, n = -1.125, 0.05, 23
    exact = analysis.two_wells_bin_weights(lo, w, n, N, h, r2)
    quad = analysis.two_wells_bin_weights_quadrature(lo, w, n, N, h, r2, barrier, grid=6000)
    m = (exact > 0) & (quad > 0)
    assert m.sum() >= 20
    d = np.log(quad[m]) - np.log(exact[m])
    d -= d[-3]
    assert np.abs(d).max() < 1.5e-3  # the midpoint rule's own error at this grid; 3e-4 at grid 8000


def test_dos_gate_recovers_a_planted_error():
    n, groups, per = 60, 4, 100
    weights = analysis.fake_bin_weights("quadratic", -0.025, 0.02, n, 3)
    rng = np.random.default_rng(3)
    folds = []
    for g in range(groups):
        S = np.where(weights > 0, np.log(np.where(weights > 0, weights, 1.0)), 0.0) + 7.0 + 0.01 * np.sin(np.arange(n))
        cnt = np.where(weights > 0, per, 0).astype(np.float64)
        cnt[40:] = 0.5 * per  # covered by half of the walkers only: not gated
        folds.append({"lnw_sum": (S + 1e-3 * rng.standard_normal(n)) * cnt, "lnw_count": cnt})
    r = analysis.dos_gate(folds, groups, per, weights)
    gated = r["mask"]
    assert gated.sum() == r["n_bins"] and not gated[40:].any() and gated[5:40].all()
    want = 0.01 * np.sin(np.arange(n))[gated]
    want = want - want.mean()
    assert abs(r["rms_all"] - np.sqrt(np.mean(want ** 2))) < 5e-4
    assert r["rms_sem"] < 1e-3


# ---- round 2: canonical cross-check and the run at the headline SAD parameters (tests/golden/lj31_cv_run_r02) -----------

def _r02(name):
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lj31_cv_run_r02", name)


def test_lj31_canonical_metropolis_on_the_engine_matches_restmc_and_the_sad_run():
    """Sampler-independent check of the physics: plain Metropolis sampling at fixed T on the same engine (18 944 walkers x
    1e7 moves per temperature, tools/lj31_canonical.py; Cv from the walkers' exact energy moments, no entropy, no bins).
    It reproduces RESTMC (`LJ31_Cv_Reference_alt.csv`, the curve with the reference's 2.5 sigma container) to 2 % and
    t-REM to 2 % for T in [0.2, 0.4], agrees with the round-1 SAD run (min_T 0.1, bin 0.1) to 1.5 %, and lies 2-13 % below
    `LJ31_Cv_Reference.csv` (REM) exactly where RESTMC / t-REM do: the gap to the CSV BASELINE.json names is between the
    literature curves (different constraining volume / constraint), not between the engine and the model."""
    import json
    canon = [c for c in json.load(open(_r02("r02_lj31_canonical.json"))) if c["T"] >= 0.2]  # T = 0.1 does not equilibrate from a quench
    T = np.array([c["T"] for c in canon])
    cv = np.array([c["Cv"] for c in canon])
    sem = np.array([c["Cv_sem"] for c in canon])
    assert len(T) == 5 and (sem / cv).max() < 1e-3
    for name, tol in (("LJ31_Cv_Reference_alt.csv", 0.02), ("tRem_Ref.csv", 0.02)):
        Tr, Cr = analysis.load_lj31_reference(_lit(name))
        err = (cv - np.interp(T, Tr, Cr)) / np.interp(T, Tr, Cr)
        assert np.abs(err).max() < tol, (name, err)
    Tr, Cr = analysis.load_lj31_reference(_lit("LJ31_Cv_Reference.csv"))
    err = (cv - np.interp(T, Tr, Cr)) / np.interp(T, Tr, Cr)
    Ta, Ca = analysis.load_lj31_reference(_lit("LJ31_Cv_Reference_alt.csv"))
    gap = (np.interp(T, Ta, Ca) - np.interp(T, Tr, Cr)) / np.interp(T, Tr, Cr)
    assert -0.16 < err.min() < -0.10 and np.abs(err - gap).max() < 0.02  # the literature's own gap, reproduced
    sad, _, _ = analysis.cv_from_grouped_folds(_cv_run("1e+08"), T)
    assert np.abs(sad / cv - 1.0).max() < 0.015


def test_lj31_sad_at_the_headline_parameters_is_still_converging_after_2e8_moves_per_walker():
    """run-lj-clusters.sh:53 parameters (min_T 0.01, energy bin 0.01: 13 365 bins), 37 888 walkers x 2e8 moves on one
    B200 (7.6e12 moves, 15 min).  Stated result, not a pass: the merged SAD entropy is converged only where every walker's
    range was established early -- Cv within 5 % of t-REM for T >= 0.32 -- and far off below (+83 % at T = 0.21), improving
    monotonically with moves per walker.  A SAD walker needs many round trips through its 13 000 bins; the reference's own
    runs use --max-iter 1e12 per walker, and no number of parallel walkers shortens one walker's learning."""
    errs = []
    for tag in ("1e+07", "3e+07", "1e+08", "2e+08"):
        d = np.load(_r02("lj31_cv_headline_%s.npz" % tag))
        T, cv, sem, ref, err = analysis.cv_error_vs_reference(d, _lit("tRem_Ref.csv"), 0.15, 0.41)
        errs.append((np.abs(err).max(), np.abs(err[T >= 0.32]).max()))
    worst = [e[0] for e in errs]
    assert worst[1] >= worst[2] >= worst[3] and worst[3] < 0.9 and worst[0] > 1.4   # 1.59 -> 1.53 -> 1.28 -> 0.835
    assert errs[3][1] < 0.05                                                          # T >= 0.32 at 2e8: 4.6 %
    d = np.load(_r02("lj31_cv_headline_2e+08.npz"))
    assert np.median(d["too_lo"]) < -132.8 and float(d["moves"]) == 2e8
    # against the CSV BASELINE.json names, same temperatures: the REM curve's own offset on top
    T, cv, sem, ref, err = analysis.cv_error_vs_reference(d, _lit("LJ31_Cv_Reference.csv"), 0.32, 0.41)
    assert np.abs(err).max() < 0.16


def test_lj31_replica_exchange_heat_capacity_matches_the_literature_curves():
    """tests/golden/lj31_tempering_r02/run.json: tools/lj31_tempering_cv.py on one B200 (1 184 simulations x 32 temperatures,
    2e7 moves per replica, 65 s).  Gates: within 2 % of t-REM and RESTMC for T in [0.045, 0.37]; within 1.5 % of
    LJ31_Cv_Reference.csv (the curve BASELINE.json names) for T in [0.045, 0.15], above which that curve departs from the
    other two by up to 12 % (profiles/r02_lj31_cv.md section 1).  The solid-solid feature below T = 0.04 is not equilibrated
    at this run length and is excluded."""
    import json
    run = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lj31_tempering_r02", "run.json")))
    T, cv, sem = np.array(run["T"]), np.array(run["Cv"]), np.array(run["Cv_sem"])
    assert run["sims"] * run["n_T"] == 37888 and run["moves_per_replica"] >= 1.9e7
    mid = (T >= 0.045) & (T <= 0.37)
    for name in ("tRem_Ref.csv", "LJ31_Cv_Reference_alt.csv"):
        ref = np.array(run["references"][name], float)
        err = np.abs(cv[mid] / ref[mid] - 1.0)
        assert err.max() <= 0.02, (name, err.max())
    low = (T >= 0.045) & (T <= 0.15)
    ref = np.array(run["references"]["LJ31_Cv_Reference.csv"], float)
    assert np.abs(cv[low] / ref[low] - 1.0).max() <= 0.015
    assert (sem[mid] / cv[mid]).max() <= 0.002  # ensemble error bars: 8 groups of 148 simulations


def test_lj31_replica_exchange_long_run_is_within_one_percent_of_the_literature():
    """tests/golden/lj31_tempering_r02/run_long.json: the same tool, 2.2e8 moves per replica (8.3e12 moves, 706 s of one B200).
    Within 1 % of t-REM and 0.5 % of RESTMC for T in [0.045, 0.37]; within 1.3 % of LJ31_Cv_Reference.csv for T in
    [0.045, 0.15].  Below T = 0.04 the ladder has not found the Mackay ground state (lowest energy seen -133.20 against
    -133.59): the solid-solid feature is not equilibrated and is excluded, as in the shorter run."""
    import json
    run = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lj31_tempering_r02", "run_long.json")))
    T, cv, sem = np.array(run["T"]), np.array(run["Cv"]), np.array(run["Cv_sem"])
    assert run["moves_per_replica"] > 2e8 and run["moves_per_s"] > 1e10
    mid = (T >= 0.045) & (T <= 0.37)
    trem = np.array(run["references"]["tRem_Ref.csv"], float)
    restmc = np.array(run["references"]["LJ31_Cv_Reference_alt.csv"], float)
    assert np.abs(cv[mid] / trem[mid] - 1.0).max() <= 0.01
    assert np.abs(cv[mid] / restmc[mid] - 1.0).max() <= 0.005
    low = (T >= 0.045) & (T <= 0.15)
    rem = np.array(run["references"]["LJ31_Cv_Reference.csv"], float)
    assert np.abs(cv[low] / rem[low] - 1.0).max() <= 0.013
    assert (sem[mid] / cv[mid]).max() <= 5e-4
