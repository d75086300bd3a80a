"""Post-processing restated from the reference's plotting scripts: SAD entropy reconstruction and heat capacity."""
import numpy as np

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
