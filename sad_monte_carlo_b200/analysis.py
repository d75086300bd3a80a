"""Post-processing of walker bins: SAD entropy reconstruction, heat capacity, exact densities of states.

Restates, for arrays that come back from the engine, what the reference's Python tools compute from a
checkpoint (the reference ships these only as plotting scripts):
  excess_entropy   plotting/parse-binning.py:150-169  (SAD: entropy outside [too_lo, too_hi] from the histogram)
  heat_capacity    plotting/final_heat_capacity.py:81-89
  exact DOS        plotting/analyze-boundaries.py:22-40 (fake systems)
  LJ31 references  plotting/final_heat_capacity.py:27-53 (CSV conventions)
Host-side numpy only; nothing here is on the hot path.
"""
import os

import numpy as np


def bin_centres(bins_min, width, n):
    """Bins::index_to_state, reference src/mc/energy.rs:366-370."""
    return bins_min + (np.arange(n) + 0.5) * width


def sad_excess_entropy(lnw, histogram, energies, too_lo, too_hi, min_T):
    """plotting/parse-binning.py:150-169 for one walker."""
    lnw = np.array(lnw, dtype=np.float64)
    hist = np.asarray(histogram, dtype=np.float64)
    E = np.asarray(energies, dtype=np.float64)
    i_lo = int(np.abs(E - too_lo).argmin())
    i_hi = int(np.abs(E - too_hi).argmin())
    mean_hist = hist[i_lo:i_hi + 1].mean()
    with np.errstate(divide="ignore", invalid="ignore"):
        lo = E < too_lo
        hi = E > too_hi
        lnw[lo] = lnw[i_lo] + (E[lo] - too_lo) / min_T + np.log(hist[lo] / mean_hist)
        lnw[hi] = lnw[i_hi] + np.log(hist[hi] / mean_hist)
    lnw[np.isnan(lnw)] = 0
    lnw[np.isinf(lnw)] = 0
    return lnw - lnw.max()


def heat_capacity(T, E, S):
    """plotting/final_heat_capacity.py:81-89: C(T) = <(E - U)^2> / T^2 with P ~ exp(S - E/T) on bin centres."""
    T = np.atleast_1d(np.asarray(T, dtype=np.float64))
    E = np.asarray(E, dtype=np.float64)
    S = np.asarray(S, dtype=np.float64)
    C = np.zeros_like(T)
    for i in range(len(T)):
        a = S - E / T[i]
        P = np.exp(a - a.max())
        P = P / P.sum()
        U = (E * P).sum()
        C[i] = ((E - U) ** 2 * P).sum() / T[i] ** 2
    return C


def merged_entropy(fold, min_walkers=1):
    """Walker-averaged, max-aligned ln w per window bin from a fold (sadmc_fold): mean, standard error, mask."""
    cnt = np.asarray(fold["lnw_count"], dtype=np.float64)
    ok = cnt >= max(1, min_walkers)
    mean = np.zeros_like(cnt)
    err = np.zeros_like(cnt)
    mean[ok] = fold["lnw_sum"][ok] / cnt[ok]
    var = np.zeros_like(cnt)
    var[ok] = np.maximum(fold["lnw_sq_sum"][ok] / cnt[ok] - mean[ok] ** 2, 0.0)
    many = cnt > 1
    err[many] = np.sqrt(var[many] / (cnt[many] - 1))
    return mean, err, ok


# ---- exact densities of states of the fake systems (plotting/analyze-boundaries.py:22-40) ------------------------

def fake_exact_dos(function, E, dimensions=3, a=None, b=None, e1=None, e2=None, sigma=None):
    E = np.asarray(E, dtype=np.float64)
    inside = (E > 0) & (E < 1)
    if function == "linear":
        return np.where(inside, 1.0, 0.0)
    if function == "quadratic":
        # volume of the d-ball below radius sqrt(E): D(E) = (d/2) E^(d/2 - 1); the reference lists d = 3 (1.5 sqrt(E))
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.where(inside, 0.5 * dimensions * np.power(np.where(inside, E, 1.0), 0.5 * dimensions - 1.0), 0.0)
    if function == "gaussian":
        with np.errstate(invalid="ignore", divide="ignore"):
            ok = (E > -1) & (E < 0)
            Es = np.where(ok, E, -0.5)
            return np.where(ok, (np.pi * sigma ** 3 * np.sqrt(32 * np.log(-1 / Es))) / -Es / (4 * np.pi / 3), 0.0)
    raise ValueError(function)


def entropy_rms_error(S, E, dos, mask):
    """RMS of S - ln D after removing the additive constant, over `mask` (bins inside the sampled range)."""
    with np.errstate(divide="ignore"):
        lnD = np.log(dos)
    m = mask & np.isfinite(lnD)
    d = S[m] - lnD[m]
    d = d - d.mean()
    return float(np.sqrt(np.mean(d ** 2))), int(m.sum())


# ---- LJ31 literature curves shipped with the reference ---------------------------------------------------------------

LJ31_REFERENCES = {
    # file: (conversion to absolute Cv, reference plotting/final_heat_capacity.py:27-53)
    "LJ31_Cv_Reference.csv": lambda c: (c - 1.5) * 31,
    "LJ31_Cv_Reference_3.csv": lambda c: (c - 1.5) * 31,
    "LJ31_Cv_Reference_4.csv": lambda c: (c - 1.5) * 31,
    "LJ31_Cv_Reference_alt.csv": lambda c: c,
    "tRem_Ref.csv": lambda c: c,
}


def load_lj31_reference(path):
    """(T, Cv) of one literature curve; `path` is one of the CSVs named in LJ31_REFERENCES."""
    T, c = np.loadtxt(path, delimiter=",", unpack=True)
    return T, LJ31_REFERENCES[os.path.basename(path)](c)


# ---- heat capacity of a folded multi-walker SAD run (tools/lj31_cv_run.py) -------------------------------------------

def cv_from_grouped_folds(folds, T, min_fraction=0.5):
    """Cv(T) per interleaved walker group from SAD-range-only folds, and the ensemble mean / standard error.

    `folds` is the mapping tools/lj31_cv_run.py saves: window_lo, width, walkers, groups and, per group g,
    lnw_sum_g / lnw_count_g (max-aligned ln w summed over the group's walkers, and how many walkers' SAD range
    covers each bin).  A bin takes part when at least `min_fraction` of the group's walkers cover it; the entropy
    is the walker mean of the aligned ln w (each walker is an independent SAD estimate of the same S(E)).
    """
    G = int(folds["groups"])
    nb = len(folds["lnw_sum_0"])
    E = float(folds["window_lo"]) + (np.arange(nb) + 0.5) * float(folds["width"])
    per_group = int(folds["walkers"]) // G
    cvs = []
    for g in range(G):
        cnt = np.asarray(folds["lnw_count_%d" % g], dtype=np.float64)
        ok = cnt >= min_fraction * per_group
        S = np.asarray(folds["lnw_sum_%d" % g], dtype=np.float64)[ok] / cnt[ok]
        cvs.append(heat_capacity(T, E[ok], S))
    cvs = np.array(cvs)
    return cvs.mean(0), cvs.std(0, ddof=1) / np.sqrt(G), cvs


def cv_error_vs_reference(folds, ref_csv, T_min, T_max=np.inf):
    """The reference's own error metric (plotting/final_heat_capacity.py:185-194): Err = (Cv - Cv_ref) / Cv_ref at
    the literature curve's own temperatures, restricted to [T_min, T_max].  Returns (T, Cv, sem, Cv_ref, Err)."""
    Tr, Cr = load_lj31_reference(ref_csv)
    m = (Tr >= T_min) & (Tr <= T_max)
    mean, sem, _ = cv_from_grouped_folds(folds, Tr[m])
    return Tr[m], mean, sem, Cr[m], (mean - Cr[m]) / Cr[m]
