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


# ---- exact bin weights of the analytic systems: what ln w of a converged flat-histogram run estimates ---------------
# A walker's ln w[i] converges to ln of the density of states INTEGRATED over bin i (plus a constant), so the gate
# compares with the integral over the bin, not with D at the bin centre (they differ in the half bins at E = 0 and
# E = 1 and wherever D varies across a bin, e.g. sqrt(E) near 0).

def fake_bin_weights(function, window_lo, width, n, dimensions=3):
    """Integral of the exact DOS of `--fake-linear` / `--fake-quadratic-dimensions d` over each window bin.

    Same densities as plotting/analyze-boundaries.py:22-25 (linear: 1 on [0, 1]; quadratic: (d/2) E^(d/2-1), the
    reference lists d = 3, 1.5 sqrt(E)), via their cumulative forms N(E) = E and N(E) = E^(d/2) on [0, 1]."""
    edges = np.clip(window_lo + np.arange(n + 1) * width, 0.0, 1.0)
    if function == "linear":
        cum = edges
    elif function == "quadratic":
        cum = edges ** (0.5 * dimensions)
    else:
        raise ValueError(function)
    return np.diff(cum)


def two_wells_energy(x1, rho2, h2_to_h1, r2, barrier_over_h1):
    """`TwoWells::find_energy` (src/system/two_wells.rs:266-315) on arrays: energy, and NaN where the move is forbidden."""
    r1 = 1.0
    rw = np.sqrt(barrier_over_h1) * 1.0 + r2 * np.sqrt(1.0 + barrier_over_h1 - 1.0 / h2_to_h1)
    x2 = x1 - r1 - r2
    xw = x1 - rw
    xi = x2 + rw
    d1, d2, dw, di = rho2 + x1 * x1, rho2 + x2 * x2, rho2 + xw * xw, rho2 + xi * xi
    e_1 = d1 / (r1 * r1) - 1.0
    e_2 = h2_to_h1 * (d2 / (r2 * r2) - 1.0)
    e_w = h2_to_h1 * (dw / (r2 * r2) - 1.0)
    e_i = di / (r1 * r1) - 1.0
    big = d1 <= r1 * r1
    small = ~big & (d2 <= r2 * r2)
    cyl = ~big & ~small & (rho2 <= r2 * r2) & (x1 > 0.0) & (x1 <= r1 + r2)
    e = np.full(np.broadcast(x1, rho2).shape, np.nan)
    e = np.where(big, np.minimum(e_1, e_w), e)
    e = np.where(small, np.where((e_i > e_2) & (e_i < 0.0), e_i, e_2), e)
    e = np.where(cyl, 0.0, e)
    return e


def two_wells_bin_weights(window_lo, width, n, N, h2_to_h1, r2):
    """Exact volume of configuration space per energy bin (E < 0) of the two-wells system, any barrier height.

    two-wells/system.py:86-90 writes D(e) as the sum of two hypersphere wells.  That sum is exact for the geometry of
    two_wells.rs:266-315, not an approximation: inside the big sphere the states below E are the UNION of the ball of
    radius sqrt(1 + E) about the origin and the ball of radius r2 sqrt(1 + E/h2) about `well_position`; inside the small
    sphere they are the INTERSECTION of the mirror-image pair (same radii, same separation), so the two lens volumes
    cancel and the cumulative volume is (1 + E)^(N/2) + r2^N (1 + E/h2)^(N/2) (times the unit N-ball's volume).  The
    zero-energy cylinder only adds weight to the bin that holds E = 0.  tests/test_analysis.py checks this against the
    direct quadrature below for barriers 0, 0.1, 0.2."""
    edges = np.minimum(window_lo + np.arange(n + 1) * width, 0.0)
    cum = np.clip(1.0 + edges, 0.0, None) ** (0.5 * N) + r2 ** N * np.clip(1.0 + edges / h2_to_h1, 0.0, None) ** (0.5 * N)
    w = np.diff(cum)
    w[window_lo + (np.arange(n) + 1) * width > 0.0] = 0.0  # the bin holding E = 0 also holds the cylinder: not gated
    return w


def two_wells_bin_weights_quadrature(window_lo, width, n, N, h2_to_h1, r2, barrier_over_h1, grid=6000):
    """The same volumes by direct quadrature of `find_energy` (for coarse bins: the midpoint rule aliases fine ones).

    The energy depends on x1 and rho^2 = sum of the other N - 1 squared coordinates only, so the N-dimensional
    volume element is rho^(N-2) d rho d x1 (constant factors drop out of an entropy).  Midpoint rule on a
    grid x grid mesh, binned exactly as `Bins::state_to_index` (energy.rs:371-373) bins energies."""
    x = -1.0 + (np.arange(grid) + 0.5) * ((2.0 + 2.0 * r2 + 1e-9) / grid)
    rho = (np.arange(grid) + 0.5) * (1.0 / grid)
    w = np.zeros(n)
    wt_rho = rho ** (N - 2)
    for i0 in range(0, grid, 500):
        xs = x[i0:i0 + 500, None]
        e = two_wells_energy(xs, (rho * rho)[None, :], h2_to_h1, r2, barrier_over_h1)
        ok = np.isfinite(e)
        idx = np.floor((e[ok] - window_lo) / width).astype(np.int64)
        wt = np.broadcast_to(wt_rho[None, :], e.shape)[ok]
        good = (idx >= 0) & (idx < n)
        w += np.bincount(idx[good], weights=wt[good], minlength=n)
    return w


def grouped_entropy(folds, groups, per_group, min_fraction=0.9):
    """Per-group walker-mean entropies from SAD-range-only folds: list of (S, mask) for g in range(groups).

    folds[g] is one `sadmc_fold` result for the interleaved group g (`sadmc_fold_select(g, groups, 1)`)."""
    out = []
    for g in range(groups):
        cnt = np.asarray(folds[g]["lnw_count"], dtype=np.float64)
        ok = cnt >= min_fraction * per_group
        S = np.zeros_like(cnt)
        S[ok] = np.asarray(folds[g]["lnw_sum"], dtype=np.float64)[ok] / cnt[ok]
        out.append((S, ok))
    return out


def dos_gate(folds, groups, per_group, weights, min_fraction=0.9, min_weight=0.0, centres=None, energy_range=None):
    """RMS of S - ln(exact bin weight) after removing the additive constant, per interleaved walker group.

    energy_range = (lo, hi) with `centres` (bin centres) restricts the gate to the bins whose centre lies inside.

    Returns dict(rms_mean, rms_sem, rms_all, n_bins, worst): `rms_mean`/`rms_sem` are the mean and the standard
    error over the groups' own RMS values (ensemble error bar); `rms_all` is the RMS of the all-walker mean entropy."""
    lnD = np.full(len(weights), -np.inf)
    pos = weights > min_weight
    lnD[pos] = np.log(weights[pos])
    per = grouped_entropy(folds, groups, per_group, min_fraction)
    common = np.isfinite(lnD)
    if energy_range is not None:
        c = np.asarray(centres, dtype=np.float64)
        common = common & (c >= energy_range[0]) & (c <= energy_range[1])
    for _, ok in per:
        common = common & ok
    rms = []
    Sall = np.zeros(len(weights))
    for S, _ in per:
        d = S[common] - lnD[common]
        d = d - d.mean()
        rms.append(float(np.sqrt(np.mean(d * d))))
        Sall[common] += S[common] / groups
    d = Sall[common] - lnD[common]
    d = d - d.mean()
    rms = np.array(rms)
    return {"rms_mean": float(rms.mean()), "rms_sem": float(rms.std(ddof=1) / np.sqrt(groups)) if groups > 1 else 0.0,
            "rms_all": float(np.sqrt(np.mean(d * d))), "n_bins": int(common.sum()),
            "worst": float(np.abs(d).max()) if common.any() else float("nan"), "residual": d, "mask": common}


def cv_low_edge_weight(folds, T, min_fraction=0.5, edge_bins=20):
    """How much of the canonical distribution at temperature T sits in the lowest `edge_bins` covered bins of the
    all-walker mean entropy: where this is not small, energies below the covered range (below the walkers' too_lo)
    would contribute and Cv(T) from the covered bins alone is not converged."""
    G = int(folds["groups"])
    cnt = sum(np.asarray(folds["lnw_count_%d" % g], dtype=np.float64) for g in range(G))
    tot = sum(np.asarray(folds["lnw_sum_%d" % g], dtype=np.float64) for g in range(G))
    ok = cnt >= min_fraction * int(folds["walkers"])
    nb = len(cnt)
    E = float(folds["window_lo"]) + (np.arange(nb) + 0.5) * float(folds["width"])
    S = tot[ok] / cnt[ok]
    Eo = E[ok]
    out = []
    for t in np.atleast_1d(T):
        a = S - Eo / t
        P = np.exp(a - a.max())
        P /= P.sum()
        out.append(P[:edge_bins].sum())
    return np.array(out), float(Eo.min())


def cv_from_production(run, T, min_count=1.0):
    """Cv(T) from a fixed-weight production run (tools/lj31_production.py): per interleaved walker group
    S_g(E) = ln w(E) + ln H_g(E) on the bins the group visited; ensemble mean, standard error, per-group curves."""
    w = np.asarray(run["weights"], dtype=np.float64)
    H = np.asarray(run["histogram_groups"], dtype=np.float64)
    nb = len(w)
    E = float(run["window_lo"]) + (np.arange(nb) + 0.5) * float(run["width"])
    cvs = []
    for g in range(H.shape[0]):
        ok = H[g] >= min_count
        cvs.append(heat_capacity(T, E[ok], w[ok] + np.log(H[g][ok])))
    cvs = np.array(cvs)
    return cvs.mean(0), cvs.std(0, ddof=1) / np.sqrt(H.shape[0]), cvs


def cv_production_vs_reference(run, ref_csv, T_min, T_max=np.inf):
    """plotting/final_heat_capacity.py:185-194 for a production run: (T, Cv, sem, Cv_ref, Err) at the curve's temperatures."""
    Tr, Cr = load_lj31_reference(ref_csv)
    m = (Tr >= T_min) & (Tr <= T_max)
    mean, sem, _ = cv_from_production(run, Tr[m])
    return Tr[m], mean, sem, Cr[m], (mean - Cr[m]) / Cr[m]
