#!/usr/bin/env python3
"""Heat capacity of the LJ31 production run against the literature curves the reference ships, as a markdown table.

    python tools/lj31_cv_report.py tests/golden/lj31_cv_run_r02/lj31_cv_headline_2e+08.npz [canonical.json]

Entropy = walker mean of the max-aligned ln w strictly inside each walker's SAD range (8 interleaved groups ->
ensemble error bars); Cv(T) = <(E - U)^2> / T^2 (plotting/final_heat_capacity.py:81-89); Err = (Cv - Cv_ref) / Cv_ref at the
reference curve's own temperatures (185-194)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sad_monte_carlo_b200 import analysis  # noqa: E402

LIT = os.path.join(ROOT, "tests", "golden", "lj31_literature")


def main():
    d = np.load(sys.argv[1])
    canon = json.load(open(sys.argv[2])) if len(sys.argv) > 2 else []
    print("run: %d walkers x %.1e moves, min_T %.3g, bin %.3g, range mode %s" % (int(d["walkers"]), float(d["moves"]), float(d["min_T"]),
                                                                               float(d["width"]), d["range_mode"] if "range_mode" in d else 1))
    print("too_lo of the sampled walkers: median %.3f, 10%% / 90%% quantiles %.3f / %.3f" % (
        np.median(d["too_lo"]), np.quantile(d["too_lo"], 0.1), np.quantile(d["too_lo"], 0.9)))
    for name in ("LJ31_Cv_Reference.csv", "tRem_Ref.csv", "LJ31_Cv_Reference_alt.csv"):
        T, cv, sem, ref, err = analysis.cv_error_vs_reference(d, os.path.join(LIT, name), 0.02, 0.45)
        edge, emin = analysis.cv_low_edge_weight(d, T)
        print("\n### against %s (lowest covered energy %.2f)\n" % (name, emin))
        print("| T | Cv (8-group mean) | s.e.m. | reference | Err | weight in the lowest 20 covered bins |")
        print("|---|---|---|---|---|---|")
        step = max(1, len(T) // 24)
        for i in range(0, len(T), step):
            print("| %.4f | %.2f | %.2f | %.2f | %+.1f %% | %.1e |" % (T[i], cv[i], sem[i], ref[i], 100 * err[i], edge[i]))
        ok = edge < 1e-3
        for lo, hi in ((0.05, 0.40), (0.10, 0.40), (0.165, 0.40)):
            m = (T >= lo) & (T <= hi) & ok
            if m.any():
                print("\nT in [%.3f, %.2f] (%d points, converged low edge): max |Err| %.1f %%, mean |Err| %.1f %%, max s.e.m./Cv %.2f %%" % (
                    lo, hi, m.sum(), 100 * np.abs(err[m]).max(), 100 * np.abs(err[m]).mean(), 100 * (sem[m] / cv[m]).max()))
    if canon:
        print("\n### canonical Metropolis cross-check (tools/lj31_canonical.py)\n")
        print("| T | Cv canonical | s.e.m. | Cv from the SAD entropy | s.e.m. | difference |")
        print("|---|---|---|---|---|---|")
        Tc = np.array([c["T"] for c in canon])
        mean, sem, _ = analysis.cv_from_grouped_folds(d, Tc)
        for c, m, s in zip(canon, mean, sem):
            print("| %.3f | %.2f | %.2f | %.2f | %.2f | %+.1f %% |" % (c["T"], c["Cv"], c["Cv_sem"], m, s, 100 * (m - c["Cv"]) / c["Cv"]))


if __name__ == "__main__":
    main()
