#!/usr/bin/env python3
"""Exact-DOS gate on the GPU (BASELINE.json config 2): 65 536 SAD walkers on the analytic test systems.

For `--fake-linear`, `--fake-quadratic-dimensions 3` (fake/run-fake.py:28-36,76-79: min_T 0.001, translation scale
0.05, energy bins 0.001 / 0.01 / 0.1) and two-wells "T-trans-1" (two-wells/run-two-wells.py:144-148) the converged
ln w of a SAD run must equal ln of the exact density of states integrated over each bin
(plotting/analyze-boundaries.py:22-40, two-wells/system.py:86-90) up to a constant.  The walkers are folded in G
interleaved groups (`sadmc_fold_select(g, G, sad_range_only=1)`), the entropy of a group is the walker mean of the
max-aligned ln w, and the gate is RMS(S - S_exact) over the bins covered by >= 90 % of every group's walkers,
reported with the ensemble spread over the groups.  One JSON line per (system, bin width, move count).

    python tools/dos_gate.py [--walkers 65536] [--schedule 1e5,1e6,1e7] [--out profiles/r02_dos_gate.jsonl]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sad_monte_carlo_b200 import WalkerEngine, make_config, _abi, analysis  # noqa: E402


def cases(which):
    out = []
    if "linear" in which:
        for de in (0.1, 0.01, 0.001):
            out.append(("fake-linear dE=%g" % de, dict(system="fake", fake_function=_abi.FAKE_LINEAR, energy_bin=de,
                                                       move_value=0.05, sad_min_T=0.001, bin_window_lo=-2 * de, bin_window_hi=1 + 2 * de),
                        lambda lo, w, n: analysis.fake_bin_weights("linear", lo, w, n)))
    if "quadratic" in which:
        for de in (0.1, 0.01, 0.001):
            out.append(("fake-quadratic d=3 dE=%g" % de, dict(system="fake", fake_function=_abi.FAKE_QUADRATIC, N=3, energy_bin=de,
                                                              move_value=0.05, sad_min_T=0.001, bin_window_lo=-2 * de, bin_window_hi=1 + 2 * de),
                        lambda lo, w, n: analysis.fake_bin_weights("quadratic", lo, w, n, 3)))
    if "two-wells-bigstep" in which:  # not a reference parameter set: a five times larger step, to see the same gate converge sooner
        out.append(("two-wells T-trans-1 barrier=0 dE=1e-3 scale=5e-2",
                    dict(system="two-wells", N=12, tw_h2_to_h1=1.1, tw_barrier_over_h1=0.0, tw_r2=0.5, sad_min_T=0.001,
                         energy_bin=1e-3, move_value=5e-2),
                    lambda lo, w, n: analysis.two_wells_bin_weights(lo, w, n, 12, 1.1, 0.5)))
    if "two-wells" in [x for x in which if x == "two-wells"]:
        for barrier in (0.0, 0.1):
            out.append(("two-wells T-trans-1 barrier=%g dE=1e-3 scale=1e-2" % barrier,
                        dict(system="two-wells", N=12, tw_h2_to_h1=1.1, tw_barrier_over_h1=barrier, tw_r2=0.5, sad_min_T=0.001,
                             energy_bin=1e-3, move_value=1e-2),
                        lambda lo, w, n: analysis.two_wells_bin_weights(lo, w, n, 12, 1.1, 0.5)))
    return out


def run_case(name, kw, weights_fn, walkers, schedule, groups, seed=0, settled=0.5, dump=None):
    kw = dict(kw)
    system = kw.pop("system")
    cfg = make_config(system, "sad", n_walkers=walkers, seed=seed, **kw)
    eng = WalkerEngine(cfg)
    lo, width, nb = eng.window()
    weights = weights_fn(lo, width, nb)
    lines = []
    done = 0
    t0 = time.time()
    for target in schedule:
        eng.run(target - done)
        done = target
        for frac in sorted({0.0, settled}):
            # only walkers whose SAD range has been unchanged since move frac * done take part (0 = all walkers)
            eng.fold_settled(int(frac * done))
            folds = []
            for g in range(groups):
                eng.fold_select(g, groups, 2)  # strictly inside each walker's SAD range
                folds.append(eng.fold())
            eng.fold_select(0, 1, False)
            eng.fold_settled(0)
            taking_part = int(sum(f["lnw_count"].max() for f in folds))
            r = analysis.dos_gate(folds, groups, taking_part / groups, weights)
            line = {"case": name, "walkers": walkers, "groups": groups, "moves_per_walker": done, "settled_since": frac,
                    "walkers_taking_part": taking_part, "rms_all_walkers": r["rms_all"], "rms_group_mean": r["rms_mean"],
                    "rms_group_sem": r["rms_sem"], "worst_bin": r["worst"], "bins_gated": r["n_bins"], "bins_in_window": int(nb),
                    "halted": list(eng.num_halted()), "wall_s": round(time.time() - t0, 2)}
            lines.append(line)
            print(json.dumps(line), flush=True)
            if dump:
                import numpy as np
                ws = [eng.walker(w) for w in range(0, walkers, max(1, walkers // 2048))]
                np.savez_compressed("%s_%s_%d_%g.npz" % (dump, name.split(" dE")[0].replace(" ", "_") + ("_dE%g" % width), done, frac),
                                    residual=r["residual"], mask=r["mask"], window_lo=lo, width=width,
                                    tL=np.array([w.tL for w in ws]), too_lo=np.array([w.too_lo for w in ws]),
                                    too_hi=np.array([w.too_hi for w in ws]), lnw_count=np.array([f["lnw_count"] for f in folds]))
    eng.close()
    return lines


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--walkers", type=int, default=65536)
    ap.add_argument("--schedule", default="1e5,1e6,1e7")
    ap.add_argument("--groups", type=int, default=8)
    ap.add_argument("--systems", default="linear,quadratic,two-wells")
    ap.add_argument("--out", default=None)
    ap.add_argument("--settled", type=float, default=0.5, help="also gate over the walkers whose SAD range is unchanged since this fraction of the run")
    ap.add_argument("--dump", default=None, help="prefix for npz files with the residuals and a sample of walker states")
    a = ap.parse_args()
    schedule = [int(float(x)) for x in a.schedule.split(",")]
    lines = []
    for name, kw, wf in cases(a.systems.split(",")):
        lines += run_case(name, kw, wf, a.walkers, schedule, a.groups, settled=a.settled, dump=a.dump)
    if a.out:
        with open(a.out, "w") as f:
            for ln in lines:
                f.write(json.dumps(ln) + "\n")


if __name__ == "__main__":
    main()
