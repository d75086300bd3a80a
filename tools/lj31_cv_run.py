#!/usr/bin/env python3
"""LJ31 SAD production run on one GPU for the heat-capacity check (BASELINE.json config 3).

Runs the bench workload (reference run-lj-clusters.sh:53 parameters) for a schedule of move counts and, at each
checkpoint, folds the walkers in G interleaved groups (SAD range only) into gpurun_out/lj31_cv_<moves>.npz.
Analysis (entropy -> Cv(T), comparison with LJ31_Cv_Reference.csv) is sad_monte_carlo_b200.analysis.cv_report.

    python tools/lj31_cv_run.py --walkers 75776 --schedule 1e6,3e6,1e7 --groups 8
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sad_monte_carlo_b200 import WalkerEngine, make_config, _abi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--walkers", type=int, default=75776)
    ap.add_argument("--schedule", default="1e6,3e6,1e7")
    ap.add_argument("--groups", type=int, default=8)
    ap.add_argument("--min-T", type=float, default=0.01)
    ap.add_argument("--energy-bin", type=float, default=0.01)
    ap.add_argument("--chunk", type=float, default=1e6)
    ap.add_argument("--exact", action="store_true")
    ap.add_argument("--range-mode", type=int, default=2, help="1: bins inside [too_lo, too_hi]; 2: strictly inside (without the half-updated end bins)")
    ap.add_argument("--settled", type=float, default=0.5, help="also fold only the walkers whose SAD range is unchanged since this fraction of the run (0: skip)")
    ap.add_argument("--tag", default="lj31_cv")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out"))
    a = ap.parse_args()
    cfg = make_config("lj", "sad", N=31, lj_radius=2.5, max_allowed_energy=0.0, sad_min_T=a.min_T, energy_bin=a.energy_bin,
                      move_value=0.05, n_walkers=a.walkers, init_mode=_abi.INIT_RANDOMIZE, lanes_per_walker=1, seed=0,
                      flags=0 if a.exact else _abi.FLAG_FAST_MATH, bin_window_lo=-133.62, bin_window_hi=0.02)
    eng = WalkerEngine(cfg)
    lo, width, nb = eng.window()
    os.makedirs(a.out, exist_ok=True)
    t0 = time.time()
    done = 0
    age_log = []  # (walker age at the end of the chunk, moves/s of the chunk's launch): throughput against walker age
    for target in [int(float(x)) for x in a.schedule.split(",")]:
        while done < target:
            n = int(min(a.chunk, target - done))
            eng.run(n)
            done += n
            age_log.append((done, a.walkers * n / (eng.last_run_ms() * 1e-3)))
        out = {"window_lo": lo, "width": width, "moves": done, "walkers": a.walkers, "groups": a.groups, "min_T": a.min_T,
               "range_mode": a.range_mode, "settled": a.settled, "age_log": np.array(age_log)}
        for frac, suffix in ((0.0, ""), (a.settled, "_settled")):
            if suffix and not a.settled:
                continue
            eng.fold_settled(int(frac * done))
            for g in range(a.groups):
                eng.fold_select(g, a.groups, a.range_mode)
                f = eng.fold()
                for k in ("histogram", "lnw_sum", "lnw_sq_sum", "lnw_count"):
                    if suffix and k == "histogram":
                        continue
                    out["%s%s_%d" % (k, suffix, g)] = f[k]
        eng.fold_settled(0)
        eng.fold_select(0, 1, False)
        ws = [eng.walker(w) for w in range(0, a.walkers, max(1, a.walkers // 256))]
        out["too_lo"] = np.array([w.too_lo for w in ws])
        out["too_hi"] = np.array([w.too_hi for w in ws])
        out["energy"] = np.array([w.energy for w in ws])
        out["status"] = np.array([w.status for w in ws])
        out["tL"] = np.array([w.tL for w in ws])
        np.savez_compressed(os.path.join(a.out, "%s_%.0e.npz" % (a.tag, done)), **out)
        print("moves/walker %.1e  wall %.1f s  %.3g moves/s (last chunk)  too_lo median %.2f min %.2f  too_hi median %.2f  settled %.0f%%  halted %s" % (
            done, time.time() - t0, age_log[-1][1], np.median(out["too_lo"]), out["too_lo"].min(), np.median(out["too_hi"]),
            100.0 * np.mean(out["tL"] <= a.settled * done), eng.num_halted()), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
