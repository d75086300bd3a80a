#!/usr/bin/env python3
"""LJ31 heat capacity by plain canonical Metropolis sampling on the GPU engine -- a sampler-independent check of the
flat-histogram (SAD) result.

The same system (`--lj-N 31 --lj-radius 2.5`, translation scale 0.05) with `Method::Canonical` (reference
src/mc/energy.rs:504-510) at fixed temperatures: Cv(T) = (<E^2> - <E>^2) / T^2 from the walkers' exact energy moments
(`energy_total`, `energy_squared_total` of the bins, summed by the device fold), second half of the run only, with the
spread over interleaved walker groups as error bar.  No entropy estimate, no reweighting, no bins involved.

    python tools/lj31_canonical.py --temperatures 0.2,0.25,0.3,0.35 --walkers 37888 --moves 2e7 --out gpurun_out/lj31_canonical.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sad_monte_carlo_b200 import WalkerEngine, make_config, _abi  # noqa: E402


def moments(eng, groups):
    out = []
    for g in range(groups):
        eng.fold_select(g, groups, 0)
        f = eng.fold()
        out.append((float(f["histogram"].sum()), float(f["energy_total"].sum()), float(f["energy_squared_total"].sum())))
    eng.fold_select(0, 1, 0)
    return np.array(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--temperatures", default="0.2,0.25,0.3,0.35")
    ap.add_argument("--walkers", type=int, default=37888)
    ap.add_argument("--moves", type=float, default=2e7, help="per walker; the first half is equilibration")
    ap.add_argument("--groups", type=int, default=8)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    half = int(a.moves) // 2
    res = []
    for T in [float(x) for x in a.temperatures.split(",")]:
        cfg = make_config("lj", "canonical", N=31, lj_radius=2.5, max_allowed_energy=0.0, canonical_T=T, energy_bin=0.1, move_value=0.05,
                          n_walkers=a.walkers, init_mode=_abi.INIT_RANDOMIZE, lanes_per_walker=1, seed=0, flags=_abi.FLAG_FAST_MATH,
                          bin_window_lo=-134.0, bin_window_hi=0.3)
        eng = WalkerEngine(cfg)
        t0 = time.time()
        eng.run(half)
        m1 = moments(eng, a.groups)
        eng.run(half)
        m2 = moments(eng, a.groups) - m1  # the second half only (the sums are cumulative)
        n, s1, s2 = m2[:, 0], m2[:, 1], m2[:, 2]
        U = s1 / n
        cv = (s2 / n - U * U) / (T * T)
        line = {"T": T, "Cv": float(cv.mean()), "Cv_sem": float(cv.std(ddof=1) / np.sqrt(a.groups)), "U": float(U.mean()),
                "U_sem": float(U.std(ddof=1) / np.sqrt(a.groups)), "walkers": a.walkers, "moves_per_walker": 2 * half,
                "measured_over": "second half", "groups": a.groups, "accepted_fraction": eng.num_accepted_moves() / (a.walkers * 2.0 * half),
                "halted": list(eng.num_halted()), "wall_s": round(time.time() - t0, 1)}
        res.append(line)
        print(json.dumps(line), flush=True)
        eng.close()
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
