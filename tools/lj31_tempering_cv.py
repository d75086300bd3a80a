#!/usr/bin/env python3
"""LJ31 heat capacity by replica exchange on the GPU engine (`sadmc_tempering_*`, the reference's `tempering` binary,
src/mc/tempering.rs) against the literature curves the reference ships.

n_sim independent tempering simulations x n_T temperatures (geometric ladder, two-wells/run-two-wells.py:36-43); every
replica collects <E>, <E^2> exactly as `Replica::run_once` does (tempering.rs:105-111); Cv(T) = (<E^2> - <E>^2) / T^2
from the moments accumulated over the second half of the run, error bar = spread over interleaved groups of simulations.
The step is set per temperature (`Replica::translation_scale` is serialised state; the constructor's 1.0 sigma accepts
nothing in a cluster): scale = step * sqrt(T / 0.3).

    python tools/lj31_tempering_cv.py --sims 1184 --n-T 32 --T-min 0.02 --T-max 0.45 --rounds 40000 --out gpurun_out/r02_lj31_tempering.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sad_monte_carlo_b200 import analysis, make_config, _abi  # noqa: E402
from sad_monte_carlo_b200.tempering import TemperingMC, geometric_spacing  # noqa: E402

LIT = os.path.join(ROOT, "tests", "golden", "lj31_literature")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sims", type=int, default=1184)
    ap.add_argument("--n-T", type=int, default=32)
    ap.add_argument("--T-min", type=float, default=0.02)
    ap.add_argument("--T-max", type=float, default=0.45)
    ap.add_argument("--canonical-steps", type=int, default=16)
    ap.add_argument("--rounds", type=int, default=20000, help="MC::run_once calls; the first half is equilibration")
    ap.add_argument("--step", type=float, default=0.1, help="translation scale at T = 0.3")
    ap.add_argument("--groups", type=int, default=8)
    ap.add_argument("--chunk", type=int, default=500, help="rounds per call (one host round trip each)")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    T = np.array(geometric_spacing(a.T_min, a.T_max, a.n_T))
    cfg = make_config("lj", N=31, lj_radius=2.5, n_walkers=a.sims, seed=0, lanes_per_walker=1, flags=_abi.FLAG_FAST_MATH)
    t0 = time.time()
    mc = TemperingMC(cfg, T, a.canonical_steps)
    mc.set_translation_scales(a.step * np.sqrt(T / 0.3))
    t_create = time.time() - t0
    half = a.rounds // 2
    ms = 0.0

    def run(n):
        nonlocal ms
        done = 0
        while done < n:
            k = min(a.chunk, n - done)
            mc.run_once(k)
            ms += mc.last_run_ms()
            done += k

    run(half)
    m1 = mc.all_replicas()
    run(a.rounds - half)
    m2 = mc.all_replicas()
    d = {k: m2[k] - m1[k] for k in m2 if k != "energy"}
    n = d["accepted_count"] + d["rejected_count"] + d["accepted_swap_count"] + d["rejected_swap_count"]
    cv_g, u_g = [], []
    for g in range(a.groups):
        sel = slice(g, None, a.groups)
        nn = n[sel].sum(axis=0)
        U = d["total_energy"][sel].sum(axis=0) / nn
        E2 = d["total_energy_squared"][sel].sum(axis=0) / nn
        u_g.append(U)
        cv_g.append((E2 - U * U) / (T * T))
    cv_g, u_g = np.array(cv_g), np.array(u_g)
    cv, sem = cv_g.mean(axis=0), cv_g.std(axis=0, ddof=1) / np.sqrt(a.groups)
    acc = d["accepted_count"].sum(axis=0) / (d["accepted_count"] + d["rejected_count"]).sum(axis=0)
    swp = d["accepted_swap_count"].sum(axis=0) / np.maximum(1.0, (d["accepted_swap_count"] + d["rejected_swap_count"]).sum(axis=0))
    moves_total = mc.moves * a.sims
    res = {"what": "LJ31 R=2.5 replica exchange", "sims": a.sims, "n_T": a.n_T, "canonical_steps": a.canonical_steps, "rounds": a.rounds,
           "moves_per_replica": mc.steps_per_round * a.rounds, "moves_total": moves_total, "device_ms": ms,
           "moves_per_s": moves_total / (ms * 1e-3), "create_s": round(t_create, 1), "wall_s": round(time.time() - t0, 1),
           "T": T.tolist(), "Cv": cv.tolist(), "Cv_sem": sem.tolist(), "U": u_g.mean(axis=0).tolist(), "move_acceptance": acc.tolist(),
           "swap_acceptance": swp.tolist(), "lowest_energy_seen": float(m2["energy"].min())}
    refs = {}
    for name in ("LJ31_Cv_Reference.csv", "tRem_Ref.csv", "LJ31_Cv_Reference_alt.csv"):
        Tr, cr = analysis.load_lj31_reference(os.path.join(LIT, name))
        order = np.argsort(Tr)
        ref = np.interp(T, Tr[order], cr[order], left=np.nan, right=np.nan)
        refs[name] = ref.tolist()
    res["references"] = refs
    print("| T | Cv (PT, GPU) | s.e.m. | REM LJ31_Cv_Reference.csv | t-REM | RESTMC _alt | move acc | swap acc |")
    print("|---|---|---|---|---|---|---|---|")
    for i in range(a.n_T):
        print("| %.4f | %.2f | %.2f | %.2f | %.2f | %.2f | %.2f | %.2f |" % (T[i], cv[i], sem[i], refs["LJ31_Cv_Reference.csv"][i], refs["tRem_Ref.csv"][i],
                                                                      refs["LJ31_Cv_Reference_alt.csv"][i], acc[i], swp[i]))
    print(json.dumps({k: res[k] for k in ("sims", "n_T", "rounds", "moves_per_replica", "moves_total", "device_ms", "moves_per_s", "wall_s",
                                          "lowest_energy_seen")}))
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
