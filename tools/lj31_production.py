#!/usr/bin/env python3
"""LJ31 heat capacity from a fixed-weight production run on the GPU engine.

SAD learns ln w(E) per walker, and a walker's estimate only settles after many round trips through the energy range --
~1e10..1e12 moves at the reference's headline parameters (energy bin 0.01, min_T 0.01; the reference's own runs use
--max-iter 1e12), which no number of parallel walkers shortens.  What parallel walkers CAN do is sample one fixed
ensemble together.  So: take the merged SAD entropy of a learning run (tools/lj31_cv_run.py, any stage), give it to
every walker as fixed weights (`sadmc_set_lnw`, Method::Samc with t0 = 0: gamma = 0), run all walkers, and reweight the
summed histogram: S(E) = ln w(E) + ln H(E) + const.  That identity is exact for fixed weights whatever their quality;
their quality only decides how evenly the energy range is visited.  Repeating the step with the improved S as weights is the
usual multicanonical recursion.

    python tools/lj31_production.py --weights tests/golden/lj31_cv_run_r02/lj31_cv_headline_2e+08.npz --moves 4e7 --out gpurun_out/lj31_production_1.npz
    python tools/lj31_production.py --weights gpurun_out/lj31_production_1.npz --moves 4e7 --out gpurun_out/lj31_production_2.npz
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sad_monte_carlo_b200 import WalkerEngine, make_config, _abi  # noqa: E402


def weights_from(path, nb, slope_T):
    """Window-aligned fixed weights from a learning run (SAD folds) or from an earlier production run (its entropy)."""
    d = np.load(path)
    if "entropy" in d:
        S, ok = np.array(d["entropy"]), np.array(d["entropy_ok"], bool)
    else:
        G = int(d["groups"])
        cnt = sum(np.asarray(d["lnw_count_%d" % g], dtype=np.float64) for g in range(G))
        tot = sum(np.asarray(d["lnw_sum_%d" % g], dtype=np.float64) for g in range(G))
        ok = cnt >= 0.5 * int(d["walkers"])
        S = np.zeros(nb)
        S[ok] = tot[ok] / cnt[ok]
    assert len(S) == nb
    idx = np.nonzero(ok)[0]
    lo, hi = idx.min(), idx.max()
    S = np.interp(np.arange(nb), idx, S[idx])  # fill gaps inside the covered range
    width = float(d["width"])
    below = np.arange(nb) < lo
    S[below] = S[lo] - (lo - np.arange(nb)[below]) * width / slope_T  # below: the slope of a canonical ensemble at slope_T
    S[np.arange(nb) > hi] = S[hi]                                     # above: flat (as SAD treats E > too_hi)
    return S - S.max(), int(lo), int(hi)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--weights", required=True)
    ap.add_argument("--walkers", type=int, default=37888)
    ap.add_argument("--moves", type=float, default=4e7, help="per walker; the first quarter is equilibration and is not counted")
    ap.add_argument("--groups", type=int, default=8)
    ap.add_argument("--slope-T", type=float, default=0.04)
    ap.add_argument("--chunk", type=float, default=2e6)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    cfg = make_config("lj", "samc", N=31, lj_radius=2.5, max_allowed_energy=0.0, samc_t0=0.0, energy_bin=0.01, move_value=0.05,
                      n_walkers=a.walkers, init_mode=_abi.INIT_RANDOMIZE, lanes_per_walker=1, seed=1000000, flags=_abi.FLAG_FAST_MATH,
                      bin_window_lo=-133.62, bin_window_hi=0.02)
    eng = WalkerEngine(cfg)
    lo, width, nb = eng.window()
    w, i_lo, i_hi = weights_from(a.weights, nb, a.slope_T)
    eng.set_lnw(w)

    def hist():
        out = []
        for g in range(a.groups):
            eng.fold_select(g, a.groups, 0)
            out.append(eng.fold()["histogram"].astype(np.float64))
        eng.fold_select(0, 1, 0)
        return np.array(out)

    def run(n):
        done = 0
        while done < n:
            k = int(min(a.chunk, n - done))
            eng.run(k)
            done += k

    t0 = time.time()
    burn = int(a.moves) // 4
    run(burn)
    h0 = hist()
    run(int(a.moves) - burn)
    H = hist() - h0  # counted part only
    Hall = H.sum(0)
    ok = Hall > 0
    S = np.zeros(nb)
    S[ok] = w[ok] + np.log(Hall[ok])
    E = lo + (np.arange(nb) + 0.5) * width
    np.savez_compressed(a.out, window_lo=lo, width=width, walkers=a.walkers, groups=a.groups, moves=int(a.moves), counted_from=burn,
                        weights=w, weights_covered=np.array([i_lo, i_hi]), histogram_groups=H, entropy=S - S[ok].max(), entropy_ok=ok,
                        energies=eng.energies()[::16])
    vis = np.nonzero(ok)[0]
    flat = Hall[ok]
    print("production: %d walkers x %.1e moves (%.1e counted) in %.0f s; bins visited %d (E from %.2f to %.2f); histogram max/median %.1f; "
          "halted %s" % (a.walkers, a.moves, a.moves - burn, time.time() - t0, ok.sum(), E[vis.min()], E[vis.max()],
                         flat.max() / np.median(flat), eng.num_halted()), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
