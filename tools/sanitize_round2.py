#!/usr/bin/env python3
"""The round-2 kernels under compute-sanitizer: the `binning` bookkeeping (histogram and linear bins, high-resolution
histogram), replica exchange (sweep + swap kernels) and the WCA lane-group kernels, with awkward walker counts.
`compute-sanitizer --tool memcheck|racecheck python tools/sanitize_round2.py`"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sad_monte_carlo_b200 import WalkerEngine, make_config, _abi  # noqa: E402
from sad_monte_carlo_b200.tempering import TemperingMC, geometric_spacing  # noqa: E402

FM, R, B, BL = _abi.FLAG_FAST_MATH, _abi.INIT_RANDOMIZE, _abi.FLAG_BINNING, _abi.FLAG_BINNING | _abi.FLAG_BINNING_LINEAR
LJ = dict(N=31, lj_radius=2.5, max_allowed_energy=0.0, sad_min_T=0.01, energy_bin=0.01, init_mode=R, bin_window_lo=-133.62, bin_window_hi=0.02)
CASES = [
    ("binning ising sad + high-res", "ising", "sad", dict(N=16, sad_min_T=1.0, energy_bin=8.0, high_resolution_de=4.0, n_walkers=70, flags=B)),
    ("binning ising wl", "ising", "wl", dict(N=8, wl_min_gamma=1e-3, min_allowed_energy=-128.0, max_allowed_energy=50.0, energy_bin=4.0, n_walkers=33, flags=B)),
    ("binning fake quadratic inv-t-wl", "fake", "inv-t-wl", dict(fake_function=_abi.FAKE_QUADRATIC, N=3, energy_bin=0.01, min_allowed_energy=0.0,
                                                                max_allowed_energy=0.99, n_walkers=45, bin_window_lo=-2.5, bin_window_hi=4.0, flags=B)),
    ("binning two-wells sad", "two-wells", "sad", dict(N=12, tw_h2_to_h1=1.1, tw_barrier_over_h1=0.1, tw_r2=0.5, sad_min_T=0.001, energy_bin=1e-3,
                                                       move_value=1e-2, n_walkers=37, flags=B)),
    ("binning lj31 fast thread", "lj", "sad", dict(LJ, n_walkers=150, lanes_per_walker=1, flags=FM | B)),
    ("binning sw sad (warp per walker)", "sw", "sad", dict(N=64, filling_fraction=0.25, sad_min_T=0.5, n_walkers=9, flags=B)),
    ("binning wca samc (warp per walker)", "wca", "samc", dict(N=40, reduced_density=0.5, energy_bin=1.0, n_walkers=9, samc_t0=1e3, max_allowed_energy=400.0,
                                                               init_mode=R, bin_window_lo=0.0, bin_window_hi=420.0, lanes_per_walker=32, flags=B)),
    ("linear fake quadratic sad", "fake", "sad", dict(fake_function=_abi.FAKE_QUADRATIC, N=3, sad_min_T=0.001, energy_bin=0.01, n_walkers=45,
                                                      bin_window_lo=-2.5, bin_window_hi=4.0, flags=BL)),
    ("linear ising wl", "ising", "wl", dict(N=8, min_allowed_energy=-128.0, max_allowed_energy=50.0, energy_bin=4.0, n_walkers=33, flags=BL)),
    ("wca group 8 fast", "wca", "samc", dict(N=40, reduced_density=0.5, energy_bin=1.0, n_walkers=21, samc_t0=1e3, max_allowed_energy=400.0,
                                             init_mode=R, bin_window_lo=0.0, bin_window_hi=420.0, lanes_per_walker=8, flags=FM)),
    ("wca group 4", "wca", "samc", dict(N=40, reduced_density=0.5, energy_bin=1.0, n_walkers=13, samc_t0=1e3, max_allowed_energy=400.0,
                                        init_mode=R, bin_window_lo=0.0, bin_window_hi=420.0, lanes_per_walker=4)),
]
moves = int(sys.argv[1]) if len(sys.argv) > 1 else 300
for name, system, method, kw in CASES:
    eng = WalkerEngine(make_config(system, method, **kw))
    eng.run(moves)
    eng.run(7)
    ok = all(eng.walker(w).status == 0 for w in range(eng.n_walkers))
    if kw.get("flags", 0) & _abi.FLAG_BINNING_LINEAR:
        total = eng.binning_bins_f64(eng.n_walkers - 1)["energy_count"].sum()
    elif kw.get("flags", 0) & B:
        total = eng.fold()["histogram"].sum() / eng.n_walkers
    else:
        total = eng.fold()["histogram"].sum() / eng.n_walkers - 1
    print("%-36s walkers %4d moves %d visits per walker %.1f ok=%s" % (name, eng.n_walkers, eng.num_moves(), float(total), ok), flush=True)
    eng.close()
for name, cfg, T in (
    ("tempering two-wells (odd ladder)", make_config("two-wells", N=12, tw_h2_to_h1=1.1, tw_barrier_over_h1=0.1, tw_r2=0.5, n_walkers=7, seed=1), geometric_spacing(0.01, 1.0, 5)),
    ("tempering ising", make_config("ising", N=8, n_walkers=5, seed=1), [1.0, 2.0, 3.0, 4.0]),
    ("tempering sw (warp per walker)", make_config("sw", N=50, filling_fraction=0.3, sw_well_width=1.3, n_walkers=3, seed=3), [0.5, 1.0, 2.0]),
):
    mc = TemperingMC(cfg, T, 2)
    mc.run_once(20)
    reps = mc.replicas(mc.n_sim - 1)
    print("%-36s simulations %d x %d replicas, moves %d, swaps tried by replica 0: %d" % (name, mc.n_sim, mc.n_T, mc.moves,
                                                                                       reps[0].accepted_swap_count + reps[0].rejected_swap_count), flush=True)
    mc.close()
from sad_monte_carlo_b200.replicas import ReplicasMC  # noqa: E402
for name, cfg, args in (
    ("replicas fake quadratic", make_config("fake", fake_function=_abi.FAKE_QUADRATIC, N=3, n_walkers=5, seed=3), (0.001, 4, 24, 256)),
    ("replicas ising", make_config("ising", N=8, n_walkers=3, seed=2), (2.0, 4, 64, 128)),
    ("replicas wca (warp per walker)", make_config("wca", N=20, reduced_density=0.3, n_walkers=2, seed=4, lanes_per_walker=32), (0.5, 4, 16, 64)),
):
    z = ReplicasMC(cfg, *args)
    z.run_once(300 if "wca" not in name else 60)
    print("%-36s simulations %d, replicas of the last %d, moves %d" % (name, z.n_sim, z.num_replicas(z.n_sim - 1), z.moves(z.n_sim - 1)), flush=True)
    z.close()
