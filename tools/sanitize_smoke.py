#!/usr/bin/env python3
"""A few hundred moves of every kernel family with awkward walker counts (partial warps / CTAs), for
`compute-sanitizer --tool memcheck|racecheck python tools/sanitize_smoke.py`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sad_monte_carlo_b200 import WalkerEngine, make_config, _abi  # noqa: E402

FM, R = _abi.FLAG_FAST_MATH, _abi.INIT_RANDOMIZE
LJ = dict(N=31, lj_radius=2.5, max_allowed_energy=0.0, sad_min_T=0.01, energy_bin=0.01, init_mode=R, bin_window_lo=-133.62, bin_window_hi=0.02)
CASES = [
    ("ising sad", "ising", "sad", dict(N=16, sad_min_T=1.0, n_walkers=70)),
    ("ising wl", "ising", "wl", dict(N=8, wl_min_gamma=1e-3, min_allowed_energy=-128.0, max_allowed_energy=50.0, n_walkers=33)),
    ("fake quadratic samc", "fake", "samc", dict(fake_function=_abi.FAKE_QUADRATIC, N=3, samc_t0=1e3, energy_bin=0.01, n_walkers=45, bin_window_lo=-2.5, bin_window_hi=4.0)),
    ("two-wells sad", "two-wells", "sad", dict(N=12, tw_h2_to_h1=1.1, tw_barrier_over_h1=0.1, tw_r2=0.5, sad_min_T=0.001, energy_bin=1e-3, move_value=1e-2, n_walkers=37)),
    ("erfinv samc", "fake-erfinv", "samc", dict(N=3, erfinv_mean_energy=0.0, samc_t0=1e3, energy_bin=0.05, n_walkers=19, bin_window_lo=-30.0, bin_window_hi=30.0)),
    ("lj31 exact thread", "lj", "sad", dict(LJ, n_walkers=150, lanes_per_walker=1)),
    ("lj31 fast thread", "lj", "sad", dict(LJ, n_walkers=150, lanes_per_walker=1, flags=FM)),
    ("lj31 fast 2 lanes", "lj", "sad", dict(LJ, n_walkers=75, lanes_per_walker=2, flags=FM)),
    ("lj31 fast 4 lanes", "lj", "sad", dict(LJ, n_walkers=41, lanes_per_walker=4, flags=FM)),
    ("lj31 warp 8 lanes", "lj", "sad", dict(LJ, n_walkers=21, lanes_per_walker=8)),
    ("lj31 warp 32 lanes", "lj", "wl", dict(LJ, n_walkers=7, lanes_per_walker=32, min_allowed_energy=-133.0)),
    ("lj38 fast thread inv-t-wl", "lj", "inv-t-wl", dict(N=38, lj_radius=3.0, min_allowed_energy=-173.0, max_allowed_energy=-100.0, energy_bin=0.01,
                                                      n_walkers=40, init_mode=R, lanes_per_walker=1, flags=FM, bin_window_lo=-174.0, bin_window_hi=0.02)),
    ("wca samc", "wca", "samc", dict(N=40, reduced_density=0.5, energy_bin=1.0, n_walkers=9, samc_t0=1e3, max_allowed_energy=400.0, init_mode=R, bin_window_lo=0.0, bin_window_hi=420.0)),
    ("sw sad", "sw", "sad", dict(N=64, filling_fraction=0.25, sad_min_T=0.5, n_walkers=9)),
]
moves = int(sys.argv[1]) if len(sys.argv) > 1 else 300
for name, system, method, kw in CASES:
    eng = WalkerEngine(make_config(system, method, **kw))
    eng.run(moves)
    eng.run(7)
    f = eng.fold()
    b = eng.bins(eng.n_walkers - 1)
    ok = all(eng.walker(w).status == 0 for w in range(eng.n_walkers))
    print("%-28s walkers %4d moves %d histogram total %d ok=%s" % (name, eng.n_walkers, eng.num_moves(), int(f["histogram"].sum()), ok), flush=True)
    eng.close()
