#!/usr/bin/env python3
"""LJ_N SAD run with the bench workload's parameters but another atom count (occupancy probe: smaller clusters leave
shared memory for more warps per SM).   python tools/profile_ljn.py N window_lo [walkers] [moves] [reps] [burn_in]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sad_monte_carlo_b200 import WalkerEngine, make_config, _abi
N = int(sys.argv[1]); lo = float(sys.argv[2])
W = int(sys.argv[3]) if len(sys.argv) > 3 else 75776
moves = int(sys.argv[4]) if len(sys.argv) > 4 else 20000
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 4
burn = int(sys.argv[6]) if len(sys.argv) > 6 else 200000
eng = WalkerEngine(make_config("lj", "sad", N=N, lj_radius=2.5, max_allowed_energy=0.0, sad_min_T=0.01, energy_bin=0.01,
                               move_value=0.05, n_walkers=W, init_mode=_abi.INIT_RANDOMIZE, lanes_per_walker=1, seed=0,
                               flags=_abi.FLAG_FAST_MATH if hasattr(_abi, "FLAG_FAST_MATH") else 4,
                               bin_window_lo=lo, bin_window_hi=0.02))
eng.run(burn)
ms = []
for _ in range(reps):
    eng.run(moves)
    ms.append(eng.last_run_ms())
halted = eng.num_halted() if hasattr(eng, "num_halted") else "?"
print("N", N, "ms", " ".join("%.1f" % x for x in ms), "| best moves/s %.4g median %.4g | halted %s" % (
    W * moves / min(ms) * 1e3, W * moves / sorted(ms)[len(ms) // 2] * 1e3, halted))
