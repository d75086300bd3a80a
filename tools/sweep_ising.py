#!/usr/bin/env python3
"""Ising 32x32 SAD throughput vs walker count (exploration tool)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sad_monte_carlo_b200 import WalkerEngine, make_config
for arg in sys.argv[1:] or ["4096", "65536", "262144"]:
    parts = arg.split(":")
    W = int(parts[0]); flags = int(parts[1]) if len(parts) > 1 else 0
    moves = int(parts[2]) if len(parts) > 2 else 20000
    eng = WalkerEngine(make_config("ising", "sad", N=32, sad_min_T=1.0, n_walkers=W, flags=flags))
    eng.run(100000)
    best = 0
    for _ in range(3):
        eng.run(moves)
        best = max(best, W * moves / (eng.last_run_ms() * 1e-3))
    print(json.dumps({"walkers": W, "flags": flags, "moves_per_s": best, "acc": eng.num_accepted_moves() / (W * eng.num_moves())}), flush=True)
    eng.close()
