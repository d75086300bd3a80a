#!/usr/bin/env python3
"""Short LJ31 SAD run for ncu: one burn-in launch, then a few short launches of the move kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sad_monte_carlo_b200 import WalkerEngine
W = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 8
moves = int(sys.argv[3]) if len(sys.argv) > 3 else 400
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
eng = WalkerEngine(bench.lj31_config(W, lanes=lanes, flags=flags))
eng.run(100000)
for _ in range(3):
    eng.run(moves)
print("ms", eng.last_run_ms(), "moves/s", W * moves / eng.last_run_ms() * 1e3)
