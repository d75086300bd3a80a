#!/usr/bin/env python3
"""Fixed cost per launch vs per-move cost of the LJ31 move kernel (exploration tool)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sad_monte_carlo_b200 import WalkerEngine
W = int(sys.argv[1]); lanes = int(sys.argv[2]); flags = int(sys.argv[3])
eng = WalkerEngine(bench.lj31_config(W, lanes=lanes, flags=flags))
eng.run(100000)
for n in (1, 10, 100, 300, 1000, 3000, 10000):
    best = 1e9
    for _ in range(3):
        eng.run(n)
        best = min(best, eng.last_run_ms())
    print(json.dumps({"moves": n, "ms": best, "us_per_move_iter": 1e3 * best / n}), flush=True)
