#!/usr/bin/env python3
"""LJ31 SAD throughput vs energy_bin (memory-footprint sensitivity probe)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sad_monte_carlo_b200 import WalkerEngine
for arg in sys.argv[1:]:
    W, lanes, flags, de = arg.split(":")
    cfg = bench.lj31_config(int(W), lanes=int(lanes), flags=int(flags))
    cfg.energy_bin = float(de)
    eng = WalkerEngine(cfg)
    eng.run(100000)
    best = 0
    for _ in range(3):
        eng.run(20000)
        best = max(best, int(W) * 20000 / (eng.last_run_ms() * 1e-3))
    print(json.dumps({"walkers": int(W), "lanes": int(lanes), "flags": int(flags), "energy_bin": float(de), "moves_per_s": best}), flush=True)
    eng.close()
