#!/usr/bin/env python3
"""Quick LJ31 SAD throughput sweep over lanes_per_walker and walker count (exploration tool, not the bench)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sad_monte_carlo_b200 import WalkerEngine

def one(W, lanes, flags=0, burn=100000, moves=20000, reps=3):
    eng = WalkerEngine(bench.lj31_config(W, lanes=lanes, flags=flags))
    eng.run(burn)
    best = 0
    for _ in range(reps):
        eng.run(moves)
        ms = eng.last_run_ms()
        best = max(best, W * moves / (ms * 1e-3))
    acc = eng.num_accepted_moves() / (W * eng.num_moves())
    eng.close()
    return best, acc

if __name__ == "__main__":
    combos = [tuple(int(x) for x in a.split(":")) for a in sys.argv[1:]] or [(8192, 32), (32768, 8)]
    for c in combos:
        W, lanes = c[0], c[1]
        flags = c[2] if len(c) > 2 else 0
        v, acc = one(W, lanes, flags)
        print(json.dumps({"walkers": W, "lanes": lanes, "flags": flags, "moves_per_s": v, "acceptance": acc}), flush=True)
