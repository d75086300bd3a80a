#!/usr/bin/env python3
"""Short LJ31 SAD run (the bench workload) for ncu and for A/B timing of kernel variants: a burn-in launch, then `reps`
launches of `moves` moves; prints every launch's CUDA-event time and the SM clock sampled right after.

    python tools/profile_lj.py [walkers] [lanes] [moves] [flags] [reps] [burn_in]
    SADMC_GPU_LIB=build_exp/x.so python tools/profile_lj.py 75776 1 20000 4 8"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sad_monte_carlo_b200 import WalkerEngine
W = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 8
moves = int(sys.argv[3]) if len(sys.argv) > 3 else 400
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
burn = int(sys.argv[6]) if len(sys.argv) > 6 else 100000
eng = WalkerEngine(bench.lj31_config(W, lanes=lanes, flags=flags))
eng.run(burn)
ms = []
for _ in range(reps):
    eng.run(moves)
    ms.append(eng.last_run_ms())
try:
    clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu", "--format=csv,noheader,nounits", "-i", "0"],
                         capture_output=True, text=True, timeout=5).stdout.strip()
except Exception:
    clk = "?"
best = min(ms)
print("ms", " ".join("%.1f" % x for x in ms), "| best moves/s %.4g median %.4g | clocks(sm MHz, W, C) %s" % (
    W * moves / best * 1e3, W * moves / sorted(ms)[len(ms) // 2] * 1e3, clk))
