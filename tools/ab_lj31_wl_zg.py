#!/usr/bin/env python3
"""LJ31 1/t-WL on [-130, -80] (bin 0.01) with z streamed from L2 (three 128-thread CTAs per SM) against the all-shared-memory layout
(two CTAs), each at one whole wave of its own shape: calibrates the engine's layout choice for the WL methods."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sad_monte_carlo_b200 import WalkerEngine, make_config, _abi
FM, R = _abi.FLAG_FAST_MATH, _abi.INIT_RANDOMIZE
for name, flag, per_sm in (("stream", _abi.FLAG_LJ_STREAM_Z, 384), ("smem", _abi.FLAG_LJ_SMEM_Z, 256)):
    W = 148 * per_sm
    eng = WalkerEngine(make_config("lj", "inv-t-wl", N=31, lj_radius=2.5, energy_bin=0.01, init_mode=R, lanes_per_walker=1, n_walkers=W, flags=FM | flag,
                                   min_allowed_energy=-130.0, max_allowed_energy=-80.0, bin_window_lo=-133.62, bin_window_hi=0.02))
    eng.run(100000)
    ms = []
    for _ in range(3):
        eng.run(20000)
        ms.append(eng.last_run_ms())
    print(json.dumps({"config": "LJ31 1/t-WL fast-math", "layout": name, "walkers": W, "moves_per_s": W * 20000 / (min(ms) * 1e-3),
                      "ms": [round(x, 1) for x in ms], "halted": eng.num_halted()}), flush=True)
    del eng
