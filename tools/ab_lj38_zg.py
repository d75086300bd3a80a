#!/usr/bin/env python3
"""LJ38 (config C4: SAD and 1/t-WL, R = 3, bin 0.01) with z streamed from L2 (one 320-thread CTA per SM) against the all-shared-memory
layout (one 224-thread CTA), each at whole waves of its own CTA shape.  One JSON line per run."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sad_monte_carlo_b200 import WalkerEngine, make_config, _abi
FM, R = _abi.FLAG_FAST_MATH, _abi.INIT_RANDOMIZE
BASE = dict(N=38, lj_radius=3.0, energy_bin=0.01, init_mode=R, lanes_per_walker=1, bin_window_lo=-174.0, bin_window_hi=0.02)
METHODS = {"sad": dict(max_allowed_energy=0.0, sad_min_T=0.01), "inv-t-wl": dict(min_allowed_energy=-173.0, max_allowed_energy=-100.0)}
burn, moves = 100000, 10000
for method, mkw in METHODS.items():
    for name, flag, per_sm in (("stream", _abi.FLAG_LJ_STREAM_Z, 320), ("smem", _abi.FLAG_LJ_SMEM_Z, 224)):
        for waves in (1, 2):
            W = 148 * per_sm * waves
            eng = WalkerEngine(make_config("lj", method, n_walkers=W, flags=FM | flag, **BASE, **mkw))
            eng.run(burn)
            ms = []
            for _ in range(3):
                eng.run(moves)
                ms.append(eng.last_run_ms())
            print(json.dumps({"config": "C4 LJ38 %s fast-math" % method, "layout": name, "walkers": W, "waves": waves, "launch_shape": eng.move_launch_shape(),
                              "moves_per_s": W * moves / (min(ms) * 1e-3), "ms": [round(x, 1) for x in ms], "halted": eng.num_halted()}), flush=True)
            del eng
