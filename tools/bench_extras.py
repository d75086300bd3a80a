#!/usr/bin/env python3
"""Throughput of the round-2 additions on one GPU: the `binning` bookkeeping (SADMC_FLAG_BINNING) beside the `histogram`
bookkeeping on the same configurations, and replica exchange (sadmc_tempering_*).  One JSON line per case; moves/s from
CUDA-event kernel time after a burn-in."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sad_monte_carlo_b200 import WalkerEngine, make_config, _abi  # noqa: E402
from sad_monte_carlo_b200.tempering import TemperingMC, geometric_spacing  # noqa: E402

FM, R, B = _abi.FLAG_FAST_MATH, _abi.INIT_RANDOMIZE, _abi.FLAG_BINNING
CASES = [
    ("fake linear SAD 65536 walkers, bin 0.001", dict(system="fake", method="sad", fake_function=_abi.FAKE_LINEAR, energy_bin=0.001, sad_min_T=0.001,
                                                      move_value=0.05, n_walkers=65536, bin_window_lo=-0.1, bin_window_hi=1.1), 200000, 100000),
    ("two-wells SAD 65536 walkers", dict(system="two-wells", method="sad", N=12, tw_h2_to_h1=1.1, tw_barrier_over_h1=0.1, tw_r2=0.5, sad_min_T=0.001,
                                         energy_bin=1e-4, move_value=1e-3, n_walkers=65536), 200000, 100000),
    ("ising32 SAD 262144 walkers", dict(system="ising", method="sad", N=32, sad_min_T=1.0, energy_bin=4.0, n_walkers=262144), 100000, 50000),
    ("LJ31 SAD fast-math 75776 walkers", dict(system="lj", method="sad", N=31, lj_radius=2.5, max_allowed_energy=0.0, sad_min_T=0.01, energy_bin=0.01,
                                              n_walkers=75776, init_mode=R, lanes_per_walker=1, flags=FM, bin_window_lo=-133.62, bin_window_hi=0.02),
     200000, 20000),
]


def main():
    for name, kw, burn, moves in CASES:
        for label, extra in (("histogram (energy.rs)", 0), ("binning (energy_binning.rs)", B)):
            k = dict(kw)
            k["flags"] = k.get("flags", 0) | extra
            try:
                eng = WalkerEngine(make_config(k.pop("system"), k.pop("method"), **k))
                eng.run(burn)
                ms = []
                for _ in range(3):
                    eng.run(moves)
                    ms.append(eng.last_run_ms())
                W = eng.n_walkers
                out = {"config": name, "bookkeeping": label, "walkers": W, "moves_per_launch": moves, "ms_per_launch": min(ms),
                       "moves_per_s": W * moves / (min(ms) * 1e-3), "halted": list(eng.num_halted())}
                eng.close()
            except Exception as ex:
                out = {"config": name, "bookkeeping": label, "error": str(ex)}
            print(json.dumps(out), flush=True)
    # replica exchange
    for name, cfg, T, steps, rounds, scales in (
        ("tempering two-wells N=12, 10 temperatures x 16384 simulations, --canonical-steps 10 (two-wells/run-two-wells.py:204)",
         make_config("two-wells", N=12, tw_h2_to_h1=1.1, tw_barrier_over_h1=0.0, tw_r2=0.5, n_walkers=16384, seed=0),
         geometric_spacing(0.001, 1.0, 10), 10, 400, None),
        ("tempering LJ31 R=2.5 fast-math, 32 temperatures x 2368 simulations, --canonical-steps 16, step 0.1 sqrt(T / 0.3)",
         make_config("lj", N=31, lj_radius=2.5, n_walkers=2368, seed=0, lanes_per_walker=1, flags=FM),
         geometric_spacing(0.02, 0.45, 32), 16, 100, "sqrt"),
    ):
        try:
            mc = TemperingMC(cfg, T, steps)
            if scales:
                mc.set_translation_scales(0.1 * np.sqrt(np.array(T) / 0.3))
            mc.run_once(rounds)
            ms = []
            for _ in range(3):
                mc.run_once(rounds)
                ms.append(mc.last_run_ms())
            moves = mc.steps_per_round * rounds * mc.n_T * mc.n_sim
            out = {"config": name, "replicas": mc.n_T * mc.n_sim, "rounds_per_call": rounds, "moves_per_round_and_replica": mc.steps_per_round,
                   "ms_per_call": min(ms), "moves_per_s": moves / (min(ms) * 1e-3), "launches_per_call": 2 * rounds}
            mc.close()
        except Exception as ex:
            out = {"config": name, "error": str(ex)}
        print(json.dumps(out), flush=True)
    # energy-ceiling replicas
    from sad_monte_carlo_b200.replicas import ReplicasMC
    for name, cfg, args, rounds in (
        ("replicas fake quadratic d=3, 16384 simulations x up to 48 replicas, min_T 0.001", make_config("fake", fake_function=_abi.FAKE_QUADRATIC, N=3, n_walkers=16384, seed=0),
         (0.001, 64, 48, 4096), 2000),
        ("replicas LJ31 R=2.5 fast-math, 1184 simulations x up to 32 replicas, min_T 0.1", make_config("lj", N=31, lj_radius=2.5, n_walkers=1184, seed=0, lanes_per_walker=1, flags=FM),
         (0.1, 16, 32, 1024), 500),
    ):
        try:
            z = ReplicasMC(cfg, *args)
            z.run_once(rounds)
            m0 = sum(z.moves(k) for k in range(0, z.n_sim, max(1, z.n_sim // 64)))
            z.run_once(rounds)
            ms = z.last_run_ms()
            sample = list(range(0, z.n_sim, max(1, z.n_sim // 64)))
            m1 = sum(z.moves(k) for k in sample)
            moves = (m1 - m0) / len(sample) * z.n_sim
            out = {"config": name, "rounds_per_call": rounds, "ms_per_call": ms, "moves_per_s": moves / (ms * 1e-3), "launches_per_call": 2 * rounds,
                   "replicas_per_simulation_now": float(np.mean([z.num_replicas(k) for k in sample]))}
            z.close()
        except Exception as ex:
            out = {"config": name, "error": str(ex)}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
