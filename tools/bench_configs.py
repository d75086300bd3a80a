#!/usr/bin/env python3
"""Throughput of the move kernel on BASELINE.json's five configs (one GPU); the headline config is bench.py's.

Prints one JSON line per config: moves/s from CUDA-event kernel time of `reps` launches after a burn-in,
plus the algorithmic-traffic figure (80 B bin traffic per move) against MEASURED_PEAKS.json's HBM bandwidth."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sad_monte_carlo_b200 import WalkerEngine, make_config, _abi  # noqa: E402

FM = _abi.FLAG_FAST_MATH
R = _abi.INIT_RANDOMIZE
CONFIGS = [
    ("C1 ising32 SAD 4096 walkers (bit-exact tier)", dict(system="ising", method="sad", N=32, sad_min_T=1.0, n_walkers=4096), 200000, 100000),
    ("C1 ising32 SAD 262144 walkers", dict(system="ising", method="sad", N=32, sad_min_T=1.0, n_walkers=262144), 100000, 50000),
    ("C1 ising32 WL (ising-wl-min-gamma.sh) 262144 walkers", dict(system="ising", method="wl", N=32, wl_min_gamma=1e-4, min_allowed_energy=-2048.0,
                                                                 max_allowed_energy=50.0, n_walkers=262144), 100000, 50000),
    ("C2 fake linear SAD 65536 walkers", dict(system="fake", method="sad", fake_function=_abi.FAKE_LINEAR, energy_bin=0.001, sad_min_T=0.001,
                                              move_value=0.05, n_walkers=65536, bin_window_lo=-0.1, bin_window_hi=1.1), 200000, 100000),
    ("C2 two-wells SAD 65536 walkers", dict(system="two-wells", method="sad", N=12, tw_h2_to_h1=1.1, tw_barrier_over_h1=0.1, tw_r2=0.5,
                                            sad_min_T=0.001, energy_bin=1e-4, move_value=1e-3, n_walkers=65536), 200000, 100000),
    ("C3 LJ31 SAD exact arithmetic 75776 walkers", dict(system="lj", method="sad", N=31, lj_radius=2.5, max_allowed_energy=0.0, sad_min_T=0.01,
                                                        energy_bin=0.01, n_walkers=75776, init_mode=R, lanes_per_walker=1,
                                                        bin_window_lo=-133.62, bin_window_hi=0.02), 100000, 10000),
    ("C4 LJ38 SAD fast-math 65536 walkers", dict(system="lj", method="sad", N=38, lj_radius=3.0, max_allowed_energy=0.0, sad_min_T=0.01,
                                                 energy_bin=0.01, n_walkers=65536, init_mode=R, lanes_per_walker=1, flags=FM,
                                                 bin_window_lo=-174.0, bin_window_hi=0.02), 100000, 10000),
    ("C4 LJ38 1/t-WL fast-math 65536 walkers", dict(system="lj", method="inv-t-wl", N=38, lj_radius=3.0, min_allowed_energy=-173.0,
                                                    max_allowed_energy=-100.0, energy_bin=0.01, n_walkers=65536, init_mode=R,
                                                    lanes_per_walker=1, flags=FM, bin_window_lo=-174.0, bin_window_hi=0.02), 100000, 10000),
    # random start + the reference's downhill relaxation to E < max_allowed_energy (energy.rs:840-851); 10 N as wca/run-wca.py's max_E
    ("C5 WCA N=256 rho=0.8 SAMC 8192 walkers", dict(system="wca", method="samc", N=256, reduced_density=0.8, samc_t0=1e7, energy_bin=1.0,
                                                    max_allowed_energy=2560.0, n_walkers=8192, init_mode=R, bin_window_lo=0.0,
                                                    bin_window_hi=2600.0), 20000, 10000),
    ("C5 SW N=100 eta=0.3 SAD 16384 walkers (bit-exact tier)", dict(system="sw", method="sad", N=100, filling_fraction=0.3, sad_min_T=0.5,
                                                                    n_walkers=16384), 20000, 10000),
]


def main():
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    only = sys.argv[1:]
    for name, kw, burn, moves in CONFIGS:
        if only and not any(o in name for o in only):
            continue
        kw = dict(kw)
        try:
            eng = WalkerEngine(make_config(kw.pop("system"), kw.pop("method"), **kw))
            eng.run(burn)
            ms = []
            for _ in range(3):
                eng.run(moves)
                ms.append(eng.last_run_ms())
            W = eng.n_walkers
            halted = sum(eng.walker(w).status != 0 for w in range(0, W, max(1, W // 64)))
            v = W * moves / (min(ms) * 1e-3)
            out = {"config": name, "moves_per_s": v, "ms_per_launch": min(ms), "moves_per_launch": moves, "walkers": W,
                   "halted_in_sample": halted, "bin_traffic_GBps_at_80B_per_move": 80.0 * v / 1e9}
            if peaks.get("hbm_gbs"):
                out["frac_of_hbm_peak"] = 80.0 * v / 1e9 / peaks["hbm_gbs"]
            eng.close()
        except Exception as ex:  # report and go on: one unsupported config must not hide the others
            out = {"config": name, "error": str(ex)}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
