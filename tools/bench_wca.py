#!/usr/bin/env python3
"""Throughput of the WCA move kernels (BASELINE.json config 5: N = 256, cell lists, SAMC) on one GPU, with the CPU
restatement timed beside it.

    python tools/bench_wca.py [--variants g8fast,g8,g4fast,g16fast,warp] [--walkers 9472] [--moves 20000] [--cpu-seconds 10]

One JSON line per kernel variant: moves/s from the CUDA-event time of the move-kernel launches after a burn-in.  The
work per move is ~2 x 30 candidate distance tests (27 subcells around the old and the new position) and ~12 pair
potentials, ~1 kflop in FP64, plus 80 B of bin traffic.
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sad_monte_carlo_b200 import WalkerEngine, make_config, _abi  # noqa: E402

VARIANTS = {
    "warp": dict(lanes_per_walker=32),
    "g16": dict(lanes_per_walker=16), "g8": dict(lanes_per_walker=8), "g4": dict(lanes_per_walker=4),
    "g16fast": dict(lanes_per_walker=16, flags=_abi.FLAG_FAST_MATH), "g8fast": dict(lanes_per_walker=8, flags=_abi.FLAG_FAST_MATH),
    "g4fast": dict(lanes_per_walker=4, flags=_abi.FLAG_FAST_MATH),
}
FLOP_PER_MOVE = 1000.0  # ~110 distance tests x 8 flop + ~12 potentials x 9 flop


def wca_config(n_walkers, method="samc", N=256, rho=0.8, **kw):
    # random start + the reference's downhill relaxation to E < max_allowed_energy (energy.rs:840-851); 10 N as wca/run-wca.py's max_E
    base = dict(N=N, reduced_density=rho, samc_t0=1e7, energy_bin=1.0, max_allowed_energy=10.0 * N, n_walkers=n_walkers,
                init_mode=_abi.INIT_RANDOMIZE, bin_window_lo=0.0, bin_window_hi=10.0 * N + 40.0)
    base.update(kw)
    return make_config("wca", method, **base)


def cpu_baseline(seconds, N, rho):
    """The CPU restatement (oracle/, test infrastructure) on all host threads, one walker per thread."""
    so = os.path.join(ROOT, "oracle", "liboracle_sadmc.so")
    L = C.CDLL(so)
    L.oracle_bench.restype = C.c_double
    L.oracle_bench.argtypes = [C.POINTER(_abi.Config), C.c_uint32, C.c_uint64, C.c_uint64]
    threads = os.cpu_count() or 1
    cfg = wca_config(threads, N=N, rho=rho)
    probe = 100000
    t = L.oracle_bench(C.byref(cfg), threads, 20000, probe)
    n = max(probe, int(probe * seconds / max(t, 1e-9)))
    t = L.oracle_bench(C.byref(cfg), threads, 20000, n)
    return {"value": threads * n / t, "unit": "moves/s", "cores": threads, "kind": "port",
            "sample": "%d threads x %d moves after 20000 warm-up moves each, %.1f s" % (threads, n, t)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="g8fast,g8,g4fast,g16fast,warp")
    ap.add_argument("--walkers", type=int, default=9472, help="default: two waves of 148 SMs x 32 resident walkers")
    ap.add_argument("--moves", type=int, default=20000)
    ap.add_argument("--burn-in", type=int, default=20000)
    ap.add_argument("--N", type=int, default=256)
    ap.add_argument("--rho", type=float, default=0.8)
    ap.add_argument("--method", default="samc")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    a = ap.parse_args()
    cpu = None
    if a.cpu_seconds > 0:
        try:
            cpu = cpu_baseline(a.cpu_seconds, a.N, a.rho)
        except Exception as ex:
            cpu = {"value": None, "error": str(ex)}
    for v in a.variants.split(","):
        try:
            eng = WalkerEngine(wca_config(a.walkers, method=a.method, N=a.N, rho=a.rho, **VARIANTS[v]))
            eng.run(a.burn_in)
            ms = []
            for _ in range(3):
                eng.run(a.moves)
                ms.append(eng.last_run_ms())
            rate = a.walkers * a.moves / (min(ms) * 1e-3)
            acc = eng.num_accepted_moves() / (a.walkers * eng.num_moves())
            out = {"config": "C5 WCA N=%d rho=%g %s" % (a.N, a.rho, a.method.upper()), "kernel": v, "walkers": a.walkers,
                   "moves_per_s": rate, "ms_per_launch": min(ms), "moves_per_launch": a.moves, "acceptance": acc,
                   "fp64_tflops_at_1kflop_per_move": rate * FLOP_PER_MOVE / 1e12, "halted": list(eng.num_halted()),
                   "cpu_baseline": cpu}
            eng.close()
        except Exception as ex:
            out = {"kernel": v, "error": str(ex)}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
