#!/usr/bin/env python3
"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/), which is the only runnable restatement of the
reference in this environment (the Rust crate cannot be built: no cargo/rustc, no vendored crates).

The fixtures pin (a) the oracle against accidental change and (b) the CUDA engine against the oracle without needing
the oracle at test time.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from sad_monte_carlo_b200 import make_config, _abi  # noqa: E402
from tests.oracle_lib import OracleMC, rng_stream  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (config kwargs, walker, moves)
    "ising32_sad": (dict(system="ising", method="sad", N=32, sad_min_T=1.0), 3, 200000),
    "ising8_wl": (dict(system="ising", method="wl", N=8, wl_min_gamma=1e-3, min_allowed_energy=-128.0,
                       max_allowed_energy=50.0, seed=3), 1, 200000),
    "fake_quadratic3_sad": (dict(system="fake", method="sad", fake_function=_abi.FAKE_QUADRATIC, N=3, sad_min_T=0.001,
                                 energy_bin=0.01, seed=3), 0, 100000),
    "two_wells_sad": (dict(system="two-wells", method="sad", N=12, tw_h2_to_h1=1.1, tw_barrier_over_h1=0.1, tw_r2=0.5,
                           sad_min_T=0.001, energy_bin=1e-3, move_value=1e-2, seed=1), 2, 100000),
    "sw100_sad": (dict(system="sw", method="sad", N=100, filling_fraction=0.3, sad_min_T=0.5), 1, 30000),
    "lj31_sad": (dict(system="lj", method="sad", N=31, lj_radius=2.5, max_allowed_energy=0.0, sad_min_T=0.01,
                      energy_bin=0.01, init_mode=_abi.INIT_RANDOMIZE, lanes_per_walker=1, bin_window_lo=-133.62,
                      bin_window_hi=0.02), 5, 60000),
}


def config_of(kw):
    kw = dict(kw)
    return make_config(kw.pop("system"), kw.pop("method"), n_walkers=8, **kw)


def main():
    for name, (kw, walker, moves) in CASES.items():
        o = OracleMC(config_of(kw), walker=walker)
        o.run(moves)
        w = o.walker()
        b = o.bins()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), moves=moves, walker=walker,
                            scalars=np.array([w.accepted_moves, w.rng_s0, w.rng_s1, w.tL, w.tF, w.num_states,
                                              w.highest_hist, w.bins_len, w.max_S_index], dtype=np.uint64),
                            floats=np.array([w.energy, w.bins_min, w.too_lo, w.too_hi, w.latest_parameter,
                                             w.acceptance_rate, w.max_S, w.wl_gamma, w.wl_num_states]),
                            histogram=b["histogram"], t_found=b["t_found"], lnw=b["lnw"], energy_total=b["energy_total"],
                            round_trips=b["round_trips"], wl_hist=b["wl_hist"], system=o.system())
        print(name, "moves", moves, "bins", w.bins_len, "E", w.energy)
    # RNG streams: the published xoroshiro128+ vector and the restated rand 0.7 / rand_distr 0.2 samplers
    st = np.array([1, 2], np.uint64)
    u64 = rng_stream(st, 0, 64)
    st = np.array([0xe220a8397b1dcdaf, 0x6e789e6aa1b965f4], np.uint64)
    normals = rng_stream(st.copy(), 4, 4096)
    ranges = rng_stream(st.copy(), 2, 1024, n_arg=32)
    uniforms = rng_stream(st.copy(), 3, 1024, n_arg=31)
    np.savez_compressed(os.path.join(HERE, "rng_streams.npz"), u64_from_1_2=u64, normal_bits_seed0=normals,
                        gen_range32_seed0=ranges, uniform31_seed0=uniforms)


if __name__ == "__main__":
    main()
