"""CPU checks of the tempering oracle (oracle/oracle_tempering.hpp) and of the pieces it takes from un-vendored crates."""
import ctypes as C

import numpy as np

from sad_monte_carlo_b200 import _abi, make_config
from tests.oracle_lib import OracleTempering, load_oracle, u64p


def _jump_python(s0, s1):
    """xoroshiro128+ (24, 16, 37) jump with the published 2^64 polynomial, in plain integers."""
    M = (1 << 64) - 1
    a0 = a1 = 0
    for word in (0xdf900294d8f554a5, 0x170865df4b3201fc):
        for b in range(64):
            if word >> b & 1:
                a0 ^= s0
                a1 ^= s1
            t = s1 ^ s0
            s0 = (((s0 << 24) | (s0 >> 40)) & M) ^ t ^ ((t << 16) & M)
            s1 = ((t << 37) | (t >> 27)) & M
    return a0, a1


def test_jump_matches_an_independent_model_and_commutes_with_stepping():
    L = load_oracle()
    for seed in (0, 1, 12345):
        s = np.zeros(2, np.uint64)
        L.oracle_rng_seed(seed, s.ctypes.data_as(u64p))
        want = _jump_python(int(s[0]), int(s[1]))
        j = s.copy()
        L.oracle_rng_jump(j.ctypes.data_as(u64p))
        assert (int(j[0]), int(j[1])) == want
        # jump is a power of the state transition: stepping once then jumping == jumping then stepping once
        out = np.zeros(1, np.uint64)
        a = s.copy()
        L.oracle_rng_stream(a.ctypes.data_as(u64p), 0, 0, 0.0, 0.0, 1, out.ctypes.data_as(u64p))
        L.oracle_rng_jump(a.ctypes.data_as(u64p))
        b = j.copy()
        L.oracle_rng_stream(b.ctypes.data_as(u64p), 0, 0, 0.0, 0.0, 1, out.ctypes.data_as(u64p))
        assert np.array_equal(a, b)


def test_replicas_start_from_clones_of_one_generator_and_the_simulation_generator_is_its_jump():
    cfg = make_config("fake", fake_function=_abi.FAKE_QUADRATIC, N=3, seed=6)
    o = OracleTempering(cfg, [0.1, 0.2, 0.4], 1)
    reps = o.replicas()
    assert len({(r.rng_s0, r.rng_s1) for r in reps}) == 1  # tempering.rs:161 rng.clone()
    assert o.rng() == _jump_python(reps[0].rng_s0, reps[0].rng_s1)  # tempering.rs:164


def test_moves_count_and_swap_bookkeeping():
    cfg = make_config("two-wells", N=12, tw_h2_to_h1=1.1, tw_barrier_over_h1=0.1, tw_r2=0.5, seed=3)
    T = [0.01 * 2 ** i for i in range(7)]
    o = OracleTempering(cfg, T, canonical_steps=10)
    o.run_once(500)
    assert o.moves == 500 * 7 * 12 * 10  # steps = min_moves_to_randomize * canonical_steps per replica (tempering.rs:274)
    reps = o.replicas()
    swaps = [r.accepted_swap_count + r.rejected_swap_count for r in reps]
    # end replicas take part in every other round on average, inner ones in every round ... for 7 replicas: pairs
    # (0,1)(2,3)(4,5) or (1,2)(3,4)(5,6)
    assert swaps[0] + swaps[1] - swaps[0] == swaps[1] and all(s == 500 for s in swaps[1:6]) and swaps[0] + swaps[6] == 500
    # a swap attempt is counted by both partners
    assert sum(r.accepted_swap_count for r in reps) % 2 == 0
    e = [r.total_energy / (r.accepted_count + r.rejected_count + s) for r, s in zip(reps, swaps)]
    assert e[0] < e[-1]  # colder replicas sit lower
