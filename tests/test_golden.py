"""Committed fixtures (tests/golden/, written by tests/golden/make_golden.py from the CPU oracle).

CPU: the oracle still reproduces them (pins the restatement against accidental change).
GPU: the CUDA engine reproduces them through the C ABI without consulting the oracle at test time."""
import os

import numpy as np
import pytest

from tests.golden.make_golden import CASES, config_of

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EXACT = {"ising32_sad", "ising8_wl", "fake_quadratic3_sad", "two_wells_sad", "sw100_sad", "lj31_sad"}


def _check(name, w, b, system, exact=True):
    g = np.load(os.path.join(HERE, name + ".npz"))
    sc = [w.accepted_moves, w.rng_s0, w.rng_s1, w.tL, w.tF, w.num_states, w.highest_hist, w.bins_len, w.max_S_index]
    wl = "_wl" in name
    keep = [0, 1, 2, 7, 8] if wl else list(range(9))  # tL, tF, num_states, highest_hist are SAD state
    assert [int(g["scalars"][k]) for k in keep] == [int(sc[k]) for k in keep], name
    fl = np.array([w.energy, w.bins_min, w.too_lo, w.too_hi, w.latest_parameter, w.acceptance_rate, w.max_S, w.wl_gamma,
                   w.wl_num_states])
    for k in ("histogram", "t_found", "round_trips"):
        assert np.array_equal(g[k], b[k]), (name, k)
    fkeep = [0, 1, 5, 6, 7, 8] if wl else list(range(7))  # too_lo/too_hi/latest_parameter: SAD; wl_*: WL
    gf, fl = g["floats"][fkeep], fl[fkeep]
    if exact:
        assert np.array_equal(gf, fl), name
        assert np.array_equal(g["lnw"], b["lnw"]) and np.array_equal(g["energy_total"], b["energy_total"]), name
        assert np.array_equal(g["system"], system), name
    else:
        assert np.allclose(gf, fl, rtol=1e-12, atol=1e-12)
        assert np.allclose(g["lnw"], b["lnw"], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_fixture(name):
    from tests.oracle_lib import OracleMC
    kw, walker, moves = CASES[name]
    o = OracleMC(config_of(kw), walker=walker)
    o.run(moves)
    _check(name, o.walker(), o.bins(), o.system())


def test_rng_fixture():
    from tests.oracle_lib import rng_stream
    g = np.load(os.path.join(HERE, "rng_streams.npz"))
    st = np.array([1, 2], np.uint64)
    assert np.array_equal(rng_stream(st, 0, 64), g["u64_from_1_2"])
    assert int(g["u64_from_1_2"][1]) == 412333834243  # rand_xoshiro's published vector
    st = np.array([0xe220a8397b1dcdaf, 0x6e789e6aa1b965f4], np.uint64)
    assert np.array_equal(rng_stream(st.copy(), 4, 4096), g["normal_bits_seed0"])
    assert np.array_equal(rng_stream(st.copy(), 2, 1024, n_arg=32), g["gen_range32_seed0"])
    assert np.array_equal(rng_stream(st.copy(), 3, 1024, n_arg=31), g["uniform31_seed0"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_engine_reproduces_fixture(name):
    from sad_monte_carlo_b200 import WalkerEngine
    kw, walker, moves = CASES[name]
    eng = WalkerEngine(config_of(kw))
    eng.run(moves // 3)
    eng.run(moves - moves // 3)
    _check(name, eng.walker(walker), eng.bins(walker), eng.system(walker), exact=name in EXACT)
