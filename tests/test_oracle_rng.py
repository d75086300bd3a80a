"""Pins the oracle's RNG / math restatement against every known answer available
without a Rust toolchain."""
import ctypes as C
import math
import os
import struct
import subprocess
import sys

import numpy as np

from tests.oracle_lib import load_oracle, rng_stream

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_xoroshiro128plus_known_answer_vector():
    # rand_xoshiro's published reference vector for Xoroshiro128Plus from state (1, 2)
    state = np.array([1, 2], np.uint64)
    got = rng_stream(state, 0, 10)
    want = [3, 412333834243, 2360170716294286339, 9295852285959843169, 2797080929874688578,
            6019711933173041966, 3076529664176959358, 3521761819100106140, 7493067640054542992,
            920801338098114767]
    assert [int(x) for x in got] == want


def test_seed_from_u64_splitmix():
    L = load_oracle()
    s = np.zeros(2, np.uint64)
    L.oracle_rng_seed(0, s.ctypes.data_as(C.POINTER(C.c_uint64)))
    assert (int(s[0]), int(s[1])) == (0xe220a8397b1dcdaf, 0x6e789e6aa1b965f4)
    L.oracle_rng_seed(10137, s.ctypes.data_as(C.POINTER(C.c_uint64)))
    assert (int(s[0]), int(s[1])) == (0x02edf8daf2766d59, 0x5d8be32458db11f2)


def _py_next(s):
    M = (1 << 64) - 1
    a, b = s
    out = (a + b) & M
    b ^= a
    rotl = lambda x, k: ((x << k) | (x >> (64 - k))) & M
    s[0] = rotl(a, 24) ^ b ^ ((b << 16) & M)
    s[1] = rotl(b, 37)
    return out


def test_gen_f64_and_int_ranges_against_python_model():
    seed = [0xe220a8397b1dcdaf, 0x6e789e6aa1b965f4]
    # gen::<f64>() = (u64 >> 11) * 2^-53
    st = np.array(seed, np.uint64)
    got = rng_stream(st, 1, 100).view(np.float64)
    s = list(seed)
    want = [(_py_next(s) >> 11) * 2.0 ** -53 for _ in range(100)]
    assert list(got) == want
    # gen_range(0, n): zone = (n << lzcnt(n)) - 1   (UniformInt::sample_single, rand 0.7)
    for n in (2, 3, 4, 31, 32, 38, 150, 1000):
        st = np.array(seed, np.uint64)
        got = rng_stream(st, 2, 500, n_arg=n)
        s = list(seed)
        want = []
        zone = ((n << (64 - n.bit_length())) - 1) & ((1 << 64) - 1)
        while len(want) < 500:
            m = _py_next(s) * n
            if (m & ((1 << 64) - 1)) <= zone:
                want.append(m >> 64)
        assert [int(x) for x in got] == want
        assert (int(st[0]), int(st[1])) == tuple(s)
        # Uniform::new(0, n).sample: zone = 2^64 - 1 - (2^64 - n) % n
        st = np.array(seed, np.uint64)
        got = rng_stream(st, 3, 500, n_arg=n)
        s = list(seed)
        want = []
        zone = (1 << 64) - 1 - ((1 << 64) - n) % n
        while len(want) < 500:
            m = _py_next(s) * n
            if (m & ((1 << 64) - 1)) <= zone:
                want.append(m >> 64)
        assert [int(x) for x in got] == want


def test_power_of_two_gen_range_rejects_half():
    # for N = 32, gen_range's conservative zone rejects ~half of the draws
    st = np.array([1, 2], np.uint64)
    rng_stream(st, 2, 2000, n_arg=32)
    st2 = np.array([1, 2], np.uint64)
    rng_stream(st2, 0, 4000)  # if exactly half were rejected the states would be near each other
    # count draws actually consumed
    s = [1, 2]
    used = 0
    got = 0
    zone = ((32 << (64 - 6)) - 1) & ((1 << 64) - 1)
    while got < 2000:
        used += 1
        if ((_py_next(s) * 32) & ((1 << 64) - 1)) <= zone:
            got += 1
    assert 3700 < used < 4300
    assert (int(st[0]), int(st[1])) == tuple(s)


def test_uniform_f64_is_52_bit_mantissa():
    seed = [0xe220a8397b1dcdaf, 0x6e789e6aa1b965f4]
    st = np.array(seed, np.uint64)
    got = rng_stream(st, 5, 200, lo=-1.0, hi=1.0).view(np.float64)
    s = list(seed)
    want = []
    for _ in range(200):
        v = _py_next(s) >> 12
        f = struct.unpack("<d", struct.pack("<Q", v | (1023 << 52)))[0]
        want.append((f - 1.0) * 2.0 + -1.0)
    assert list(got) == want
    assert got.min() >= -1.0 and got.max() < 1.0


def test_zig_tables_match_generator_and_survey_constants():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_ziggurat_tables as g
    x, f = g.tables()
    L = load_oracle()
    X = np.zeros(257)
    F = np.zeros(257)
    L.oracle_zig_tables(X.ctypes.data_as(C.POINTER(C.c_double)), F.ctypes.data_as(C.POINTER(C.c_double)))
    assert list(X) == x and list(F) == f
    # constants recalled from rand_distr's ziggurat_tables.rs
    assert ["%.18f" % v for v in X[:6]] == ["3.910757959537090045", "3.654152885361008796", "3.449278298560964462",
                                           "3.320244733839166074", "3.224575052047029100", "3.147889289517149969"]
    assert "%.18f" % X[255] == "0.215241895913273806" and X[256] == 0.0
    assert ["%.18f" % v for v in F[:3]] == ["0.000477467764586655", "0.001260285930498598", "0.002609072746106363"]
    assert F[256] == 1.0
    assert np.all(np.diff(X) < 0) and np.all(np.diff(F) > 0)


def test_zig_table_header_is_exactly_what_the_generator_writes():
    """include/sadmc_zig_tables.h is shared by the oracle and the kernels, so no parity test can see an edit to it:
    this one regenerates the header text from tools/gen_ziggurat_tables.py and compares it byte for byte."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_ziggurat_tables as g
    with open(os.path.join(ROOT, "include", "sadmc_zig_tables.h")) as f:
        assert f.read() == g.render()


def test_gen_range_f64_redraws_with_the_same_scale():
    """rand 0.7.3 UniformFloat::sample_single shrinks `scale` only when high - low overflowed; for finite bounds a
    result that rounds up to `high` is redrawn with the SAME scale.  [1, 1 + 2 ulp) makes that case common:
    value0_1 * scale + low rounds to `high` for about a quarter of the draws."""
    lo, hi = 1.0, 1.0 + 2 * 2.220446049250313e-16
    st = np.array([0xe220a8397b1dcdaf, 0x6e789e6aa1b965f4], np.uint64)
    s = [int(st[0]), int(st[1])]
    got = rng_stream(st, 7, 400, lo=lo, hi=hi).view(np.float64)
    want = []
    while len(want) < 400:
        v = _py_next(s) >> 12
        f = struct.unpack("<d", struct.pack("<Q", v | (1023 << 52)))[0]
        res = (f - 1.0) * (hi - lo) + lo
        if res < hi:
            want.append(res)
    assert list(got) == want
    assert (int(st[0]), int(st[1])) == (s[0], s[1])  # the same number of words was consumed


def test_standard_normal_moments_and_tail():
    st = np.array([0xe220a8397b1dcdaf, 0x6e789e6aa1b965f4], np.uint64)
    z = rng_stream(st, 4, 2_000_000).view(np.float64)
    n = z.size
    assert abs(z.mean()) < 4 / math.sqrt(n)
    assert abs(z.var() - 1.0) < 5 * math.sqrt(2.0 / n)
    assert abs((z ** 4).mean() - 3.0) < 0.03
    # tail beyond R = 3.654 exercises zero_case: P(|z| > R) = 2.58e-4
    tail = np.mean(np.abs(z) > 3.654152885361008796)
    assert 1.9e-4 < tail < 3.3e-4
    # symmetric
    assert abs(np.mean(z > 0) - 0.5) < 3 / math.sqrt(n)
    # KS-like check on the CDF at a few points
    from math import erf
    for q in (-2.0, -1.0, -0.3, 0.0, 0.7, 1.5, 2.5):
        cdf = 0.5 * (1 + erf(q / math.sqrt(2)))
        assert abs(np.mean(z < q) - cdf) < 4 * math.sqrt(cdf * (1 - cdf) / n) + 1e-4


def _ulp_diff(a, b):
    ia = struct.unpack("<q", struct.pack("<d", a))[0]
    ib = struct.unpack("<q", struct.pack("<d", b))[0]
    return abs(ia - ib)


def test_shared_exp_log_within_one_ulp_of_libm():
    L = load_oracle()
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.uniform(-745, 709, 20000), rng.uniform(-40, 0, 40000), rng.uniform(-1, 1, 20000),
                         -np.exp(rng.uniform(-40, 3, 20000)), [0.0, -0.0, 1.0, -1.0, 709.78, -745.1, 1e-300, -1e-10]])
    worst = 0
    for x in xs:
        worst = max(worst, _ulp_diff(L.oracle_exp(float(x)), math.exp(float(x))))
    assert worst <= 1
    assert L.oracle_exp(710.0) == math.inf and L.oracle_exp(-746.0) == 0.0
    ys = np.concatenate([np.exp(rng.uniform(-700, 700, 30000)), rng.uniform(0.5, 2.0, 30000), rng.uniform(0, 1, 30000),
                         [1.0, 2.0, 0.5, 5e-324, 1e-310, 1.7e308]])
    worst = 0
    for y in ys:
        if y > 0:
            worst = max(worst, _ulp_diff(L.oracle_log(float(y)), math.log(float(y))))
    assert worst <= 1
    assert L.oracle_log(0.0) == -math.inf and math.isnan(L.oracle_log(-1.0))


def test_erf_inv_round_trip():
    L = load_oracle()
    for x in np.linspace(-0.999, 0.999, 401):
        y = L.oracle_erf_inv(float(x))
        assert abs(math.erf(y) - x) < 4e-16
