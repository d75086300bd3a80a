// oracle_rng.hpp -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// CPU restatement of the random-number path the reference's hot loop uses.
// The arithmetic lives in un-vendored crates (Cargo.toml:21-41, no Cargo.lock):
//   rand_xoshiro ^0.4  Xoroshiro128Plus (`MyRng`, src/rng.rs:27; algorithm visible in
//                      the unused in-tree twin src/rng.rs:48-57), SplitMix64 seeding
//   rand ^0.7.2        Standard f64, gen_range / Uniform for usize and f64
//   rand_distr ^0.2.2  StandardNormal (256-layer ziggurat), Open01
// so each routine restates the crate's published algorithm; the pinning this
// container allows is the generator's public known-answer vector and the seed
// constants listed in SURVEY.md section 8(c) (tests/test_oracle_rng.py).
// PARITY UNPINNED against the Rust binary for gen_range/Uniform zone rules and
// the ziggurat control flow (no Rust toolchain here).
//
// Written independently of sad_monte_carlo_b200/csrc/rng.cuh on purpose: the
// GPU parity tests compare two implementations, not one implementation twice.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../include/sadmc_math.h"       // shared exp/log (documented there)
#include "../include/sadmc_zig_tables.h" // generated tables (data, not code)

namespace oracle {

static const double ZIG_X[257] = SADMC_ZIG_NORM_X_INIT;
static const double ZIG_F[257] = SADMC_ZIG_NORM_F_INIT;

// Switchable so a test can show that libm vs shared exp/log give the same trajectories.
struct MathMode {
  static bool& use_libm() {
    static bool v = false;
    return v;
  }
};
inline double o_exp(double x) { return MathMode::use_libm() ? std::exp(x) : sadmc_exp(x); }
inline double o_log(double x) { return MathMode::use_libm() ? std::log(x) : sadmc_log(x); }

struct Rng {
  uint64_t s0 = 0, s1 = 0;

  // rand_xoshiro::SplitMix64::next_u64
  static uint64_t splitmix64(uint64_t& x) {
    x += 0x9e3779b97f4a7c15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
  }
  // Xoroshiro128Plus::seed_from_u64: two SplitMix64 outputs, little-endian into s0, s1
  // (call sites: energy.rs:835, ising.rs:45, lj.rs:128, wca.rs:402, optsquare.rs:408).
  static Rng seed_from_u64(uint64_t seed) {
    Rng r;
    uint64_t x = seed;
    r.s0 = splitmix64(x);
    r.s1 = splitmix64(x);
    return r;
  }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  // src/rng.rs:48-57
  uint64_t next_u64() {
    const uint64_t a = s0;
    uint64_t b = s1;
    const uint64_t out = a + b;
    b ^= a;
    s0 = rotl(a, 24) ^ b ^ (b << 16);
    s1 = rotl(b, 37);
    return out;
  }
  // rand 0.7 Standard for f64: 53 high bits times 2^-53  (accept test, energy.rs:465,489,498,508)
  double gen_f64() { return (double)(next_u64() >> 11) * (1.0 / 9007199254740992.0); }

  static void wmul(uint64_t a, uint64_t b, uint64_t& hi, uint64_t& lo) {
    const unsigned __int128 p = (unsigned __int128)a * b;
    hi = (uint64_t)(p >> 64);
    lo = (uint64_t)p;
  }
  // rand 0.7 `rng.gen_range(0, n)` for usize == UniformInt::sample_single
  // (ising.rs:105-106, fake.rs:129, two_wells.rs:453, erfinv.rs:102).
  uint64_t gen_range_usize(uint64_t low, uint64_t high) {
    const uint64_t range = high - low;
    const uint64_t zone = (range << __builtin_clzll(range)) - 1;
    for (;;) {
      uint64_t hi, lo;
      wmul(next_u64(), range, hi, lo);
      if (lo <= zone) return low + hi;
    }
  }
  // rand 0.7 `rng.sample(Uniform::new(0, n))` for usize == UniformInt::new + sample
  // (lj.rs:368, wca.rs:345, optsquare.rs:276, 414-417).  Different rejection zone!
  uint64_t uniform_usize(uint64_t low, uint64_t high) {
    const uint64_t range = high - low; // new(low,high) == new_inclusive(low, high-1)
    const uint64_t ints_to_reject = (0ull - range) % range; // (2^64 - range) % range
    const uint64_t zone = ~0ull - ints_to_reject;
    for (;;) {
      uint64_t hi, lo;
      wmul(next_u64(), range, hi, lo);
      if (lo <= zone) return low + hi;
    }
  }
  static double bits_to_f64(uint64_t b) {
    double d;
    std::memcpy(&d, &b, 8);
    return d;
  }
  // IntoFloat::into_float_with_exponent
  static double float_with_exponent(uint64_t fraction52, int exponent) {
    return bits_to_f64(fraction52 | ((uint64_t)(1023 + exponent) << 52));
  }
  // rand 0.7 `rng.sample(Uniform::new(low, high))` for f64 (lj.rs:138-140, wca.rs:258-260, 456-465)
  double uniform_f64(double low, double high) {
    double scale = high - low;
    const double max_rand = float_with_exponent(~0ull >> 12, 0) - 1.0;
    while (scale * max_rand + low >= high) scale = bits_to_f64(sadmc_f64_bits(scale) - 1);
    const double v12 = float_with_exponent(next_u64() >> 12, 0);
    return (v12 - 1.0) * scale + low;
  }
  // rand 0.7 `rng.gen_range(low, high)` for f64 == UniformFloat::sample_single
  // (fake.rs:107, erfinv.rs:80, two_wells.rs:177)
  double gen_range_f64(double low, double high) {
    double scale = high - low;
    for (;;) {
      const double v12 = float_with_exponent(next_u64() >> 12, 0);
      const double res = (v12 - 1.0) * scale + low;
      if (res < high) return res;
      // rand 0.7.3 uniform.rs, UniformFloat::sample_single: `let mask = !scale.finite_mask(); if mask.any() { scale =
      // scale.decrease_masked(mask) }` -- the scale only shrinks when high - low overflowed; a finite scale whose
      // rounding produced res >= high simply draws again.
      if (!std::isfinite(scale)) scale = bits_to_f64(sadmc_f64_bits(scale) - 1);
    }
  }
  // rand 0.7 Open01 for f64: (0,1)
  double open01() {
    const double v12 = float_with_exponent(next_u64() >> 12, 0);
    return v12 - (1.0 - 2.220446049250313e-16 / 2.0);
  }
  // rand_distr 0.2 StandardNormal: ziggurat(symmetric = true)
  // (src/rng.rs:111-117 `vector`, fake.rs:131, erfinv.rs:104)
  double standard_normal() {
    const double R = SADMC_ZIG_NORM_R;
    for (;;) {
      const uint64_t bits = next_u64();
      const unsigned i = (unsigned)(bits & 0xff);
      const double u = float_with_exponent(bits >> 12, 1) - 3.0;
      const double x = u * ZIG_X[i];
      const double test_x = x < 0.0 ? -x : x;
      if (test_x < ZIG_X[i + 1]) return x;
      if (i == 0) {
        // tail
        double xx = 1.0, yy = 0.0;
        while (-2.0 * yy < xx * xx) {
          const double a = open01();
          const double b = open01();
          xx = o_log(a) / R;
          yy = o_log(b);
        }
        return u < 0.0 ? xx - R : R - xx;
      }
      if (ZIG_F[i + 1] + (ZIG_F[i] - ZIG_F[i + 1]) * gen_f64() < o_exp(-x * x / 2.0)) return x;
    }
  }
};

} // namespace oracle
