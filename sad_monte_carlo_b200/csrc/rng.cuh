// rng.cuh -- per-walker random streams on the device.
//
// The reference's `MyRng` is rand_xoshiro::Xoroshiro128Plus (src/rng.rs:27; the
// algorithm is spelled out in the unused in-tree twin, src/rng.rs:48-57), driven
// through rand 0.7 (`gen::<f64>`, `gen_range`, `Uniform`) and rand_distr 0.2
// (`StandardNormal`).  Each walker keeps its 16-byte state in registers for the
// whole launch; there is no shared generator and no counter-based substitute --
// stream parity with a reference process run with `--seed w` is the point.
#pragma once
#include <stdint.h>

#include "../../include/sadmc_math.h"
#include "../../include/sadmc_zig_tables.h"
#include "fastmath.cuh"

namespace sadmc {

#if defined(__CUDA_ARCH__)
#define SADMC_UMUL64HI(a, b) __umul64hi((a), (b))
#else
#define SADMC_UMUL64HI(a, b) ((uint64_t)(((unsigned __int128)(a) * (unsigned __int128)(b)) >> 64))
#endif

// One implementation for device code and for the host-side constructors
// (host_ctor.hpp); the oracle has its own, independent one.
struct Rng {
  uint64_t s0, s1;

  __host__ __device__ __forceinline__ uint64_t next() { // src/rng.rs:48-57
    const uint64_t a = s0;
    uint64_t b = s1;
    const uint64_t r = a + b;
    b ^= a;
    s0 = ((a << 24) | (a >> 40)) ^ b ^ (b << 16);
    s1 = (b << 37) | (b >> 27);
    return r;
  }
  // rand 0.7 Standard f64: accept test of energy.rs:465,489,498,508
  __host__ __device__ __forceinline__ double gen_f64() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }

  // rand 0.7 UniformInt::sample_single == `rng.gen_range(0, n)`
  // (ising.rs:105-106, fake.rs:129, two_wells.rs:453, erfinv.rs:102).
  // `zone` = (n << clz(n)) - 1, precomputed by zone_single().
  __host__ __device__ __forceinline__ uint32_t below(uint32_t n, uint64_t zone) {
    for (;;) {
      const uint64_t v = next();
      const uint64_t lo = v * (uint64_t)n;
      if (lo <= zone) return (uint32_t)SADMC_UMUL64HI(v, (uint64_t)n);
    }
  }
  // [1,2) from the top 52 bits, IntoFloat::into_float_with_exponent(0)
  __host__ __device__ __forceinline__ double f12() { return sadmc_bits_f64((next() >> 12) | 0x3ff0000000000000ull); }
  // rand 0.7 UniformFloat::sample with precomputed scale (lj.rs:138-140, wca.rs:258-260)
  __host__ __device__ __forceinline__ double uniform_f64(double low, double scale) { return (f12() - 1.0) * scale + low; }
  // rand 0.7 UniformFloat::sample_single == `rng.gen_range(lo, hi)` for f64 (fake.rs:107, erfinv.rs:80)
  __host__ __device__ __forceinline__ double gen_range_f64(double low, double high) {
    double scale = high - low;
    for (;;) {
      const double res = (f12() - 1.0) * scale + low;
      if (res < high) return res;
      scale = sadmc_bits_f64(sadmc_f64_bits(scale) - 1);
    }
  }
  __host__ __device__ __forceinline__ double open01() { return f12() - (1.0 - 2.220446049250313e-16 / 2.0); }

  // rand_distr 0.2 StandardNormal (ziggurat, symmetric); zx/zf are the 257-entry
  // tables staged in shared memory.  (src/rng.rs:111-117, fake.rs:131, erfinv.rs:104)
  // The rare part (tail + wedge tests, ~1.2 % of draws) is one out-of-line copy that takes and
  // returns the generator state BY VALUE: its exp/log bodies stay out of the hot loop's
  // instruction footprint and the state never has its address taken (it stays in registers).
  struct SlowOut {
    double x;
    uint64_t s0, s1;
  };
  static __host__ __device__ __noinline__ SlowOut normal_slow(uint64_t s0_, uint64_t s1_, const double* zx, const double* zf, uint32_t i,
                                                            double u, double x) {
    Rng r;
    r.s0 = s0_;
    r.s1 = s1_;
    SlowOut o;
    for (;;) {
      if (i == 0) {
        double xx = 1.0, yy = 0.0;
        while (-2.0 * yy < xx * xx) {
          const double a = r.open01();
          const double b = r.open01();
          xx = sadmc_log(a) / SADMC_ZIG_NORM_R;
          yy = sadmc_log(b);
        }
        o.x = u < 0.0 ? xx - SADMC_ZIG_NORM_R : SADMC_ZIG_NORM_R - xx;
        break;
      }
      if (exp_cmp(zf[i + 1] + (zf[i] - zf[i + 1]) * r.gen_f64(), -x * x / 2.0) < 0) { // lhs < exp(-x^2/2), decided exactly
        o.x = x;
        break;
      }
      const uint64_t bits = r.next();
      i = (uint32_t)(bits & 0xff);
      u = sadmc_bits_f64((bits >> 12) | 0x4000000000000000ull) - 3.0;
      x = u * zx[i];
      if (fabs(x) < zx[i + 1]) {
        o.x = x;
        break;
      }
    }
    o.s0 = r.s0;
    o.s1 = r.s1;
    return o;
  }
  __host__ __device__ __forceinline__ double normal(const double* zx, const double* zf) {
    const uint64_t bits = next();
    const uint32_t i = (uint32_t)(bits & 0xff);
    const double u = sadmc_bits_f64((bits >> 12) | 0x4000000000000000ull) - 3.0;
    const double x = u * zx[i];
    if (fabs(x) < zx[i + 1]) return x;
    const SlowOut o = normal_slow(s0, s1, zx, zf, i, u, x);
    s0 = o.s0;
    s1 = o.s1;
    return o.x;
  }
  // Three successive StandardNormal draws (crate::rng::vector, src/rng.rs:111-117) with the same
  // stream semantics as three normal() calls.  The three raw words are drawn first and the three
  // ziggurat fast paths evaluated side by side (independent chains); only when one of them leaves
  // the fast path (3.6 % of calls) is the stream rewound to just after that word and finished
  // serially: ONE out-of-line slow call, then the remaining draws.
  __host__ __device__ __forceinline__ void normal3(const double* zx, const double* zf, double& v0, double& v1, double& v2) {
    const uint64_t b0 = next();
    const uint64_t p0 = s0, q0 = s1;
    const uint64_t b1 = next();
    const uint64_t p1 = s0, q1 = s1;
    const uint64_t b2 = next();
    const uint32_t i0 = (uint32_t)(b0 & 0xff), i1 = (uint32_t)(b1 & 0xff), i2 = (uint32_t)(b2 & 0xff);
    const double u0 = sadmc_bits_f64((b0 >> 12) | 0x4000000000000000ull) - 3.0;
    const double u1 = sadmc_bits_f64((b1 >> 12) | 0x4000000000000000ull) - 3.0;
    const double u2 = sadmc_bits_f64((b2 >> 12) | 0x4000000000000000ull) - 3.0;
    v0 = u0 * zx[i0];
    v1 = u1 * zx[i1];
    v2 = u2 * zx[i2];
    const bool ok0 = fabs(v0) < zx[i0 + 1], ok1 = fabs(v1) < zx[i1 + 1], ok2 = fabs(v2) < zx[i2 + 1];
    if (ok0 && ok1 && ok2) return;
    const int f = !ok0 ? 0 : (!ok1 ? 1 : 2); // first draw that needs the slow path
    const SlowOut o = normal_slow(f == 0 ? p0 : (f == 1 ? p1 : s0), f == 0 ? q0 : (f == 1 ? q1 : s1), zx, zf,
                                  f == 0 ? i0 : (f == 1 ? i1 : i2), f == 0 ? u0 : (f == 1 ? u1 : u2),
                                  f == 0 ? v0 : (f == 1 ? v1 : v2));
    s0 = o.s0;
    s1 = o.s1;
    if (f == 0) {
      v0 = o.x;
      v1 = normal(zx, zf);
    } else if (f == 1) {
      v1 = o.x;
    } else {
      v2 = o.x;
    }
    if (f < 2) v2 = normal(zx, zf);
  }
};

// zone of UniformInt::sample_single (conservative power-of-two approximation)
__host__ __device__ inline uint64_t zone_single(uint64_t n) {
  int lz = 0;
  for (uint64_t t = n; !(t >> 63); t <<= 1) lz++;
  return (n << lz) - 1;
}
// zone of Uniform::new(0, n).sample (exact): 2^64 - 1 - (2^64 - n) % n
__host__ __device__ inline uint64_t zone_uniform(uint64_t n) { return ~0ull - ((0ull - n) % n); }

// SplitMix64 seeding of Xoroshiro128Plus::seed_from_u64 (energy.rs:835)
__host__ __device__ inline void seed_from_u64(uint64_t seed, uint64_t* s0, uint64_t* s1) {
  uint64_t x = seed;
  uint64_t out[2];
  for (int k = 0; k < 2; k++) {
    x += 0x9e3779b97f4a7c15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    out[k] = z ^ (z >> 31);
  }
  *s0 = out[0];
  *s1 = out[1];
}

} // namespace sadmc
