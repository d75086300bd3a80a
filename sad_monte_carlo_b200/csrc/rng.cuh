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
      // the scale shrinks only when high - low overflowed (rand 0.7.3 `decrease_masked(!scale.finite_mask())`);
      // a finite scale draws again
      if (!(fabs(scale) <= 1.7976931348623157e308)) scale = sadmc_bits_f64(sadmc_f64_bits(scale) - 1);
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
  // Three successive StandardNormal draws (crate::rng::vector, src/rng.rs:111-117) with the same stream
  // semantics as three normal() calls.
  //
  // 3.6 % of the calls leave the ziggurat's fast path somewhere, i.e. 69 % of all warps per move: a
  // divergent serial slow path costs every warp what it costs the slowest lane (measured: 13 % of the LJ
  // kernel).  So the stream is evaluated speculatively instead.  Words w0..w4 of the generator are taken
  // as they come; each is turned into a fast-path normal n_k (valid when ok_k).  If word f is the first
  // to fail, the reference's next steps are fixed: w_{f+1} is the uniform of the wedge test
  // (zf[i+1] + (zf[i] - zf[i+1]) u < exp(-x^2/2)); if it accepts, x_f stands and the later draws shift by
  // one word; if it rejects, w_{f+2} is the redraw and they shift by two.  Everything is selected from
  // registers without a branch.  Only a second irregular event in the same call (tail layer i == 0, two
  // failures, a failing redraw: ~0.1 % of calls) rewinds the stream and replays it serially.
  struct Slow3Out {
    double v0, v1, v2;
    uint64_t s0, s1;
  };
  // What normal3 does once one of its three words has left the fast path.  (a0, a1) = generator state before the
  // call, (b0, b1) = after the third word.  Out of line on request (SADMC_ZIG_SLOW_NOINLINE): 69 % of the warps
  // come here every move for a lane or two, but its ~150 instructions then sit outside the move loop's
  // instruction-cache footprint.
#ifdef SADMC_ZIG_SLOW_NOINLINE
  static __host__ __device__ __noinline__ Slow3Out
#else
  static __host__ __device__ __forceinline__ Slow3Out
#endif
  normal3_slow(uint64_t a0, uint64_t a1, uint64_t b0, uint64_t b1, uint64_t w0, uint64_t w1, uint64_t w2, const double* zx, const double* zf) {
    Rng r;
    r.s0 = b0;
    r.s1 = b1;
#define SADMC_ZIG_FAST(k)                                                                 \
  const uint32_t i##k = (uint32_t)(w##k & 0xff);                                           \
  const double n##k = (sadmc_bits_f64((w##k >> 12) | 0x4000000000000000ull) - 3.0) * zx[i##k]; \
  const bool ok##k = fabs(n##k) < zx[i##k + 1];
    SADMC_ZIG_FAST(0)
    SADMC_ZIG_FAST(1)
    SADMC_ZIG_FAST(2)
    const uint64_t w3 = r.next();
    const uint64_t p3 = r.s0, q3 = r.s1;
    const uint64_t w4 = r.next();
    const uint64_t p4 = r.s0, q4 = r.s1;
    SADMC_ZIG_FAST(3)
    SADMC_ZIG_FAST(4)
#undef SADMC_ZIG_FAST
    Slow3Out o;
    const int f = !ok0 ? 0 : (!ok1 ? 1 : 2); // first word that left the fast path
    const uint32_t fi = f == 0 ? i0 : (f == 1 ? i1 : i2);
    const double fx = f == 0 ? n0 : (f == 1 ? n1 : n2);
    const uint64_t fu = f == 0 ? w1 : (f == 1 ? w2 : w3); // the word the wedge test draws its uniform from
    const double u01 = (double)(fu >> 11) * (1.0 / 9007199254740992.0);
    const bool acc = exp_cmp(zf[fi + 1] + (zf[fi] - zf[fi + 1]) * u01, -fx * fx / 2.0) < 0;
    // fast-path validity of the words that the shifted stream turns into normals
    const bool need_ok = f == 0 ? (acc ? (ok2 && ok3) : (ok2 && ok3 && ok4)) : (f == 1 ? (acc ? ok3 : (ok3 && ok4)) : (acc ? true : ok4));
    if (fi != 0 && need_ok) {
      o.v0 = f == 0 ? (acc ? n0 : n2) : n0;
      o.v1 = f == 0 ? (acc ? n2 : n3) : (f == 1 ? (acc ? n1 : n3) : n1);
      o.v2 = f == 2 ? (acc ? n2 : n4) : (acc ? n3 : n4);
      o.s0 = acc ? p3 : p4;
      o.s1 = acc ? q3 : q4;
      return o;
    }
    r.s0 = a0; // replay serially from the start of the call
    r.s1 = a1;
    o.v0 = r.normal(zx, zf);
    o.v1 = r.normal(zx, zf);
    o.v2 = r.normal(zx, zf);
    o.s0 = r.s0;
    o.s1 = r.s1;
    return o;
  }
  __host__ __device__ __forceinline__ void normal3(const double* zx, const double* zf, double& v0, double& v1, double& v2) {
    const uint64_t a0 = s0, a1 = s1; // for the replay
    const uint64_t w0 = next();
    const uint64_t w1 = next();
    const uint64_t w2 = next();
#define SADMC_ZIG_FAST(k)                                                                 \
  const uint32_t i##k = (uint32_t)(w##k & 0xff);                                           \
  const double n##k = (sadmc_bits_f64((w##k >> 12) | 0x4000000000000000ull) - 3.0) * zx[i##k]; \
  const bool ok##k = fabs(n##k) < zx[i##k + 1];
    SADMC_ZIG_FAST(0)
    SADMC_ZIG_FAST(1)
    SADMC_ZIG_FAST(2)
#undef SADMC_ZIG_FAST
    v0 = n0;
    v1 = n1;
    v2 = n2;
    if (ok0 && ok1 && ok2) return;
#ifdef SADMC_ABL_NOSLOW /* ablation experiment only: wrong statistics */
    return;
#endif
    const Slow3Out o = normal3_slow(a0, a1, s0, s1, w0, w1, w2, zx, zf);
    v0 = o.v0;
    v1 = o.v1;
    v2 = o.v2;
    s0 = o.s0;
    s1 = o.s1;
  }
};

// ---- the NEXT proposal's draws, evaluated one move ahead (move_kernel.cuh, Sys::PREDRAW) -----------------------
// A proposal of the cluster systems is `Uniform::new(0, N)` + three StandardNormals (lj.rs:368-369): words
// W, Z0, Z1, Z2 of the stream, Z3 / Z4 when one ziggurat draw leaves the fast path (normal3 above).  Which word
// the proposal starts at is only known once the current move's accept test has or has not drawn its uniform
// (energy.rs:465) -- and that test waits ~1 600 cycles for the bin record.  So both candidates (start at word 0,
// start at word 1) are evaluated in the shadow of that load from the seven words the two of them can touch, and the
// move loop picks one afterwards.  `ok == false` (integer rejection zone, ziggurat tail layer, two irregular draws
// in one call: ~0.1 % of proposals) means "not evaluated": the caller then draws the proposal the ordinary way from
// the unchanged generator state, so the stream is the reference's in every case.
struct PreDraw {
  double v0, v1, v2;
  uint64_t s0, s1; // generator state after the proposal's draws
  uint32_t which;
  bool ok;
};
// One candidate: ww = the integer word, z0..z4 = the following five words, (e2s0,e2s1) / (e3s0,e3s1) / (e4s0,e4s1) =
// generator state after z2 / z3 / z4.
__device__ __forceinline__ PreDraw predraw_one(uint64_t ww, uint64_t z0, uint64_t z1, uint64_t z2, uint64_t z3, uint64_t z4, uint64_t e2s0,
                                               uint64_t e2s1, uint64_t e3s0, uint64_t e3s1, uint64_t e4s0, uint64_t e4s1, uint32_t n, uint64_t zone,
                                               const double* zx, const double* zf) {
  PreDraw r;
  r.which = (uint32_t)SADMC_UMUL64HI(ww, (uint64_t)n);
  const bool which_ok = ww * (uint64_t)n <= zone; // Rng::below's acceptance
#define SADMC_ZIG_FAST(k)                                                                 \
  const uint32_t i##k = (uint32_t)(z##k & 0xff);                                           \
  const double n##k = (sadmc_bits_f64((z##k >> 12) | 0x4000000000000000ull) - 3.0) * zx[i##k]; \
  const bool ok##k = fabs(n##k) < zx[i##k + 1];
  SADMC_ZIG_FAST(0)
  SADMC_ZIG_FAST(1)
  SADMC_ZIG_FAST(2)
  if (ok0 && ok1 && ok2) {
    r.v0 = n0;
    r.v1 = n1;
    r.v2 = n2;
    r.s0 = e2s0;
    r.s1 = e2s1;
    r.ok = which_ok;
    return r;
  }
  SADMC_ZIG_FAST(3)
  SADMC_ZIG_FAST(4)
#undef SADMC_ZIG_FAST
  // exactly Rng::normal3's single-irregular-draw case
  const int f = !ok0 ? 0 : (!ok1 ? 1 : 2);
  const uint32_t fi = f == 0 ? i0 : (f == 1 ? i1 : i2);
  const double fx = f == 0 ? n0 : (f == 1 ? n1 : n2);
  const uint64_t fu = f == 0 ? z1 : (f == 1 ? z2 : z3);
  const double u01 = (double)(fu >> 11) * (1.0 / 9007199254740992.0);
  const bool acc = exp_cmp(zf[fi + 1] + (zf[fi] - zf[fi + 1]) * u01, -fx * fx / 2.0) < 0;
  const bool need_ok = f == 0 ? (acc ? (ok2 && ok3) : (ok2 && ok3 && ok4)) : (f == 1 ? (acc ? ok3 : (ok3 && ok4)) : (acc ? true : ok4));
  r.v0 = f == 0 ? (acc ? n0 : n2) : n0;
  r.v1 = f == 0 ? (acc ? n2 : n3) : (f == 1 ? (acc ? n1 : n3) : n1);
  r.v2 = f == 2 ? (acc ? n2 : n4) : (acc ? n3 : n4);
  r.s0 = acc ? e3s0 : e4s0;
  r.s1 = acc ? e3s1 : e4s1;
  r.ok = which_ok && fi != 0 && need_ok;
  return r;
}
// Both candidates from generator state `g` (taken by value: the caller's generator is not advanced).
__device__ __forceinline__ void predraw_both(Rng g, uint32_t n, uint64_t zone, const double* zx, const double* zf, PreDraw& a, PreDraw& b) {
  const uint64_t w0 = g.next();
  const uint64_t w1 = g.next();
  const uint64_t w2 = g.next();
  const uint64_t w3 = g.next();
  const uint64_t t3s0 = g.s0, t3s1 = g.s1; // after w3
  const uint64_t w4 = g.next();
  const uint64_t t4s0 = g.s0, t4s1 = g.s1;
  const uint64_t w5 = g.next();
  const uint64_t t5s0 = g.s0, t5s1 = g.s1;
  const uint64_t w6 = g.next();
  a = predraw_one(w0, w1, w2, w3, w4, w5, t3s0, t3s1, t4s0, t4s1, t5s0, t5s1, n, zone, zx, zf);
  b = predraw_one(w1, w2, w3, w4, w5, w6, t4s0, t4s1, t5s0, t5s1, g.s0, g.s1, n, zone, zx, zf);
}

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
