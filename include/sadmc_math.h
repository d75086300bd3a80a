/* sadmc_math.h -- exp() and log() with ONE definition for host and device.
 *
 * Why: the reference's accept test is `u > (lnw1 - lnw2).exp()`
 * (src/mc/energy.rs:465,489,498,508) and its normal sampler calls `exp`/`ln`
 * on the ziggurat's rare paths (rand_distr 0.2).  Both resolve to the platform
 * libm, which is not bit-reproducible across platforms (the reference itself is
 * built for glibc and musl, lj/run-lj.py:5-6), and CUDA's exp() is a different
 * 1-ulp implementation again.  To make "kernel == oracle, bit for bit" a
 * meaningful statement, the kernels AND the CPU oracle both evaluate these two
 * functions with the routines below: the classic Sun fdlibm argument
 * reductions and minimax polynomials (e_exp.c / e_log.c; < 1 ulp), restated
 * with only IEEE +,-,*,/ -- no FMA, no tables -- so every operation rounds
 * identically on x86 and on sm_100a.  Build flags that keep it so:
 * `-fmad=false` (nvcc) and `-ffp-contract=off` (gcc).
 * tests/test_oracle_rng.py bounds the distance to libm (<= 1 ulp) on the CPU; the
 * bit-exact GPU trajectory tests (tests/test_gpu_ising.py, test_gpu_lj.py, ...) pin
 * CPU == GPU, because one differing accept decision makes the trajectories diverge.
 */
#ifndef SADMC_MATH_H
#define SADMC_MATH_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define SADMC_HD __host__ __device__ __forceinline__
#else
#define SADMC_HD static inline
#endif

SADMC_HD uint64_t sadmc_f64_bits(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u;
  memcpy(&u, &x, 8);
  return u;
#endif
}
SADMC_HD double sadmc_bits_f64(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x;
  memcpy(&x, &u, 8);
  return x;
#endif
}

/* e^x.  |error| < 1 ulp.  Overflow -> +inf, underflow -> 0 (through subnormals). */
SADMC_HD double sadmc_exp(double x) {
  const double ln2_hi = 6.93147180369123816490e-01; /* 0x3fe62e42fee00000 */
  const double ln2_lo = 1.90821492927058770002e-10; /* 0x3dea39ef35793c76 */
  const double inv_ln2 = 1.44269504088896338700e+00;
  const double P1 = 1.66666666666666019037e-01;
  const double P2 = -2.77777777770155933842e-03;
  const double P3 = 6.61375632143793436117e-05;
  const double P4 = -1.65339022054652515390e-06;
  const double P5 = 4.13813679705723846039e-08;
  if (x != x) return x;
  if (x > 7.09782712893383973096e+02) return sadmc_bits_f64(0x7ff0000000000000ull);
  if (x < -7.45133219101941108420e+02) return 0.0;
  const double ax = x < 0.0 ? -x : x;
  double hi = 0.0, lo = 0.0;
  int k = 0;
  if (ax > 0.34657359027997264) { /* 0.5 ln2 */
    if (ax < 1.0397207708399179) { /* 1.5 ln2 */
      if (x > 0.0) {
        hi = x - ln2_hi;
        lo = ln2_lo;
        k = 1;
      } else {
        hi = x + ln2_hi;
        lo = -ln2_lo;
        k = -1;
      }
    } else {
      k = (int)(inv_ln2 * x + (x < 0.0 ? -0.5 : 0.5));
      const double t = (double)k;
      hi = x - t * ln2_hi; /* t*ln2_hi is exact: ln2_hi has 21 trailing zero bits */
      lo = t * ln2_lo;
    }
    x = hi - lo;
  } else if (ax < 3.7252902984619141e-09) { /* 2^-28 */
    return 1.0 + x;
  }
  const double t = x * x;
  const double c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
  if (k == 0) return 1.0 - ((x * c) / (c - 2.0) - x);
  const double y = 1.0 - ((lo - (x * c) / (2.0 - c)) - hi);
  /* scale by 2^k; y is in [0.5, 2). */
  if (k >= -1021) {
    return sadmc_bits_f64(sadmc_f64_bits(y) + ((uint64_t)(int64_t)k << 52));
  }
  const double tiny = 9.33263618503218878990e-302; /* 2^-1000 */
  return sadmc_bits_f64(sadmc_f64_bits(y) + ((uint64_t)(int64_t)(k + 1000) << 52)) * tiny;
}

/* ln(x).  |error| < 1 ulp.  x < 0 -> NaN, x == 0 -> -inf. */
SADMC_HD double sadmc_log(double x) {
  const double ln2_hi = 6.93147180369123816490e-01;
  const double ln2_lo = 1.90821492927058770002e-10;
  const double Lg1 = 6.666666666666735130e-01;
  const double Lg2 = 3.999999999940941908e-01;
  const double Lg3 = 2.857142874366239149e-01;
  const double Lg4 = 2.222219843214978396e-01;
  const double Lg5 = 1.818357216161805012e-01;
  const double Lg6 = 1.531383769920937332e-01;
  const double Lg7 = 1.479819860511658591e-01;
  if (x != x) return x;
  if (x < 0.0) return sadmc_bits_f64(0x7ff8000000000000ull);
  if (x == 0.0) return sadmc_bits_f64(0xfff0000000000000ull);
  uint64_t bits = sadmc_f64_bits(x);
  if (bits >= 0x7ff0000000000000ull) return x; /* +inf */
  int k = 0;
  if (bits < 0x0010000000000000ull) { /* subnormal: scale up by 2^54 */
    x = x * 1.80143985094819840000e+16;
    bits = sadmc_f64_bits(x);
    k = -54;
  }
  int32_t hx = (int32_t)(bits >> 32);
  const uint32_t lx = (uint32_t)bits;
  k += (hx >> 20) - 1023;
  hx &= 0x000fffff;
  const int32_t i0 = (hx + 0x95f64) & 0x100000; /* mantissa >= sqrt(2) ? */
  /* normalise x to [sqrt(2)/2, sqrt(2)) */
  x = sadmc_bits_f64(((uint64_t)(uint32_t)(hx | (i0 ^ 0x3ff00000)) << 32) | lx);
  k += (i0 >> 20);
  const double f = x - 1.0;
  const double dk = (double)k;
  if ((0x000fffff & (2 + hx)) < 3) { /* |f| < 2^-20 */
    if (f == 0.0) {
      if (k == 0) return 0.0;
      return dk * ln2_hi + dk * ln2_lo;
    }
    const double R = f * f * (0.5 - 0.33333333333333333 * f);
    if (k == 0) return f - R;
    return dk * ln2_hi - ((R - dk * ln2_lo) - f);
  }
  const double s = f / (2.0 + f);
  const double z = s * s;
  const double w = z * z;
  const double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
  const double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
  const double R = t2 + t1;
  const int32_t i = (hx - 0x6147a) | (0x6b851 - hx);
  if (i > 0) {
    const double hfsq = 0.5 * f * f;
    if (k == 0) return f - (hfsq - s * (hfsq + R));
    return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
  }
  if (k == 0) return f - s * (f - R);
  return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
}

#endif /* SADMC_MATH_H */
