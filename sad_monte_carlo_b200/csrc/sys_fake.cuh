// sys_fake.cuh -- the analytic test systems, one thread per walker, state in registers/local memory.
//
// Device forms of
//   `Fake`     src/system/fake.rs      plan_move 128-144, Function::energy 47-60, energy 96-99,
//                                      confirm 122-124, randomize 103-112
//   `TwoWells` src/system/two_wells.rs plan_move 451-464, find_energy 266-315, find_which 317-371,
//                                      confirm 437-447, data_to_collect 408-418
//   `ErfInv`   src/system/erfinv.rs    plan_move 101-110, find_energy 60-71, randomize 78-83
// These exist to check the flat-histogram bookkeeping against exact densities of
// states; the proposal is a handful of flops, so the kernels are bound by the
// bookkeeping (HBM sectors), like Ising.  Every floating-point operation is in the
// reference's order with no FMA, so Fake and TwoWells are bit-identical to the CPU
// oracle; ErfInv goes through erf()/exp()/log() of the platform and is a
// tolerance-tier system (as it is between two builds of the reference itself).
#pragma once
#include "book.cuh"
#include "sys_limits.hpp"
#include "rng.cuh"

// resident CTAs per SM the one-thread-per-walker kernels of the small systems are compiled for (register cap = 65536 / (128 * this))
#ifndef SADMC_SMALL_MIN_BLOCKS
#define SADMC_SMALL_MIN_BLOCKS 4
#endif

namespace sadmc {


struct FakeSys {
  static constexpr int G = 1;
  static constexpr bool FAST_BOOK = false;
  static constexpr int BLOCK = 128;
  static constexpr int MIN_BLOCKS = SADMC_SMALL_MIN_BLOCKS;
  static constexpr bool COOP = false;
  __device__ __forceinline__ void set_cooperative(bool) {}
  __device__ __forceinline__ void finish_move() {}
  double pos[FAKE_MAX_DIM], cand[FAKE_MAX_DIM];
  int dim, fn;
  double a, b, e1, e2, sigma;
  double E, ch_e;
  unsigned long long zone;

  static __host__ __device__ size_t smem_bytes(const DevParams&, int) { return 0; }
  __device__ FakeSys(const DevParams& P, uint32_t, int, unsigned, unsigned char*)
      : dim(P.fake_dim), fn(P.fake_fn), a(P.fake_a), b(P.fake_b), e1(P.fake_e1), e2(P.fake_e2), sigma(P.fake_sigma), zone(P.zone_a) {}

  __device__ __forceinline__ double f(double r) const { // fake.rs:47-60
    switch (fn) {
      case SADMC_FAKE_LINEAR: return r;
      case SADMC_FAKE_QUADRATIC: return r * r;
      case SADMC_FAKE_GAUSSIAN: return -sadmc_exp(-r * r / (2.0 * sigma * sigma));
      default:
        if (r < a) return (r * r) / (a * a) * e1 - e1;
        return ((r - b) / (b - a)) * ((r - b) / (b - a)) * e2 - e2;
    }
  }
  __device__ __forceinline__ double radius(const double* p) const { // iter().map(x*x).sum::<f64>().sqrt()
    double s = 0.0;
    for (int k = 0; k < dim; k++) s += p[k] * p[k];
    return sqrt(s);
  }
  __device__ void load(const DevParams& P, uint32_t w, const WalkerRec&) {
    const double* g = P.sys + (size_t)w * P.sys_stride;
    for (int k = 0; k < dim; k++) pos[k] = cand[k] = g[k];
    E = f(radius(pos)); // energy() is computed from the position (fake.rs:96-99)
  }
  __device__ void store(const DevParams& P, uint32_t w, WalkerRec& r, bool) {
    double* g = P.sys + (size_t)w * P.sys_stride;
    for (int k = 0; k < dim; k++) g[k] = pos[k];
    r.E = E;
    r.err = 0.0;
  }
  __device__ __forceinline__ double energy() const { return E; }
  __device__ __forceinline__ bool plan_move(Rng& rng, double d, const double* zx, const double* zf, double& e_out) {
    const int i = (int)rng.below((uint32_t)dim, zone); // gen_range(0, dim), fake.rs:129
    const double v = rng.normal(zx, zf);
    for (int k = 0; k < dim; k++) cand[k] = pos[k];
    for (int k = 0; k < dim; k++)
      if (k == i) cand[k] += v * d;
    const double r = radius(cand);
    if (r > 1.0) return false;
    ch_e = f(r);
    e_out = ch_e;
    return true;
  }
  __device__ __forceinline__ void confirm() { // fake.rs:122-124
    for (int k = 0; k < dim; k++) pos[k] = cand[k];
    E = f(radius(pos));
  }
  __device__ double compute_energy() const { return f(radius(pos)); }
  __device__ double randomize(Rng& rng) { // fake.rs:103-112
    double r = 5.0;
    while (r >= 1.0) {
      for (int k = 0; k < dim; k++) pos[k] = rng.gen_range_f64(0.0, 1.0);
      r = radius(pos);
    }
    for (int k = 0; k < dim; k++) cand[k] = pos[k];
    E = f(radius(pos));
    return E;
  }
  __device__ bool verify_energy() const { return true; }
  __device__ __forceinline__ bool extra(unsigned long long, double&) const { return false; }
  __device__ void get_pending(double* p, bool writer, bool) const {
    if (!writer) return;
    p[0] = 1.0; // possible_change is always overwritten by plan_move, even when it returns None
    for (int k = 0; k < dim; k++) p[1 + k] = cand[k];
  }
  __device__ bool set_pending(const double* p) {
    if (p[0] == 0.0) return false;
    for (int k = 0; k < dim; k++) cand[k] = p[1 + k];
    return true;
  }
};


struct TwoWellsSys {
  static constexpr int G = 1;
  static constexpr bool FAST_BOOK = false;
  static constexpr int BLOCK = 128;
  static constexpr int MIN_BLOCKS = SADMC_SMALL_MIN_BLOCKS;
  static constexpr bool COOP = false;
  static constexpr bool HAS_EXTRA = true; // `which` well, two_wells.rs:408-418
  __device__ __forceinline__ void set_cooperative(bool) {}
  __device__ __forceinline__ void finish_move() {}
  double pos[TW_MAX_DIM];
  double d_squared;
  int N;
  double h2h1, r2, rw;
  int ch_index;
  double cx, cy, cz;
  unsigned long long zone;

  static __host__ __device__ size_t smem_bytes(const DevParams&, int) { return 0; }
  __device__ TwoWellsSys(const DevParams& P, uint32_t, int, unsigned, unsigned char*)
      : N((int)P.N), h2h1(P.tw_h2h1), r2(P.tw_r2), rw(P.tw_rw), ch_index(0), cx(0), cy(0), cz(0), zone(P.zone_a) {}

  struct Regions {
    double e_1, e_2, e_w, e_i, d_1_squared, d_2_squared;
  };
  __device__ __forceinline__ Regions regions(double x1, double d_orthog_squared) const { // two_wells.rs:266-293
    const double r1 = 1.0;
    const double x2 = x1 - r1 - r2;
    const double xw = x1 - rw;
    const double xi = x2 + rw;
    Regions g;
    g.d_1_squared = d_orthog_squared + x1 * x1;
    g.d_2_squared = d_orthog_squared + x2 * x2;
    const double d_w_squared = d_orthog_squared + xw * xw;
    const double d_i_squared = d_orthog_squared + xi * xi;
    g.e_1 = 1.0 * (g.d_1_squared / (r1 * r1) - 1.0);
    g.e_2 = h2h1 * (g.d_2_squared / (r2 * r2) - 1.0);
    g.e_w = h2h1 * (d_w_squared / (r2 * r2) - 1.0);
    g.e_i = 1.0 * (d_i_squared / (r1 * r1) - 1.0);
    return g;
  }
  __device__ __forceinline__ bool find_energy(double x1, double d_orthog_squared, double& e) const { // two_wells.rs:266-315
    const Regions g = regions(x1, d_orthog_squared);
    if (g.d_1_squared <= 1.0) {
      e = g.e_1 < g.e_w ? g.e_1 : g.e_w;
      return true;
    } else if (g.d_2_squared <= r2 * r2) {
      e = (g.e_i > g.e_2 && g.e_i < 0.0) ? g.e_i : g.e_2;
      return true;
    } else if (d_orthog_squared <= r2 * r2 && x1 > 0.0 && x1 <= 1.0 + r2) {
      e = 0.0;
      return true;
    }
    return false;
  }
  __device__ __forceinline__ double find_which(double x1, double d_orthog_squared) const { // two_wells.rs:317-371
    const Regions g = regions(x1, d_orthog_squared);
    if (g.d_1_squared <= 1.0) return g.e_1 < g.e_w ? 0.0 : 1.0;
    if (g.d_2_squared <= r2 * r2) return (g.e_i > g.e_2 && g.e_i < 0.0) ? 0.0 : 1.0;
    return 0.0;
  }
  __device__ void load(const DevParams& P, uint32_t w, const WalkerRec&) {
    const double* g = P.sys + (size_t)w * P.sys_stride;
    for (int k = 0; k < N; k++) pos[k] = g[k];
    d_squared = g[N];
  }
  __device__ void store(const DevParams& P, uint32_t w, WalkerRec& r, bool) {
    double* g = P.sys + (size_t)w * P.sys_stride;
    for (int k = 0; k < N; k++) g[k] = pos[k];
    g[N] = d_squared;
    r.E = energy();
    r.err = 0.0;
    r.d_squared = d_squared;
  }
  __device__ __forceinline__ double energy() const { // two_wells.rs:375-384 (recomputed on every call)
    double e = 0.0;
    find_energy(pos[0], d_squared - pos[0] * pos[0], e);
    return e;
  }
  __device__ __forceinline__ bool plan_move(Rng& rng, double d, const double* zx, const double* zf, double& e_out) { // 451-464
    const int index = 3 * (int)rng.below((uint32_t)(N / 3), zone);
    const double vx = rng.normal(zx, zf), vy = rng.normal(zx, zf), vz = rng.normal(zx, zf);
    double ox = 0, oy = 0, oz = 0;
    for (int k = 0; k < N; k += 3)
      if (k == index) {
        ox = pos[k];
        oy = pos[k + 1];
        oz = pos[k + 2];
      }
    cx = vx * d + ox; // vector(rng) * d + old_r
    cy = vy * d + oy;
    cz = vz * d + oz;
    const double dsq = d_squared - (ox * ox + oy * oy + oz * oz) + (cx * cx + cy * cy + cz * cz);
    const double x1 = index == 0 ? cx : pos[0];
    ch_index = index;
    return find_energy(x1, dsq - x1 * x1, e_out);
  }
  __device__ __forceinline__ void confirm() { // two_wells.rs:437-447
    double s = 0.0;
    for (int k = 0; k < N; k += 3)
      if (k == ch_index) {
        s = 0.0 + pos[k] * pos[k];
        s += pos[k + 1] * pos[k + 1];
        s += pos[k + 2] * pos[k + 2];
        pos[k] = cx;
        pos[k + 1] = cy;
        pos[k + 2] = cz;
      }
    d_squared -= s;
    d_squared += cx * cx + cy * cy + cz * cz;
  }
  __device__ double compute_energy() const { return energy(); }
  __device__ double randomize(Rng&) { return energy(); } // the InvCdf sampler is host-side set-up (two_wells.rs:25-180): unsupported
  __device__ bool verify_energy() const { return true; }
  __device__ __forceinline__ bool extra(unsigned long long, double& v) const { // data_to_collect "which", every move
    v = find_which(pos[0], d_squared - pos[0] * pos[0]);
    return true;
  }
  __device__ void get_pending(double* p, bool writer, bool) const {
    if (!writer) return;
    p[0] = 1.0; // `change` is overwritten by every plan_move
    p[1] = (double)ch_index;
    p[2] = cx;
    p[3] = cy;
    p[4] = cz;
  }
  __device__ bool set_pending(const double* p) {
    if (p[0] == 0.0) return false;
    ch_index = (int)p[1];
    cx = p[2];
    cy = p[3];
    cz = p[4];
    return true;
  }
};

// erf_inv: Winitzki's starting guess refined by four Halley steps on erf(); a few ulp.
// (statrs 0.7's erf_inv is an un-vendored dependency; the oracle restates it the same way.)
__device__ __forceinline__ double erf_inv_dev(double x) {
  if (x <= -1.0) return -__longlong_as_double(0x7ff0000000000000ll);
  if (x >= 1.0) return __longlong_as_double(0x7ff0000000000000ll);
  if (x == 0.0) return 0.0;
  const double a = 0.147;
  const double ln1mx2 = log(1.0 - x * x);
  const double t = 2.0 / (3.14159265358979323846 * a) + 0.5 * ln1mx2;
  double y = sqrt(sqrt(t * t - ln1mx2 / a) - t);
  if (x < 0) y = -y;
  for (int it = 0; it < 4; it++) {
    const double err = erf(y) - x;
    const double d = 2.0 / sqrt(3.14159265358979323846) * exp(-y * y);
    y -= err / (d + y * err);
  }
  return y;
}


struct ErfInvSys {
  static constexpr int G = 1;
  static constexpr bool FAST_BOOK = false;
  static constexpr int BLOCK = 128;
  static constexpr int MIN_BLOCKS = SADMC_SMALL_MIN_BLOCKS;
  static constexpr bool COOP = false;
  __device__ __forceinline__ void set_cooperative(bool) {}
  __device__ __forceinline__ void finish_move() {}
  double pos[ERFINV_MAX_DIM], terms[ERFINV_MAX_DIM];
  int N, ch_i;
  double mean, E, ch_e, ch_x, ch_term;
  unsigned long long zone;

  static __host__ __device__ size_t smem_bytes(const DevParams&, int) { return 0; }
  __device__ ErfInvSys(const DevParams& P, uint32_t, int, unsigned, unsigned char*) : N((int)P.N), ch_i(0), mean(P.erfinv_mean), zone(P.zone_a) {}
  // erfinv.rs:60-71: sum over coordinates of mean + erf_inv(x), in coordinate order.  The terms are cached
  // per coordinate (erf_inv of an unchanged x is the same number), the sum is always rebuilt in order.
  __device__ __forceinline__ double total() const {
    double s = 0.0;
    for (int k = 0; k < N; k++) s += terms[k];
    return s;
  }
  __device__ void load(const DevParams& P, uint32_t w, const WalkerRec&) {
    const double* g = P.sys + (size_t)w * P.sys_stride;
    for (int k = 0; k < N; k++) {
      pos[k] = g[k];
      terms[k] = mean + erf_inv_dev(pos[k]);
    }
    E = total();
  }
  __device__ void store(const DevParams& P, uint32_t w, WalkerRec& r, bool) {
    double* g = P.sys + (size_t)w * P.sys_stride;
    for (int k = 0; k < N; k++) g[k] = pos[k];
    r.E = E;
    r.err = 0.0;
  }
  __device__ __forceinline__ double energy() const { return E; }
  __device__ __forceinline__ bool plan_move(Rng& rng, double d, const double* zx, const double* zf, double& e_out) { // 101-110
    const int i = (int)rng.below((uint32_t)N, zone);
    const double v = rng.normal(zx, zf);
    double xi = 0.0;
    for (int k = 0; k < N; k++)
      if (k == i) xi = pos[k];
    xi += v * d;
    ch_i = i;
    ch_x = xi;
    if (xi >= 1.0 || xi <= -1.0) return false;
    ch_term = mean + erf_inv_dev(xi);
    double s = 0.0;
    for (int k = 0; k < N; k++) s += k == i ? ch_term : terms[k];
    ch_e = s;
    e_out = s;
    return true;
  }
  __device__ __forceinline__ void confirm() {
    for (int k = 0; k < N; k++)
      if (k == ch_i) {
        pos[k] = ch_x;
        terms[k] = ch_term;
      }
    E = ch_e;
  }
  __device__ double compute_energy() const {
    double s = 0.0;
    for (int k = 0; k < N; k++) s += mean + erf_inv_dev(pos[k]);
    return s;
  }
  __device__ double randomize(Rng& rng) { // erfinv.rs:78-83
    for (int k = 0; k < N; k++) {
      pos[k] = rng.gen_range_f64(-1.0, 1.0);
      terms[k] = mean + erf_inv_dev(pos[k]);
    }
    E = total();
    return E;
  }
  __device__ bool verify_energy() const { return true; }
  __device__ __forceinline__ bool extra(unsigned long long, double&) const { return false; }
  __device__ void get_pending(double* p, bool writer, bool some) const {
    if (!writer) return;
    p[0] = some ? 1.0 : 2.0; // 2: a `None` proposal still replaced possible_change (erfinv.rs:103-107)
    p[1] = (double)ch_i;
    p[2] = ch_x;
    p[3] = ch_term;
    p[4] = ch_e;
  }
  __device__ bool set_pending(const double* p) {
    if (p[0] == 0.0) return false;
    ch_i = (int)p[1];
    ch_x = p[2];
    ch_term = p[3];
    ch_e = p[4];
    return true;
  }
};

} // namespace sadmc
