// sys_lj.cuh -- Lennard-Jones cluster in a hard spherical container.
//
// Device form of `Lj` (src/system/lj.rs): move_atom 86-105, potential 78-81,
// plan_move 365-374, confirm 339-346, set_energy 110-123, compute_energy 236-244,
// randomize 262-279, verify_energy 249-261.
//
// Mapping: G lanes of a warp cooperate on one walker (G = 32 is "one warp per
// walker"; G = 16/8/4 pack 2/4/8 walkers into a warp so that the scalar
// bookkeeping is shared SIMD-wise between walkers).  Atom a lives in the
// registers of lane a % G, slot a / G (A = ceil(N / G) slots per lane).  A move
// broadcasts the chosen atom's old position by shuffle, every lane evaluates its
// <= A pairs against the old and the new position, and an xor-butterfly over the
// G lanes sums the terms.  Positions and the hard-wall test use exactly the
// reference's arithmetic (no FMA) so configurations stay bit-identical to the
// oracle's; the pair sum is FMA-contracted and tree-ordered, which is where the
// 1e-12 relative tolerance of the floating-point tier comes from.
#pragma once
#include "book.cuh"
#include "rng.cuh"

namespace sadmc {

template <int G_, int A_>
struct LjSys {
  static constexpr int G = G_;
  static constexpr bool FAST_BOOK = false;
  static constexpr int A = A_;
  static constexpr int BLOCK = 128;
  static constexpr int MIN_BLOCKS = 4;
  static constexpr bool COOP = false;
  static constexpr bool VERIFIES = true; // overrides System::verify_energy: run at the cadence of energy.rs:907-911
  __device__ __forceinline__ void set_cooperative(bool) {}
  __device__ __forceinline__ void finish_move() {}
  double px[A], py[A], pz[A];
  double E, err;
  int N, lane;
  unsigned gmask;
  double R, R2;
  unsigned long long zone;
  // pending change (Change::Move, lj.rs:60-64)
  int ch_which;
  double tx, ty, tz, ch_e;

  static __host__ __device__ size_t smem_bytes(const DevParams&, int) { return 0; }

  __device__ LjSys(const DevParams& P, uint32_t, int lane_in_group, unsigned mask, unsigned char*)
      : N((int)P.N), lane(lane_in_group), gmask(mask), R(P.lj_R), R2(P.lj_R2), zone(P.zone_b), ch_which(-1) {}

  __device__ void load(const DevParams& P, uint32_t w, const WalkerRec& r) {
    const double* g = P.sys + (size_t)w * P.sys_stride;
#pragma unroll
    for (int s = 0; s < A; s++) {
      const int a = s * G + lane;
      const bool ok = a < N;
      px[s] = ok ? g[3 * a] : 1e150; // far away: contributes exactly 0 - 0
      py[s] = ok ? g[3 * a + 1] : 1e150;
      pz[s] = ok ? g[3 * a + 2] : 1e150;
    }
    E = r.E;
    err = r.err;
  }
  __device__ void store(const DevParams& P, uint32_t w, WalkerRec& r, bool writer) {
    double* g = P.sys + (size_t)w * P.sys_stride;
#pragma unroll
    for (int s = 0; s < A; s++) {
      const int a = s * G + lane;
      if (a < N) {
        g[3 * a] = px[s];
        g[3 * a + 1] = py[s];
        g[3 * a + 2] = pz[s];
      }
    }
    if (writer) {
      g[3 * N] = E;
      g[3 * N + 1] = err;
      r.E = E;
      r.err = err;
    }
  }
  __device__ __forceinline__ double energy() const { return E; }

  // s^6 - s^3 with s = 1/r^2 (sigma = 1); the factor 4 epsilon is applied to the sum.
  static __device__ __forceinline__ double pair(double dx, double dy, double dz) {
    const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
    const double s = 1.0 / r2;
    const double s3 = s * s * s;
    return fma(s3, s3, -s3);
  }
  __device__ __forceinline__ double group_sum(double v) const {
#pragma unroll
    for (int off = G / 2; off >= 1; off >>= 1) v += __shfl_xor_sync(gmask, v, off, G);
    return v;
  }
  __device__ __forceinline__ void fetch(int a, double& x, double& y, double& z) const {
    const int slot = a / G, owner = a % G;
    double sx = px[0], sy = py[0], sz = pz[0];
#pragma unroll
    for (int s = 1; s < A; s++)
      if (slot == s) {
        sx = px[s];
        sy = py[s];
        sz = pz[s];
      }
    x = __shfl_sync(gmask, sx, owner, G);
    y = __shfl_sync(gmask, sy, owner, G);
    z = __shfl_sync(gmask, sz, owner, G);
  }

  __device__ __forceinline__ bool plan_move(Rng& rng, double scale, const double* zx, const double* zf, double& e2) {
    const int which = (int)rng.below((uint32_t)N, zone); // Uniform::new(0, N), lj.rs:368
    const double vx = rng.normal(zx, zf);                // crate::rng::vector, rng.rs:111-117
    const double vy = rng.normal(zx, zf);
    const double vz = rng.normal(zx, zf);
    double ox, oy, oz;
    fetch(which, ox, oy, oz);
    tx = ox + vx * scale; // lj.rs:369 (no FMA: -fmad=false)
    ty = oy + vy * scale;
    tz = oz + vz * scale;
    const double new_r2 = tx * tx + ty * ty + tz * tz;
    const double prev_r2 = ox * ox + oy * oy + oz * oz;
    const bool none = new_r2 > R2 && new_r2 > prev_r2; // lj.rs:87-90
    double acc = 0.0;
#pragma unroll
    for (int s = 0; s < A; s++) {
      const int a = s * G + lane;
      if (a != which) acc += pair(px[s] - tx, py[s] - ty, pz[s] - tz) - pair(px[s] - ox, py[s] - oy, pz[s] - oz);
    }
    acc = group_sum(acc);
    ch_which = which;
    ch_e = E + 4.0 * acc; // lj.rs:91-102
    e2 = ch_e;
    return !none;
  }

  __device__ double compute_energy() const { // lj.rs:236-244, lane-parallel
    double acc = 0.0;
    for (int b = 1; b < N; b++) {
      double bx, by, bz;
      fetch(b, bx, by, bz);
#pragma unroll
      for (int s = 0; s < A; s++) {
        const int a = s * G + lane;
        if (a < b) acc += pair(px[s] - bx, py[s] - by, pz[s] - bz);
      }
    }
    return 4.0 * group_sum(acc);
  }
  __device__ __forceinline__ double expected_accuracy(double newe) const { return fabs(newe) * 1e-14 * (double)N * (double)N; } // lj.rs:106-108

  __device__ __forceinline__ void confirm() { // lj.rs:339-346 + set_energy 110-123
    const int slot = ch_which / G, owner = ch_which % G;
    if (lane == owner) {
#pragma unroll
      for (int s = 0; s < A; s++)
        if (slot == s) {
          px[s] = tx;
          py[s] = ty;
          pz[s] = tz;
        }
    }
    const double new_e = ch_e;
    const double new_error = fabs(new_e) > fabs(E) ? fabs(new_e) * 1e-15 * (double)N : fabs(E) * 1e-15 * (double)N;
    err = new_error + err;
    if (err > expected_accuracy(new_e)) {
      err *= 0.0;
      E = compute_energy();
    } else {
      E = new_e;
    }
  }

  __device__ double randomize(Rng& rng) { // lj.rs:262-279
    for (int a = 0; a < N; a++) {
      double x, y, z;
      for (;;) {
        x = rng.uniform_f64(-1.0, 2.0);
        y = rng.uniform_f64(-1.0, 2.0);
        z = rng.uniform_f64(-1.0, 2.0);
        if (x * x + y * y + z * z < 1.0) break;
      }
      const int slot = a / G, owner = a % G;
      if (lane == owner) {
#pragma unroll
        for (int s = 0; s < A; s++)
          if (slot == s) {
            px[s] = x * R;
            py[s] = y * R;
            pz[s] = z * R;
          }
      }
    }
    E = compute_energy(); // `error` is left as it was, as in the reference
    return E;
  }
  __device__ bool verify_energy() const { // lj.rs:249-261
    const double egood = compute_energy();
    if (fabs(egood - E) > expected_accuracy(E)) return egood == E;
    return true;
  }
  __device__ __forceinline__ bool extra(unsigned long long, double&) const { return false; }
  // pending change across the trait shims (possible_change, lj.rs:38)
  __device__ void get_pending(double* p, bool writer, bool some) const {
    if (!writer || !some) return;
    p[0] = 1.0;
    p[1] = (double)ch_which;
    p[2] = tx;
    p[3] = ty;
    p[4] = tz;
    p[5] = ch_e;
  }
  __device__ bool set_pending(const double* p) {
    if (p[0] == 0.0) return false;
    ch_which = (int)p[1];
    tx = p[2];
    ty = p[3];
    tz = p[4];
    ch_e = p[5];
    return true;
  }
};

} // namespace sadmc
