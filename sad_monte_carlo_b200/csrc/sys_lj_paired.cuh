// sys_lj_paired.cuh -- EXPERIMENT (SADMC_FLAG_HELPER_WARPS): the LJ thread-per-walker kernel of sys_lj_thread.cuh
// with a HELPER warp per bookkeeping warp for the pair loop.
//
// A walker's move is a serial chain -- draws, pair loop, bin record, bookkeeping -- and an SM holds only ~300 LJ31
// walkers (shared memory), so throughput = resident walkers / chain latency.  The pair loop is
// the one link that parallelises without duplicating the scalar tail: a CTA is 4 bookkeeping warps (threads 0-127,
// one walker each, exactly LjThreadSys<fast, NT, 1>) and 4 helper warps (threads 128-255); lane i of helper warp k
// works for the walker of lane i of bookkeeping warp k and reads the same shared-memory columns.  Per move the
// bookkeeping thread parks the moved atom, publishes (old, new position) in 6 doubles of shared memory, both sum
// half of the rows (rows [0, HALF) / [HALF, NT)), the helper posts its partial sum; two 64-thread named barriers
// (`bar.sync id, 64`) order the hand-over.  Registers: the kernel is compiled for 128 per thread
// (__launch_bounds__(256, 2)); the bookkeeping warpgroup raises itself to 200 with setmaxnreg, the helpers drop to 56
// (2 CTAs x (128 x 200 + 128 x 56) = 65 536).
//
// Same arithmetic as the fast one-lane kernel except that the pair sum is split in two halves (tolerance tier,
// <= 1e-12 relative per move); the generator stream and every decision rule are untouched.
#pragma once
#include "sys_lj_thread.cuh"

namespace sadmc {

__device__ __forceinline__ void named_barrier_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

template <int NT>
struct LjPairedSys : LjThreadSys<true, NT, 1, 128> {
  static_assert(NT > 1, "compile-time atom counts only");
  typedef LjThreadSys<true, NT, 1, 128> Base;
  static constexpr bool HELPERS = true;
  static constexpr int BLOCK = 256;     // threads per CTA at launch: 128 bookkeeping + 128 helpers
  static constexpr int MIN_BLOCKS = 2;
  static constexpr int WALKERS_PER_BLOCK = 128;
#ifndef SADMC_PAIR_HALF
#define SADMC_PAIR_HALF ((NT + 1) / 2)
#endif
#ifndef SADMC_PAIR_MAIN_REGS
#define SADMC_PAIR_MAIN_REGS 200
#define SADMC_PAIR_HELPER_REGS 56
#endif
#ifndef SADMC_PAIR_UNROLL
#define SADMC_PAIR_UNROLL 2
#endif
  static constexpr int HALF = SADMC_PAIR_HALF; // rows [0, HALF) stay with the bookkeeping thread
  static constexpr int MAIN_REGS = SADMC_PAIR_MAIN_REGS, HELPER_REGS = SADMC_PAIR_HELPER_REGS;
  static constexpr int PAIR_UNROLL = SADMC_PAIR_UNROLL; // two-atom bodies in flight
  static constexpr int EX_SLOTS = 7;        // ox, oy, oz, tx, ty, tz, partial

  double* ex; // this walker's exchange column

  static __host__ __device__ size_t smem_bytes(const DevParams& P, int) {
    return Base::smem_bytes(P, WALKERS_PER_BLOCK) + (size_t)EX_SLOTS * WALKERS_PER_BLOCK * sizeof(double);
  }
  __device__ LjPairedSys(const DevParams& P, uint32_t w, int lane_in_group, unsigned group_mask_, unsigned char* smem)
      : Base(P, w, lane_in_group, group_mask_, smem),
        ex(reinterpret_cast<double*>(smem + Base::smem_bytes(P, WALKERS_PER_BLOCK)) + (threadIdx.x & (WALKERS_PER_BLOCK - 1))) {}

  // rows [K0, K1) of column `col` against the old and the new position: LjThreadSys::plan_move_drawn's loop body
  template <int K0, int K1>
  static __device__ __forceinline__ double half_sum(const double* col, double ox, double oy, double oz, double tx, double ty, double tz) {
    constexpr int S = WALKERS_PER_BLOCK;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll(PAIR_UNROLL)
    for (int k = K0; k + 1 < K1; k += 2) {
      const double xa = col[(0 * NT + k) * S], ya = col[(1 * NT + k) * S], za = col[(2 * NT + k) * S];
      const double xb = col[(0 * NT + k + 1) * S], yb = col[(1 * NT + k + 1) * S], zb = col[(2 * NT + k + 1) * S];
      const double aax = xa - tx, aay = ya - ty, aaz = za - tz;
      const double abx = xa - ox, aby = ya - oy, abz = za - oz;
      const double bax = xb - tx, bay = yb - ty, baz = zb - tz;
      const double bbx = xb - ox, bby = yb - oy, bbz = zb - oz;
      const double rna = fma(aaz, aaz, fma(aay, aay, aax * aax));
      const double roa = fma(abz, abz, fma(aby, aby, abx * abx));
      const double rnb = fma(baz, baz, fma(bay, bay, bax * bax));
      const double rob = fma(bbz, bbz, fma(bby, bby, bbx * bbx));
      const double pa = rna * roa, pb = rnb * rob;
      const double inv = rcp_newton(pa * pb);
      const double ia = inv * pb, ib = inv * pa;
      const double sna = ia * roa, soa = ia * rna, snb = ib * rob, sob = ib * rnb;
      const double sna3 = sna * sna * sna, soa3 = soa * soa * soa, snb3 = snb * snb * snb, sob3 = sob * sob * sob;
      acc[k & 3] += fma(sna3, sna3, -sna3) - fma(soa3, soa3, -soa3);
      acc[(k + 1) & 3] += fma(snb3, snb3, -snb3) - fma(sob3, sob3, -sob3);
    }
    if ((K1 - K0) & 1) {
      constexpr int k = K1 - 1;
      const double x = col[(0 * NT + k) * S], y = col[(1 * NT + k) * S], z = col[(2 * NT + k) * S];
      const double ax = x - tx, ay = y - ty, az = z - tz;
      const double bx = x - ox, by = y - oy, bz = z - oz;
      const double rn = fma(az, az, fma(ay, ay, ax * ax));
      const double ro = fma(bz, bz, fma(by, by, bx * bx));
      const double inv = rcp_newton(rn * ro);
      const double sn = inv * ro, so = inv * rn;
      const double sn3 = sn * sn * sn, so3 = so * so * so;
      acc[k & 3] += fma(sn3, sn3, -sn3) - fma(so3, so3, -so3);
    }
    return (acc[0] + acc[1]) + (acc[2] + acc[3]);
  }

  // Called by EVERY bookkeeping thread once per move (barriers inside); `active` = the walker really proposes.
  __device__ __forceinline__ bool plan_move_paired(bool active, Rng& rng, double scale, const double* zx, const double* zf, double& e2) {
    int which = 0;
    double vx = 0.0, vy = 0.0, vz = 0.0;
    if (active) {
      which = (int)rng.below((uint32_t)NT, this->zone); // Uniform::new(0, N), lj.rs:368
      rng.normal3(zx, zf, vx, vy, vz);                  // rng.rs:111-117
    }
    const double ox = this->pos(0, which), oy = this->pos(1, which), oz = this->pos(2, which);
    this->tx = ox + vx * scale; // lj.rs:369 (an inactive walker "moves" by zero: every term is u - u = 0)
    this->ty = oy + vy * scale;
    this->tz = oz + vz * scale;
    const double tx = this->tx, ty = this->ty, tz = this->tz;
    const double new_r2 = tx * tx + ty * ty + tz * tz;
    const double prev_r2 = ox * ox + oy * oy + oz * oz;
    const bool none = new_r2 > this->R2 && new_r2 > prev_r2; // lj.rs:87-90
    this->own(0, which) = Base::FAR; // parked: its own term is exactly 0 - 0 in whichever half it sits
    ex[0 * WALKERS_PER_BLOCK] = ox;
    ex[1 * WALKERS_PER_BLOCK] = oy;
    ex[2 * WALKERS_PER_BLOCK] = oz;
    ex[3 * WALKERS_PER_BLOCK] = tx;
    ex[4 * WALKERS_PER_BLOCK] = ty;
    ex[5 * WALKERS_PER_BLOCK] = tz;
    const int bar = 1 + (int)((threadIdx.x >> 5) & 3);
    named_barrier_sync(bar, 64); // proposal published, atom parked
    const double mine = half_sum<0, HALF>(this->sp, ox, oy, oz, tx, ty, tz);
    named_barrier_sync(bar + 4, 64); // the helper's partial sum is posted, it no longer reads the columns
    const double theirs = ex[6 * WALKERS_PER_BLOCK];
    this->own(0, which) = ox;
    const double e = this->E + 4.0 * (mine + theirs);
    this->ch_which = which;
    this->ch_e = e;
    e2 = e;
    return active && !none;
  }

  // The helper threads' whole kernel: n_moves hand-overs.
  static __device__ void helper_loop(const DevParams& P, unsigned char* smem_sys, unsigned long long n_moves) {
    const int colidx = (int)(threadIdx.x & (WALKERS_PER_BLOCK - 1));
    const double* col = reinterpret_cast<const double*>(smem_sys) + colidx;
    double* exc = reinterpret_cast<double*>(smem_sys + Base::smem_bytes(P, WALKERS_PER_BLOCK)) + colidx;
    const int bar = 1 + (int)((threadIdx.x >> 5) & 3);
#pragma unroll 1
    for (unsigned long long m = 0; m < n_moves; m++) {
      named_barrier_sync(bar, 64);
      const double ox = exc[0 * WALKERS_PER_BLOCK], oy = exc[1 * WALKERS_PER_BLOCK], oz = exc[2 * WALKERS_PER_BLOCK];
      const double tx = exc[3 * WALKERS_PER_BLOCK], ty = exc[4 * WALKERS_PER_BLOCK], tz = exc[5 * WALKERS_PER_BLOCK];
      exc[6 * WALKERS_PER_BLOCK] = half_sum<HALF, NT>(col, ox, oy, oz, tx, ty, tz);
      named_barrier_sync(bar + 4, 64);
    }
  }
};

} // namespace sadmc
