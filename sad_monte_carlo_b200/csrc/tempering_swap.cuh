// tempering_swap.cuh -- the swap step of `MC::run_once` (src/mc/tempering.rs:285-320); included by engine.cu only
// (tempering.cuh holds the per-system move kernel and explains the layout).
#pragma once
#include "book.cuh"
#include "kernel_set.cuh"
#include "rng.cuh"

namespace sadmc {

// One block per simulation; warp v handles pairs v, v + n_warps, ...  mc_rng: [n_sim][2] generator of `MC` (tempering.rs:127).
__global__ void __launch_bounds__(128) temper_swap_kernel(const DevParams P, TemperRec* reps, unsigned long long* mc_rng, uint32_t n_T) {
  __shared__ int first_pair;
  const uint32_t sim = blockIdx.x;
  if (threadIdx.x == 0) {
    Rng g;
    g.s0 = mc_rng[2 * sim];
    g.s1 = mc_rng[2 * sim + 1];
    // rand 0.7 `Standard` for bool: (next_u32() as i32) < 0, next_u32 = upper half of next_u64: the top bit of the word
    const bool odd_ones = (g.next() >> 63) != 0;
    mc_rng[2 * sim] = g.s0;
    mc_rng[2 * sim + 1] = g.s1;
    first_pair = odd_ones ? 0 : 1; // chunks_exact_mut(2) of replicas[..] or of replicas[1..] (tempering.rs:285-291)
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const uint32_t start = (uint32_t)first_pair;
  if (n_T < start + 2) return;
  const uint32_t n_pairs = (n_T - start) / 2;
  for (uint32_t p = warp; p < n_pairs; p += n_warps) {
    const uint32_t s0 = sim * n_T + start + 2 * p, s1 = s0 + 1;
    WalkerRec& w0 = P.walkers[s0];
    WalkerRec& w1 = P.walkers[s1];
    int do_swap = 0;
    if (lane == 0) {
      TemperRec r0 = reps[s0], r1 = reps[s1];
      const double de_db = (w0.E - w1.E) * (1.0 / r0.T - 1.0 / r1.T); // tempering.rs:296
      bool acc = de_db >= 0.0;
      if (!acc) {
        Rng g;
        g.s0 = w1.s0;
        g.s1 = w1.s1;
        acc = exp_cmp(g.gen_f64(), de_db) < 0; // r1.rng.gen::<f64>() < de_db.exp()
        w1.s0 = g.s0;
        w1.s1 = g.s1;
      }
      double e0 = w0.E, e1 = w1.E;
      if (acc) {
        r0.accepted_swap += 1;
        r1.accepted_swap += 1;
        const double t = e0;
        e0 = e1;
        e1 = t;
      } else {
        r0.rejected_swap += 1;
        r1.rejected_swap += 1;
      }
      r0.total_energy += e0; // tempering.rs:307-319
      r1.total_energy += e1;
      r0.total_energy_squared += e0 * e0;
      r1.total_energy_squared += e1 * e1;
      if (e0 >= 0.0) r0.ignored += 1;
      if (e1 >= 0.0) r1.ignored += 1;
      reps[s0] = r0;
      reps[s1] = r1;
      do_swap = acc ? 1 : 0;
    }
    do_swap = __shfl_sync(0xffffffffu, do_swap, 0);
    if (do_swap) { // std::mem::swap(&mut r0.system, &mut r1.system)
      if (P.sys_stride) {
        double* a = P.sys + (size_t)s0 * P.sys_stride;
        double* b = P.sys + (size_t)s1 * P.sys_stride;
        for (uint32_t k = lane; k < P.sys_stride; k += 32) {
          const double t = a[k];
          a[k] = b[k];
          b[k] = t;
        }
      }
      if (P.ising_words) {
        uint32_t* a = P.sys_words + (size_t)s0 * P.ising_words;
        uint32_t* b = P.sys_words + (size_t)s1 * P.ising_words;
        for (uint32_t k = lane; k < P.ising_words; k += 32) {
          const uint32_t t = a[k];
          a[k] = b[k];
          b[k] = t;
        }
      }
      if (lane == 0) { // the part of the system that lives in the walker record: cached energy, its error, two-wells d^2
        double t = w0.E;
        w0.E = w1.E;
        w1.E = t;
        t = w0.err;
        w0.err = w1.err;
        w1.err = t;
        t = w0.d_squared;
        w0.d_squared = w1.d_squared;
        w1.d_squared = t;
      }
    }
    __syncwarp();
  }
}

} // namespace sadmc
