// replicas.cuh -- the `replicas` binary on the move kernels' systems: `Replica::run_once` / `occasional_update` and the
// set-up sweep of `MC::from_params` in src/mc/energy_replicas.rs.  (The round logic -- swaps, median estimator, splitting
// off a new replica -- is replicas_round.cuh.)
//
//   Replica::run_once            energy_replicas.rs:206-248   a bounded replica accepts every proposal below its max_energy (no
//                                                             random number); the unbounded top replica re-randomizes its system
//                                                             once per round; energy moments above / below the cutoff
//   Replica::occasional_update   energy_replicas.rs:249-290   step size from the accepted / rejected ratio
//   MC::from_params              energy_replicas.rs:346-399   MAX_INIT randomized energies, the two first replicas
//   MC::run_once                 energy_replicas.rs:504-525   min_moves_to_randomize moves per bounded replica and round
//
// Batch dimension: n_sim independent simulations (simulation k = `replicas --seed seed + k`) x R_MAX replica slots each;
// slot s = k * R_MAX + r is one walker of the engine, replica r of simulation k, in use while r < n_rep[k].
#pragma once
#include "move_kernel.cuh"

namespace sadmc {

__device__ __forceinline__ void xoroshiro_jump_device(Rng& g) { // rand_xoshiro 0.4 Xoroshiro128Plus::jump
  const unsigned long long J[2] = {0xdf900294d8f554a5ull, 0x170865df4b3201fcull};
  unsigned long long s0 = 0, s1 = 0;
  for (int i = 0; i < 2; i++)
    for (int b = 0; b < 64; b++) {
      if (J[i] & (1ull << b)) {
        s0 ^= g.s0;
        s1 ^= g.s1;
      }
      g.next();
    }
  g.s0 = s0;
  g.s1 = s1;
}

// MC::from_params up to the sort (energy_replicas.rs:346-368): MAX_INIT x system.randomize, high_system = clone + randomize,
// then system.randomize until its energy is not above energies[MAX_INIT / 2] (of the UNSORTED list, as the reference reads it).
// One walker-sized group of threads per simulation; `system` ends in slot 1, `high_system` in slot 0, both slots and the
// simulation share the generator state reached here (the host sorts the energies, fills the replica records and jumps the
// simulation's generator).
template <class Sys>
__global__ void __launch_bounds__(Sys::BLOCK, Sys::MIN_BLOCKS) replica_init_kernel(const DevParams P, unsigned long long seed0, uint32_t n_sim, uint32_t r_max,
                                                                                  double* energies, uint32_t max_init, unsigned long long* sim_rng) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int G = Sys::G;
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t sim = tid / G;
  const int lane = (int)(tid % G);
  if (sim >= n_sim) return;
  const unsigned gmask = group_mask<G>();
  const uint32_t s_sys = sim * r_max + 1, s_high = sim * r_max;
  Sys sys(P, s_sys, lane, gmask, smem + zig_smem_bytes<Sys>());
  sys.load(P, s_sys, P.walkers[s_sys]);
  sys.set_cooperative(false);
  Rng rng;
  seed_from_u64(seed0 + sim, (uint64_t*)&rng.s0, (uint64_t*)&rng.s1);
  double* en = energies + (size_t)sim * max_init;
  for (uint32_t k = 0; k < max_init; k++) {
    const double e = sys.randomize(rng);
    if (lane == 0) en[k] = e;
  }
  if (G > 1) __syncwarp(gmask);
  sys.store(P, s_high, P.walkers[s_high], lane == 0); // high_system = system.clone()
  sys.store(P, s_sys, P.walkers[s_sys], lane == 0);
  if (G > 1) __syncwarp(gmask);
  sys.load(P, s_high, P.walkers[s_high]);
  sys.randomize(rng);
  sys.store(P, s_high, P.walkers[s_high], lane == 0);
  if (G > 1) __syncwarp(gmask);
  sys.load(P, s_sys, P.walkers[s_sys]);
  const double threshold = en[max_init / 2];
  while (sys.energy() > threshold) sys.randomize(rng);
  sys.store(P, s_sys, P.walkers[s_sys], lane == 0);
  if (lane == 0) {
    P.walkers[s_high].s0 = rng.s0; // rng.clone() for both replicas (energy_replicas.rs:374, 380)
    P.walkers[s_high].s1 = rng.s1;
    P.walkers[s_sys].s0 = rng.s0;
    P.walkers[s_sys].s1 = rng.s1;
    xoroshiro_jump_device(rng); // 383
    sim_rng[2 * sim] = rng.s0;
    sim_rng[2 * sim + 1] = rng.s1;
  }
}

// One round's moves for every replica slot (the rayon par_iter_mut of energy_replicas.rs:513-525), preceded by the
// occasional_update that the reference runs at the end of the previous round (591-593; nothing in between touches what it
// reads or writes).  steps = min_moves_to_randomize.
template <class Sys>
__global__ void __launch_bounds__(Sys::BLOCK, Sys::MIN_BLOCKS) replica_move_kernel(const DevParams P, ReplicaRec* reps, const ReplicaSim* sims, uint32_t r_max,
                                                                                  unsigned long long steps, double* slot_energy) {
  extern __shared__ __align__(16) unsigned char smem[];
  const double* zx = stage_zig<Sys>(P, smem);
  const double* zf = zx + SADMC_ZIG_TABLE_LEN;
  constexpr int G = Sys::G;
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t slot = tid / G;
  const int lane = (int)(tid % G);
  if (slot >= P.n_walkers) return;
  const uint32_t sim = slot / r_max, r = slot % r_max;
  const ReplicaSim S = sims[sim];
  if ((int)r >= S.n_rep) return;
  const unsigned gmask = group_mask<G>();
  WalkerRec& wr = P.walkers[slot];
  Sys sys(P, slot, lane, gmask, smem + zig_smem_bytes<Sys>());
  sys.load(P, slot, wr);
  sys.set_cooperative(false);
  Rng rng;
  rng.s0 = wr.s0;
  rng.s1 = wr.s1;
  ReplicaRec q = reps[slot];
  const bool bounded = isfinite(q.max_energy);
  // occasional_update of the previous round (energy_replicas.rs:249-290)
  if (q.rejected > 128 && q.accepted > 128 && bounded) {
    const double ratio = (double)q.accepted / (double)q.rejected;
    const double max_ratio = (double)steps;
    if (ratio < 0.5 || ratio > 2.0 * max_ratio) {
      double adjustment = ratio < 0.5 ? ratio / sqrt(max_ratio) : ratio * sqrt(max_ratio);
      if (adjustment > 2.0)
        adjustment = 2.0;
      else if (adjustment < 0.5)
        adjustment = 0.5;
      q.tscale *= adjustment;
      q.accepted = 0;
      q.rejected = 0;
    }
  }
  const double very_lowest = reps[(size_t)sim * r_max + (S.n_rep - 1)].max_energy;
  const unsigned long long n = bounded ? steps : 1ull; // the unbounded replica randomizes once per round (519-524)
#pragma unroll 1
  for (unsigned long long i = 0; i < n; i++) {
    if (bounded) {
      double e;
      if (sys.plan_move(rng, q.tscale, zx, zf, e)) {
        if (e < q.max_energy) {
          sys.confirm();
          q.accepted += 1;
        } else {
          q.rejected += 1;
        }
      } else {
        q.rejected += 1;
      }
    } else {
      sys.randomize(rng);
      q.lowest_max = q.max_energy;
    }
    const double e = sys.energy();
    if (q.collecting) {
      if (e > q.cutoff) {
        q.above_count += 1;
        q.above_total += e;
        q.above_sq += e * e;
        double xv;
        if (sys.extra(S.moves + i, xv)) {
          q.xtot += xv;
          q.xcnt += 1;
        }
      } else {
        q.below_count += 1;
        q.below_total += e;
        q.below_sq += e * e;
      }
      if (q.lowest_max == very_lowest) q.upwelling += 1;
    }
  }
  sys.store(P, slot, wr, lane == 0);
  if (lane == 0) {
    wr.s0 = rng.s0;
    wr.s1 = rng.s1;
    reps[slot] = q;
    slot_energy[slot] = sys.energy();
  }
}

} // namespace sadmc
