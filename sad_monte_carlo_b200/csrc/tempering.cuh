// tempering.cuh -- replica exchange on the move kernels' systems: `Replica::run_once` and the swap step of
// `MC::run_once` in src/mc/tempering.rs (the `tempering` binary, two-wells/run-two-wells.py:45-61).
//
//   Replica::run_once  tempering.rs:96-113   canonical move at the replica's own temperature, fixed step 1.0 (88),
//                                            energy moments collected only when plan_move returned Some
//   MC::run_once       tempering.rs:272-342  `steps` moves per replica (274), then ONE swap attempt per neighbouring
//                                            pair: pairs (0,1),(2,3).. or (1,2),(3,4).. chosen by gen::<bool>() of the
//                                            simulation's own generator (285-291), accepted when
//                                            dE dbeta >= 0 or r1.rng.gen::<f64>() < exp(dE dbeta) (295-303)
//
// Batch dimension: `n_sim` independent tempering simulations (simulation k = the reference process run with
// `--seed seed + k`) x `n_T` temperatures; replica slot s = k * n_T + r is one walker of the engine.  A launch of
// temper_move_kernel is the rayon `par_iter_mut` of tempering.rs:279-284 for every simulation at once; the swap
// exchanges the SYSTEMS of two slots (configuration image, cached energy and error), while temperature, generator
// and counters stay with the slot, as `std::mem::swap(&mut r0.system, &mut r1.system)` does (299).
#pragma once
#include "move_kernel.cuh"

namespace sadmc {

template <class Sys>
__global__ void __launch_bounds__(Sys::BLOCK, Sys::MIN_BLOCKS) temper_move_kernel(const DevParams P, TemperRec* reps, unsigned long long steps) {
  extern __shared__ __align__(16) unsigned char smem[];
  const double* zx = stage_zig<Sys>(P, smem);
  const double* zf = zx + SADMC_ZIG_TABLE_LEN;
  constexpr int G = Sys::G;
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t w_raw = tid / G;
  const int lane = (int)(tid % G);
  const bool ghost = w_raw >= P.n_walkers;
  if (ghost && !Sys::COOP) return;
  const uint32_t w = ghost ? P.n_walkers - 1 : w_raw;
  const unsigned gmask = group_mask<G>();
  WalkerRec& wr = P.walkers[w];
  Sys sys(P, w, lane, gmask, smem + zig_smem_bytes<Sys>());
  sys.load(P, w, wr);
  sys.set_cooperative(true);
  Rng rng;
  rng.s0 = wr.s0;
  rng.s1 = wr.s1;
  TemperRec r = reps[w];
#pragma unroll 1
  for (unsigned long long m = 0; m < steps; m++) {
    double e = 0.0;
    bool some = false;
    if (!ghost) {
      some = sys.plan_move(rng, r.tscale, zx, zf, e); // translation_scale: Length::new(1.0) unless set (tempering.rs:88)
      if (some) {
        const double beta_delta_e = (e - sys.energy()) / r.T;
        // beta_delta_e < 0.0 || rng.gen::<f64>() < (-beta_delta_e).exp()  (tempering.rs:99): the uniform is drawn for
        // every proposal that does not lower the energy, also for beta_delta_e == 0
        if (beta_delta_e < 0.0 || exp_cmp(rng.gen_f64(), -beta_delta_e) < 0) {
          sys.confirm();
          r.accepted += 1;
        } else {
          r.rejected += 1;
        }
      }
    }
    if (Sys::COOP) sys.finish_move(); // converged point: a cooperative re-summation of the energy belongs to confirm()
    if (some) { // tempering.rs:105-111
      const double en = sys.energy();
      r.total_energy += en;
      r.total_energy_squared += en * en;
      if (en >= 0.0) r.ignored += 1;
    }
  }
  if (ghost) return;
  sys.store(P, w, wr, lane == 0);
  if (lane == 0) {
    wr.s0 = rng.s0;
    wr.s1 = rng.s1;
    reps[w] = r;
  }
}

} // namespace sadmc
