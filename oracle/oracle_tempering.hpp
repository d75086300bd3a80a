// oracle_tempering.hpp -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// CPU restatement of the `tempering` binary: `Replica` and `MC` of src/mc/tempering.rs (replica exchange over a
// list of temperatures; two-wells/run-two-wells.py:45-61 is its user).  Function by function, in the reference's order.
// From un-vendored crates (restated from their published algorithms, not pinned by a reference run in this image):
// rand_xoshiro 0.4 `Xoroshiro128Plus::jump` (the 2^64 jump polynomial of xoroshiro128+ 2018) and rand 0.7
// `Standard` for bool (`(next_u32() as i32) < 0` with next_u32 = upper half of next_u64).
#pragma once
#include <cmath>
#include <cstdint>
#include <memory>
#include <utility>
#include <vector>

#include "oracle_mc.hpp"

namespace oracle {
namespace tempering {

inline void jump(Rng& g) {
  static const uint64_t JUMP[2] = {0xdf900294d8f554a5ull, 0x170865df4b3201fcull};
  uint64_t s0 = 0, s1 = 0;
  for (int i = 0; i < 2; i++)
    for (int b = 0; b < 64; b++) {
      if (JUMP[i] & (1ull << b)) {
        s0 ^= g.s0;
        s1 ^= g.s1;
      }
      g.next_u64();
    }
  g.s0 = s0;
  g.s1 = s1;
}
inline bool gen_bool(Rng& g) { return (int32_t)(uint32_t)(g.next_u64() >> 32) < 0; }

struct Replica { // tempering.rs:46-73
  double T = 0;
  uint64_t rejected_count = 0, accepted_count = 0, rejected_swap_count = 0, accepted_swap_count = 0, ignored_count = 0;
  std::unique_ptr<System> system;
  Rng rng;
  double total_energy = 0, total_energy_squared = 0;
  double translation_scale = 1.0; // Length::new(1.0), tempering.rs:88

  double energy() const { return system->energy(); }
  void run_once() { // tempering.rs:96-113
    double e;
    if (system->plan_move(rng, translation_scale, &e)) {
      const double beta_delta_e = (e - energy()) / T;
      if (beta_delta_e < 0.0 || rng.gen_f64() < o_exp(-beta_delta_e)) {
        system->confirm();
        accepted_count += 1;
      } else {
        rejected_count += 1;
      }
      const double en = energy();
      total_energy += en;
      total_energy_squared += en * en;
      if (en >= 0.0) ignored_count += 1;
    }
  }
};

struct MC { // tempering.rs:123-145
  Rng rng;
  uint64_t moves = 0;
  std::vector<Replica> replicas;
  uint64_t canonical_steps = 1;
  uint64_t min_moves_to_randomize = 1;

  // from_params, tempering.rs:152-175; `systems` = system.clone() for every temperature
  MC(uint64_t seed, const std::vector<double>& T, uint64_t can_steps, uint64_t min_moves, std::vector<std::unique_ptr<System>> systems)
      : canonical_steps(can_steps), min_moves_to_randomize(min_moves) {
    rng = Rng::seed_from_u64(seed);
    for (size_t i = 0; i < T.size(); i++) {
      Replica r;
      r.T = T[i];
      r.system = std::move(systems[i]);
      r.rng = rng; // rng.clone(): every replica starts with the SAME generator state
      replicas.push_back(std::move(r));
    }
    jump(rng);
  }

  void run_once() { // tempering.rs:272-342 (movie / report / save belong to the host)
    const uint64_t steps = min_moves_to_randomize * canonical_steps;
    uint64_t these_moves = 0;
    for (auto& r : replicas) { // par_iter_mut: replicas are independent here
      these_moves += steps;
      for (uint64_t k = 0; k < steps; k++) r.run_once();
    }
    const size_t first = gen_bool(rng) ? 0 : 1; // chunks_exact_mut(2) of replicas[..] or replicas[1..]
    for (size_t i = first; i + 1 < replicas.size(); i += 2) {
      Replica& r0 = replicas[i];
      Replica& r1 = replicas[i + 1];
      const double de_db = (r0.energy() - r1.energy()) * (1.0 / r0.T - 1.0 / r1.T);
      if (de_db >= 0.0 || r1.rng.gen_f64() < o_exp(de_db)) {
        r0.accepted_swap_count += 1;
        r1.accepted_swap_count += 1;
        std::swap(r0.system, r1.system);
      } else {
        r0.rejected_swap_count += 1;
        r1.rejected_swap_count += 1;
      }
      const double e0 = r0.energy(), e1 = r1.energy();
      r0.total_energy += e0;
      r1.total_energy += e1;
      r0.total_energy_squared += e0 * e0;
      r1.total_energy_squared += e1 * e1;
      if (e0 >= 0.0) r0.ignored_count += 1;
      if (e1 >= 0.0) r1.ignored_count += 1;
    }
    moves += these_moves;
  }
};

} // namespace tempering
} // namespace oracle
