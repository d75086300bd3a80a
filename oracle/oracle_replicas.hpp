// oracle_replicas.hpp -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// CPU restatement of the `replicas` binary: `MedianEstimator`, `Replica` and `MC` of src/mc/energy_replicas.rs (energy
// ceilings instead of temperatures: a replica accepts every move that stays below its max_energy; neighbours swap systems
// when the upper one has come below the lower one's ceiling; a new, lower replica is split off at the median of the
// energies seen below the lowest cutoff).  fake/run-fake.py:16-23 and two-wells/run-two-wells.py run it.  Function by
// function, in the reference's order.  `jump` and `gen::<bool>` as in oracle_tempering.hpp.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "oracle_tempering.hpp"

namespace oracle {
namespace replicas {

constexpr size_t ESTIMATOR_SIZE = 4096; // energy_replicas.rs:49

struct MedianEstimator { // energy_replicas.rs:45-99
  std::vector<double> energies;
  explicit MedianEstimator(double e = 0) { energies.push_back(e); }
  void reset(double e) {
    energies.clear();
    energies.push_back(e);
  }
  void add_energy(double e, Rng& rng) { // 60-69
    if (energies.size() < ESTIMATOR_SIZE) {
      energies.push_back(e);
    } else if (rng.gen_f64() < 1.0 / ((double)energies.size() + 1.0)) {
      const size_t i = rng.gen_range_usize(0, energies.size());
      energies[i] = e;
    }
  }
  double median() { // 70-98
    std::sort(energies.begin(), energies.end());
    const size_t middle = energies.size() / 2;
    const double e_middle = energies[middle];
    for (size_t k = middle + 1; k < energies.size(); k++)
      if (energies[k] != e_middle) return 0.5 * (e_middle + energies[k]);
    for (size_t k = middle; k-- > 0;)
      if (energies[k] != e_middle) return 0.5 * (e_middle + energies[k]);
    return e_middle;
  }
};

struct Replica { // energy_replicas.rs:103-145
  double max_energy = INFINITY, cutoff_energy = 0;
  uint64_t rejected_count = 0, accepted_count = 0, above_count = 0, below_count = 0, upwelling_count = 0;
  double above_total = 0, below_total = 0, above_total_squared = 0, below_total_squared = 0;
  std::map<std::string, std::pair<double, uint64_t>> above_extra;
  double lowest_max_energy = INFINITY;
  std::unique_ptr<System> system;
  uint64_t unique_visitors = 1;
  bool collecting_data = true;
  Rng rng;
  double translation_scale = 1.0;

  void decimate() { // 171-196
    upwelling_count = 0;
    if (above_count > 1) {
      above_total /= (double)above_count;
      above_total_squared /= (double)above_count;
      above_count = 1;
    }
    if (below_count > 1) {
      below_total /= (double)below_count;
      below_total_squared /= (double)below_count;
      below_count = 1;
    }
    for (auto& kv : above_extra)
      if (kv.second.second > 1) {
        kv.second.first /= (double)kv.second.second;
        kv.second.second = 1;
      }
    lowest_max_energy = max_energy;
    accepted_count = 1;
    rejected_count = 1;
    unique_visitors = 1;
  }
  double energy() const { return system->energy(); }
  void run_once(uint64_t moves, double very_lowest_max_energy) { // 206-248
    if (std::isfinite(max_energy)) {
      double e;
      if (system->plan_move(rng, translation_scale, &e)) {
        if (e < max_energy) {
          system->confirm();
          accepted_count += 1;
        } else {
          rejected_count += 1;
        }
      } else {
        rejected_count += 1;
      }
    } else {
      system->randomize(rng);
      lowest_max_energy = max_energy;
    }
    const double e = system->energy();
    if (collecting_data) {
      if (e > cutoff_energy) {
        above_count += 1;
        above_total += e;
        above_total_squared += e * e;
        std::string key;
        double value;
        if (system->data_to_collect(moves, &key, &value)) {
          auto it = above_extra.find(key);
          if (it != above_extra.end()) {
            it->second.first += value;
            it->second.second += 1;
          } else {
            above_extra.emplace(key, std::make_pair(value, (uint64_t)1));
          }
        }
      } else {
        below_count += 1;
        below_total += e;
        below_total_squared += e * e;
      }
      if (lowest_max_energy == very_lowest_max_energy) upwelling_count += 1;
    }
  }
  void occasional_update(uint64_t min_moves_to_randomize) { // 249-290
    if (rejected_count > 128 && accepted_count > 128 && std::isfinite(max_energy)) {
      const double acceptance_ratio = (double)accepted_count / (double)rejected_count;
      const double max_acceptance_ratio = (double)min_moves_to_randomize;
      if (acceptance_ratio < 0.5 || acceptance_ratio > 2.0 * max_acceptance_ratio) {
        double adjustment = acceptance_ratio < 0.5 ? acceptance_ratio / std::sqrt(max_acceptance_ratio) : acceptance_ratio * std::sqrt(max_acceptance_ratio);
        if (adjustment > 2.0)
          adjustment = 2.0;
        else if (adjustment < 0.5)
          adjustment = 0.5;
        translation_scale *= adjustment;
        accepted_count = 0;
        rejected_count = 0;
      }
    }
  }
};

struct SystemTraits { // what MC needs from MovableSystem besides the System interface (system/mod.rs:93-95, 119)
  uint64_t min_moves_to_randomize = 1, dimensionality = 1;
  double max_size = 1.0;
};

struct MC { // energy_replicas.rs:307-333
  double min_T = 0.2;
  Rng rng;
  uint64_t moves = 0, independent_systems_before_new_bin = 64;
  MedianEstimator median;
  std::vector<Replica> replicas;
  SystemTraits traits;
  std::unique_ptr<System> (*clone)(const System&, const void*) = nullptr;
  const void* clone_ctx = nullptr;

  // from_params, energy_replicas.rs:346-399.  max_init: MAX_INIT = 1 << 15 in the reference (a smaller number in fast tests)
  MC(uint64_t seed, double min_T_, uint64_t indep, SystemTraits tr, std::unique_ptr<System> system, std::unique_ptr<System> (*cl)(const System&, const void*),
     const void* ctx, size_t max_init = (size_t)1 << 15)
      : min_T(min_T_), independent_systems_before_new_bin(indep), traits(tr), clone(cl), clone_ctx(ctx) {
    rng = Rng::seed_from_u64(seed);
    std::vector<double> energies;
    energies.reserve(max_init);
    for (size_t k = 0; k < max_init; k++) energies.push_back(system->randomize(rng));
    std::unique_ptr<System> high_system = clone(*system, clone_ctx);
    high_system->randomize(rng);
    while (system->energy() > energies[energies.size() / 2]) system->randomize(rng); // the UNSORTED list, as the reference has it
    std::sort(energies.begin(), energies.end());
    Replica r0, r1;
    r0.max_energy = INFINITY;
    r0.cutoff_energy = energies[energies.size() / 2];
    r0.lowest_max_energy = r0.max_energy;
    r0.translation_scale = traits.max_size;
    r0.system = std::move(high_system);
    r0.rng = rng;
    r1.max_energy = energies[energies.size() / 2];
    r1.cutoff_energy = energies[energies.size() / 4];
    r1.lowest_max_energy = r1.max_energy;
    r1.translation_scale = traits.max_size;
    r1.system = std::move(system);
    r1.rng = rng;
    replicas.push_back(std::move(r0));
    replicas.push_back(std::move(r1));
    tempering::jump(rng);
    median = MedianEstimator(energies[energies.size() / 4]);
  }

  void run_once() { // energy_replicas.rs:504-642 (movie / report / save belong to the host)
    const uint64_t moves0 = moves;
    const uint64_t steps = traits.min_moves_to_randomize;
    uint64_t these_moves = 0;
    const double lowest_max_energy = replicas.back().max_energy;
    for (auto& r : replicas) {
      these_moves += steps;
      if (std::isfinite(r.max_energy)) {
        for (uint64_t i = 0; i < steps; i++) r.run_once(moves0 + i, lowest_max_energy);
      } else {
        r.run_once(moves0, lowest_max_energy);
      }
    }
    const size_t first = tempering::gen_bool(rng) ? 0 : 1;
    for (size_t i = first; i + 1 < replicas.size(); i += 2) {
      Replica& r0 = replicas[i];
      Replica& r1 = replicas[i + 1];
      if (r0.energy() < r1.max_energy) {
        std::swap(r0.system, r1.system);
        std::swap(r0.lowest_max_energy, r1.lowest_max_energy);
        r0.collecting_data = true;
        r1.collecting_data = true;
        if (r1.lowest_max_energy > r1.max_energy) {
          r1.unique_visitors += 1;
          r1.lowest_max_energy = r1.max_energy;
        }
      }
    }
    {
      const double last_energy = replicas.back().energy();
      if (last_energy < replicas.back().cutoff_energy) median.add_energy(last_energy, rng);
    }
    {
      Replica& r = replicas.back();
      if (r.unique_visitors >= independent_systems_before_new_bin && r.lowest_max_energy == r.max_energy) {
        const double mean_below = r.below_total / (double)r.below_count;
        if (mean_below + min_T < r.cutoff_energy && r.energy() < r.cutoff_energy) {
          const double median_below = median.median();
          median.reset(median_below);
          Replica n; // r.clone()
          n.max_energy = r.max_energy;
          n.cutoff_energy = r.cutoff_energy;
          n.rejected_count = r.rejected_count;
          n.accepted_count = r.accepted_count;
          n.above_count = r.above_count;
          n.below_count = r.below_count;
          n.upwelling_count = r.upwelling_count;
          n.above_total = r.above_total;
          n.below_total = r.below_total;
          n.above_total_squared = r.above_total_squared;
          n.below_total_squared = r.below_total_squared;
          n.above_extra = r.above_extra;
          n.lowest_max_energy = r.lowest_max_energy;
          n.system = clone(*r.system, clone_ctx);
          n.unique_visitors = r.unique_visitors;
          n.collecting_data = r.collecting_data;
          n.rng = r.rng;
          n.translation_scale = r.translation_scale;
          n.max_energy = r.cutoff_energy;
          n.cutoff_energy = median_below;
          n.decimate();
          n.lowest_max_energy = -INFINITY;
          n.translation_scale = r.translation_scale * std::pow(0.5, 1.0 / (double)traits.dimensionality);
          tempering::jump(n.rng);
          replicas.push_back(std::move(n));
        }
      }
    }
    for (auto& r : replicas) r.occasional_update(traits.min_moves_to_randomize);
    moves += these_moves;
  }
};

} // namespace replicas
} // namespace oracle
