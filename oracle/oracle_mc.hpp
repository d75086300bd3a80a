// oracle_mc.hpp -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// CPU restatement of `EnergyMC<S>` (src/mc/energy.rs): the flat-histogram
// bookkeeping half of the hot path -- SAD, SAMC, WL, 1/t-WL and canonical on a
// growable 1-D energy histogram.  Function by function, in the reference's
// order, with the reference's growable vectors (front inserts included).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "oracle_systems.hpp"

namespace oracle {

// Rust `x as usize` for f64: truncates toward zero, saturates, NaN -> 0.
inline size_t f64_as_usize(double x) {
  if (!(x > 0.0)) return 0; // negative, -0, NaN
  if (x >= 18446744073709551615.0) return ~(size_t)0;
  return (size_t)x;
}

struct BinCounts { // energy.rs:136-142
  std::vector<double> total;
  std::vector<uint64_t> count;
};

struct Bins { // energy.rs:146-163
  double min = 0, width = 1;
  std::vector<uint64_t> histogram, t_found;
  std::vector<double> lnw, energy_total, energy_squared_total;
  std::map<std::string, BinCounts> extra;

  double index_to_energy(size_t i) const { return min + ((double)i + 0.5) * width; } // energy.rs:366-370
  size_t energy_to_index(double e) const { return f64_as_usize((e - min) / width); } // energy.rs:371-373
  void accumulate_extra(const std::string& k, size_t idx, double value) {            // energy.rs:374-386
    auto it = extra.find(k);
    if (it == extra.end()) {
      BinCounts b;
      b.count.assign(lnw.size(), 0);
      b.total.assign(lnw.size(), 0.0);
      it = extra.emplace(k, b).first;
    }
    it->second.count[idx] += 1;
    it->second.total[idx] += value;
  }
};

enum MethodKind { M_SAD = 1, M_SAMC = 2, M_WL = 3, M_CANONICAL = 5 };

struct Method { // energy.rs:212-244
  int kind = M_SAD;
  // Sad
  double min_T = 0, too_lo = 0, too_hi = 0;
  uint64_t tL = 0, tF = 0, num_states = 0, highest_hist = 0;
  double latest_parameter = 0;
  // Samc
  double t0 = 0;
  // WL
  double gamma = 0;
  uint64_t wl_lowest_hist = 0, wl_highest_hist = 0, wl_total_hist = 0;
  double wl_num_states = 0;
  std::vector<uint64_t> hist;
  double min_energy = 0;
  bool inv_t = false;
  bool has_min_gamma = false;
  double min_gamma = 0;
  // Canonical
  double temperature = 0;
};

struct MCParams { // EnergyMCParams, energy.rs:82-97 (+ MethodParams 44-69, MoveParams 73-78)
  int method = 1; // sadmc_method_kind: 1 Sad, 2 Samc, 3 WL, 4 Inv_t_WL, 5 Canonical
  double sad_min_T = 0.2, samc_t0 = 0, wl_min_gamma = NAN, canonical_T = 0;
  uint64_t seed = 0;
  double energy_bin = NAN, min_allowed_energy = NAN, max_allowed_energy = NAN;
  bool acceptance_rate_plan = false;
  double move_value = 0.05;
  bool randomize_first = false; // SADMC_INIT_RANDOMIZE: system.randomize(mc rng) before the relaxation
  uint64_t max_relax = 100000000ull; // the reference's 1e8 (energy.rs:841)
};

struct EnergyMC { // energy.rs:167-210
  std::unique_ptr<System> system;
  Method method;
  uint64_t moves = 0, accepted_moves = 0;
  bool has_min = false, has_max = false;
  double min_allowed_energy = 0, max_allowed_energy = 0;
  bool acceptance_rate_plan = false;
  double move_plan_value = 0;
  double translation_scale = 0.05, acceptance_rate = 0.5;
  Rng rng;
  Bins bins;
  std::vector<uint8_t> have_visited_since_maxentropy;
  std::vector<uint64_t> round_trips;
  double max_S = 0;
  size_t max_S_index = 0;
  uint64_t verify_failures = 0;

  // Method::new, energy.rs:246-313
  static Method new_method(const MCParams& p, double E, double dE, bool has_min, double mine, bool has_max, double maxe) {
    Method m;
    switch (p.method) {
      case 1:
        m.kind = M_SAD;
        m.min_T = p.sad_min_T;
        m.too_lo = E;
        m.too_hi = E;
        m.tL = 0;
        m.tF = 0;
        m.num_states = 1;
        m.highest_hist = 1;
        m.latest_parameter = 0.0;
        break;
      case 2:
        m.kind = M_SAMC;
        m.t0 = p.samc_t0;
        break;
      case 3:
      case 4:
        m.kind = M_WL;
        m.gamma = 1.0;
        m.wl_lowest_hist = (has_min && has_max) ? 0 : 1;
        m.wl_highest_hist = 1;
        m.wl_total_hist = 0;
        m.wl_num_states = (has_min && has_max) ? (maxe - mine) / dE : 1.0;
        m.min_energy = E;
        m.inv_t = p.method == 4;
        m.has_min_gamma = p.method == 3 && !std::isnan(p.wl_min_gamma);
        m.min_gamma = p.wl_min_gamma;
        break;
      case 5:
        m.kind = M_CANONICAL;
        m.temperature = p.canonical_T;
        break;
    }
    return m;
  }

  // MonteCarlo::from_params, energy.rs:830-898
  EnergyMC(const MCParams& p, std::unique_ptr<System> sys) : system(std::move(sys)) {
    double native;
    const double ewidth = !std::isnan(p.energy_bin) ? p.energy_bin : (system->has_delta_energy(&native) ? native : 1.0);
    rng = Rng::seed_from_u64(p.seed);
    if (p.randomize_first) system->randomize(rng);
    has_min = !std::isnan(p.min_allowed_energy);
    has_max = !std::isnan(p.max_allowed_energy);
    min_allowed_energy = p.min_allowed_energy;
    max_allowed_energy = p.max_allowed_energy;
    if (has_max) { // energy.rs:840-851
      for (uint64_t it = 0; it < p.max_relax; it++) {
        double newe;
        if (system->plan_move(rng, 0.05, &newe)) {
          if (newe < system->energy()) system->confirm();
          if (system->energy() < max_allowed_energy) break;
        }
      }
    }
    const double e0 = system->energy();
    const double emin = (std::round(e0 / ewidth) - 0.5) * ewidth; // energy.rs:852 (f64::round: half away from zero)
    method = new_method(p, e0, ewidth, has_min, min_allowed_energy, has_max, max_allowed_energy);
    bins.histogram = {1};
    bins.t_found = {0};
    bins.lnw = {0.0};
    bins.energy_total = {e0};
    bins.energy_squared_total = {e0 * e0};
    bins.min = emin;
    bins.width = ewidth;
    have_visited_since_maxentropy = {0};
    round_trips = {1};
    acceptance_rate_plan = p.acceptance_rate_plan;
    move_plan_value = p.move_value;
    translation_scale = p.acceptance_rate_plan ? 0.05 : p.move_value; // energy.rs:884-887
  }

  // energy.rs:400-434
  void prepare_for_state(double e) {
    while (e < bins.min) {
      bins.histogram.insert(bins.histogram.begin(), 0);
      bins.t_found.insert(bins.t_found.begin(), 0);
      bins.lnw.insert(bins.lnw.begin(), 0.0);
      bins.energy_total.insert(bins.energy_total.begin(), 0.0);
      bins.energy_squared_total.insert(bins.energy_squared_total.begin(), 0.0);
      for (auto& kv : bins.extra) {
        kv.second.count.insert(kv.second.count.begin(), 0);
        kv.second.total.insert(kv.second.total.begin(), 0.0);
      }
      have_visited_since_maxentropy.insert(have_visited_since_maxentropy.begin(), 1);
      round_trips.insert(round_trips.begin(), 1);
      bins.min -= bins.width;
    }
    while (e >= bins.min + bins.width * (double)bins.lnw.size()) {
      bins.lnw.push_back(0.0);
      bins.histogram.push_back(0);
      bins.t_found.push_back(0);
      for (auto& kv : bins.extra) {
        kv.second.count.push_back(0);
        kv.second.total.push_back(0.0);
      }
      bins.energy_total.push_back(0.0);
      bins.energy_squared_total.push_back(0.0);
      have_visited_since_maxentropy.push_back(1);
      round_trips.push_back(1);
    }
  }

  // SadVersion::compute_gamma, energy.rs:21-39
  static double sad_gamma(double latest_parameter, double t, double tF, double num_states) {
    if (latest_parameter * tF * num_states == 0.0) return 0.0;
    return (latest_parameter + t / tF) / (latest_parameter + t / num_states * (t / tF));
  }
  // energy.rs:799-824
  double gamma() const {
    switch (method.kind) {
      case M_CANONICAL: return 0.0;
      case M_SAD: return sad_gamma(method.latest_parameter, (double)moves, (double)method.tF, (double)method.num_states);
      case M_SAMC: {
        const double t = (double)moves;
        return t > method.t0 ? method.t0 / t : 1.0;
      }
      default: return method.gamma;
    }
  }

  // energy.rs:440-512
  bool reject_move(double e1, double e2) {
    const size_t i1 = bins.energy_to_index(e1);
    const size_t i2 = bins.energy_to_index(e2);
    const std::vector<double>& lnw = bins.lnw;
    switch (method.kind) {
      case M_SAD: {
        const double too_lo = method.too_lo, too_hi = method.too_hi, min_T = method.min_T;
        const double lnw1 = e1 < too_lo   ? lnw[bins.energy_to_index(too_lo)] + (e1 - too_lo) / min_T
                            : e1 > too_hi ? lnw[bins.energy_to_index(too_hi)]
                                          : lnw[i1];
        const double lnw2 = e2 < too_lo   ? lnw[bins.energy_to_index(too_lo)] + (e2 - too_lo) / min_T
                            : e2 > too_hi ? lnw[bins.energy_to_index(too_hi)]
                                          : lnw[i2];
        const bool rejected = lnw2 > lnw1 && rng.gen_f64() > o_exp(lnw1 - lnw2);
        if (!rejected && bins.histogram[i2] == 0 && e2 < too_hi && e2 > too_lo) {
          method.latest_parameter = (too_hi - too_lo) / min_T;
          method.num_states += 1;
          method.tL = moves;
        }
        return rejected;
      }
      case M_SAMC: {
        const double lnw1 = lnw[i1], lnw2 = lnw[i2];
        return lnw2 > lnw1 && rng.gen_f64() > o_exp(lnw1 - lnw2);
      }
      case M_WL: {
        const double lnw1 = lnw[i1], lnw2 = lnw[i2];
        const bool rejected = lnw2 > lnw1 && rng.gen_f64() > o_exp(lnw1 - lnw2);
        if (!rejected && bins.histogram[i2] == 0 && method.wl_lowest_hist > 0) method.wl_num_states += 1.0;
        return rejected;
      }
      default: // canonical
        if (e1 >= e2) return false;
        return rng.gen_f64() > o_exp((e1 - e2) / method.temperature);
    }
  }

  // energy.rs:514-761
  void update_weights(double energy) {
    const size_t i = bins.energy_to_index(energy);
    const double g = gamma();
    const double old_lnw = bins.lnw[i];
    bins.lnw[i] += g;
    bool switch_to_samc = false;
    double samc_t0 = 0;
    if (method.kind == M_SAD) {
      Method& m = method;
      const std::vector<uint64_t>& histogram = bins.histogram;
      if (m.too_lo > m.too_hi || energy < m.too_lo || energy > m.too_hi) bins.lnw[i] = old_lnw;
      if (histogram[i] > m.highest_hist) {
        m.highest_hist = histogram[i];
        if (energy > m.too_hi) {
          const size_t ihi = bins.energy_to_index(m.too_hi);
          for (size_t j = 0; j < histogram.size(); j++) {
            const double ej = bins.index_to_energy(j);
            if (ej > m.too_hi && ej <= energy) {
              if (histogram[j] != 0) {
                bins.lnw[j] = bins.lnw[ihi];
                m.num_states += 1;
              } else {
                bins.lnw[j] = 0.0;
              }
            }
          }
          m.latest_parameter = (energy - m.too_lo) / m.min_T;
          m.tL = moves;
          m.too_hi = bins.index_to_energy(bins.energy_to_index(energy));
        } else if (energy < m.too_lo) {
          const size_t ilo = bins.energy_to_index(m.too_lo);
          for (size_t j = 0; j < histogram.size(); j++) {
            const double ej = bins.index_to_energy(j);
            if (ej < m.too_lo && ej >= energy) {
              if (histogram[j] != 0) {
                bins.lnw[j] = bins.lnw[ilo] + (ej - m.too_lo) / m.min_T;
                if (bins.lnw[j] < 0.0) bins.lnw[j] = 0.0;
                m.num_states += 1;
              } else {
                bins.lnw[j] = 0.0;
              }
            }
          }
          m.latest_parameter = (m.too_hi - energy) / m.min_T;
          m.tL = moves;
          m.too_lo = bins.index_to_energy(bins.energy_to_index(energy));
        }
      }
      if (m.tL == moves) {
        const size_t ilo = bins.energy_to_index(m.too_lo);
        const size_t ihi = bins.energy_to_index(m.too_hi);
        const uint64_t old_tF = m.tF;
        uint64_t mx = 0;
        for (size_t j = ilo; j < ihi + 1; j++) mx = std::max(mx, bins.t_found[j]);
        m.tF = mx;
        if (old_tF != m.tF && acceptance_rate_plan) {
          double s = acceptance_rate / move_plan_value;
          s = s < 0.8 ? 0.8 : (s > 1.2 ? 1.2 : s);
          translation_scale *= s;
        }
      }
    } else if (method.kind == M_WL) {
      Method& m = method;
      if (m.has_min_gamma && m.gamma < m.min_gamma) { // production run
        m.hist[i] += 1;
        return;
      }
      if (m.hist.size() != bins.lnw.size()) {
        if (m.hist.empty() || (m.gamma != 1.0 && m.wl_lowest_hist > 0)) {
          m.gamma = 1.0;
          m.wl_lowest_hist = 0;
          m.wl_highest_hist = 0;
          m.wl_total_hist = 0;
          m.hist.assign(bins.lnw.size(), 0);
          m.min_energy = bins.min;
        } else {
          while (m.min_energy > bins.min) {
            m.min_energy -= bins.width;
            m.hist.insert(m.hist.begin(), 0);
          }
          while (m.hist.size() < bins.lnw.size()) m.hist.push_back(0);
          m.wl_lowest_hist = *std::min_element(m.hist.begin(), m.hist.end());
        }
      }
      m.hist[i] += 1;
      if (m.hist[i] > m.wl_highest_hist) m.wl_highest_hist = m.hist[i];
      m.wl_total_hist += 1;
      const std::vector<uint64_t>& histogram = bins.histogram;
      const double max_energy = m.min_energy + (double)m.hist.size() * bins.width;
      bool cond = m.hist[i] == m.wl_lowest_hist + 1 && m.hist.size() > 1 && (!has_min || min_allowed_energy >= m.min_energy) &&
                  (!has_max || max_allowed_energy <= max_energy);
      if (cond) {
        bool any = false;
        uint64_t mn = 0;
        for (size_t j = 0; j < m.hist.size(); j++)
          if (histogram[j] != 0) {
            if (!any || m.hist[j] < mn) mn = m.hist[j];
            any = true;
          }
        cond = any && mn == m.wl_lowest_hist + 1;
      }
      if (cond) {
        m.wl_lowest_hist = m.hist[i];
        if ((m.inv_t && m.wl_lowest_hist > 0) || (double)m.wl_lowest_hist >= 0.8 * (double)m.wl_total_hist / m.wl_num_states) {
          m.gamma *= 0.5;
          for (auto& h : m.hist) h = 0;
          m.wl_total_hist = 0;
          m.wl_lowest_hist = 0;
          m.wl_highest_hist = 0;
          if (m.has_min_gamma && m.gamma < m.min_gamma) m.gamma = 0.0;
        }
        if (m.inv_t && m.gamma < m.wl_num_states / (double)moves) {
          switch_to_samc = true;
          samc_t0 = m.wl_num_states;
        }
      }
    }
    if (switch_to_samc) { // energy.rs:754-756
      method.kind = M_SAMC;
      method.t0 = samc_t0;
    }
  }

  // energy.rs:904-974 (the plugin tick, 967-973, belongs to the host)
  void move_once() {
    moves += 1;
    {
      const uint64_t len = bins.histogram.size();
      if (moves % (len * len * 1000) == 0)
        if (!system->verify_energy()) verify_failures++;
    }
    const double e1 = system->energy();
    const double recent_scale = std::sqrt(1.0 / (double)moves);
    acceptance_rate *= 1.0 - recent_scale;
    double e2;
    if (system->plan_move(rng, translation_scale, &e2)) {
      bool out_of_bounds = false;
      if (has_max) out_of_bounds = e2 > max_allowed_energy && e2 > e1;
      if (has_min) out_of_bounds = out_of_bounds || (e2 < min_allowed_energy && e2 < e1);
      if (!out_of_bounds) {
        prepare_for_state(e2);
        if (!reject_move(e1, e2)) {
          accepted_moves += 1;
          acceptance_rate += recent_scale;
          system->confirm();
        }
      }
    }
    const double energy = system->energy();
    const size_t i = bins.energy_to_index(energy);
    if (bins.histogram[i] == 0) bins.t_found[i] = moves;
    bins.histogram[i] += 1;
    bins.energy_total[i] += energy;
    bins.energy_squared_total[i] += energy * energy;
    {
      std::string key;
      double value;
      if (system->data_to_collect(moves, &key, &value)) bins.accumulate_extra(key, i, value);
    }
    update_weights(energy);

    if (bins.lnw[i] > max_S) {
      max_S = bins.lnw[i];
      max_S_index = i;
      std::fill(have_visited_since_maxentropy.begin(), have_visited_since_maxentropy.end(), 1);
    } else if (i == max_S_index) {
      if (bins.energy_to_index(e1) != i) std::fill(have_visited_since_maxentropy.begin(), have_visited_since_maxentropy.end(), 0);
    } else if (!have_visited_since_maxentropy[i]) {
      have_visited_since_maxentropy[i] = 1;
      round_trips[i] += 1;
    }
  }
};

} // namespace oracle
