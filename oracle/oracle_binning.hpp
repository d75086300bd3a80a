// oracle_binning.hpp -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// CPU restatement of the `binning` binary's Monte Carlo: `EnergyMC<S>` of src/mc/energy_binning.rs on the
// `Binning` trait (src/mc/binning.rs:259-334) with its histogram implementation
// (src/mc/binning/histogram.rs) -- what fake/run-fake.py:25-26, wca-transposed/run.py:52 and
// wca/start-simulations.sh:13-15 run (`--histogram-bin`).  Function by function, in the reference's order,
// including the lazily maintained aggregates of `BinCounts` (min_total, max_total, min_count, ...), which the
// sampler reads in three places only: lnw.max_count (SAD), t_found.max_total (SAD) and hist.min_count (WL).
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "oracle_mc.hpp"

namespace oracle {
namespace binning {

inline double max_of(const std::vector<double>& v) { // histogram.rs:366-368: fold(NaN, f64::max)
  double m = std::numeric_limits<double>::quiet_NaN();
  for (double x : v) m = std::fmax(m, x);
  return m;
}
inline double min_of(const std::vector<double>& v) { // histogram.rs:370-372
  double m = std::numeric_limits<double>::quiet_NaN();
  for (double x : v) m = std::fmin(m, x);
  return m;
}

struct BinCounts { // histogram.rs:12-32
  std::vector<double> total;
  double min_total = 0, max_total = 0, e_max_total = -INFINITY;
  std::vector<uint64_t> count;
  uint64_t min_count = 0, max_count = 0;
  double e_max_count = -INFINITY;
  uint64_t total_count = 0;

  explicit BinCounts(size_t sz = 0) : total(sz, 0.0), count(sz, 0) {} // histogram.rs:35-47
  void insert_zero() {                                                // histogram.rs:48-53
    total.insert(total.begin(), 0.0);
    count.insert(count.begin(), 0);
    min_total = 0;
    min_count = 0;
  }
  void push_zero() { // histogram.rs:54-59
    total.push_back(0.0);
    count.push_back(0);
    min_total = 0;
    min_count = 0;
  }
  void increment_count(double e, size_t idx, double value) { // histogram.rs:60-80
    total_count += 1;
    const uint64_t old_count = count[idx];
    count[idx] += 1;
    const double old_total = total[idx];
    total[idx] += value;
    if (total[idx] > max_total && max_of(total) == total[idx]) {
      max_total = total[idx];
      e_max_total = e;
    }
    if (old_total == min_total) min_total = min_of(total);
    if (old_count == min_count) {
      uint64_t m = count[0];
      for (uint64_t c : count) m = c < m ? c : m;
      min_count = m;
    }
    if (count[idx] > max_count) {
      max_count = count[idx];
      e_max_count = e;
    }
  }
  double get_total(size_t idx) const { return idx < total.size() ? total[idx] : 0.0; }  // histogram.rs:82-88
  uint64_t get_count(size_t idx) const { return idx < count.size() ? count[idx] : 0; } // histogram.rs:89-95
};

struct Bins { // histogram.rs:99-111
  double min = INFINITY, min_e = INFINITY, max_e = -INFINITY, width = 1.0;
  BinCounts lnw;
  std::map<std::string, BinCounts> extra;

  Bins() {}
  Bins(double e, double w) : min((std::round(e / w) - 0.5) * w), min_e(e), max_e(e), width(w) {} // Binning::new, histogram.rs:170-180

  double index_to_energy(size_t i) const { return min + ((double)i + 0.5) * width; } // histogram.rs:132-134
  size_t energy_to_index(double e) const {                                          // histogram.rs:135-146
    if (e < min) return ~(size_t)0;
    const double i = (e - min) / width;
    if (i == (double)lnw.total.size()) return f64_as_usize(i) - 1;
    return f64_as_usize(i);
  }
  void prep_for_e(double e) { // histogram.rs:147-167
    if (lnw.count.empty()) min = std::floor(e / width) * width;
    while (e < min) {
      lnw.insert_zero();
      for (auto& kv : extra) kv.second.insert_zero();
      min -= width;
    }
    while (e >= min + width * (double)lnw.count.size()) {
      lnw.push_zero();
      for (auto& kv : extra) kv.second.push_zero();
    }
  }
  void increment_count(double e, double gamma) { // histogram.rs:181-191
    if (e > max_e) max_e = e;
    if (e < min_e) min_e = e;
    prep_for_e(e);
    lnw.increment_count(e, energy_to_index(e), gamma);
  }
  double get_lnw(double e) const { return lnw.get_total(energy_to_index(e)); }                   // histogram.rs:215-217
  double get_count(double e) const { return (double)lnw.get_count(energy_to_index(e)) / width; } // histogram.rs:218-221
  template <class F>
  void set_lnw(F f) { // histogram.rs:192-200
    for (size_t i = 0; i < lnw.count.size(); i++) {
      const double e = index_to_energy(i);
      double v;
      if (f(e, get_count(e), &v)) {
        lnw.total[i] = v;
        lnw.count[i] = 0;
      }
    }
  }
  template <class F>
  size_t count_states(F f) const { // histogram.rs:201-210
    size_t n = 0;
    for (size_t i = 0; i < lnw.count.size(); i++) {
      const double e = index_to_energy(i);
      if (f(e, get_count(e))) n++;
    }
    return n;
  }
  size_t num_states() const { return lnw.count.size(); }              // histogram.rs:211-213
  double max_count() const { return (double)lnw.max_count / width; } // histogram.rs:228-230

  void accumulate_extra(const std::string& name, double e, double value) { // histogram.rs:235-261
    prep_for_e(e);
    const size_t idx = energy_to_index(e);
    auto it = extra.find(name);
    if (it == extra.end()) it = extra.emplace(name, BinCounts(lnw.total.size())).first;
    BinCounts& d = it->second;
    d.total_count += 1;
    d.count[idx] += 1;
    if (d.count[idx] == d.min_count + 1) {
      uint64_t m = d.count[0];
      for (uint64_t c : d.count) m = c < m ? c : m;
      d.min_count = m;
    }
    if (d.count[idx] > d.max_count) d.max_count = d.count[idx];
    const double old_total = d.total[idx];
    d.total[idx] = old_total + value;
    if (d.total[idx] > d.max_total) d.max_total = d.total[idx];
    if (old_total == d.min_total) d.min_total = min_of(d.total);
  }
  void zero_out_extra(const std::string& name) { // histogram.rs:262-276
    auto it = extra.find(name);
    if (it == extra.end()) return;
    BinCounts& d = it->second;
    for (auto& v : d.count) v = 0;
    for (auto& v : d.total) v = 0.0;
    d.min_total = 0.0;
    d.max_total = -INFINITY;
    d.min_count = 0;
    d.max_count = 0;
    d.total_count = 0;
  }
  double mean_extra(const std::string& name, double e) const { // histogram.rs:277-288
    auto it = extra.find(name);
    if (it == extra.end()) return 0.0;
    const size_t idx = energy_to_index(e);
    return it->second.count[idx] > 0 ? it->second.get_total(idx) / (double)it->second.count[idx] : 0.0;
  }
  double total_extra(const std::string& name, double e) const { // histogram.rs:289-296
    auto it = extra.find(name);
    return it == extra.end() ? 0.0 : it->second.get_total(energy_to_index(e));
  }
  double max_total_extra(const std::string& name) const { // histogram.rs:297-303
    auto it = extra.find(name);
    return it == extra.end() ? 0.0 : it->second.max_total;
  }
  double mean_count_extra(const std::string& name) const { // histogram.rs:330-336
    auto it = extra.find(name);
    return it == extra.end() ? 0.0 : (double)it->second.total_count / (width * (double)num_states());
  }
  double min_count_extra(const std::string& name) const { // histogram.rs:337-343
    auto it = extra.find(name);
    return it == extra.end() ? 0.0 : (double)it->second.min_count / width;
  }
};


// ---- binning::linear (src/mc/binning/linear.rs): ln w and every accumulator interpolated linearly between bin points ----
struct LBinCounts { // linear.rs:12-32: counts are f64
  std::vector<double> total;
  double min_total = 0, max_total = 0, e_max_total = -INFINITY;
  std::vector<double> count;
  double min_count = 0, max_count = 0, e_max_count = -INFINITY;
  uint64_t total_count = 0;

  explicit LBinCounts(size_t sz = 0) : total(sz, 0.0), count(sz, 0.0) {} // linear.rs:60-72
  void insert_zero() {                                                   // linear.rs:73-78
    total.insert(total.begin(), 0.0);
    count.insert(count.begin(), 0.0);
    min_total = 0;
    min_count = 0;
  }
  void push_zero() { // linear.rs:79-84
    total.push_back(0.0);
    count.push_back(0.0);
    min_total = 0;
    min_count = 0;
  }
  // interpret_float_index (linear.rs:85-98) + the Idx iterator (40-57): sum of v[i] f over the one or two points
  double interpolate(const std::vector<double>& v, double fidx) const {
    const double len = (double)total.size();
    double acc = 0.0;
    if (fidx < -1.0) return acc;                                     // Idx::None
    if (fidx < 0.0) return acc + v[0] * (1.0 - (-fidx));             // Idx::One(0, -fidx)
    if (fidx < len - 1.0) {                                          // Idx::Two(i, fidx - i): (i + 1, o) first, then (i, 1 - o)
      const size_t i = f64_as_usize(fidx);
      const double o = fidx - (double)i;
      acc += v[i + 1] * o;
      acc += v[i] * (1.0 - o);
      return acc;
    }
    if (fidx < len) return acc + v[total.size() - 1] * (1.0 - (fidx - (len - 1.0))); // Idx::One(len - 1, fidx - (len - 1))
    return acc;
  }
  void increment_count(double elo, double ehi, double fidx, double value) { // linear.rs:99-134
    const size_t idx = f64_as_usize(fidx);
    const double offset = fidx - (double)idx;
    total_count += 1;
    const double old_count = count[idx], old_plus_count = count[idx + 1];
    count[idx] += 1.0 - offset;
    count[idx + 1] += offset;
    const double old_total = total[idx], old_plus_total = total[idx + 1];
    total[idx] += value * (1.0 - offset);
    total[idx + 1] += value * offset;
    if (total[idx] > max_total && max_of(total) == total[idx]) {
      max_total = total[idx];
      e_max_total = elo;
    }
    if (total[idx + 1] > max_total && max_of(total) == total[idx + 1]) {
      max_total = total[idx + 1];
      e_max_total = ehi;
    }
    if (old_total == min_total || old_plus_total == min_total) min_total = min_of(total);
    if (old_count == min_count || old_plus_count == min_count) min_count = min_of(count);
    if (count[idx] > max_count) {
      max_count = count[idx];
      e_max_count = elo;
    }
    if (count[idx + 1] > max_count) {
      max_count = count[idx + 1];
      e_max_count = ehi;
    }
  }
  double get_total(double fidx) const { return interpolate(total, fidx); } // linear.rs:136-142
  double get_count(double fidx) const { return interpolate(count, fidx); } // linear.rs:143-149
};

struct LinearBins { // linear.rs:153-165
  double min = INFINITY, min_e = INFINITY, max_e = -INFINITY, width = 1.0;
  LBinCounts lnw;
  std::map<std::string, LBinCounts> extra;

  LinearBins() {}
  LinearBins(double e, double w) : min((std::round(e / w) - 0.5) * w), min_e(e), max_e(e), width(w) {} // linear.rs:230-240

  double index_to_energy(size_t i) const { return min + ((double)i + 0.5) * width; } // linear.rs:200-202
  double energy_to_index(double e) const { return (e - min) / width; }               // linear.rs:203-205
  void prep_for_e(double e) {                                                        // linear.rs:206-226
    if (lnw.count.empty()) min = std::floor(e / width) * width;
    while (e < min) {
      lnw.insert_zero();
      for (auto& kv : extra) kv.second.insert_zero();
      min -= width;
    }
    while (e >= min + width * ((double)lnw.count.size() - 1.0)) {
      lnw.push_zero();
      for (auto& kv : extra) kv.second.push_zero();
    }
  }
  void increment_count(double e, double gamma) { // linear.rs:241-258
    if (e > max_e) max_e = e;
    if (e < min_e) min_e = e;
    prep_for_e(e);
    const double idx = energy_to_index(e);
    const size_t int_idx = f64_as_usize(idx);
    const double rescaled_gamma = gamma * 1.0 / width;
    lnw.increment_count(index_to_energy(int_idx), index_to_energy(int_idx + 1), idx, rescaled_gamma);
  }
  double get_lnw(double e) const { return lnw.get_total(energy_to_index(e)); }           // linear.rs:283-285
  double get_count(double e) const { return lnw.get_count(energy_to_index(e)) / width; } // linear.rs:286-289
  template <class F>
  void set_lnw(F f) { // linear.rs:259-267
    for (size_t i = 0; i < lnw.count.size(); i++) {
      const double e = index_to_energy(i);
      double v;
      if (f(e, get_count(e), &v)) {
        lnw.total[i] = v;
        lnw.count[i] = 0.0;
      }
    }
  }
  template <class F>
  size_t count_states(F f) const { // linear.rs:268-277
    size_t n = 0;
    for (size_t i = 0; i < lnw.count.size(); i++) {
      const double e = index_to_energy(i);
      if (f(e, get_count(e))) n++;
    }
    return n;
  }
  size_t num_states() const { return lnw.count.size(); }      // linear.rs:279-281
  double max_count() const { return lnw.max_count / width; } // linear.rs:296-298

  void accumulate_extra(const std::string& name, double e, double value) { // linear.rs:303-345
    prep_for_e(e);
    const double idx = energy_to_index(e);
    auto it = extra.find(name);
    if (it == extra.end()) it = extra.emplace(name, LBinCounts(lnw.total.size())).first;
    LBinCounts& d = it->second;
    const size_t int_idx = f64_as_usize(idx);
    const double offset = idx - (double)int_idx;
    d.total_count += 1;
    const double old_count = d.count[int_idx], old_plus_count = d.count[int_idx + 1];
    d.count[int_idx] += 1.0 - offset;
    d.count[int_idx + 1] += offset;
    if (old_count == d.min_count || old_plus_count == d.min_count) d.min_count = min_of(d.count);
    if (d.count[int_idx] > d.max_count) d.max_count = d.count[int_idx];
    if (d.count[int_idx + 1] > d.max_count) d.max_count = d.count[int_idx + 1];
    const double old_total = d.total[int_idx], old_plus_total = d.total[int_idx + 1];
    d.total[int_idx] = old_total + value * (1.0 - offset);
    d.total[int_idx + 1] = old_plus_total + value * offset;
    if (d.total[int_idx] > d.max_total) d.max_total = d.total[int_idx];
    if (d.total[int_idx + 1] > d.max_total) d.max_total = d.total[int_idx + 1];
    if (old_total == d.min_total || old_plus_total == d.min_total) d.min_total = min_of(d.total);
  }
  void zero_out_extra(const std::string& name) { // linear.rs:346-360
    auto it = extra.find(name);
    if (it == extra.end()) return;
    LBinCounts& d = it->second;
    for (auto& v : d.count) v = 0.0;
    for (auto& v : d.total) v = 0.0;
    d.min_total = 0.0;
    d.max_total = -INFINITY;
    d.min_count = 0.0;
    d.max_count = 0.0;
    d.total_count = 0;
  }
  double mean_extra(const std::string& name, double e) const { // linear.rs:361-373
    auto it = extra.find(name);
    if (it == extra.end()) return 0.0;
    const double idx = energy_to_index(e);
    return it->second.get_count(idx) > 0.0 ? it->second.get_total(idx) / it->second.get_count(idx) : 0.0;
  }
  double total_extra(const std::string& name, double e) const { // linear.rs:374-381
    auto it = extra.find(name);
    return it == extra.end() ? 0.0 : it->second.get_total(energy_to_index(e));
  }
  double max_total_extra(const std::string& name) const { // linear.rs:382-388
    auto it = extra.find(name);
    return it == extra.end() ? 0.0 : it->second.max_total;
  }
  double mean_count_extra(const std::string& name) const { // linear.rs:415-421
    auto it = extra.find(name);
    return it == extra.end() ? 0.0 : (double)it->second.total_count / (width * (double)num_states());
  }
  double min_count_extra(const std::string& name) const { // linear.rs:422-428
    auto it = extra.find(name);
    return it == extra.end() ? 0.0 : it->second.min_count / width;
  }
};

// `test_linear` (linear.rs:167-186) on this restatement; 0 = passes
inline int reference_test_linear() {
  {
    LinearBins b; // test_binning::<linear::Bins> first (binning.rs:336-364)
    const double eps = 1.0;
    if (b.get_count(eps) != 0.0 / eps) return 1;
    b.increment_count(eps, 1.0);
    if (!(b.get_count(eps) > 0.0 / eps)) return 2;
    b.accumulate_extra("datum", eps, 7.0);
    if (b.mean_extra("datum", eps) != 7.0) return 3;
    if (b.total_extra("datum", eps) != 7.0) return 4;
    if (b.max_total_extra("datum") != 7.0) return 5;
  }
  LinearBins bins;
  bins.increment_count(1.9, 1.0);
  if (!(bins.get_lnw(1.95) > bins.get_lnw(1.85))) return 6;
  if (!(bins.get_lnw(1.0) > 0.0)) return 7;
  if (!(bins.get_lnw(0.51) > 0.0)) return 8;
  if (!(bins.get_lnw(2.49) > 0.0)) return 9;
  if (bins.get_lnw(3.01) != 0.0) return 10;
  if (bins.get_lnw(-0.01) != 0.0) return 11;
  return 0;
}

// `test_binning::<histogram::Bins>` (binning.rs:336-364, histogram.rs:113-116) on this restatement; 0 = passes,
// otherwise the number of the assertion that failed
inline int reference_test_binning() {
  Bins b; // Default: min = +inf, width = 1, empty (histogram.rs:118-129)
  const double eps = 1.0;
  if (b.get_count(eps) != 0.0 / eps) return 1;
  b.increment_count(eps, 1.0);
  if (!(b.get_count(eps) > 0.0 / eps)) return 2;
  b.accumulate_extra("datum", eps, 7.0);
  if (b.mean_extra("datum", eps) != 7.0) return 3;
  if (b.total_extra("datum", eps) != 7.0) return 4;
  if (b.max_total_extra("datum") != 7.0) return 5;
  return 0;
}

enum { B_SAD = 1, B_SAMC = 2, B_WL = 3 };

struct Method { // energy_binning.rs:128-148
  int kind = B_SAD;
  // Sad
  uint64_t num_states = 0;
  double min_T = 0, too_lo = 0, too_hi = 0;
  uint64_t tL = 0;
  double tF = 0, latest_parameter = 0;
  // Samc
  double t0 = 0;
  // WL
  double gamma = 1.0;
  bool inv_t = false, has_min_gamma = false;
  double min_gamma = 0;
};

template <class BinsT>
struct EnergyMCT { // energy_binning.rs:92-126 (BinsT: the variant of binning::Bins, binning.rs:71-78)
  std::unique_ptr<System> system;
  Method method;
  uint64_t moves = 0, accepted_moves = 0;
  bool has_min = false, has_max = false;
  double min_allowed_energy = 0, max_allowed_energy = 0;
  bool acceptance_rate_plan = false;
  double move_plan_value = 0;
  double translation_scale = 0.05, acceptance_rate = 0.5;
  Rng rng;
  BinsT bins;
  bool has_high_resolution = false;
  Bins high_resolution;
  uint64_t verify_failures = 0;

  // from_params, energy_binning.rs:539-586 (+ Method::new 150-174)
  EnergyMCT(const MCParams& p, std::unique_ptr<System> sys, double high_resolution_de = NAN) : system(std::move(sys)) {
    rng = Rng::seed_from_u64(p.seed);
    if (p.randomize_first) system->randomize(rng); // SADMC_INIT_RANDOMIZE (engine-side start, not the reference's)
    has_min = !std::isnan(p.min_allowed_energy);
    has_max = !std::isnan(p.max_allowed_energy);
    min_allowed_energy = p.min_allowed_energy;
    max_allowed_energy = p.max_allowed_energy;
    if (has_max) { // energy_binning.rs:546-557
      for (uint64_t it = 0; it < p.max_relax; it++) {
        double newe;
        if (system->plan_move(rng, 0.05, &newe)) {
          if (newe < system->energy()) system->confirm();
          if (system->energy() < max_allowed_energy) break;
        }
      }
    }
    const double e0 = system->energy();
    switch (p.method) {
      case 1:
        method.kind = B_SAD;
        method.num_states = 0;
        method.min_T = p.sad_min_T;
        method.too_lo = e0;
        method.too_hi = e0;
        break;
      case 2:
        method.kind = B_SAMC;
        method.t0 = p.samc_t0;
        break;
      case 3:
        method.kind = B_WL;
        method.has_min_gamma = !std::isnan(p.wl_min_gamma);
        method.min_gamma = p.wl_min_gamma;
        break;
      case 4:
        method.kind = B_WL;
        method.inv_t = true;
        break;
      default: throw std::invalid_argument("energy_binning.rs has no such method");
    }
    // BinningParams::Histogram { bin } (binning.rs:50-69, default bin 1.0)
    const double width = !std::isnan(p.energy_bin) ? p.energy_bin : 1.0;
    bins = BinsT(e0, width);
    if (!std::isnan(high_resolution_de)) {
      has_high_resolution = true;
      high_resolution = Bins(e0, high_resolution_de);
    }
    acceptance_rate_plan = p.acceptance_rate_plan;
    move_plan_value = p.move_value;
    translation_scale = p.acceptance_rate_plan ? 0.05 : p.move_value; // energy_binning.rs:572-575
  }

  double gamma() const { // energy_binning.rs:507-533
    switch (method.kind) {
      case B_SAD: {
        const double num_states = (double)method.num_states;
        if (method.latest_parameter * method.tF * num_states == 0.0) return 0.0;
        const double t = (double)moves;
        return (method.latest_parameter + t / method.tF) / (method.latest_parameter + t / num_states * (t / method.tF));
      }
      case B_SAMC: {
        const double t = (double)moves;
        return t > method.t0 ? method.t0 / t : 1.0;
      }
      default: return method.gamma;
    }
  }

  bool reject_move(double e1, double e2) { // energy_binning.rs:276-321
    double lnw1 = bins.get_lnw(e1);
    double lnw2 = bins.get_lnw(e2);
    if (method.kind == B_SAD) {
      const double too_lo = method.too_lo, too_hi = method.too_hi, min_T = method.min_T;
      lnw1 = e1 < too_lo ? bins.get_lnw(too_lo) + (e1 - too_lo) / min_T : (e1 > too_hi ? bins.get_lnw(too_hi) : lnw1);
      lnw2 = e2 < too_lo ? bins.get_lnw(too_lo) + (e2 - too_lo) / min_T : (e2 > too_hi ? bins.get_lnw(too_hi) : lnw2);
      const bool rejected = lnw2 > lnw1 && rng.gen_f64() > o_exp(lnw1 - lnw2);
      if (!rejected && bins.get_count(e2) == 0.0 && e2 < too_hi && e2 > too_lo) method.tL = moves;
      return rejected;
    }
    return lnw2 > lnw1 && rng.gen_f64() > o_exp(lnw1 - lnw2);
  }

  void update_weights(double energy) { // energy_binning.rs:323-503
    const double g = gamma();
    const double old_highest_hist = bins.max_count();
    const double old_hist_here = bins.get_count(energy);
    bins.increment_count(energy, g);
    if (has_high_resolution) high_resolution.increment_count(energy, 0.0);
    bool switch_to_samc = false;
    double samc_t0 = 0;
    if (method.kind == B_SAD) {
      Method& m = method;
      const double hist_here = bins.get_count(energy);
      if (old_hist_here == 0.0) bins.accumulate_extra("t_found", energy, (double)moves);
      if (hist_here > old_highest_hist) {
        if (energy > m.too_hi) {
          const double lnw_too_hi = bins.get_lnw(m.too_hi);
          const double too_hi = m.too_hi;
          bins.set_lnw([&](double e, double count, double* v) {
            if (e > too_hi && count > 0.0) {
              *v = lnw_too_hi;
              return true;
            }
            return false;
          });
          m.latest_parameter = (energy - m.too_lo) / m.min_T;
          m.tL = moves;
          m.too_hi = energy;
          m.num_states = bins.count_states([&](double e, double count) { return e >= m.too_lo && e <= m.too_hi && count >= 0.0; });
        } else if (energy < m.too_lo) {
          const double lnw_too_lo = bins.get_lnw(m.too_lo);
          const double too_lo = m.too_lo, min_T = m.min_T;
          bins.set_lnw([&](double e, double count, double* v) {
            if (e < too_lo && count > 0.0) {
              *v = lnw_too_lo + (e - too_lo) / min_T;
              return true;
            }
            return false;
          });
          m.latest_parameter = (m.too_hi - energy) / m.min_T;
          m.tL = moves;
          m.too_lo = energy;
          m.num_states = bins.count_states([&](double e, double count) { return e >= m.too_lo && e <= m.too_hi && count >= 0.0; });
        }
      }
      if (m.tL == moves) {
        const double old_tF = m.tF;
        m.tF = bins.max_total_extra("t_found");
        if (old_tF != m.tF && acceptance_rate_plan) {
          double s = acceptance_rate / move_plan_value;
          s = s < 0.8 ? 0.8 : (s > 1.2 ? 1.2 : s);
          translation_scale *= s;
        }
      }
    } else if (method.kind == B_WL) {
      Method& m = method;
      const double old_lowest_hist = bins.min_count_extra("hist");
      bins.accumulate_extra("hist", energy, 0.0);
      if (m.has_min_gamma && m.gamma < m.min_gamma) return; // production run
      const double lowest_hist = bins.min_count_extra("hist");
      if (lowest_hist > old_lowest_hist && (!has_min || bins.get_count(min_allowed_energy) > 0.0) &&
          (!has_max || bins.get_count(max_allowed_energy) > 0.0)) {
        if ((m.inv_t && lowest_hist > 0.0) || lowest_hist >= 0.8 * bins.mean_count_extra("hist")) {
          m.gamma *= 0.5;
          bins.zero_out_extra("hist");
          if (m.has_min_gamma && m.gamma < m.min_gamma) m.gamma = 0.0;
        }
        if (m.inv_t && m.gamma < (double)bins.num_states() / (double)moves) {
          switch_to_samc = true;
          samc_t0 = (double)bins.num_states();
        }
      }
    }
    if (switch_to_samc) {
      method.kind = B_SAMC;
      method.t0 = samc_t0;
    }
  }

  void move_once() { // energy_binning.rs:592-633 (the plugin tick belongs to the host)
    moves += 1;
    if (moves % 100000000ull == 0)
      if (!system->verify_energy()) verify_failures++;
    const double e1 = system->energy();
    bins.accumulate_extra("energy", e1, e1);
    {
      std::string key;
      double value;
      if (system->data_to_collect(moves, &key, &value)) bins.accumulate_extra(key, e1, value);
    }
    const double recent_scale = std::sqrt(1.0 / (double)moves);
    acceptance_rate *= 1.0 - recent_scale;
    double e2;
    if (system->plan_move(rng, translation_scale, &e2)) {
      bool out_of_bounds = false;
      if (has_max) out_of_bounds = e2 > max_allowed_energy && e2 > e1;
      if (has_min) out_of_bounds = out_of_bounds || (e2 < min_allowed_energy && e2 < e1);
      if (!out_of_bounds) {
        if (!reject_move(e1, e2)) {
          accepted_moves += 1;
          acceptance_rate += recent_scale;
          system->confirm();
        }
      }
    }
    update_weights(system->energy());
  }
};

using EnergyMC = EnergyMCT<Bins>;             // BinningParams::Histogram
using EnergyMCLinear = EnergyMCT<LinearBins>; // BinningParams::Linear

} // namespace binning
} // namespace oracle
