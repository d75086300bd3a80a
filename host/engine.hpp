// engine.hpp -- C++ face of the C ABI (include/sadmc_gpu.h): what a compiled host binds.
//
// `GpuEnergyMC` stands where the reference has `EnergyMC<Any>` (src/mc/energy.rs:167-210) for MANY walkers: names
// follow the `MonteCarlo` and `System` traits (mc/mod.rs:37-143, system/mod.rs:54-120) -- num_moves,
// num_accepted_moves, system, energy, compute_energy, plan_move, confirm, verify_energy -- and `run(n)` is
// `n x move_once()` for every walker in one kernel launch.  Panics of the reference (invalid configuration,
// failed verification) arrive as `EngineError` carrying the negative status and `sadmc_last_error()`.
// The library is loaded with dlopen so that argument parsing and --dry-run work on hosts without CUDA; there is
// no CPU fallback: creating an engine without a device fails (SADMC_ERR_CUDA).
#pragma once
#include <dlfcn.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "../include/sadmc_gpu.h"

namespace sadmc_host {

struct EngineError : std::runtime_error {
  int code;
  EngineError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

struct Api {
  void* handle = nullptr;
#define SADMC_FN(name) decltype(&::name) name = nullptr;
  SADMC_FN(sadmc_create)
  SADMC_FN(sadmc_destroy)
  SADMC_FN(sadmc_last_error)
  SADMC_FN(sadmc_abi_version)
  SADMC_FN(sadmc_reference_system)
  SADMC_FN(sadmc_start)
  SADMC_FN(sadmc_run)
  SADMC_FN(sadmc_num_moves)
  SADMC_FN(sadmc_num_accepted_moves)
  SADMC_FN(sadmc_accepted_moves_range)
  SADMC_FN(sadmc_num_halted)
  SADMC_FN(sadmc_get_walker)
  SADMC_FN(sadmc_get_bins)
  SADMC_FN(sadmc_system_len)
  SADMC_FN(sadmc_get_system)
  SADMC_FN(sadmc_set_system)
  SADMC_FN(sadmc_set_walker_bins)
  SADMC_FN(sadmc_resume)
  SADMC_FN(sadmc_window)
  SADMC_FN(sadmc_cell_box)
  SADMC_FN(sadmc_launch_count)
  SADMC_FN(sadmc_sys_energy)
  SADMC_FN(sadmc_sys_compute_energy)
  SADMC_FN(sadmc_sys_plan_move)
  SADMC_FN(sadmc_sys_confirm)
  SADMC_FN(sadmc_sys_verify_energy)
#undef SADMC_FN

  static Api& get(const std::string& path_hint = "") {
    static Api api;
    if (api.handle) return api;
    std::vector<std::string> candidates;
    if (const char* env = getenv("SADMC_GPU_LIB")) candidates.push_back(env);
    if (!path_hint.empty()) candidates.push_back(path_hint);
    candidates.push_back("libsadmc_gpu.so");
    std::string errs;
    for (auto& c : candidates) {
      api.handle = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL);
      if (api.handle) break;
      errs += std::string(dlerror()) + "; ";
    }
    if (!api.handle) throw EngineError(SADMC_ERR_CUDA, "cannot load libsadmc_gpu.so (there is no CPU fallback): " + errs);
#define SADMC_FN(name)                                                     \
  api.name = reinterpret_cast<decltype(api.name)>(dlsym(api.handle, #name)); \
  if (!api.name) throw EngineError(SADMC_ERR_INVALID, "libsadmc_gpu.so does not export " #name);
    SADMC_FN(sadmc_create)
    SADMC_FN(sadmc_destroy)
    SADMC_FN(sadmc_last_error)
    SADMC_FN(sadmc_abi_version)
    SADMC_FN(sadmc_reference_system)
    SADMC_FN(sadmc_start)
    SADMC_FN(sadmc_run)
    SADMC_FN(sadmc_num_moves)
    SADMC_FN(sadmc_num_accepted_moves)
    SADMC_FN(sadmc_accepted_moves_range)
    SADMC_FN(sadmc_num_halted)
    SADMC_FN(sadmc_get_walker)
    SADMC_FN(sadmc_get_bins)
    SADMC_FN(sadmc_system_len)
    SADMC_FN(sadmc_get_system)
    SADMC_FN(sadmc_set_system)
    SADMC_FN(sadmc_set_walker_bins)
    SADMC_FN(sadmc_resume)
    SADMC_FN(sadmc_window)
    SADMC_FN(sadmc_cell_box)
    SADMC_FN(sadmc_launch_count)
    SADMC_FN(sadmc_sys_energy)
    SADMC_FN(sadmc_sys_compute_energy)
    SADMC_FN(sadmc_sys_plan_move)
    SADMC_FN(sadmc_sys_confirm)
    SADMC_FN(sadmc_sys_verify_energy)
#undef SADMC_FN
    if (api.sadmc_abi_version() != SADMC_ABI_VERSION) throw EngineError(SADMC_ERR_INVALID, "libsadmc_gpu.so has another ABI version");
    return api;
  }
};

// per-bin vectors of one walker in the reference's index order (Bins, energy.rs:146-163, + round-trip vectors 203-205)
struct WalkerBins {
  std::vector<uint64_t> histogram, t_found, round_trips, wl_hist, extra_count;
  std::vector<double> lnw, energy_total, energy_squared_total, extra_total;
  std::vector<uint8_t> have_visited;
  void resize(size_t n) {
    histogram.assign(n, 0);
    t_found.assign(n, 0);
    round_trips.assign(n, 0);
    wl_hist.assign(n, 0);
    extra_count.assign(n, 0);
    lnw.assign(n, 0.0);
    energy_total.assign(n, 0.0);
    energy_squared_total.assign(n, 0.0);
    extra_total.assign(n, 0.0);
    have_visited.assign(n, 0);
  }
};

class GpuEnergyMC {
  Api& api;
  sadmc_engine* h = nullptr;
  void check(int rc) const {
    if (rc != 0) throw EngineError(rc, api.sadmc_last_error());
  }

 public:
  sadmc_config cfg;
  size_t system_len = 0;

  GpuEnergyMC(const sadmc_config& c, const std::string& lib_hint = "") : api(Api::get(lib_hint)), cfg(c) { // from_params, energy.rs:830-898
    check(api.sadmc_create(&cfg, &h));
    check(api.sadmc_system_len(h, &system_len));
  }
  GpuEnergyMC(const GpuEnergyMC&) = delete;
  GpuEnergyMC& operator=(const GpuEnergyMC&) = delete;
  ~GpuEnergyMC() {
    if (h) api.sadmc_destroy(h);
  }
  uint32_t n_walkers() const { return cfg.n_walkers; }

  void run(uint64_t n_moves) { check(api.sadmc_run(h, n_moves)); } // n_moves x move_once (energy.rs:904-974) for every walker
  void start() { check(api.sadmc_start(h)); }
  void resume(uint64_t moves) { check(api.sadmc_resume(h, moves)); }
  uint64_t num_moves() const { // MonteCarlo::num_moves, energy.rs:981
    uint64_t m = 0;
    check(api.sadmc_num_moves(h, &m));
    return m;
  }
  uint64_t num_accepted_moves() const { // energy.rs:984, summed over walkers
    uint64_t a = 0;
    check(api.sadmc_num_accepted_moves(h, &a));
    return a;
  }
  uint64_t min_accepted_moves() const { // the slowest walker's count: every walker is one reference run
    uint64_t lo = 0, hi = 0;
    check(api.sadmc_accepted_moves_range(h, &lo, &hi));
    return lo;
  }
  void num_halted(uint64_t* left_window, uint64_t* failed_verify) const { check(api.sadmc_num_halted(h, left_window, failed_verify)); }
  uint64_t launch_count() const {
    uint64_t n = 0;
    check(api.sadmc_launch_count(h, &n));
    return n;
  }
  sadmc_walker_state walker(uint32_t w) const {
    sadmc_walker_state s;
    check(api.sadmc_get_walker(h, w, &s));
    return s;
  }
  WalkerBins bins(uint32_t w) const {
    const sadmc_walker_state s = walker(w);
    WalkerBins b;
    b.resize(s.bins_len);
    check(api.sadmc_get_bins(h, w, s.bins_len, b.histogram.data(), b.t_found.data(), b.lnw.data(), b.energy_total.data(),
                             b.energy_squared_total.data(), b.round_trips.data(), b.have_visited.data(), b.wl_hist.data(), b.extra_total.data(),
                             b.extra_count.data()));
    return b;
  }
  void set_walker_bins(uint32_t w, const sadmc_walker_state& s, const WalkerBins& b) {
    check(api.sadmc_set_walker_bins(h, w, &s, b.histogram.data(), b.t_found.data(), b.lnw.data(), b.energy_total.data(),
                                    b.energy_squared_total.data(), b.round_trips.data(), b.have_visited.data(), b.wl_hist.data(),
                                    b.extra_total.data(), b.extra_count.data()));
  }
  std::vector<double> system(uint32_t w) const { // MonteCarlo::system(), as the ABI's f64 image
    std::vector<double> img(system_len);
    check(api.sadmc_get_system(h, w, img.data(), img.size()));
    return img;
  }
  void set_system(uint32_t w, const std::vector<double>& img) { check(api.sadmc_set_system(h, w, img.data(), img.size())); }
  void cell_box(double box[3], double* r_cutoff) const { check(api.sadmc_cell_box(h, box, r_cutoff)); }

  // trait-shaped shims, system/mod.rs:54-120
  double energy(uint32_t w) const {
    double e = 0;
    check(api.sadmc_sys_energy(h, w, &e));
    return e;
  }
  double compute_energy(uint32_t w) const {
    double e = 0;
    check(api.sadmc_sys_compute_energy(h, w, &e));
    return e;
  }
  bool plan_move(uint32_t w, double mean_distance, double* e_new) { // false == None
    int some = 0;
    check(api.sadmc_sys_plan_move(h, w, mean_distance, &some, e_new));
    return some != 0;
  }
  void confirm(uint32_t w) { check(api.sadmc_sys_confirm(h, w)); }
  bool verify_energy(uint32_t w) const {
    const int rc = api.sadmc_sys_verify_energy(h, w);
    if (rc == SADMC_ERR_VERIFY) return false;
    check(rc);
    return true;
  }
};

} // namespace sadmc_host
