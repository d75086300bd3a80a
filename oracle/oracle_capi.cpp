// oracle_capi.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// extern "C" face of the CPU oracle, loaded by tests/ (ctypes), by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
// Nothing under sad_monte_carlo_b200/ may link or load this library.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>

#include "../include/sadmc_gpu.h" // config / state structs only (shared vocabulary with the engine)
#include "oracle_binning.hpp"
#include "oracle_mc.hpp"
#include "oracle_replicas.hpp"
#include "oracle_tempering.hpp"

namespace oracle {

// statrs 0.7 `erf_inv` (erfinv.rs:66) is an un-vendored dependency whose
// coefficient tables are not reproducible from memory; restated instead as
// Newton iterations on libm erf from the Winitzki starting guess -- accurate to
// a few ulp, which is inside the floating-point tier (1e-12) of this system.
double erf_inv(double x) {
  if (x <= -1.0) return -INFINITY;
  if (x >= 1.0) return INFINITY;
  if (x == 0.0) return 0.0;
  const double a = 0.147;
  const double ln1mx2 = std::log(1.0 - x * x);
  const double t = 2.0 / (M_PI * a) + 0.5 * ln1mx2;
  double y = std::sqrt(std::sqrt(t * t - ln1mx2 / a) - t);
  if (x < 0) y = -y;
  for (int it = 0; it < 4; it++) {
    const double err = std::erf(y) - x;
    const double d = 2.0 / std::sqrt(M_PI) * std::exp(-y * y);
    // Halley step
    y -= err / (d + y * err);
  }
  return y;
}

// Test-only: pair sum of Lj::move_atom in the order of the kernel's
// lanes_per_walker = G layout (atom a -> lane a % G, slot a / G; per-lane
// sequential over slots, then an xor butterfly), with the kernel's pair
// arithmetic (sad_monte_carlo_b200/csrc/sys_lj.cuh: fma-contracted r^2).
double Lj::move_atom_tree(size_t which, Vec3 r) const {
  const int G = tree_lanes;
  const Vec3 from = positions[which];
  std::vector<double> lane(G, 0.0);
  auto r2 = [](const Vec3& a, const Vec3& b) {
    const double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return std::fma(dz, dz, std::fma(dy, dy, dx * dx));
  };
  auto pot = [](double rr) {
    const double s = 1.0 / rr;
    const double s3 = s * s * s;
    return std::fma(s3, s3, -s3);
  };
  for (size_t k = 0; k < positions.size(); k++) {
    if (k == which) continue;
    lane[k % G] += pot(r2(positions[k], r)) - pot(r2(positions[k], from));
  }
  for (int off = G / 2; off >= 1; off >>= 1) {
    std::vector<double> nxt(G);
    for (int l = 0; l < G; l++) nxt[l] = lane[l] + lane[l ^ off];
    lane.swap(nxt);
  }
  return E + 4.0 * lane[0];
}

// Test-only: compute_energy in the kernel's order (sys_lj.cuh compute_energy).
double Lj::compute_energy_tree() const {
  const int G = tree_lanes;
  const size_t n = positions.size();
  std::vector<double> lane(G, 0.0);
  for (size_t b = 1; b < n; b++)
    for (size_t a = 0; a < b; a++) { // per lane: b ascending, then slot (a / G) ascending
      const double dx = positions[a].x - positions[b].x, dy = positions[a].y - positions[b].y, dz = positions[a].z - positions[b].z;
      const double rr = std::fma(dz, dz, std::fma(dy, dy, dx * dx));
      const double s = 1.0 / rr;
      const double s3 = s * s * s;
      lane[a % G] += std::fma(s3, s3, -s3);
    }
  for (int off = G / 2; off >= 1; off >>= 1) {
    std::vector<double> nxt(G);
    for (int l = 0; l < G; l++) nxt[l] = lane[l] + lane[l ^ off];
    lane.swap(nxt);
  }
  return 4.0 * lane[0];
}

static thread_local std::string g_err;

static Vec3 box_from_config(const sadmc_config& c, bool sw) {
  if (c.cell_width[0] > 0) return Vec3(std::fabs(c.cell_width[0]), std::fabs(c.cell_width[1]), std::fabs(c.cell_width[2]));
  double vol;
  if (sw)
    vol = (double)c.N * (M_PI * 1.0 * 1.0 * 1.0 / 6.0) / c.filling_fraction; // optsquare.rs:365-367
  else
    vol = (double)c.N / c.reduced_density; // wca.rs:399-401
  const double w = std::cbrt(vol);         // optcell.rs:47-50
  return Vec3(w, w, w);
}

static std::unique_ptr<System> make_system(const sadmc_config& c, uint64_t attempts_override) {
  const bool ref = c.init_mode == SADMC_INIT_REFERENCE;
  switch (c.system) {
    case SADMC_SYS_ISING: return std::unique_ptr<System>(new Ising(c.N));
    case SADMC_SYS_LJ:
      if (ref) return std::unique_ptr<System>(attempts_override ? new Lj(c.N, c.lj_radius, attempts_override, 100000000ull) : new Lj(c.N, c.lj_radius));
      {
        Lj* lj = new Lj(c.N, c.lj_radius, Lj::Empty());
        if (c.flags & SADMC_FLAG_SUM_TREE) lj->tree_lanes = c.lanes_per_walker ? c.lanes_per_walker : 8;
        return std::unique_ptr<System>(lj);
      }
    case SADMC_SYS_WCA:
      if (ref) return std::unique_ptr<System>(new Wca(Wca::from_n(c.N, box_from_config(c, false), attempts_override ? attempts_override : ~0ull)));
      {
        Wca* w = new Wca(Cell(box_from_config(c, false), Wca::r_cutoff()));
        w->cell.positions.assign(c.N, Vec3());
        w->cell.update_caches();
        return std::unique_ptr<System>(w);
      }
    case SADMC_SYS_SW:
      if (ref) return std::unique_ptr<System>(new SquareWell(SquareWell::from_n(c.N, box_from_config(c, true), c.sw_well_width)));
      {
        SquareWell* s = new SquareWell(Cell(box_from_config(c, true), c.sw_well_width));
        s->cell.positions.assign(c.N, Vec3());
        s->cell.update_caches();
        return std::unique_ptr<System>(s);
      }
    case SADMC_SYS_FAKE:
      return std::unique_ptr<System>(new Fake((Fake::Kind)c.fake_function, c.N, c.fake_a, c.fake_b, c.fake_e1, c.fake_e2, c.fake_sigma));
    case SADMC_SYS_TWO_WELLS: return std::unique_ptr<System>(new TwoWells(c.N, c.tw_h2_to_h1, c.tw_barrier_over_h1, c.tw_r2));
    case SADMC_SYS_FAKE_ERFINV: return std::unique_ptr<System>(new ErfInv(c.N, c.erfinv_mean_energy));
  }
  throw std::invalid_argument("unknown system kind");
}

static MCParams mc_params(const sadmc_config& c, uint32_t walker) {
  MCParams p;
  p.method = c.method;
  p.sad_min_T = c.sad_min_T;
  p.samc_t0 = c.samc_t0;
  p.wl_min_gamma = c.wl_min_gamma;
  p.canonical_T = c.canonical_T;
  p.seed = c.seed + walker;
  p.energy_bin = c.energy_bin;
  p.min_allowed_energy = c.min_allowed_energy;
  p.max_allowed_energy = c.max_allowed_energy;
  p.acceptance_rate_plan = c.move_plan == SADMC_MOVE_ACCEPTANCE_RATE;
  p.move_value = c.move_value;
  p.randomize_first = c.init_mode == SADMC_INIT_RANDOMIZE;
  return p;
}

} // namespace oracle

using namespace oracle;

struct oracle_mc {
  std::unique_ptr<EnergyMC> mc;
  sadmc_config cfg;
};

extern "C" {

const char* oracle_last_error(void) { return g_err.c_str(); }

void oracle_set_math_mode(int use_libm) { MathMode::use_libm() = use_libm != 0; }

// Walker `walker` (global index) of the configuration `cfg`: the reference process
// `histogram <flags> --seed (cfg->seed + walker)`.  If `system_state` is non-NULL the
// system is overwritten with it before from_params runs (SADMC_INIT_EXTERNAL).
// `attempts_override` != 0 shrinks the constructor's attempt count (tests only).
oracle_mc* oracle_create(const sadmc_config* cfg, uint32_t walker, const double* system_state, size_t n_state,
                         uint64_t attempts_override) {
  try {
    std::unique_ptr<System> sys = make_system(*cfg, attempts_override);
    if (system_state) sys->set_state(std::vector<double>(system_state, system_state + n_state));
    oracle_mc* o = new oracle_mc;
    o->cfg = *cfg;
    MCParams p = mc_params(*cfg, walker);
    if (system_state) p.randomize_first = false;
    o->mc.reset(new EnergyMC(p, std::move(sys)));
    return o;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return nullptr;
  }
}
void oracle_destroy(oracle_mc* o) { delete o; }

int oracle_set_lj_tree_lanes(oracle_mc* o, int lanes) {
  Lj* lj = dynamic_cast<Lj*>(o->mc->system.get());
  if (!lj) return -1;
  lj->tree_lanes = lanes;
  return 0;
}

int oracle_run(oracle_mc* o, uint64_t n) {
  try {
    for (uint64_t k = 0; k < n; k++) o->mc->move_once();
    return 0;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return -1;
  }
}

int oracle_get_walker(oracle_mc* o, sadmc_walker_state* s) {
  const EnergyMC& m = *o->mc;
  std::memset(s, 0, sizeof(*s));
  s->moves = m.moves;
  s->accepted_moves = m.accepted_moves;
  s->acceptance_rate = m.acceptance_rate;
  s->translation_scale = m.translation_scale;
  s->rng_s0 = m.rng.s0;
  s->rng_s1 = m.rng.s1;
  s->energy = m.system->energy();
  s->bins_min = m.bins.min;
  s->bins_width = m.bins.width;
  s->bins_len = (uint32_t)m.bins.lnw.size();
  s->window_first = 0;
  const Method& me = m.method;
  s->method = me.kind == M_WL ? (me.inv_t ? SADMC_METHOD_INV_T_WL : SADMC_METHOD_WL) : me.kind;
  s->status = m.verify_failures ? SADMC_ERR_VERIFY : 0;
  s->too_lo = me.too_lo;
  s->too_hi = me.too_hi;
  s->latest_parameter = me.latest_parameter;
  s->tL = me.tL;
  s->tF = me.tF;
  s->num_states = me.num_states;
  s->highest_hist = me.highest_hist;
  s->samc_t0 = me.t0;
  s->wl_gamma = me.gamma;
  s->wl_num_states = me.wl_num_states;
  s->wl_min_energy = me.min_energy;
  s->wl_lowest_hist = me.wl_lowest_hist;
  s->wl_highest_hist = me.wl_highest_hist;
  s->wl_total_hist = me.wl_total_hist;
  s->wl_hist_len = (uint32_t)me.hist.size();
  s->wl_inv_t = me.inv_t;
  s->max_S = m.max_S;
  s->max_S_index = (uint32_t)m.max_S_index;
  return 0;
}

int oracle_get_bins(oracle_mc* o, uint32_t cap, uint64_t* histogram, uint64_t* t_found, double* lnw, double* energy_total,
                    double* energy_squared_total, uint64_t* round_trips, uint8_t* have_visited, uint64_t* wl_hist,
                    double* extra_total, uint64_t* extra_count) {
  const EnergyMC& m = *o->mc;
  const size_t n = m.bins.lnw.size();
  if (cap < n) {
    g_err = "capacity too small";
    return -1;
  }
  for (size_t i = 0; i < n; i++) {
    if (histogram) histogram[i] = m.bins.histogram[i];
    if (t_found) t_found[i] = m.bins.t_found[i];
    if (lnw) lnw[i] = m.bins.lnw[i];
    if (energy_total) energy_total[i] = m.bins.energy_total[i];
    if (energy_squared_total) energy_squared_total[i] = m.bins.energy_squared_total[i];
    if (round_trips) round_trips[i] = m.round_trips[i];
    if (have_visited) have_visited[i] = m.have_visited_since_maxentropy[i];
    if (wl_hist) wl_hist[i] = i < m.method.hist.size() ? m.method.hist[i] : 0;
    if (extra_total) extra_total[i] = 0;
    if (extra_count) extra_count[i] = 0;
  }
  if (!m.bins.extra.empty()) {
    const BinCounts& b = m.bins.extra.begin()->second;
    for (size_t i = 0; i < n && i < b.total.size(); i++) {
      if (extra_total) extra_total[i] = b.total[i];
      if (extra_count) extra_count[i] = b.count[i];
    }
  }
  return 0;
}

size_t oracle_system_len(oracle_mc* o) { return o->mc->system->get_state().size(); }
int oracle_get_system(oracle_mc* o, double* buf, size_t n) {
  const std::vector<double> s = o->mc->system->get_state();
  if (n < s.size()) return -1;
  std::memcpy(buf, s.data(), s.size() * sizeof(double));
  return 0;
}
int oracle_set_system(oracle_mc* o, const double* buf, size_t n) {
  o->mc->system->set_state(std::vector<double>(buf, buf + n));
  return 0;
}
int oracle_set_rng(oracle_mc* o, uint64_t s0, uint64_t s1) {
  o->mc->rng.s0 = s0;
  o->mc->rng.s1 = s1;
  return 0;
}

// trait-shaped shims, src/system/mod.rs:54-120
double oracle_sys_energy(oracle_mc* o) { return o->mc->system->energy(); }
double oracle_sys_compute_energy(oracle_mc* o) { return o->mc->system->compute_energy(); }
int oracle_sys_plan_move(oracle_mc* o, double mean_distance, int* some, double* e_new) {
  double e = 0;
  *some = o->mc->system->plan_move(o->mc->rng, mean_distance, &e) ? 1 : 0;
  *e_new = e;
  return 0;
}
void oracle_sys_confirm(oracle_mc* o) { o->mc->system->confirm(); }
// System::randomize (system/mod.rs:59) driven by the walker's own generator; returns 0 and the new energy, or
// SADMC_ERR_UNSUPPORTED where the reference is todo!() / not restated (optsquare.rs:202-204, two-wells)
int oracle_sys_randomize(oracle_mc* o, double* energy) {
  try {
    *energy = o->mc->system->randomize(o->mc->rng);
    return 0;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return SADMC_ERR_UNSUPPORTED;
  }
}
int oracle_sys_verify_energy(oracle_mc* o) { return o->mc->system->verify_energy() ? 0 : SADMC_ERR_VERIFY; }
// SquareWell only: the reference's slow all-image recount (optsquare.rs:108-152)
double oracle_sw_compute_energy_slowly(oracle_mc* o) {
  SquareWell* s = dynamic_cast<SquareWell*>(o->mc->system.get());
  return s ? s->compute_energy_slowly() : NAN;
}

// ---- the `binning` binary: energy_binning.rs over binning::histogram / binning::linear (oracle_binning.hpp) ----
struct oracle_bmc {
  std::unique_ptr<binning::EnergyMC> mc;        // BinningParams::Histogram
  std::unique_ptr<binning::EnergyMCLinear> mcl; // BinningParams::Linear (cfg->flags & SADMC_FLAG_BINNING_LINEAR)
};
oracle_bmc* oracle_binning_create(const sadmc_config* cfg, uint32_t walker, const double* system_state, size_t n_state,
                                  uint64_t attempts_override) {
  try {
    std::unique_ptr<System> sys = make_system(*cfg, attempts_override);
    if (system_state) sys->set_state(std::vector<double>(system_state, system_state + n_state));
    MCParams p = mc_params(*cfg, walker);
    if (system_state) p.randomize_first = false;
    oracle_bmc* o = new oracle_bmc;
    const double hr = cfg->high_resolution_de > 0 ? cfg->high_resolution_de : NAN;
    if (cfg->flags & SADMC_FLAG_BINNING_LINEAR)
      o->mcl.reset(new binning::EnergyMCLinear(p, std::move(sys), hr));
    else
      o->mc.reset(new binning::EnergyMC(p, std::move(sys), hr));
    return o;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return nullptr;
  }
}
void oracle_binning_destroy(oracle_bmc* o) { delete o; }
int oracle_binning_reference_test(void) { return binning::reference_test_binning(); }
int oracle_binning_reference_test_linear(void) { return binning::reference_test_linear(); }
int oracle_binning_run(oracle_bmc* o, uint64_t n) {
  try {
    if (o->mcl)
      for (uint64_t k = 0; k < n; k++) o->mcl->move_once();
    else
      for (uint64_t k = 0; k < n; k++) o->mc->move_once();
    return 0;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return -1;
  }
}
extern "C++" {
template <class B>
static const auto* find_extra(const B& b, const char* name) {
  auto it = b.extra.find(name);
  return it == b.extra.end() ? nullptr : &it->second;
}
}
extern "C++" {
template <class MC>
static void fill_binning_state(const MC& m, sadmc_binning_state* s) {
  std::memset(s, 0, sizeof(*s));
  s->moves = m.moves;
  s->accepted_moves = m.accepted_moves;
  s->acceptance_rate = m.acceptance_rate;
  s->translation_scale = m.translation_scale;
  s->rng_s0 = m.rng.s0;
  s->rng_s1 = m.rng.s1;
  s->energy = m.system->energy();
  s->bins_min = m.bins.min;
  s->bins_width = m.bins.width;
  s->bins_min_e = m.bins.min_e;
  s->bins_max_e = m.bins.max_e;
  s->bins_len = (uint32_t)m.bins.lnw.total.size();
  const binning::Method& me = m.method;
  s->method = me.kind == binning::B_WL ? (me.inv_t ? SADMC_METHOD_INV_T_WL : SADMC_METHOD_WL) : me.kind;
  s->status = m.verify_failures ? SADMC_ERR_VERIFY : 0;
  s->too_lo = me.too_lo;
  s->too_hi = me.too_hi;
  s->latest_parameter = me.latest_parameter;
  s->tF = me.tF;
  s->tL = me.tL;
  s->num_states = me.num_states;
  s->samc_t0 = me.t0;
  s->wl_gamma = me.gamma;
  s->wl_inv_t = me.inv_t;
  s->lnw_max_count = (uint64_t)m.bins.lnw.max_count;
  s->lnw_max_count_f64 = (double)m.bins.lnw.max_count;
  s->lnw_total_count = m.bins.lnw.total_count;
  if (const auto* t = find_extra(m.bins, "t_found")) s->t_found_max_total = t->max_total;
  if (const auto* h = find_extra(m.bins, "hist")) {
    s->hist_min_count = (uint64_t)h->min_count;
    s->hist_min_count_f64 = (double)h->min_count;
    s->hist_total_count = h->total_count;
  }
}
}
int oracle_binning_get_walker(oracle_bmc* o, sadmc_binning_state* s) {
  if (o->mcl)
    fill_binning_state(*o->mcl, s);
  else
    fill_binning_state(*o->mc, s);
  return 0;
}
// counts as f64 (exact for the histogram variant below 2^53; the linear variant's counts ARE f64, linear.rs:22)
extern "C++" {
template <class B>
static int fill_binning_bins(const B& b, uint32_t cap, double* lnw_total, double* lnw_count, double* energy_total, double* energy_count,
                             double* t_found_total, double* t_found_count, double* hist_count, double* extra_total, double* extra_count) {
  const size_t n = b.lnw.total.size();
  if (cap < n) {
    g_err = "capacity too small";
    return -1;
  }
  const auto* en = find_extra(b, "energy");
  const auto* tf = find_extra(b, "t_found");
  const auto* hi = find_extra(b, "hist");
  decltype(en) sx = nullptr; // the system's own data_to_collect key
  for (const auto& kv : b.extra)
    if (kv.first != "energy" && kv.first != "t_found" && kv.first != "hist") sx = &kv.second;
  for (size_t i = 0; i < n; i++) {
    if (lnw_total) lnw_total[i] = b.lnw.total[i];
    if (lnw_count) lnw_count[i] = (double)b.lnw.count[i];
    if (energy_total) energy_total[i] = en ? en->total[i] : 0.0;
    if (energy_count) energy_count[i] = en ? (double)en->count[i] : 0.0;
    if (t_found_total) t_found_total[i] = tf ? tf->total[i] : 0.0;
    if (t_found_count) t_found_count[i] = tf ? (double)tf->count[i] : 0.0;
    if (hist_count) hist_count[i] = hi ? (double)hi->count[i] : 0.0;
    if (extra_total) extra_total[i] = sx ? sx->total[i] : 0.0;
    if (extra_count) extra_count[i] = sx ? (double)sx->count[i] : 0.0;
  }
  return 0;
}
}
int oracle_binning_get_bins_f64(oracle_bmc* o, uint32_t cap, double* lnw_total, double* lnw_count, double* energy_total, double* energy_count,
                                double* t_found_total, double* t_found_count, double* hist_count, double* extra_total, double* extra_count) {
  if (o->mcl)
    return fill_binning_bins(o->mcl->bins, cap, lnw_total, lnw_count, energy_total, energy_count, t_found_total, t_found_count, hist_count, extra_total,
                             extra_count);
  return fill_binning_bins(o->mc->bins, cap, lnw_total, lnw_count, energy_total, energy_count, t_found_total, t_found_count, hist_count, extra_total,
                           extra_count);
}
int oracle_binning_get_bins(oracle_bmc* o, uint32_t cap, double* lnw_total, uint64_t* lnw_count, double* energy_total, uint64_t* energy_count,
                            double* t_found_total, uint64_t* t_found_count, uint64_t* hist_count, double* extra_total, uint64_t* extra_count) {
  if (o->mcl) {
    g_err = "the linear variant's counts are f64: use oracle_binning_get_bins_f64";
    return -1;
  }
  const binning::Bins& b = o->mc->bins;
  const size_t n = b.lnw.total.size();
  std::vector<double> c0(n), c1(n), c2(n), c3(n), c4(n);
  const int rc = fill_binning_bins(b, cap, lnw_total, c0.data(), energy_total, c1.data(), t_found_total, c2.data(), c3.data(), extra_total, c4.data());
  if (rc) return rc;
  for (size_t i = 0; i < n; i++) {
    if (lnw_count) lnw_count[i] = (uint64_t)c0[i];
    if (energy_count) energy_count[i] = (uint64_t)c1[i];
    if (t_found_count) t_found_count[i] = (uint64_t)c2[i];
    if (hist_count) hist_count[i] = (uint64_t)c3[i];
    if (extra_count) extra_count[i] = (uint64_t)c4[i];
  }
  return 0;
}
// the lazily maintained aggregates of one BinCounts (histogram.rs:12-32); name "" = bins.lnw.  out: min_total, max_total,
// e_max_total, min_count, max_count, e_max_count, total_count (counts as doubles)
extern "C++" {
template <class B>
static int fill_aggregates(const B& bins, const char* name, double* out) {
  const auto* c = name[0] ? find_extra(bins, name) : &bins.lnw;
  if (!c) return -1;
  out[0] = c->min_total;
  out[1] = c->max_total;
  out[2] = c->e_max_total;
  out[3] = (double)c->min_count;
  out[4] = (double)c->max_count;
  out[5] = c->e_max_count;
  out[6] = (double)c->total_count;
  return 0;
}
}
int oracle_binning_get_aggregates(oracle_bmc* o, const char* name, double* out) {
  return o->mcl ? fill_aggregates(o->mcl->bins, name, out) : fill_aggregates(o->mc->bins, name, out);
}
// the optional high-resolution histogram (energy_binning.rs:124-125): min, number of bins, counts
int oracle_binning_get_high_resolution(oracle_bmc* o, uint32_t cap, double* bins_min, uint32_t* len, uint64_t* count) {
  const bool has = o->mcl ? o->mcl->has_high_resolution : o->mc->has_high_resolution;
  if (!has) return -1;
  const binning::Bins& b = o->mcl ? o->mcl->high_resolution : o->mc->high_resolution;
  *bins_min = b.min;
  *len = (uint32_t)b.lnw.count.size();
  if (count) {
    if (cap < *len) return -2;
    for (size_t i = 0; i < b.lnw.count.size(); i++) count[i] = b.lnw.count[i];
  }
  return 0;
}
static System* binning_system(oracle_bmc* o) { return o->mcl ? o->mcl->system.get() : o->mc->system.get(); }
size_t oracle_binning_system_len(oracle_bmc* o) { return binning_system(o)->get_state().size(); }
int oracle_binning_get_system(oracle_bmc* o, double* buf, size_t n) {
  const std::vector<double> s = binning_system(o)->get_state();
  if (n < s.size()) return -1;
  std::memcpy(buf, s.data(), s.size() * sizeof(double));
  return 0;
}

// ---- the `tempering` binary (oracle_tempering.hpp) ----
struct oracle_tmc {
  std::unique_ptr<tempering::MC> mc;
};
static uint64_t min_moves_to_randomize(const sadmc_config& c) {
  switch (c.system) {
    case SADMC_SYS_ISING: return (uint64_t)c.N * c.N; // ising.rs:86-88
    case SADMC_SYS_FAKE: return c.fake_function == SADMC_FAKE_LINEAR ? 1 : (c.fake_function == SADMC_FAKE_QUADRATIC ? c.N : 3);
    default: return c.N;
  }
}
// simulation `sim` = the reference process `tempering <flags> --seed (cfg->seed + sim)`; system_state (optional):
// the image every replica starts from instead of the reference constructor's
oracle_tmc* oracle_tempering_create(const sadmc_config* cfg, uint32_t sim, const double* T, uint32_t n_T, uint64_t canonical_steps,
                                    const double* system_state, size_t n_state, uint64_t attempts_override) {
  try {
    std::vector<double> img;
    if (system_state) {
      img.assign(system_state, system_state + n_state);
    } else {
      sadmc_config c0 = *cfg;
      c0.init_mode = SADMC_INIT_REFERENCE;
      img = make_system(c0, attempts_override)->get_state();
    }
    std::vector<std::unique_ptr<System>> systems;
    for (uint32_t r = 0; r < n_T; r++) { // system.clone()
      sadmc_config c1 = *cfg;
      c1.init_mode = SADMC_INIT_EXTERNAL;
      std::unique_ptr<System> s = make_system(c1, attempts_override);
      s->set_state(img);
      systems.push_back(std::move(s));
    }
    oracle_tmc* o = new oracle_tmc;
    o->mc.reset(new tempering::MC(cfg->seed + sim, std::vector<double>(T, T + n_T), canonical_steps, min_moves_to_randomize(*cfg), std::move(systems)));
    return o;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return nullptr;
  }
}
void oracle_tempering_destroy(oracle_tmc* o) { delete o; }
int oracle_tempering_run(oracle_tmc* o, uint64_t n_rounds) {
  try {
    for (uint64_t k = 0; k < n_rounds; k++) o->mc->run_once();
    return 0;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return -1;
  }
}
uint64_t oracle_tempering_num_moves(oracle_tmc* o) { return o->mc->moves; }
void oracle_tempering_set_translation_scales(oracle_tmc* o, const double* scale) {
  for (size_t r = 0; r < o->mc->replicas.size(); r++) o->mc->replicas[r].translation_scale = scale[r];
}
void oracle_tempering_get_rng(oracle_tmc* o, uint64_t* s) {
  s[0] = o->mc->rng.s0;
  s[1] = o->mc->rng.s1;
}
int oracle_tempering_get_replicas(oracle_tmc* o, sadmc_replica_state* out) {
  for (size_t r = 0; r < o->mc->replicas.size(); r++) {
    const tempering::Replica& q = o->mc->replicas[r];
    out[r].T = q.T;
    out[r].rejected_count = q.rejected_count;
    out[r].accepted_count = q.accepted_count;
    out[r].rejected_swap_count = q.rejected_swap_count;
    out[r].accepted_swap_count = q.accepted_swap_count;
    out[r].ignored_count = q.ignored_count;
    out[r].total_energy = q.total_energy;
    out[r].total_energy_squared = q.total_energy_squared;
    out[r].translation_scale = q.translation_scale;
    out[r].rng_s0 = q.rng.s0;
    out[r].rng_s1 = q.rng.s1;
    out[r].energy = q.energy();
  }
  return 0;
}
size_t oracle_tempering_system_len(oracle_tmc* o) { return o->mc->replicas[0].system->get_state().size(); }
int oracle_tempering_get_system(oracle_tmc* o, uint32_t replica, double* buf, size_t n) {
  const std::vector<double> s = o->mc->replicas[replica].system->get_state();
  if (n < s.size()) return -1;
  std::memcpy(buf, s.data(), s.size() * sizeof(double));
  return 0;
}
// xoroshiro128+ jump applied to a state (test probe)
void oracle_rng_jump(uint64_t* s) {
  Rng g;
  g.s0 = s[0];
  g.s1 = s[1];
  tempering::jump(g);
  s[0] = g.s0;
  s[1] = g.s1;
}

// ---- the `replicas` binary (oracle_replicas.hpp) ----
struct oracle_zmc {
  sadmc_config cfg;
  std::unique_ptr<replicas::MC> mc;
};
static std::unique_ptr<System> clone_system(const System& s, const void* ctx) {
  sadmc_config c = *static_cast<const sadmc_config*>(ctx);
  c.init_mode = SADMC_INIT_EXTERNAL;
  std::unique_ptr<System> n = make_system(c, 0);
  n->set_state(s.get_state());
  return n;
}
oracle_zmc* oracle_replicas_create(const sadmc_config* cfg, uint32_t sim, double min_T, uint64_t indep, uint32_t max_init, uint64_t attempts_override) {
  try {
    oracle_zmc* o = new oracle_zmc;
    o->cfg = *cfg;
    sadmc_config c0 = *cfg;
    c0.init_mode = SADMC_INIT_REFERENCE;
    replicas::SystemTraits tr;
    tr.min_moves_to_randomize = min_moves_to_randomize(*cfg);
    switch (cfg->system) { // MovableSystem::max_size, System::dimensionality
      case SADMC_SYS_LJ: tr.max_size = cfg->lj_radius; tr.dimensionality = 3ull * cfg->N; break;
      case SADMC_SYS_WCA: {
        const Vec3 b = box_from_config(*cfg, false);
        tr.max_size = std::sqrt(b.x * b.x + b.y * b.y + b.z * b.z);
        tr.dimensionality = 3ull * cfg->N;
        break;
      }
      case SADMC_SYS_ISING: tr.max_size = 0.5; tr.dimensionality = (uint64_t)cfg->N * cfg->N; break;
      case SADMC_SYS_FAKE: tr.max_size = 0.5; tr.dimensionality = tr.min_moves_to_randomize; break;
      case SADMC_SYS_FAKE_ERFINV: tr.max_size = 0.5; tr.dimensionality = 3ull * cfg->N; break;
      default: throw std::invalid_argument("replicas: this system has no System::randomize restated");
    }
    o->mc.reset(new replicas::MC(cfg->seed + sim, min_T, indep, tr, make_system(c0, attempts_override), clone_system, &o->cfg,
                                 max_init ? (size_t)max_init : ((size_t)1 << 15)));
    return o;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return nullptr;
  }
}
void oracle_replicas_destroy(oracle_zmc* o) { delete o; }
int oracle_replicas_run(oracle_zmc* o, uint64_t n_rounds) {
  try {
    for (uint64_t k = 0; k < n_rounds; k++) o->mc->run_once();
    return 0;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return -1;
  }
}
uint64_t oracle_replicas_num_moves(oracle_zmc* o) { return o->mc->moves; }
uint32_t oracle_replicas_num_replicas(oracle_zmc* o) { return (uint32_t)o->mc->replicas.size(); }
void oracle_replicas_get_rng(oracle_zmc* o, uint64_t* s) {
  s[0] = o->mc->rng.s0;
  s[1] = o->mc->rng.s1;
}
uint32_t oracle_replicas_get_median(oracle_zmc* o, uint32_t cap, double* energies) {
  const std::vector<double>& e = o->mc->median.energies;
  for (size_t i = 0; i < e.size() && i < cap; i++) energies[i] = e[i];
  return (uint32_t)e.size();
}
int oracle_replicas_get_replicas(oracle_zmc* o, sadmc_zeno_replica_state* out) {
  for (size_t r = 0; r < o->mc->replicas.size(); r++) {
    const replicas::Replica& q = o->mc->replicas[r];
    sadmc_zeno_replica_state& s = out[r];
    std::memset(&s, 0, sizeof s);
    s.max_energy = q.max_energy;
    s.cutoff_energy = q.cutoff_energy;
    s.lowest_max_energy = q.lowest_max_energy;
    s.translation_scale = q.translation_scale;
    s.rejected_count = q.rejected_count;
    s.accepted_count = q.accepted_count;
    s.above_count = q.above_count;
    s.below_count = q.below_count;
    s.upwelling_count = q.upwelling_count;
    s.unique_visitors = q.unique_visitors;
    s.above_total = q.above_total;
    s.below_total = q.below_total;
    s.above_total_squared = q.above_total_squared;
    s.below_total_squared = q.below_total_squared;
    if (!q.above_extra.empty()) {
      s.above_extra_total = q.above_extra.begin()->second.first;
      s.above_extra_count = q.above_extra.begin()->second.second;
    }
    s.collecting_data = q.collecting_data ? 1 : 0;
    s.rng_s0 = q.rng.s0;
    s.rng_s1 = q.rng.s1;
    s.energy = q.energy();
  }
  return 0;
}
size_t oracle_replicas_system_len(oracle_zmc* o) { return o->mc->replicas[0].system->get_state().size(); }
int oracle_replicas_get_system(oracle_zmc* o, uint32_t replica, double* buf, size_t n) {
  const std::vector<double> s = o->mc->replicas[replica].system->get_state();
  if (n < s.size()) return -1;
  std::memcpy(buf, s.data(), s.size() * sizeof(double));
  return 0;
}

// ---- RNG / math probes for tests ----
void oracle_rng_seed(uint64_t seed, uint64_t* s) {
  Rng r = Rng::seed_from_u64(seed);
  s[0] = r.s0;
  s[1] = r.s1;
}
// kind: 0 next_u64, 1 gen_f64 (as bits), 2 gen_range_usize(0,n), 3 uniform_usize(0,n), 4 standard_normal (bits),
//       5 uniform_f64(lo,hi) bits, 6 open01 bits, 7 gen_range_f64(lo,hi) bits
void oracle_rng_stream(uint64_t* state, int kind, uint64_t n_arg, double lo, double hi, uint64_t count, uint64_t* out) {
  Rng r;
  r.s0 = state[0];
  r.s1 = state[1];
  for (uint64_t k = 0; k < count; k++) {
    double d = 0;
    switch (kind) {
      case 0: out[k] = r.next_u64(); continue;
      case 1: d = r.gen_f64(); break;
      case 2: out[k] = r.gen_range_usize(0, n_arg); continue;
      case 3: out[k] = r.uniform_usize(0, n_arg); continue;
      case 4: d = r.standard_normal(); break;
      case 5: d = r.uniform_f64(lo, hi); break;
      case 6: d = r.open01(); break;
      case 7: d = r.gen_range_f64(lo, hi); break;
    }
    std::memcpy(&out[k], &d, 8);
  }
  state[0] = r.s0;
  state[1] = r.s1;
}
double oracle_exp(double x) { return sadmc_exp(x); }
double oracle_log(double x) { return sadmc_log(x); }
double oracle_erf_inv(double x) { return erf_inv(x); }
void oracle_zig_tables(double* x, double* f) {
  std::memcpy(x, ZIG_X, sizeof(ZIG_X));
  std::memcpy(f, ZIG_F, sizeof(ZIG_F));
}

// ---- CPU baseline: one independent walker per thread, all on this host ----
// Returns wall seconds for `n_moves` moves on each of `n_threads` walkers
// (construction excluded).  Used by bench.py only.
double oracle_bench(const sadmc_config* cfg, uint32_t n_threads, uint64_t warmup_moves, uint64_t n_moves) {
  std::vector<oracle_mc*> w(n_threads, nullptr);
  {
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < n_threads; t++)
      th.emplace_back([&, t] {
        w[t] = oracle_create(cfg, cfg->walker_offset + t, nullptr, 0, 0);
        if (w[t]) oracle_run(w[t], warmup_moves);
      });
    for (auto& x : th) x.join();
  }
  for (auto* p : w)
    if (!p) return -1.0;
  const auto t0 = std::chrono::steady_clock::now();
  {
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < n_threads; t++) th.emplace_back([&, t] { oracle_run(w[t], n_moves); });
    for (auto& x : th) x.join();
  }
  const auto t1 = std::chrono::steady_clock::now();
  for (auto* p : w) oracle_destroy(p);
  return std::chrono::duration<double>(t1 - t0).count();
}

} // extern "C"
