// checkpoint.hpp -- checkpoints in the reference's serde schema, one document per walker.
//
// `walker_document` builds the mapping serde writes for `EnergyMC<Any>` (src/mc/energy.rs:167-210: externally
// tagged `system` / `method` enums, `bins` sub-map, Option::None = null, unit newtypes = bare f64) from the C ABI
// getters; `restore_walker` feeds one back through the resume entry points; `config_from_document` rebuilds the
// engine configuration from a document alone (--resume-from, mc/mod.rs:92-106).  Files are written through a
// temporary name and renamed (src/atomicfile.rs).  Same documents as the Python host (checkpoint.py); the two are
// compared file against file in tests/test_gpu_host_cpp.py.
#pragma once
#include <sys/stat.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <fstream>
#include <functional>
#include <sstream>
#include <string>
#include <vector>

#include "args.hpp"
#include "engine.hpp"
#include "value.hpp"

namespace sadmc_host {

inline Value vec3(double x, double y, double z) {
  Value v = Value::map();
  v.set("x", Value::number(x)).set("y", Value::number(y)).set("z", Value::number(z));
  return v;
}

// `SystemInvCdf::new` (two_wells.rs:46-137): cumulative distributions used only by TwoWells::randomize, which no
// EnergyMC run calls; derived data, written so that the reference can deserialise the document.
inline Value two_wells_invcdf(size_t dim, double r2) {
  const size_t num_points = 10000, mult = 100;
  const double r1 = 1.0;
  auto V = [](size_t n) { return std::pow(M_PI, 0.5 * (double)n) / std::tgamma((double)n * 0.5 + 1.0); }; // two_wells.rs:211-213
  auto lin = [](double a, double b, size_t n, size_t i) { // linspace, two_wells.rs:36-43
    const double last = (double)n - 1.0;
    return ((last - (double)i) * a + (double)i * b) * (1.0 / last);
  };
  std::vector<double> st(num_points * dim, 0.0);
  auto fill = [&](double* stencil, double a, double b, const std::function<double(double)>& pdf) {
    double val = 0.0;
    for (size_t w = 0; w + 1 < num_points; w++) {
      const double x0 = lin(a, b, num_points, w), x1 = lin(a, b, num_points, w + 1);
      const double du = lin(x0, x1, mult, 1) - lin(x0, x1, mult, 0);
      for (size_t i = 0; i + 1 < mult; i++) val += du * pdf(0.5 * (lin(x0, x1, mult, i + 1) + lin(x0, x1, mult, i)));
      stencil[w + 1] = val;
    }
    for (size_t w = 0; w < num_points; w++) stencil[w] /= val;
  };
  const double vd1 = V(dim - 1);
  fill(st.data(), -r1, r1 + 2.0 * r2, [&](double x) {
    if (x <= std::sqrt(r1 * r1 - r2 * r2)) return std::pow(std::sqrt(r1 * r1 - x * x), (double)dim - 1.0) * vd1;
    if (x < r1 + r2) return std::pow(r2, (double)dim - 1.0) * vd1;
    const double t = r2 * r2 - (x - r1 - r2) * (x - r1 - r2);
    return std::pow(std::sqrt(t > 0 ? t : 0.0), (double)dim - 1.0) * vd1;
  });
  for (size_t which = 1; which < dim; which++) {
    const size_t d = dim - which;
    const double ratio = V(d) / V(d + 1);
    fill(st.data() + which * num_points, -1.0, 1.0, [&](double x) {
      const double t = 1.0 - x * x;
      return std::pow(t > 0 ? t : 0.0, 0.5 * (double)d) * ratio;
    });
  }
  Value v = Value::map();
  v.set("num_points", Value::uinteger(num_points)).set("dim", Value::uinteger(dim)).set("r1", Value::number(r1)).set("r2", Value::number(r2));
  v.set("dx1_ball1", Value::number(lin(-r1, r1 + 2.0 * r2, num_points, 1) - lin(-r1, r1 + 2.0 * r2, num_points, 0)));
  Value s = Value::array();
  for (double x : st) s.push(Value::number(x));
  v.set("stencils", std::move(s));
  return v;
}

inline Value system_document(const GpuEnergyMC& mc, uint32_t w) { // the `Any` variant of this walker
  const sadmc_config& c = mc.cfg;
  const std::vector<double> img = mc.system(w);
  Value body = Value::map();
  auto positions = [&](uint32_t n) {
    Value p = Value::array();
    for (uint32_t k = 0; k < n; k++) p.push(vec3(img[3 * k], img[3 * k + 1], img[3 * k + 2]));
    return p;
  };
  std::string tag;
  switch (c.system) {
    case SADMC_SYS_LJ: // lj.rs:32-45
      tag = "Lj";
      body.set("E", Value::number(img[3 * c.N])).set("error", Value::number(img[3 * c.N + 1])).set("possible_change", Value::string("None"));
      body.set("positions", positions(c.N));
      body.set("max_radius_squared", Value::number(c.lj_radius * c.lj_radius)).set("max_radius", Value::number(c.lj_radius));
      break;
    case SADMC_SYS_ISING: { // ising.rs:20-29
      tag = "Ising";
      Value s = Value::array();
      for (uint32_t k = 0; k < c.N * c.N; k++) s.push(Value::integer((int64_t)img[k]));
      body.set("E", Value::number(img[(size_t)c.N * c.N])).set("N", Value::uinteger(c.N)).set("S", std::move(s)).set("possible_change", Value::null());
      break;
    }
    case SADMC_SYS_FAKE: { // fake.rs:76-83, Function 12-36
      tag = "Fake";
      Value fn;
      uint32_t dim = 3;
      if (c.fake_function == SADMC_FAKE_LINEAR) {
        fn = Value::string("Linear");
        dim = 1;
      } else if (c.fake_function == SADMC_FAKE_QUADRATIC) {
        fn = Value::map().set("Quadratic", Value::map().set("dimensions", Value::uinteger(c.N)));
        dim = c.N;
      } else if (c.fake_function == SADMC_FAKE_PIECES) {
        fn = Value::map().set("Pieces", Value::map().set("a", Value::number(c.fake_a)).set("b", Value::number(c.fake_b))
                                            .set("e1", Value::number(c.fake_e1)).set("e2", Value::number(c.fake_e2)));
      } else {
        fn = Value::map().set("Gaussian", Value::map().set("sigma", Value::number(c.fake_sigma)));
      }
      Value pos = Value::array(), pc = Value::array();
      for (uint32_t k = 0; k < dim; k++) {
        pos.push(Value::number(img[k]));
        pc.push(Value::number(0.0));
      }
      body.set("position", std::move(pos)).set("function", std::move(fn)).set("possible_change", std::move(pc));
      break;
    }
    case SADMC_SYS_WCA:
    case SADMC_SYS_SW: { // wca.rs:23-33, optsquare.rs:24-31 around optcell.rs:27-40 (subcells: #[serde(skip)])
      const bool sw = c.system == SADMC_SYS_SW;
      tag = sw ? "Sw" : "Wca";
      double box[3], rc;
      mc.cell_box(box, &rc);
      Value cell = Value::map();
      cell.set("box_diagonal", vec3(box[0], box[1], box[2])).set("r_cutoff", Value::number(rc)).set("positions", positions(c.N));
      body.set("E", Value::number(img[3 * c.N]));
      if (!sw) body.set("error", Value::number(img[3 * c.N + 1]));
      body.set("cell", std::move(cell)).set("possible_change", Value::string("None"));
      break;
    }
    case SADMC_SYS_TWO_WELLS: { // two_wells.rs:219-232
      tag = "TwoWells";
      Value pos = Value::array();
      for (uint32_t k = 0; k < c.N; k++) pos.push(Value::number(img[k]));
      Value params = Value::map();
      params.set("N", Value::uinteger(c.N)).set("h2_to_h1", Value::number(c.tw_h2_to_h1)).set("barrier_over_h1", Value::number(c.tw_barrier_over_h1))
          .set("r2", Value::number(c.tw_r2));
      const double well = std::sqrt(c.tw_barrier_over_h1) * 1.0 + c.tw_r2 * std::sqrt(1.0 + c.tw_barrier_over_h1 - 1.0 / c.tw_h2_to_h1);
      body.set("position", std::move(pos)).set("d_squared", Value::number(img[c.N])).set("parameters", std::move(params));
      body.set("change", Value::map().set("index", Value::uinteger(0)).set("values", vec3(0.0, 0.0, 0.0)));
      body.set("well_position", Value::number(well)).set("invcdf", two_wells_invcdf(c.N, c.tw_r2));
      break;
    }
    case SADMC_SYS_FAKE_ERFINV: { // erfinv.rs:29-38
      tag = "FakeErfinv";
      Value pos = Value::array();
      for (uint32_t k = 0; k < c.N; k++) pos.push(Value::number(img[k]));
      body.set("position", std::move(pos)).set("parameters", Value::map().set("mean_energy", Value::number(c.erfinv_mean_energy)));
      body.set("possible_change", Value::array());
      break;
    }
    default: throw std::runtime_error("no checkpoint document for this system kind");
  }
  return Value::map().set(tag, std::move(body));
}

template <class T, class F>
Value array_of(const std::vector<T>& v, F make, size_t n = (size_t)-1) {
  Value a = Value::array();
  for (size_t k = 0; k < v.size() && k < n; k++) a.push(make(v[k]));
  return a;
}

inline Value method_document(const sadmc_config& c, const sadmc_walker_state& st, const WalkerBins& b) {
  auto u = [](uint64_t x) { return Value::uinteger(x); };
  if (st.method == SADMC_METHOD_SAD) // energy.rs:215-226
    return Value::map().set("Sad", Value::map().set("min_T", Value::number(c.sad_min_T)).set("too_lo", Value::number(st.too_lo))
                                       .set("too_hi", Value::number(st.too_hi)).set("tL", u(st.tL)).set("tF", u(st.tF))
                                       .set("num_states", u(st.num_states)).set("highest_hist", u(st.highest_hist))
                                       .set("version", Value::string("Sad")).set("latest_parameter", Value::number(st.latest_parameter)));
  if (st.method == SADMC_METHOD_SAMC) return Value::map().set("Samc", Value::map().set("t0", Value::number(st.samc_t0))); // 227-228, 754-756
  if (st.method == SADMC_METHOD_WL || st.method == SADMC_METHOD_INV_T_WL) // 229-240
    return Value::map().set("WL", Value::map().set("gamma", Value::number(st.wl_gamma)).set("lowest_hist", u(st.wl_lowest_hist))
                                      .set("highest_hist", u(st.wl_highest_hist)).set("total_hist", u(st.wl_total_hist))
                                      .set("num_states", Value::number(st.wl_num_states))
                                      .set("hist", array_of(b.wl_hist, u, st.wl_hist_len)).set("min_energy", Value::number(st.wl_min_energy))
                                      .set("inv_t", Value::boolean(st.wl_inv_t != 0)).set("min_gamma", Value::optional(c.wl_min_gamma)));
  return Value::map().set("Canonical", Value::map().set("temperature", Value::number(c.canonical_T)));
}

// The serde document of walker `w` as the reference would write it (energy.rs:167-210).
inline Value walker_document(const GpuEnergyMC& mc, uint32_t w, const std::string& save_as, const Value& report, const Value& movies, const Value& save) {
  const sadmc_config& c = mc.cfg;
  const sadmc_walker_state st = mc.walker(w);
  if (st.status != 0) throw std::runtime_error("walker " + std::to_string(w) + " is halted (status " + std::to_string(st.status) + ")");
  const WalkerBins b = mc.bins(w);
  auto u = [](uint64_t x) { return Value::uinteger(x); };
  auto f = [](double x) { return Value::number(x); };
  Value extra = Value::map();
  const char* label = c.system == SADMC_SYS_WCA ? "pressure" : (c.system == SADMC_SYS_TWO_WELLS ? "which" : nullptr);
  bool any = false;
  for (uint64_t x : b.extra_count) any = any || x != 0;
  if (label && any) extra.set(label, Value::map().set("total", array_of(b.extra_total, f)).set("count", array_of(b.extra_count, u)));
  Value d = Value::map();
  d.set("system", system_document(mc, w));
  d.set("method", method_document(c, st, b));
  d.set("moves", u(st.moves)).set("time_L", u(0)).set("accepted_moves", u(st.accepted_moves));
  d.set("min_allowed_energy", Value::optional(c.min_allowed_energy)).set("max_allowed_energy", Value::optional(c.max_allowed_energy));
  d.set("move_plan", Value::map().set(c.move_plan == SADMC_MOVE_TRANSLATION_SCALE ? "TranslationScale" : "AcceptanceRate", f(c.move_value)));
  d.set("translation_scale", f(st.translation_scale)).set("acceptance_rate", f(st.acceptance_rate));
  d.set("rng", Value::map().set("s0", u(st.rng_s0)).set("s1", u(st.rng_s1))); // rand_xoshiro "serde1"
  d.set("save_as", Value::string(save_as));
  d.set("report", report).set("movies", movies).set("save", save).set("manager", Value::map());
  Value bins = Value::map();
  bins.set("min", f(st.bins_min)).set("width", f(st.bins_width)).set("histogram", array_of(b.histogram, u)).set("t_found", array_of(b.t_found, u));
  bins.set("lnw", array_of(b.lnw, f)).set("energy_total", array_of(b.energy_total, f)).set("energy_squared_total", array_of(b.energy_squared_total, f));
  bins.set("extra", std::move(extra));
  d.set("bins", std::move(bins));
  d.set("have_visited_since_maxentropy", array_of(b.have_visited, [](uint8_t x) { return Value::boolean(x != 0); }));
  d.set("round_trips", array_of(b.round_trips, u));
  d.set("max_S", f(st.max_S)).set("max_S_index", u(st.max_S_index));
  return d;
}

inline std::vector<double> system_image(const Value& sys, size_t length) { // inverse of system_document
  std::vector<double> img(length, 0.0);
  const std::string tag = sys.tag();
  const Value& body = sys.body();
  auto put_positions = [&](const Value& pos) {
    for (size_t k = 0; k < pos.a.size(); k++) {
      img[3 * k] = pos.a[k].at("x").as_f64();
      img[3 * k + 1] = pos.a[k].at("y").as_f64();
      img[3 * k + 2] = pos.a[k].at("z").as_f64();
    }
    return pos.a.size();
  };
  if (tag == "Lj") {
    const size_t n = put_positions(body.at("positions"));
    img[3 * n] = body.at("E").as_f64();
    img[3 * n + 1] = body.at("error").as_f64();
  } else if (tag == "Ising") {
    const Value& s = body.at("S");
    for (size_t k = 0; k < s.a.size(); k++) img[k] = (double)s.a[k].as_i64();
    img[s.a.size()] = body.at("E").as_f64();
  } else if (tag == "Fake" || tag == "FakeErfinv") {
    const Value& p = body.at("position");
    for (size_t k = 0; k < p.a.size(); k++) img[k] = p.a[k].as_f64();
  } else if (tag == "Wca" || tag == "Sw") {
    const size_t n = put_positions(body.at("cell").at("positions"));
    img[3 * n] = body.at("E").as_f64();
    const Value* e = body.find("error");
    img[3 * n + 1] = e ? e->as_f64() : 0.0;
  } else if (tag == "TwoWells") {
    const Value& p = body.at("position");
    for (size_t k = 0; k < p.a.size(); k++) img[k] = p.a[k].as_f64();
    img[p.a.size()] = body.at("d_squared").as_f64();
  } else {
    throw std::runtime_error("cannot restore system variant " + tag);
  }
  return img;
}

// Feed a document back into walker `w` of an engine created with SADMC_INIT_EXTERNAL (then mc.resume(moves)).
inline void restore_walker(GpuEnergyMC& mc, uint32_t w, const Value& doc) {
  mc.set_system(w, system_image(doc.at("system"), mc.system_len));
  sadmc_walker_state st;
  memset(&st, 0, sizeof st);
  st.moves = doc.at("moves").as_u64();
  st.accepted_moves = doc.at("accepted_moves").as_u64();
  st.acceptance_rate = doc.at("acceptance_rate").as_f64();
  st.translation_scale = doc.at("translation_scale").as_f64();
  st.rng_s0 = doc.at("rng").at("s0").as_u64();
  st.rng_s1 = doc.at("rng").at("s1").as_u64();
  const Value& bins = doc.at("bins");
  const size_t n = bins.at("lnw").a.size();
  st.bins_min = bins.at("min").as_f64();
  st.bins_width = bins.at("width").as_f64();
  st.bins_len = (uint32_t)n;
  st.max_S = doc.at("max_S").as_f64();
  st.max_S_index = (uint32_t)doc.at("max_S_index").as_u64();
  WalkerBins b;
  b.resize(n);
  const std::string mtag = doc.at("method").tag();
  const Value& m = doc.at("method").body();
  if (mtag == "Sad") {
    st.method = SADMC_METHOD_SAD;
    st.too_lo = m.at("too_lo").as_f64();
    st.too_hi = m.at("too_hi").as_f64();
    st.latest_parameter = m.at("latest_parameter").as_f64();
    st.tL = m.at("tL").as_u64();
    st.tF = m.at("tF").as_u64();
    st.num_states = m.at("num_states").as_u64();
    st.highest_hist = m.at("highest_hist").as_u64();
  } else if (mtag == "Samc") {
    st.method = SADMC_METHOD_SAMC;
    st.samc_t0 = m.at("t0").as_f64();
  } else if (mtag == "WL") {
    st.wl_inv_t = m.at("inv_t").as_bool() ? 1 : 0;
    st.method = st.wl_inv_t ? SADMC_METHOD_INV_T_WL : SADMC_METHOD_WL;
    st.wl_gamma = m.at("gamma").as_f64();
    st.wl_num_states = m.at("num_states").as_f64();
    st.wl_min_energy = m.at("min_energy").as_f64();
    st.wl_lowest_hist = m.at("lowest_hist").as_u64();
    st.wl_highest_hist = m.at("highest_hist").as_u64();
    st.wl_total_hist = m.at("total_hist").as_u64();
    const Value& h = m.at("hist");
    st.wl_hist_len = (uint32_t)h.a.size();
    for (size_t k = 0; k < h.a.size() && k < n; k++) b.wl_hist[k] = h.a[k].as_u64();
  } else {
    st.method = SADMC_METHOD_CANONICAL;
  }
  const std::string stag = doc.at("system").tag();
  const Value* e = doc.at("system").body().find("E");
  st.energy = e ? e->as_f64() : NAN; // the cached System::energy travels inside the system variant
  for (size_t k = 0; k < n; k++) {
    b.histogram[k] = bins.at("histogram").a[k].as_u64();
    b.t_found[k] = bins.at("t_found").a[k].as_u64();
    b.lnw[k] = bins.at("lnw").a[k].as_f64();
    b.energy_total[k] = bins.at("energy_total").a[k].as_f64();
    b.energy_squared_total[k] = bins.at("energy_squared_total").a[k].as_f64();
    b.round_trips[k] = doc.at("round_trips").a[k].as_u64();
    b.have_visited[k] = doc.at("have_visited_since_maxentropy").a[k].as_bool() ? 1 : 0;
  }
  if (const Value* ex = bins.find("extra"))
    for (auto& kv : ex->m) {
      for (size_t k = 0; k < n; k++) {
        b.extra_total[k] = kv.second.at("total").a[k].as_f64();
        b.extra_count[k] = kv.second.at("count").a[k].as_u64();
      }
    }
  // these keep no cached energy: System::energy evaluates the function (fake.rs:96-99) -- on the device
  if (stag == "Fake" || stag == "FakeErfinv" || stag == "TwoWells") st.energy = mc.compute_energy(w);
  mc.set_walker_bins(w, st, b);
}

// The sadmc_config a checkpoint document implies -- what --resume-from needs.
inline sadmc_config config_from_document(const Value& doc, uint32_t n_walkers) {
  sadmc_config c = default_config();
  c.n_walkers = n_walkers;
  c.init_mode = SADMC_INIT_EXTERNAL;
  const std::string tag = doc.at("system").tag();
  const Value& body = doc.at("system").body();
  if (tag == "Lj") {
    c.system = SADMC_SYS_LJ;
    c.N = (uint32_t)body.at("positions").a.size();
    c.lj_radius = body.at("max_radius").as_f64();
  } else if (tag == "Ising") {
    c.system = SADMC_SYS_ISING;
    c.N = (uint32_t)body.at("N").as_u64();
  } else if (tag == "Fake") {
    c.system = SADMC_SYS_FAKE;
    const Value& fn = body.at("function");
    const std::string ft = fn.tag();
    if (ft == "Linear") {
      c.fake_function = SADMC_FAKE_LINEAR;
      c.N = 1;
    } else if (ft == "Quadratic") {
      c.fake_function = SADMC_FAKE_QUADRATIC;
      c.N = (uint32_t)fn.body().at("dimensions").as_u64();
    } else if (ft == "Pieces") {
      c.fake_function = SADMC_FAKE_PIECES;
      c.N = 3;
      c.fake_a = fn.body().at("a").as_f64();
      c.fake_b = fn.body().at("b").as_f64();
      c.fake_e1 = fn.body().at("e1").as_f64();
      c.fake_e2 = fn.body().at("e2").as_f64();
    } else {
      c.fake_function = SADMC_FAKE_GAUSSIAN;
      c.N = 3;
      c.fake_sigma = fn.body().at("sigma").as_f64();
    }
  } else if (tag == "Wca" || tag == "Sw") {
    c.system = tag == "Wca" ? SADMC_SYS_WCA : SADMC_SYS_SW;
    const Value& cell = body.at("cell");
    c.N = (uint32_t)cell.at("positions").a.size();
    c.cell_width[0] = cell.at("box_diagonal").at("x").as_f64();
    c.cell_width[1] = cell.at("box_diagonal").at("y").as_f64();
    c.cell_width[2] = cell.at("box_diagonal").at("z").as_f64();
    if (tag == "Sw") c.sw_well_width = cell.at("r_cutoff").as_f64();
  } else if (tag == "TwoWells") {
    c.system = SADMC_SYS_TWO_WELLS;
    const Value& p = body.at("parameters");
    c.N = (uint32_t)p.at("N").as_u64();
    c.tw_h2_to_h1 = p.at("h2_to_h1").as_f64();
    c.tw_barrier_over_h1 = p.at("barrier_over_h1").as_f64();
    c.tw_r2 = p.at("r2").as_f64();
  } else if (tag == "FakeErfinv") {
    c.system = SADMC_SYS_FAKE_ERFINV;
    c.N = (uint32_t)body.at("position").a.size();
    c.erfinv_mean_energy = body.at("parameters").at("mean_energy").as_f64();
  } else {
    throw std::runtime_error("system variant " + tag + " has no device kernel");
  }
  const std::string mtag = doc.at("method").tag();
  const Value& m = doc.at("method").body();
  if (mtag == "Sad") {
    c.method = SADMC_METHOD_SAD;
    c.sad_min_T = m.at("min_T").as_f64();
  } else if (mtag == "Samc") {
    c.method = SADMC_METHOD_SAMC;
    c.samc_t0 = m.at("t0").as_f64();
  } else if (mtag == "WL") {
    c.method = m.at("inv_t").as_bool() ? SADMC_METHOD_INV_T_WL : SADMC_METHOD_WL;
    const Value* g = m.find("min_gamma");
    if (g && !g->is_null()) c.wl_min_gamma = g->as_f64();
  } else {
    c.method = SADMC_METHOD_CANONICAL;
    c.canonical_T = m.at("temperature").as_f64();
  }
  const Value* lo = doc.find("min_allowed_energy");
  const Value* hi = doc.find("max_allowed_energy");
  if (lo && !lo->is_null()) c.min_allowed_energy = lo->as_f64();
  if (hi && !hi->is_null()) c.max_allowed_energy = hi->as_f64();
  const Value& plan = doc.at("move_plan");
  c.move_plan = plan.tag() == "TranslationScale" ? SADMC_MOVE_TRANSLATION_SCALE : SADMC_MOVE_ACCEPTANCE_RATE;
  c.move_value = plan.body().as_f64();
  const Value& bins = doc.at("bins");
  c.energy_bin = bins.at("width").as_f64();
  // The device keeps a fixed bin window: where neither the bounds nor the system give one side of it, leave room
  // for as many new bins as the document holds.
  const double n = (double)bins.at("lnw").a.size(), width = c.energy_bin, bmin = bins.at("min").as_f64();
  const double span = (n > 64 ? n : 64) * width;
  const bool open_above = c.system == SADMC_SYS_LJ || c.system == SADMC_SYS_WCA || c.system == SADMC_SYS_FAKE_ERFINV;
  if (open_above && std::isnan(c.max_allowed_energy)) c.bin_window_hi = bmin + n * width + span;
  if (c.system == SADMC_SYS_FAKE_ERFINV && std::isnan(c.min_allowed_energy)) c.bin_window_lo = bmin - span;
  return c;
}

// ---- files ---------------------------------------------------------------------------------------------------
inline std::string read_file(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("error reading file \"" + path + "\"");
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}
inline bool file_exists(const std::string& path) {
  struct stat st;
  return stat(path.c_str(), &st) == 0;
}
inline void make_dirs(const std::string& dir) {
  if (dir.empty() || file_exists(dir)) return;
  const size_t slash = dir.find_last_of('/');
  if (slash != std::string::npos && slash > 0) make_dirs(dir.substr(0, slash));
  mkdir(dir.c_str(), 0777);
}
// AtomicFile (src/atomicfile.rs): the file appears under its name only when it is complete.
inline void write_atomic(const std::string& path, const std::string& data) {
  const size_t slash = path.find_last_of('/');
  if (slash != std::string::npos) make_dirs(path.substr(0, slash));
  const std::string tmp = path + ".tmp." + std::to_string((long)getpid());
  {
    std::ofstream f(tmp, std::ios::binary | std::ios::trunc);
    if (!f) throw std::runtime_error("error creating file \"" + path + "\"");
    f.write(data.data(), (std::streamsize)data.size());
    f.flush();
    if (!f) {
      unlink(tmp.c_str());
      throw std::runtime_error("error writing checkpoint \"" + path + "\"");
    }
  }
  if (rename(tmp.c_str(), path.c_str()) != 0) {
    unlink(tmp.c_str());
    throw std::runtime_error("error renaming checkpoint \"" + path + "\"");
  }
}
// One file per walker: `name.ext` for a single walker (as the reference), `name-w000017.ext` otherwise.
inline std::string walker_path(const std::string& save_as, uint32_t w, uint32_t n_walkers) {
  if (n_walkers == 1) return save_as;
  const size_t slash = save_as.find_last_of('/');
  const size_t dot = save_as.find_last_of('.');
  const bool has_ext = dot != std::string::npos && (slash == std::string::npos || dot > slash);
  char tag[16];
  snprintf(tag, sizeof tag, "-w%06u", w);
  return has_ext ? save_as.substr(0, dot) + tag + save_as.substr(dot) : save_as + tag;
}
inline Value load(const std::string& path) { return loads(read_file(path), extension_of(path)); }

inline std::string partial_marker(const std::string& save_as) {
  const size_t slash = save_as.find_last_of('/');
  const size_t dot = save_as.find_last_of('.');
  const bool has_ext = dot != std::string::npos && (slash == std::string::npos || dot > slash);
  return (has_ext ? save_as.substr(0, dot) : save_as) + ".partial";
}

// MonteCarlo::checkpoint (mc/mod.rs:110-120) for walkers [0, n_save).  The set of per-walker files is written all or
// nothing: halted walkers stop the save before any file is touched, every document goes to a temporary file beside its
// target, and only then are the temporaries renamed into place.  A partial set (n_save < n_walkers) is marked by
// `name.partial` so that a later resume refuses it with a clear message.
inline void save(const GpuEnergyMC& mc, const std::string& save_as, uint32_t n_save, const Value& report, const Value& movies, const Value& save_doc) {
  const std::string ext = extension_of(save_as);
  uint64_t left = 0, failed = 0;
  mc.num_halted(&left, &failed);
  if (left || failed)
    throw std::runtime_error("no checkpoint written: " + std::to_string(left) + " walker(s) left the bin window and " + std::to_string(failed) +
                             " failed verify_energy");
  const uint32_t n = n_save < mc.n_walkers() ? n_save : mc.n_walkers();
  std::vector<std::pair<std::string, std::string>> staged; // (temporary, target)
  try {
    for (uint32_t w = 0; w < n; w++) {
      const std::string p = walker_path(save_as, w, mc.n_walkers());
      const std::string tmp = p + ".tmp." + std::to_string((long)getpid());
      const size_t slash = p.find_last_of('/');
      if (slash != std::string::npos) make_dirs(p.substr(0, slash));
      const std::string data = dumps(walker_document(mc, w, p, report, movies, save_doc), ext);
      std::ofstream f(tmp, std::ios::binary | std::ios::trunc);
      if (!f) throw std::runtime_error("error creating file \"" + p + "\"");
      staged.emplace_back(tmp, p);
      f.write(data.data(), (std::streamsize)data.size());
      f.flush();
      if (!f) throw std::runtime_error("error writing checkpoint \"" + p + "\"");
    }
    for (auto& tp : staged)
      if (rename(tp.first.c_str(), tp.second.c_str()) != 0) throw std::runtime_error("error renaming checkpoint \"" + tp.second + "\"");
  } catch (...) {
    for (auto& tp : staged) unlink(tp.first.c_str());
    throw;
  }
  const std::string marker = partial_marker(save_as);
  if (n < mc.n_walkers())
    write_atomic(marker, std::to_string(n) + " of " + std::to_string(mc.n_walkers()) + " walkers\n");
  else if (file_exists(marker))
    unlink(marker.c_str());
}

// Before any engine is created: the checkpoint set must be complete and (cfg != nullptr) must describe the configuration
// the command line asks for.  Throws std::runtime_error saying what is wrong.
inline sadmc_config config_from_document(const Value& doc, uint32_t n_walkers);
inline void check_resumable(const sadmc_config* cfg, const std::string& save_as, uint32_t n_walkers) {
  const std::string marker = partial_marker(save_as);
  if (file_exists(marker)) {
    std::string what = read_file(marker);
    while (!what.empty() && (what.back() == '\n' || what.back() == ' ')) what.pop_back();
    throw std::runtime_error(save_as + " holds only " + what + " (written with --checkpoint-walkers): it cannot be resumed");
  }
  uint32_t missing = 0;
  std::string first_missing;
  for (uint32_t w = 0; w < n_walkers; w++)
    if (!file_exists(walker_path(save_as, w, n_walkers))) {
      if (!missing) first_missing = walker_path(save_as, w, n_walkers);
      missing++;
    }
  if (missing)
    throw std::runtime_error("checkpoint set " + save_as + " is incomplete: " + std::to_string(missing) + " of " + std::to_string(n_walkers) +
                             " walker files are missing (first: " + first_missing + ")");
  if (!cfg) return;
  const sadmc_config want = config_from_document(load(walker_path(save_as, 0, n_walkers)), n_walkers);
  if (cfg->system != want.system) throw std::runtime_error("checkpoint " + save_as + " was written for another system");
  if (cfg->N != want.N) throw std::runtime_error("checkpoint " + save_as + " was written for another system size");
  if (cfg->energy_bin == cfg->energy_bin && cfg->energy_bin != want.energy_bin)
    throw std::runtime_error("checkpoint " + save_as + " was written with another bin width");
  const bool same = cfg->method == want.method || (cfg->method == SADMC_METHOD_INV_T_WL && want.method == SADMC_METHOD_SAMC);
  if (!same) throw std::runtime_error("checkpoint " + save_as + " was written by another method");
}

// `--save-as` on an existing file (mc/mod.rs:70-84): restore every walker of a fresh INIT_EXTERNAL engine
inline void resume_into(GpuEnergyMC& mc, const std::string& save_as) {
  bool have = false;
  uint64_t moves = 0;
  for (uint32_t w = 0; w < mc.n_walkers(); w++) {
    const Value doc = load(walker_path(save_as, w, mc.n_walkers()));
    restore_walker(mc, w, doc);
    const uint64_t m = doc.at("moves").as_u64();
    if (have && m != moves)
      throw std::runtime_error("walker checkpoints disagree on `moves`: the set " + save_as + " was interrupted while it was being replaced");
    moves = m;
    have = true;
  }
  mc.resume(moves);
}

} // namespace sadmc_host
