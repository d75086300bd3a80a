// args.hpp -- the command line of the reference's `histogram` binary.
//
// `auto_args` derives the flags from the parameter structs: struct fields -> --kebab-case, enum
// variants -> prefixes, `_fields` flattened (mc/mod.rs:22-32 Params, mc/energy.rs:41-97
// MethodParams / MoveParams / EnergyMCParams, mc/plugin.rs:159-167, 322-326, 411-415, system/any.rs:10-27
// and the per-system parameter structs).  `--flag value` or `--flag=value`; numeric values are
// expressions ('10^(1/8)', 1e9, 1/3).  The table and the rules are the same as the Python host's
// (sad_monte_carlo_b200/histogram.py); tests/test_host_cpp.py checks the two parsers against each other.
#pragma once
#include <cmath>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/sadmc_gpu.h"

namespace sadmc_host {

struct UsageError : std::runtime_error {
  explicit UsageError(const std::string& m) : std::runtime_error(m) {}
};

// ---- expression-valued numbers -------------------------------------------------------------------
class Expr {
  const std::string& s;
  size_t p = 0;
  char peek() {
    while (p < s.size() && (s[p] == ' ' || s[p] == '\t')) p++;
    return p < s.size() ? s[p] : 0;
  }
  [[noreturn]] void fail(const std::string& what) { throw UsageError(what + " in '" + s + "'"); }
  double number() {
    const size_t b = p;
    while (p < s.size() && ((s[p] >= '0' && s[p] <= '9') || s[p] == '.')) p++;
    if (p < s.size() && (s[p] == 'e' || s[p] == 'E') && p > b) {
      size_t q = p + 1;
      if (q < s.size() && (s[q] == '+' || s[q] == '-')) q++;
      if (q < s.size() && s[q] >= '0' && s[q] <= '9') {
        p = q;
        while (p < s.size() && s[p] >= '0' && s[p] <= '9') p++;
      }
    }
    char* end = nullptr;
    const std::string t = s.substr(b, p - b);
    const double v = strtod(t.c_str(), &end);
    if (!end || *end) fail("bad number");
    return v;
  }
  double atom() {
    const char c = peek();
    if (c == '(') {
      p++;
      const double v = expr();
      if (peek() != ')') fail("missing )");
      p++;
      return v;
    }
    if ((c >= '0' && c <= '9') || c == '.') return number();
    if ((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z')) {
      const size_t b = p;
      while (p < s.size() && (isalnum((unsigned char)s[p]) || s[p] == '_')) p++;
      const std::string name = s.substr(b, p - b);
      if (peek() == '(') {
        p++;
        const double v = expr();
        if (peek() != ')') fail("missing )");
        p++;
        if (name == "sqrt") return std::sqrt(v);
        if (name == "abs") return std::fabs(v);
        if (name == "exp") return std::exp(v);
        if (name == "ln") return std::log(v);
        if (name == "log") return std::log10(v);
        if (name == "sin") return std::sin(v);
        if (name == "cos") return std::cos(v);
        if (name == "tan") return std::tan(v);
        if (name == "floor") return std::floor(v);
        if (name == "ceil") return std::ceil(v);
        if (name == "round") return std::nearbyint(v);
        fail("unknown function '" + name + "'");
      }
      if (name == "pi") return M_PI;
      if (name == "e") return M_E;
      if (name == "inf" || name == "infinity") return INFINITY;
      fail("unknown name '" + name + "'");
    }
    fail("cannot parse");
  }
  double power() {
    const double base = atom();
    if (peek() == '^') {
      p++;
      return std::pow(base, unary());
    }
    return base;
  }
  double unary() {
    const char c = peek();
    if (c == '-') {
      p++;
      return -unary();
    }
    if (c == '+') {
      p++;
      return unary();
    }
    return power();
  }
  double term() {
    double v = unary();
    for (;;) {
      const char c = peek();
      if (c != '*' && c != '/' && c != '%') return v;
      p++;
      const double r = unary();
      if (c == '*')
        v *= r;
      else if (c == '/') {
        if (r == 0.0) fail("division by zero");
        v /= r;
      } else
        v = std::fmod(v, r);
    }
  }
  double expr() {
    double v = term();
    for (;;) {
      const char c = peek();
      if (c != '+' && c != '-') return v;
      p++;
      const double r = term();
      v = c == '+' ? v + r : v - r;
    }
  }

 public:
  explicit Expr(const std::string& text) : s(text) {}
  double eval() {
    if (peek() == 0) fail("empty number");
    const double v = expr();
    if (peek() != 0) fail("trailing characters");
    return v;
  }
};
inline double evaluate(const std::string& text) { return Expr(text).eval(); }

// ---- the flag table ----------------------------------------------------------------------------
enum FlagKind { F64, INT, FLAG, PATH, VEC3 };
struct FlagSpec {
  const char* name;
  FlagKind kind;
  const char* group; // system name, "method", "mc" or "gpu"
};
static const FlagSpec FLAGS[] = {
    {"fake-linear", FLAG, "fake"}, {"fake-quadratic-dimensions", INT, "fake"}, {"fake-pieces-a", F64, "fake"}, {"fake-pieces-b", F64, "fake"},
    {"fake-pieces-e1", F64, "fake"}, {"fake-pieces-e2", F64, "fake"}, {"fake-gaussian-sigma", F64, "fake"},                // fake.rs:12-36,66-74
    {"fake-erfinv-mean-energy", F64, "fake-erfinv"}, {"fake-erfinv-N", INT, "fake-erfinv"},                                  // erfinv.rs:11-26
    {"wca-cell-width", VEC3, "wca"}, {"wca-cell-volume", F64, "wca"}, {"wca-reduced-density", F64, "wca"}, {"wca-N", INT, "wca"},
    {"wca-fcc", FLAG, "wca"},                                                                                                // wca.rs:361-380
    {"lj-N", INT, "lj"}, {"lj-radius", F64, "lj"},                                                                           // lj.rs:14-24
    {"ising-N", INT, "ising"},                                                                                               // ising.rs:10-17
    {"sw-well-width", F64, "sw"}, {"sw-cell-width", VEC3, "sw"}, {"sw-cell-volume", F64, "sw"}, {"sw-filling-fraction", F64, "sw"},
    {"sw-N", INT, "sw"},                                                                                                     // optsquare.rs:326-345
    {"two-wells-N", INT, "two-wells"}, {"two-wells-h2-to-h1", F64, "two-wells"}, {"two-wells-barrier-over-h1", F64, "two-wells"},
    {"two-wells-r2", F64, "two-wells"},                                                                                      // two_wells.rs:13-22
    {"water-N", INT, "water"},
    {"sad-min-T", F64, "method"}, {"samc-t0", F64, "method"}, {"wl", FLAG, "method"}, {"wl-min-gamma", F64, "method"},
    {"Inv-t-WL", FLAG, "method"}, {"inv-t-wl", FLAG, "method"}, {"T", F64, "method"}, {"canonical-T", F64, "method"},        // energy.rs:41-69
    {"seed", INT, "mc"}, {"energy-bin", F64, "mc"}, {"min-allowed-energy", F64, "mc"}, {"max-allowed-energy", F64, "mc"},  // energy.rs:81-97
    {"translation-scale", F64, "mc"}, {"acceptance-rate", F64, "mc"},                                                        // MoveParams 71-78
    {"max-iter", INT, "mc"}, {"max-independent-samples", INT, "mc"}, {"quiet", FLAG, "mc"},                                 // plugin.rs:159-167
    {"movie-time", F64, "mc"}, {"save-time", F64, "mc"},                                                                     // plugin.rs:411-415, 322-326
    {"save-as", PATH, "mc"}, {"num-threads", INT, "mc"}, {"resume-from", PATH, "mc"},                                       // mc/mod.rs:22-32
    {"num-walkers", INT, "gpu"}, {"gpu-device", INT, "gpu"}, {"bin-window-lo", F64, "gpu"}, {"bin-window-hi", F64, "gpu"},
    {"lanes-per-walker", INT, "gpu"}, {"fast-math", FLAG, "gpu"}, {"checkpoint-walkers", INT, "gpu"}, {"dry-run", FLAG, "gpu"},
    {"lj-stream-z", FLAG, "gpu"}, {"lj-smem-z", FLAG, "gpu"}, // SADMC_FLAG_LJ_STREAM_Z / _SMEM_Z: force one LJ31 / LJ38 layout
    {"max-launch", INT, "gpu"}, {"help", FLAG, "gpu"}, {"convert", PATH, "gpu"}, {"convert-to", PATH, "gpu"},
};

struct FlagValue {
  FlagKind kind = FLAG;
  double f = 0.0;
  uint64_t u = 0;
  double v3[3] = {0, 0, 0};
  std::string path;
};
typedef std::map<std::string, FlagValue> Flags;

inline const FlagSpec* find_flag(const std::string& name) {
  for (const FlagSpec& f : FLAGS)
    if (name == f.name) return &f;
  return nullptr;
}

inline Flags parse_flags(const std::vector<std::string>& argv) {
  Flags out;
  size_t i = 0;
  while (i < argv.size()) {
    const std::string& a = argv[i];
    if (a.rfind("--", 0) != 0) throw UsageError("unexpected argument '" + a + "'");
    const size_t eq = a.find('=');
    const std::string name = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
    const FlagSpec* spec = find_flag(name);
    if (!spec) throw UsageError("unknown flag --" + name);
    if (out.count(name)) throw UsageError("--" + name + " given twice");
    FlagValue v;
    v.kind = spec->kind;
    if (spec->kind == FLAG) {
      if (eq != std::string::npos) throw UsageError("--" + name + " takes no value");
      out[name] = v;
      i++;
      continue;
    }
    const size_t need = spec->kind == VEC3 ? 3 : 1;
    std::vector<std::string> vals;
    if (eq != std::string::npos) {
      std::string inl = a.substr(eq + 1);
      if (need == 1) {
        vals.push_back(inl);
      } else {
        for (char& c : inl)
          if (c == ',') c = ' ';
        size_t p = 0;
        while (p < inl.size()) {
          while (p < inl.size() && inl[p] == ' ') p++;
          size_t e = p;
          while (e < inl.size() && inl[e] != ' ') e++;
          if (e > p) vals.push_back(inl.substr(p, e - p));
          p = e;
        }
      }
      i++;
    } else {
      for (size_t k = 1; k <= need && i + k < argv.size(); k++) vals.push_back(argv[i + k]);
      for (auto& x : vals)
        if (x.rfind("--", 0) == 0) { // the next flag, not a value ("-1.5" is a value)
          vals.clear();
          break;
        }
      i += 1 + need;
    }
    if (vals.size() != need) throw UsageError("--" + name + " needs " + std::to_string(need) + " value" + (need > 1 ? "s" : ""));
    try {
      if (spec->kind == PATH) {
        v.path = vals[0];
      } else if (spec->kind == F64) {
        v.f = evaluate(vals[0]);
      } else if (spec->kind == INT) {
        const double x = evaluate(vals[0]);
        if (std::isnan(x) || x < 0 || x != std::floor(x) || x >= 18446744073709551616.0) throw UsageError("needs a non-negative integer");
        v.u = (uint64_t)x;
        v.f = x;
      } else {
        for (int k = 0; k < 3; k++) v.v3[k] = evaluate(vals[(size_t)k]);
      }
    } catch (const UsageError& e) {
      throw UsageError("--" + name + ": " + e.what());
    }
    out[name] = v;
  }
  return out;
}

inline bool has(const Flags& f, const char* name) { return f.count(name) != 0; }
inline double num(const Flags& f, const char* name) { return f.at(name).f; }

// exactly one of the named groups may be present; returns its name ("" when none and !required)
inline std::string one_of(const Flags& f, const std::vector<std::pair<std::string, std::vector<std::string>>>& groups, const std::string& what,
                          bool required = true) {
  std::vector<std::string> present;
  for (auto& g : groups)
    for (auto& n : g.second)
      if (f.count(n)) {
        present.push_back(g.first);
        break;
      }
  if (present.size() > 1) {
    std::string l;
    for (auto& p : present) l += (l.empty() ? "" : ", ") + p;
    throw UsageError("more than one " + what + " given: " + l);
  }
  if (present.empty()) {
    if (required) throw UsageError("no " + what + " given");
    return "";
  }
  return present[0];
}

inline sadmc_config default_config() { // EnergyMCParams::default (energy.rs:99-115) + per-system defaults
  sadmc_config c;
  memset(&c, 0, sizeof c);
  c.abi_version = SADMC_ABI_VERSION;
  c.sad_min_T = 0.2;
  c.wl_min_gamma = c.energy_bin = c.min_allowed_energy = c.max_allowed_energy = c.bin_window_lo = c.bin_window_hi = NAN;
  c.high_resolution_de = NAN; // energy_binning.rs only (SADMC_FLAG_BINNING); `histogram` has no such parameter
  c.move_plan = SADMC_MOVE_TRANSLATION_SCALE;
  c.move_value = 0.05;
  c.n_walkers = 1;
  c.init_mode = SADMC_INIT_REFERENCE;
  c.reduced_density = 1.0;
  c.filling_fraction = 0.3;
  c.sw_well_width = 1.3;
  return c;
}

inline void require(const Flags& f, std::initializer_list<const char*> names) {
  for (const char* n : names)
    if (!f.count(n)) throw UsageError(std::string("--") + n + " is required");
}

// `AnyParams` + `EnergyMCParams` -> sadmc_config
inline sadmc_config config_from_flags(const Flags& f) {
  std::vector<std::pair<std::string, std::vector<std::string>>> sys_groups;
  for (const char* sysname : {"fake", "fake-erfinv", "wca", "lj", "ising", "sw", "two-wells", "water"}) {
    std::vector<std::string> names;
    for (const FlagSpec& s : FLAGS)
      if (std::string(s.group) == sysname) names.push_back(s.name);
    sys_groups.emplace_back(sysname, names);
  }
  const std::string system = one_of(f, sys_groups, "system");
  sadmc_config c = default_config();
  if (system == "water") throw UsageError("--water-*: the water model has no device kernel (out of scope, SURVEY.md section 8)");
  if (system == "lj") {
    require(f, {"lj-N", "lj-radius"});
    c.system = SADMC_SYS_LJ;
    c.N = (uint32_t)f.at("lj-N").u;
    c.lj_radius = num(f, "lj-radius");
  } else if (system == "ising") {
    c.system = SADMC_SYS_ISING;
    c.N = (uint32_t)f.at("ising-N").u;
  } else if (system == "fake") {
    c.system = SADMC_SYS_FAKE;
    const std::string fn = one_of(f, {{"linear", {"fake-linear"}}, {"quadratic", {"fake-quadratic-dimensions"}},
                                      {"pieces", {"fake-pieces-a", "fake-pieces-b", "fake-pieces-e1", "fake-pieces-e2"}},
                                      {"gaussian", {"fake-gaussian-sigma"}}}, "fake function");
    if (fn == "linear") {
      c.fake_function = SADMC_FAKE_LINEAR;
      c.N = 1;
    } else if (fn == "quadratic") {
      c.fake_function = SADMC_FAKE_QUADRATIC;
      c.N = (uint32_t)f.at("fake-quadratic-dimensions").u;
    } else if (fn == "pieces") {
      require(f, {"fake-pieces-a", "fake-pieces-b", "fake-pieces-e1", "fake-pieces-e2"});
      c.fake_function = SADMC_FAKE_PIECES;
      c.N = 3;
      c.fake_a = num(f, "fake-pieces-a");
      c.fake_b = num(f, "fake-pieces-b");
      c.fake_e1 = num(f, "fake-pieces-e1");
      c.fake_e2 = num(f, "fake-pieces-e2");
    } else {
      c.fake_function = SADMC_FAKE_GAUSSIAN;
      c.N = 3;
      c.fake_sigma = num(f, "fake-gaussian-sigma");
    }
  } else if (system == "fake-erfinv") {
    require(f, {"fake-erfinv-N", "fake-erfinv-mean-energy"});
    c.system = SADMC_SYS_FAKE_ERFINV;
    c.N = (uint32_t)f.at("fake-erfinv-N").u;
    c.erfinv_mean_energy = num(f, "fake-erfinv-mean-energy");
  } else if (system == "wca" || system == "sw") {
    const bool sw = system == "sw";
    const std::string pre = system + "-";
    const std::string third = sw ? "filling-fraction" : "reduced-density";
    const std::string d = one_of(f, {{"cell-width", {pre + "cell-width"}}, {"cell-volume", {pre + "cell-volume"}}, {third, {pre + third}}},
                                 "cell dimension");
    if (!f.count(pre + "N")) throw UsageError("--" + pre + "N is required");
    c.system = sw ? SADMC_SYS_SW : SADMC_SYS_WCA;
    c.N = (uint32_t)f.at(pre + "N").u;
    if (d == "cell-width") {
      for (int k = 0; k < 3; k++) c.cell_width[k] = f.at(pre + "cell-width").v3[k];
    } else if (d == "cell-volume") {
      const double w = std::cbrt(f.at(pre + "cell-volume").f); // Cell::new: CellVolume(v) -> v.cbrt() per side (optcell.rs:47-50)
      c.cell_width[0] = c.cell_width[1] = c.cell_width[2] = w;
    } else if (sw) {
      c.filling_fraction = f.at("sw-filling-fraction").f;
    } else {
      c.reduced_density = f.at("wca-reduced-density").f;
    }
    if (sw) {
      require(f, {"sw-well-width"});
      c.sw_well_width = num(f, "sw-well-width");
    }
    if (f.count("wca-fcc")) throw UsageError("--wca-fcc: the fcc start (rand's choose_multiple over the stretched grid, wca.rs:406-446) is not restated");
  } else { // two-wells
    require(f, {"two-wells-N", "two-wells-h2-to-h1", "two-wells-barrier-over-h1", "two-wells-r2"});
    c.system = SADMC_SYS_TWO_WELLS;
    c.N = (uint32_t)f.at("two-wells-N").u;
    c.tw_h2_to_h1 = num(f, "two-wells-h2-to-h1");
    c.tw_barrier_over_h1 = num(f, "two-wells-barrier-over-h1");
    c.tw_r2 = num(f, "two-wells-r2");
  }

  const std::string method = one_of(f, {{"sad", {"sad-min-T"}}, {"samc", {"samc-t0"}}, {"wl", {"wl", "wl-min-gamma"}},
                                        {"inv-t-wl", {"Inv-t-WL", "inv-t-wl"}}, {"canonical", {"T", "canonical-T"}}}, "method");
  if (method == "sad") {
    c.method = SADMC_METHOD_SAD;
    c.sad_min_T = num(f, "sad-min-T");
  } else if (method == "samc") {
    c.method = SADMC_METHOD_SAMC;
    c.samc_t0 = num(f, "samc-t0");
  } else if (method == "wl") {
    c.method = SADMC_METHOD_WL;
    if (f.count("wl-min-gamma")) c.wl_min_gamma = num(f, "wl-min-gamma");
  } else if (method == "inv-t-wl") {
    c.method = SADMC_METHOD_INV_T_WL;
  } else {
    c.method = SADMC_METHOD_CANONICAL;
    c.canonical_T = f.count("T") ? num(f, "T") : num(f, "canonical-T");
  }
  const std::string moves = one_of(f, {{"translation-scale", {"translation-scale"}}, {"acceptance-rate", {"acceptance-rate"}}}, "move plan", false);
  if (moves == "acceptance-rate") {
    c.move_plan = SADMC_MOVE_ACCEPTANCE_RATE;
    c.move_value = num(f, "acceptance-rate");
  } else if (moves == "translation-scale") {
    c.move_value = num(f, "translation-scale");
  }
  if (f.count("energy-bin")) c.energy_bin = num(f, "energy-bin");
  if (f.count("min-allowed-energy")) c.min_allowed_energy = num(f, "min-allowed-energy");
  if (f.count("max-allowed-energy")) c.max_allowed_energy = num(f, "max-allowed-energy");
  if (f.count("bin-window-lo")) c.bin_window_lo = num(f, "bin-window-lo");
  if (f.count("bin-window-hi")) c.bin_window_hi = num(f, "bin-window-hi");
  if (f.count("lanes-per-walker")) c.lanes_per_walker = (int32_t)f.at("lanes-per-walker").u;
  if (f.count("gpu-device")) c.device = (int32_t)f.at("gpu-device").u;
  c.seed = f.count("seed") ? f.at("seed").u : 0; // energy.rs:835: params.seed.unwrap_or(0)
  c.n_walkers = f.count("num-walkers") ? (uint32_t)f.at("num-walkers").u : 1;
  if (f.count("fast-math")) c.flags |= SADMC_FLAG_FAST_MATH;
  if (f.count("lj-stream-z")) c.flags |= SADMC_FLAG_LJ_STREAM_Z;
  if (f.count("lj-smem-z")) c.flags |= SADMC_FLAG_LJ_SMEM_Z;
  return c;
}

// ReportParams / SaveParams / MovieParams as given on the command line (plugin.rs:159-167, 322-326, 411-415)
struct PluginParams {
  bool has_max_iter = false, has_max_samples = false, quiet = false, has_movie_time = false, has_save_time = true;
  uint64_t max_iter = 0, max_independent_samples = 0;
  double save_time = 1.0, movie_time = 0.0; // SaveParams::default: one hour
};
inline PluginParams plugin_params(const Flags& f) {
  PluginParams p;
  if (f.count("max-iter")) {
    p.has_max_iter = true;
    p.max_iter = f.at("max-iter").u;
  }
  if (f.count("max-independent-samples")) {
    p.has_max_samples = true;
    p.max_independent_samples = f.at("max-independent-samples").u;
  }
  p.quiet = f.count("quiet") != 0; // a bool field is false unless its flag is given
  if (f.count("save-time")) p.save_time = num(f, "save-time");
  if (f.count("movie-time")) {
    p.has_movie_time = true;
    p.movie_time = num(f, "movie-time");
  }
  return p;
}

} // namespace sadmc_host
