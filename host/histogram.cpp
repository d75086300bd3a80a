// histogram.cpp -- the reference's main binary over the device engine (compiled host).
//
//   src/bin/histogram.rs:   let mut mc = EnergyMC::<Any>::from_args::<AnyParams>();  loop { mc.move_once(); }
//
// `MonteCarlo::from_args` (src/mc/mod.rs:55-107): --save-as on an existing file resumes it (report and save
// parameters refreshed from the flags, energy.rs:899-902); --resume-from FILE continues a checkpoint as it is;
// otherwise from_params.  The loop runs one kernel launch per plugin period (plugins.hpp) and ends when `Report`
// says so, after a final checkpoint (plugin.rs:104-124; the reference calls std::process::exit(0) there).
// Panics of the reference become a message on stderr and exit status 1 (usage: 2).
//
// Added for the GPU: --num-walkers W (walker w == the reference run with --seed seed+w; one file per walker when
// W > 1), --gpu-device, --bin-window-lo/-hi, --lanes-per-walker, --fast-math, --checkpoint-walkers K, --max-launch,
// --dry-run (print the parsed configuration as JSON; needs no GPU), and `--convert IN --convert-to OUT`
// (re-encode a checkpoint between yaml / json / cbor; needs no GPU).
#include <cstdio>
#include <memory>

#include "args.hpp"
#include "checkpoint.hpp"
#include "engine.hpp"
#include "plugins.hpp"
#include "value.hpp"

using namespace sadmc_host;

static Value config_summary(const sadmc_config& c) {
  Value v = Value::map();
  auto opt = [](double x) { return Value::optional(x); };
  v.set("abi_version", Value::uinteger(c.abi_version)).set("system", Value::integer(c.system)).set("N", Value::uinteger(c.N));
  v.set("lj_radius", Value::number(c.lj_radius)).set("reduced_density", Value::number(c.reduced_density));
  v.set("filling_fraction", Value::number(c.filling_fraction));
  v.set("cell_width", Value::array().push(Value::number(c.cell_width[0])).push(Value::number(c.cell_width[1])).push(Value::number(c.cell_width[2])));
  v.set("sw_well_width", Value::number(c.sw_well_width)).set("fake_function", Value::integer(c.fake_function));
  v.set("fake_a", Value::number(c.fake_a)).set("fake_b", Value::number(c.fake_b)).set("fake_e1", Value::number(c.fake_e1));
  v.set("fake_e2", Value::number(c.fake_e2)).set("fake_sigma", Value::number(c.fake_sigma));
  v.set("tw_h2_to_h1", Value::number(c.tw_h2_to_h1)).set("tw_barrier_over_h1", Value::number(c.tw_barrier_over_h1)).set("tw_r2", Value::number(c.tw_r2));
  v.set("erfinv_mean_energy", Value::number(c.erfinv_mean_energy)).set("method", Value::integer(c.method)).set("move_plan", Value::integer(c.move_plan));
  v.set("sad_min_T", Value::number(c.sad_min_T)).set("samc_t0", Value::number(c.samc_t0)).set("wl_min_gamma", opt(c.wl_min_gamma));
  v.set("canonical_T", Value::number(c.canonical_T)).set("seed", Value::uinteger(c.seed)).set("energy_bin", opt(c.energy_bin));
  v.set("min_allowed_energy", opt(c.min_allowed_energy)).set("max_allowed_energy", opt(c.max_allowed_energy)).set("move_value", Value::number(c.move_value));
  v.set("n_walkers", Value::uinteger(c.n_walkers)).set("walker_offset", Value::uinteger(c.walker_offset)).set("device", Value::integer(c.device));
  v.set("init_mode", Value::integer(c.init_mode)).set("bin_window_lo", opt(c.bin_window_lo)).set("bin_window_hi", opt(c.bin_window_hi));
  v.set("lanes_per_walker", Value::integer(c.lanes_per_walker)).set("flags", Value::uinteger(c.flags));
  v.set("high_resolution_de", opt(c.high_resolution_de));
  return v;
}
static Value plugin_summary(const PluginParams& p) {
  Value v = Value::map();
  v.set("max_iter", p.has_max_iter ? Value::uinteger(p.max_iter) : Value::null());
  v.set("max_independent_samples", p.has_max_samples ? Value::uinteger(p.max_independent_samples) : Value::null());
  v.set("quiet", Value::boolean(p.quiet)).set("save_time", p.has_save_time ? Value::number(p.save_time) : Value::null());
  v.set("movie_time", p.has_movie_time ? Value::number(p.movie_time) : Value::null());
  return v;
}
static void print_json(const Value& v) {
  std::string s;
  to_json(v, s);
  printf("%s\n", s.c_str());
}
static std::string self_dir(const char* argv0) {
  char buf[4096];
  const ssize_t n = readlink("/proc/self/exe", buf, sizeof buf - 1);
  std::string p = n > 0 ? std::string(buf, (size_t)n) : std::string(argv0);
  const size_t slash = p.find_last_of('/');
  return slash == std::string::npos ? "." : p.substr(0, slash);
}

static int real_main(int argc, char** argv) {
  std::vector<std::string> args(argv + 1, argv + argc);
  const Flags flags = parse_flags(args);
  if (has(flags, "help")) {
    printf("histogram: flat-histogram Monte Carlo of sad-monte-carlo on a B200 (see host/histogram.cpp).\nFlags:\n");
    for (const FlagSpec& f : FLAGS)
      printf("  --%s%s\n", f.name, f.kind == F64 ? " <f64 expression>" : f.kind == INT ? " <integer expression>" : f.kind == PATH ? " <path>" : f.kind == VEC3 ? " <x> <y> <z>" : "");
    return 0;
  }
  if (has(flags, "convert")) { // codec utility: IN -> OUT by extension
    if (!has(flags, "convert-to")) throw UsageError("--convert IN needs --convert-to OUT");
    const std::string in = flags.at("convert").path, out = flags.at("convert-to").path;
    write_atomic(out, dumps(load(in), extension_of(out)));
    return 0;
  }
  PluginParams pp = plugin_params(flags);
  // One process per GPU (mpirun / torchrun export WORLD_SIZE, RANK, LOCAL_RANK): --num-walkers is the total, rank r runs
  // global walkers [r W/G, (r+1) W/G) on device LOCAL_RANK -- walker w is still the reference run with --seed seed+w,
  // whatever G is -- and writes its own files `name.rankRofG[-wNNNNNN].ext`.  No traffic between ranks.
  auto env_int = [](const char* name, long fallback) {
    const char* v = getenv(name);
    return v && *v ? strtol(v, nullptr, 10) : fallback;
  };
  const long world = env_int("WORLD_SIZE", 1), rank = env_int("RANK", 0), local_rank = env_int("LOCAL_RANK", rank);
  const uint32_t total_walkers = has(flags, "num-walkers") ? (uint32_t)flags.at("num-walkers").u : 1;
  if (world < 1 || total_walkers % (uint32_t)world) throw UsageError("--num-walkers " + std::to_string(total_walkers) + " does not divide over " + std::to_string(world) + " processes");
  const uint32_t n_walkers = total_walkers / (uint32_t)world;
  auto rank_path_of = [&](const std::string& path, long r) {
    if (world == 1) return path;
    const size_t slash = path.find_last_of('/');
    const size_t dot = path.find_last_of('.');
    const bool has_ext = dot != std::string::npos && (slash == std::string::npos || dot > slash);
    const std::string tag = ".rank" + std::to_string(r) + "of" + std::to_string(world);
    return has_ext ? path.substr(0, dot) + tag + path.substr(dot) : path + tag;
  };
  auto rank_path = [&](const std::string& path) { return rank_path_of(path, rank); };
  auto place = [&](sadmc_config& c) {
    if (world > 1) {
      c.n_walkers = n_walkers;
      c.walker_offset = (uint32_t)rank * n_walkers;
      if (!has(flags, "gpu-device")) c.device = (int32_t)local_rank;
    }
  };
  const std::string lib_hint = self_dir(argv[0]) + "/../libsadmc_gpu.so";
  auto known_ext = [](const std::string& p) {
    const std::string e = extension_of(p);
    return e == "yaml" || e == "json" || e == "cbor";
  };

  sadmc_config cfg;
  std::string save_as;
  std::unique_ptr<GpuEnergyMC> mc;
  Movie movies;
  bool restore_movies = false, resumed = false;
  Value movie_state;
  if (has(flags, "resume-from")) { // Params::ResumeFrom, mc/mod.rs:92-106: nothing else is read from the command line
    const std::string path = rank_path(flags.at("resume-from").path);
    if (!known_ext(path)) throw UsageError("I don't know how to read file \"" + path + "\"");
    const Value doc0 = load(walker_path(path, 0, n_walkers));
    cfg = config_from_document(doc0, n_walkers);
    if (has(flags, "bin-window-lo")) cfg.bin_window_lo = num(flags, "bin-window-lo");
    if (has(flags, "bin-window-hi")) cfg.bin_window_hi = num(flags, "bin-window-hi");
    if (has(flags, "gpu-device")) cfg.device = (int32_t)flags.at("gpu-device").u;
    place(cfg);
    const Value* sa = doc0.find("save_as");
    save_as = (n_walkers == 1 && sa && sa->kind == Value::String) ? sa->s : path;
    Report r;
    if (const Value* rep = doc0.find("report")) r.restore(*rep);
    pp = PluginParams();
    pp.has_max_iter = r.max_iter.kind == TimeToRun::TotalMoves;
    pp.max_iter = r.max_iter.n;
    pp.has_max_samples = r.has_max_samples;
    pp.max_independent_samples = r.max_independent_samples;
    pp.quiet = r.quiet;
    const Value* sv = doc0.find("save");
    const Value* sts = sv ? sv->find("save_time_seconds") : nullptr;
    pp.save_time = (sts && !sts->is_null() ? sts->as_f64() : 3600.0) / 3600.0;
    if (const Value* mv = doc0.find("movies")) {
      movie_state = *mv;
      restore_movies = true;
      const Value* mt = mv->find("movie_time");
      pp.has_movie_time = mt && !mt->is_null();
      pp.movie_time = pp.has_movie_time ? mt->as_f64() : 0.0;
    }
    if (has(flags, "dry-run")) {
      print_json(Value::map().set("resume_from", Value::string(path)).set("config", config_summary(cfg)).set("plugins", plugin_summary(pp)));
      return 0;
    }
    check_resumable(nullptr, path, n_walkers);
    mc.reset(new GpuEnergyMC(cfg, lib_hint));
    resume_into(*mc, path);
    printf("Resuming from file \"%s\"\n", path.c_str());
    resumed = true;
  } else {
    cfg = config_from_flags(flags);
    place(cfg);
    save_as = rank_path(has(flags, "save-as") ? flags.at("save-as").path : "resume.yaml"); // mc/mod.rs:88
    if (!known_ext(save_as)) throw UsageError("I don't know how to create file \"" + save_as + "\""); // mc/mod.rs:118
    const std::string first = walker_path(save_as, 0, n_walkers);
    // every rank takes the same decision from the files of ALL ranks (one node, one file system): a run resumes when
    // some rank has a checkpoint, and then every rank's set must be complete and written for this command line
    bool resuming = false;
    std::vector<std::string> all_sets;
    for (long r = 0; r < world; r++) all_sets.push_back(rank_path_of(has(flags, "save-as") ? flags.at("save-as").path : "resume.yaml", r));
    if (has(flags, "save-as"))
      for (auto& p : all_sets) resuming = resuming || file_exists(walker_path(p, 0, n_walkers)) || file_exists(partial_marker(p));
    if (resuming) {
      if (has(flags, "checkpoint-walkers") && flags.at("checkpoint-walkers").u < n_walkers)
        throw UsageError("--checkpoint-walkers writes a partial set that cannot be resumed; " + first + " exists: remove it or drop --checkpoint-walkers");
      for (auto& p : all_sets) check_resumable(&cfg, p, n_walkers);
    }
    if (has(flags, "dry-run")) {
      print_json(Value::map().set("config", config_summary(cfg)).set("plugins", plugin_summary(pp)).set("save_as", Value::string(save_as))
                     .set("resuming", Value::boolean(resuming)));
      return 0;
    }
    resumed = resuming;
    if (resuming) { // mc/mod.rs:70-84, then update_from_params (energy.rs:899-902): report + save come from the flags
      cfg.init_mode = SADMC_INIT_EXTERNAL;
      mc.reset(new GpuEnergyMC(cfg, lib_hint));
      resume_into(*mc, save_as);
      printf("Resuming from file \"%s\"\n", save_as.c_str());
      const Value doc0 = load(first);
      if (const Value* mv = doc0.find("movies")) {
        movie_state = *mv;
        restore_movies = true;
      }
    } else {
      mc.reset(new GpuEnergyMC(cfg, lib_hint));
    }
  }

  Report report;
  report.max_iter = pp.has_max_iter ? TimeToRun::total_moves(pp.max_iter) : TimeToRun::never();
  report.has_max_samples = pp.has_max_samples;
  report.max_independent_samples = pp.max_independent_samples;
  report.quiet = pp.quiet;
  Save save;
  save.has_save_time = pp.has_save_time;
  save.save_time_seconds = 3600.0 * pp.save_time;
  if (resumed) {
    report.set_resumed();
    save.set_resumed();
  }
  if (pp.has_movie_time) movies.set_movie_time(pp.movie_time);
  if (restore_movies) movies.restore(movie_state);
  const uint32_t n_save = has(flags, "checkpoint-walkers") ? (uint32_t)flags.at("checkpoint-walkers").u : mc->n_walkers();

  McView view;
  view.num_moves = [&] { return mc->num_moves(); };
  view.num_accepted_moves = [&] { return mc->num_accepted_moves() / (mc->n_walkers() ? mc->n_walkers() : 1); }; // mean per walker
  view.min_accepted_moves = [&] { return mc->min_accepted_moves(); };
  view.verify_energy = [&] { // PluginManager::run calls sys.verify_energy() before logging (plugin.rs:102-103)
    if (!mc->verify_energy(0)) throw EngineError(SADMC_ERR_VERIFY, "verify_energy failed for walker 0");
  };
  view.checkpoint = [&] { sadmc_host::save(*mc, save_as, n_save, report.document(), movies.document(), save.document()); };
  view.save_movie_frame = [&](uint64_t moves) { // Movie::save_frame, plugin.rs:434-444: <save_as stem>/<moves:014>.cbor
    const size_t dot = save_as.find_last_of('.');
    char name[32];
    snprintf(name, sizeof name, "/%014llu.cbor", (unsigned long long)moves);
    sadmc_host::save(*mc, save_as.substr(0, dot) + name, n_save, report.document(), movies.document(), save.document());
  };
  std::vector<Plugin*> plugins = {&report, &save, &movies};
  PluginManager manager;
  uint64_t launches = 0;
  const uint64_t max_launch = has(flags, "max-launch") ? flags.at("max-launch").u : ~0ull;
  for (;;) { // loop { mc.move_once() }
    uint64_t n = manager.moves_until_next_action();
    if (n > max_launch) n = max_launch;
    mc->run(n);
    launches++;
    if (manager.run(view, plugins, n) == Action::Exit) break;
  }
  if (!pp.quiet) printf("%llu moves per walker, %u walkers, %llu launches\n", (unsigned long long)mc->num_moves(), mc->n_walkers(), (unsigned long long)launches);
  return 0;
}

int main(int argc, char** argv) {
  try {
    return real_main(argc, argv);
  } catch (const UsageError& e) {
    fprintf(stderr, "error: %s\n(see --help)\n", e.what());
    return 2;
  } catch (const EngineError& e) {
    fprintf(stderr, "error (%d): %s\n", e.code, e.what());
    return 1;
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
}
