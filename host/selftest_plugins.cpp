// selftest_plugins.cpp -- the C++ PluginManager / Report / Save / Movie driven by a scripted Monte Carlo (no engine, no
// GPU): prints one line per manager tick so that tests/test_host_cpp.py can hold the schedule against the Python host's
// (which is itself held against a literal per-move restatement of src/mc/plugin.rs:93-144).
//   selftest_plugins MAX_ITER MOVIE_TIME|none SAVE_DOUBLING(0|1) ACCEPT_EVERY [MAX_SAMPLES]
//   selftest_plugins resumed RESUMED_AT MOVES_PER_SECOND SAVE_TIME_SECONDS N_SAVES
//       a run resumed at RESUMED_AT moves under a scripted clock: prints the moves of the first N_SAVES checkpoints
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "plugins.hpp"

using namespace sadmc_host;

static int resumed_schedule(char** argv) {
  uint64_t moves = strtoull(argv[2], nullptr, 10);
  const double rate = atof(argv[3]);
  double t = 0.0;
  plugin_clock() = [&] { return t; };
  Report report;
  report.quiet = true;
  report.max_iter = TimeToRun::never();
  report.set_resumed();
  Save save;
  save.save_time_seconds = atof(argv[4]);
  save.set_resumed();
  int saves = 0;
  const int want = atoi(argv[5]);
  McView mc;
  mc.num_moves = [&] { return moves; };
  mc.num_accepted_moves = [&] { return moves / 2; };
  mc.verify_energy = [] {};
  mc.checkpoint = [&] {
    printf("checkpoint %llu\n", (unsigned long long)moves);
    saves++;
  };
  mc.save_movie_frame = [](uint64_t) {};
  std::vector<Plugin*> plugins = {&report, &save};
  PluginManager manager;
  for (int guard = 0; guard < 1000 && saves < want; guard++) {
    const uint64_t n = manager.moves_until_next_action();
    moves += n;
    t += (double)n / rate;
    manager.run(mc, plugins, n);
  }
  return saves == want ? 0 : 1;
}

int main(int argc, char** argv) {
  if (argc == 6 && strcmp(argv[1], "resumed") == 0) return resumed_schedule(argv);
  if (argc < 5) return 2;
  const uint64_t max_iter = strtoull(argv[1], nullptr, 10);
  uint64_t moves = 0;
  const uint64_t accept_every = strtoull(argv[4], nullptr, 10);
  Report report;
  report.quiet = true;
  report.max_iter = max_iter ? TimeToRun::total_moves(max_iter) : TimeToRun::never();
  if (argc > 5) {
    report.has_max_samples = true;
    report.max_independent_samples = strtoull(argv[5], nullptr, 10);
  }
  Save save;
  save.has_save_time = atoi(argv[3]) == 0; // 1: no save_time -> checkpoints at 1, 2, 4, ... (plugin.rs:353-383)
  save.save_time_seconds = 1e-12;          // with a save_time: "every period" from the first tick on (deterministic here)
  Movie movies;
  if (strcmp(argv[2], "none") != 0) movies.set_movie_time(atof(argv[2]));
  McView mc;
  mc.num_moves = [&] { return moves; };
  mc.num_accepted_moves = [&] { return moves / accept_every; };
  mc.verify_energy = [&] { printf("verify %llu\n", (unsigned long long)moves); };
  mc.checkpoint = [&] { printf("checkpoint %llu\n", (unsigned long long)moves); };
  mc.save_movie_frame = [&](uint64_t m) { printf("frame %llu\n", (unsigned long long)m); };
  std::vector<Plugin*> plugins = {&report, &save, &movies};
  PluginManager manager;
  for (int guard = 0; guard < 100000; guard++) {
    const uint64_t n = manager.moves_until_next_action();
    moves += n; // "the engine ran n moves"
    const Action a = manager.run(mc, plugins, n);
    printf("tick %llu action %d period %llu frame %d\n", (unsigned long long)moves, (int)a, (unsigned long long)manager.period, movies.which_frame);
    if (a == Action::Exit) return 0;
  }
  return 1;
}
