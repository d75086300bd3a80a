// plugins.hpp -- Report / Save / Movie scheduling around the device engine.
//
// In the reference `move_once` calls `PluginManager::run` after EVERY move (src/mc/energy.rs:967-973); the manager
// only counts until `period` moves have passed, then asks every plugin what to do and recomputes the period as
// the minimum over the plugins' `run_period()` (src/mc/plugin.rs:93-144).  Nothing observable happens in between,
// so the host runs exactly `period` moves per kernel launch and the plugins see the state they would have seen.
// Restated: Action 54-64, TimeToRun 43-51, Report 149-310, Save 313-400, Movie 402-477.
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

#include "value.hpp"

namespace sadmc_host {

enum class Action { None = 0, Log = 1, Save = 2, Exit = 3 }; // ordered: the maximum wins (plugin.rs:54-72)

struct TimeToRun { // plugin.rs:43-51
  enum Kind { Never, TotalMoves, Period } kind = Never;
  uint64_t n = 0;
  static TimeToRun never() { return TimeToRun(); }
  static TimeToRun total_moves(uint64_t n) {
    TimeToRun t;
    t.kind = TotalMoves;
    t.n = n;
    return t;
  }
  Value document() const {
    if (kind == Never) return Value::string("Never");
    Value v = Value::map();
    v.set(kind == TotalMoves ? "TotalMoves" : "Period", Value::uinteger(n));
    return v;
  }
  static TimeToRun from_document(const Value& d) {
    if (d.kind == Value::String) return never();
    TimeToRun t;
    t.kind = d.tag() == "TotalMoves" ? TotalMoves : Period;
    t.n = d.body().as_u64();
    return t;
  }
  bool operator==(const TimeToRun& o) const { return kind == o.kind && (kind == Never || n == o.n); }
};

// what the plugins call on `MonteCarlo` (mc/mod.rs:37-143)
// With many walkers the per-run quantities are per-walker quantities (every walker is one reference run):
// num_accepted_moves = mean per walker (progress line), independent_samples = the slowest walker's accepted moves.
struct McView {
  std::function<uint64_t()> num_moves, num_accepted_moves, min_accepted_moves;
  std::function<void()> checkpoint, verify_energy;
  std::function<void(uint64_t)> save_movie_frame;
  uint64_t independent_samples() const { return min_accepted_moves ? min_accepted_moves() : num_accepted_moves(); } // mc/mod.rs:134-136
};

struct Plugin {
  virtual ~Plugin() {}
  virtual Action run(McView&) { return Action::None; }
  virtual TimeToRun run_period() const { return TimeToRun::never(); }
  virtual void log(McView&) {}
  virtual void save(McView&) {}
};

inline double steady_seconds() {
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}
// the clock the plugins read; tests replace it (host/selftest_plugins.cpp)
inline std::function<double()>& plugin_clock() {
  static std::function<double()> c = steady_seconds;
  return c;
}
inline double now_seconds() { return plugin_clock()(); }

struct Report : Plugin { // plugin.rs:149-310
  TimeToRun max_iter;
  bool has_max_samples = false;
  uint64_t max_independent_samples = 0;
  bool quiet = false;
  // `start` is #[serde(skip, default)] (plugin.rs:153-155): a resumed Report has None and takes (now, moves) at its
  // first log instead of printing (262-264) -- set_resumed()
  bool has_start = true;
  double start_time = now_seconds();
  uint64_t start_moves = 0;
  void set_resumed() { has_start = false; }

  Value document() const {
    Value v = Value::map();
    v.set("max_iter", max_iter.document());
    v.set("max_independent_samples", has_max_samples ? Value::uinteger(max_independent_samples) : Value::null());
    v.set("quiet", Value::boolean(quiet));
    return v;
  }
  void restore(const Value& d) {
    max_iter = TimeToRun::from_document(d.at("max_iter"));
    const Value* s = d.find("max_independent_samples");
    has_max_samples = s && !s->is_null();
    if (has_max_samples) max_independent_samples = s->as_u64();
    quiet = d.at("quiet").as_bool();
  }
  bool am_all_done(uint64_t moves, uint64_t samples) const { // 271-283
    if (max_iter.kind == TimeToRun::TotalMoves && moves >= max_iter.n) return true;
    if (has_max_samples) return samples >= max_independent_samples;
    return false;
  }
  Action run(McView& mc) override { return am_all_done(mc.num_moves(), mc.independent_samples()) ? Action::Exit : Action::None; }
  TimeToRun run_period() const override { return max_iter; }
  void print(McView& mc) { // Report::print 206-269 (wording kept, durations in seconds)
    if (quiet) return;
    const uint64_t moves = mc.num_moves();
    if (!has_start) { // plugin.rs:262-264
      has_start = true;
      start_time = now_seconds();
      start_moves = moves;
      return;
    }
    const double dt = now_seconds() - start_time;
    const double per_move = dt / (double)(moves > start_moves ? moves - start_moves : 1);
    if (max_iter.kind == TimeToRun::TotalMoves) {
      const double left = max_iter.n > moves ? (double)(max_iter.n - moves) * per_move : 0.0;
      printf("[%14llu] %5.1f%% complete after %.1f s (%.1f s left, %.3g us per move)\n", (unsigned long long)moves,
             100.0 * (double)moves / (double)max_iter.n, dt, left, 1e6 * per_move);
    } else {
      printf("[%14llu] after %.1f s (%.3g us per move)\n", (unsigned long long)moves, dt, 1e6 * per_move);
    }
  }
  void log(McView& mc) override { print(mc); }
  void save(McView& mc) override { // 295-309
    if (quiet) return;
    const uint64_t acc = mc.num_accepted_moves(), moves = mc.num_moves();
    printf("        Accepted %.3g/%.3g = %.0f%% of the moves\n", (double)acc, (double)moves, 100.0 * (double)acc / (double)(moves ? moves : 1));
  }
};

struct Save : Plugin { // plugin.rs:313-400
  uint64_t next_output = 1;
  bool has_save_time = true;
  double save_time_seconds = 3600.0;
  // next_output and start are #[serde(skip, default)] (plugin.rs:315-320): a resumed run saves at its first tick
  // (next_output = 0), takes (now, moves) as its start there and saves next 2^20 moves later (373-376)
  bool has_start = true;
  double start_time = now_seconds();
  uint64_t start_moves = 0;
  void set_resumed() {
    has_start = false;
    next_output = 0;
  }

  Value document() const {
    Value v = Value::map();
    v.set("save_time_seconds", has_save_time ? Value::number(save_time_seconds) : Value::null());
    return v;
  }
  bool shall_i_save(uint64_t moves) { // 353-383
    if (moves < next_output) return false;
    if (has_save_time) {
      if (!has_start) { // plugin.rs:373-376
        has_start = true;
        start_time = now_seconds();
        start_moves = moves;
        next_output = moves + (1ull << 20);
        return true;
      }
      double per_move = (now_seconds() - start_time) / (double)(moves > start_moves ? moves - start_moves : 1);
      if (per_move < 1e-30) per_move = 1e-30;
      const double mpp = 1.0 + std::floor(save_time_seconds / per_move);
      const uint64_t moves_per_period = mpp < 1.8e19 ? (uint64_t)mpp : ~0ull >> 1;
      if (moves_per_period < moves)
        next_output = moves + moves_per_period;
      else if ((double)moves + 1.0 < 1.0 / per_move)
        next_output = (uint64_t)(1.0 / per_move);
      else
        next_output = moves * 2;
    } else {
      next_output *= 2;
    }
    return true;
  }
  Action run(McView& mc) override { return mc.num_moves() >= next_output ? Action::Save : Action::None; }
  TimeToRun run_period() const override { return TimeToRun::total_moves(next_output); }
  void save(McView& mc) override { shall_i_save(mc.num_moves()); }
};

struct Movie : Plugin { // plugin.rs:402-477: frame k is due at move round(movie_time ** k)
  bool has_movie_time = false;
  double movie_time = 0.0;
  int32_t which_frame = 0;
  TimeToRun period;

  void set_movie_time(double t) {
    has_movie_time = true;
    movie_time = t;
    period = TimeToRun::total_moves(1);
  }
  Value document() const {
    Value v = Value::map();
    v.set("movie_time", has_movie_time ? Value::number(movie_time) : Value::null());
    v.set("which_frame", Value::integer(which_frame));
    v.set("period", period.document());
    return v;
  }
  void restore(const Value& d) { // all three fields are serialised (403-408): a resumed run keeps its schedule
    const Value* t = d.find("movie_time");
    has_movie_time = t && !t->is_null();
    movie_time = has_movie_time ? t->as_f64() : 0.0;
    which_frame = (int32_t)d.at("which_frame").as_i64();
    period = TimeToRun::from_document(d.at("period"));
  }
  bool shall_i_save(uint64_t moves) { // 446-463
    if (has_movie_time && period == TimeToRun::total_moves(moves)) {
      int32_t which = which_frame + 1;
      uint64_t nxt = (uint64_t)(std::pow(movie_time, (double)which) + 0.5);
      while (nxt <= moves) {
        which += 1;
        nxt = (uint64_t)(std::pow(movie_time, (double)which) + 0.5);
      }
      which_frame = which;
      period = TimeToRun::total_moves(nxt);
      return true;
    }
    return false;
  }
  Action run(McView& mc) override {
    if (shall_i_save(mc.num_moves())) {
      mc.save_movie_frame(mc.num_moves());
      return Action::Save;
    }
    return Action::None;
  }
  TimeToRun run_period() const override { return period; }
};

struct PluginManager { // plugin.rs:74-144; `period` and `moves` are not serialised: a resumed run ticks after its first move
  uint64_t period = 1, moves = 0;
  uint64_t moves_until_next_action() const { return period > moves ? period - moves : 1; }
  Action run(McView& mc, const std::vector<Plugin*>& plugins, uint64_t moves_made) {
    moves += moves_made;
    if (moves < period) return Action::None;
    moves = 0;
    Action todo = Action::None;
    for (Plugin* p : plugins) {
      const Action a = p->run(mc);
      if ((int)a > (int)todo) todo = a;
    }
    if ((int)todo >= (int)Action::Log) {
      mc.verify_energy();
      for (Plugin* p : plugins) p->log(mc);
    }
    if ((int)todo >= (int)Action::Save) {
      mc.checkpoint();
      for (Plugin* p : plugins) p->save(mc);
    }
    if (todo == Action::Exit) return Action::Exit;
    uint64_t new_period = 1ull << 40; // run plugins every trillion iterations minimum
    const uint64_t now = mc.num_moves();
    for (Plugin* p : plugins) {
      const TimeToRun t = p->run_period();
      if (t.kind == TimeToRun::TotalMoves) {
        if (t.n > now && t.n - now < new_period) new_period = t.n - now;
      } else if (t.kind == TimeToRun::Period) {
        if (t.n < new_period) new_period = t.n;
      }
    }
    period = new_period;
    return todo;
  }
};

} // namespace sadmc_host
