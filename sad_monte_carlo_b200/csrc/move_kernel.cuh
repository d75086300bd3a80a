// move_kernel.cuh -- the hot loop: n_moves x `EnergyMC::move_once`
// (src/mc/energy.rs:904-965) for every walker of one GPU, in one launch.
//
// One launch == one plugin period of the reference (`PluginManager::run`,
// src/mc/plugin.rs:93-144): walker state (system, RNG, method scalars, current
// bin record) is loaded into registers / shared memory once, advanced n_moves
// times, and stored once.  Only bin records travel to and from HBM in between.
#pragma once
#include "book.cuh"
#include "kernel_set.cuh"
#include "rng.cuh"

namespace sadmc {

constexpr int ZIG_SMEM_BYTES = 4128; // 2 * 257 doubles, padded to 32 B

// Systems that let the move loop evaluate the next proposal's draws one move ahead declare PREDRAW (rng.cuh).
template <class S, class = void>
struct HasPredraw {
  static constexpr bool value = false;
};
template <class S>
struct HasPredraw<S, decltype((void)S::PREDRAW)> {
  static constexpr bool value = S::PREDRAW;
};

// Systems with helper warps for the pair loop (sys_lj_paired.cuh) declare HELPERS.
template <class S, class = void>
struct HasHelpers {
  static constexpr bool value = false;
};
template <class S>
struct HasHelpers<S, decltype((void)S::HELPERS)> {
  static constexpr bool value = S::HELPERS;
};

// Systems that override `System::verify_energy` (lj.rs:249-261, wca.rs:237-251, optsquare.rs:199-201) declare
// VERIFIES; for the others the trait default is a no-op (system/mod.rs:85) and the cadence check compiles away.
template <class S, class = void>
struct HasVerify {
  static constexpr bool value = false;
};
template <class S>
struct HasVerify<S, decltype((void)S::VERIFIES)> {
  static constexpr bool value = S::VERIFIES;
};

// Systems that report `data_to_collect` values (energy.rs:944-946: WCA pressure, two-wells `which`) declare HAS_EXTRA:
// the move loop then requests the proposed bin's `extra` accumulators together with its record.
template <class S, class = void>
struct HasExtra {
  static constexpr bool value = false;
};
template <class S>
struct HasExtra<S, decltype((void)S::HAS_EXTRA)> {
  static constexpr bool value = S::HAS_EXTRA;
};

// Systems whose shared memory is better spent on walkers read the 4 KB ziggurat tables straight from global memory
// (they stay L1-resident) and declare ZIG_GLOBAL.
template <class S, class = void>
struct HasZigGlobal {
  static constexpr bool value = false;
};
template <class S>
struct HasZigGlobal<S, decltype((void)S::ZIG_GLOBAL)> {
  static constexpr bool value = S::ZIG_GLOBAL;
};
template <class S>
__host__ __device__ constexpr int zig_smem_bytes();

template <int G>
__device__ __forceinline__ unsigned group_mask() {
  if (G >= 32) return 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;
  return ((1u << (G & 31)) - 1u) << (lane / G * G);
}

template <class S>
__host__ __device__ constexpr int zig_smem_bytes() {
  return HasZigGlobal<S>::value ? 0 : ZIG_SMEM_BYTES;
}
template <class S>
__device__ __forceinline__ const double* stage_zig(const DevParams& P, unsigned char* smem) {
  if constexpr (HasZigGlobal<S>::value) {
    return P.zig;
  } else {
    double* z = reinterpret_cast<double*>(smem);
    for (int i = threadIdx.x; i < 2 * SADMC_ZIG_TABLE_LEN; i += blockDim.x) z[i] = P.zig[i];
    __syncthreads();
    return z;
  }
}

// Sad/WL method state at construction, `Method::new` (energy.rs:246-313), and the
// first bin (energy.rs:852, 868-882).  k_base: window bin j covers
// [(k_base + j - 0.5) width, (k_base + j + 0.5) width).
template <int G>
__device__ void first_bin(const DevParams& P, uint32_t w, WalkerRec& r, double e0, long long k_base, int method_param, bool writer) {
  const double k0 = round(e0 / P.width); // f64::round: half away from zero
  const double emin = (k0 - 0.5) * P.width;
  const long long lo = (long long)k0 - k_base;
  if (!writer) return;
  r.accepted = 0;
  r.acc_rate = 0.5;
  r.tscale = P.move_plan == SADMC_MOVE_TRANSLATION_SCALE ? P.move_value : 0.05; // energy.rs:884-887
  r.bmin = emin;
  r.len = 1;
  r.status = 0;
  if (lo < 0 || lo >= (long long)P.cap || !(e0 == e0)) {
    r.lo = 0;
    r.status = SADMC_ERR_WINDOW;
    atomicAdd(&P.halted[0], 1u);
    return;
  }
  r.lo = (int)lo;
  BinRec b;
  b.lo.lnw = 0.0;
  b.lo.hist = 1;
  b.lo.etot = e0;
  b.lo.e2tot = e0 * e0;
  b.hi.t_found = 0;
  b.hi.rt_stamp = 0;
  b.hi.round_trips = 0;
  b.hi.wl_hist = 0;
  P.rec[(size_t)w * P.cap + lo] = b;
  r.method = method_param == SADMC_METHOD_INV_T_WL ? SADMC_METHOD_WL : method_param;
  r.too_lo = e0;
  r.too_hi = e0;
  r.latest_parameter = 0.0;
  r.tL = 0;
  r.tF = 0;
  r.num_states = 1;
  r.highest_hist = 1;
  r.tfmax = 0;
  r.ilo = (int)lo;
  r.ihi = (int)lo;
  r.samc_t0 = 0.0;
  r.wl_gamma = 1.0;
  const bool both = P.has_min && P.has_max;
  r.wl_lowest = both ? 0 : 1;
  r.wl_highest = 1;
  r.wl_total = 0;
  r.wl_num_states = both ? (P.max_allowed - P.min_allowed) / P.width : 1.0;
  r.wl_min_energy = e0;
  r.wl_low_count = 0;
  r.wl_hist_len = 0;
  r.max_S = 0.0;
  r.max_S_index = 0;
  r.rt_fill_val = 0; // have_visited_since_maxentropy = [false], energy.rs:879
  r.rt_fill_time = 0;
  r.rt_fill_lo = (int)lo;
  r.rt_fill_hi = (int)lo + 1;
  r.verify_fail = 0;
  r.t_range = 0;
}

// energy_binning.rs / histogram.rs counterpart of first_bin (book_binning.cuh)
__device__ inline void first_bin_binning(const DevParams& P, uint32_t w, WalkerRec& r, double e0, long long kb_base, int method_param,
                                         bool writer);
__device__ inline void first_bin_linear(const DevParams& P, uint32_t w, WalkerRec& r, double e0, long long kb_base, int method_param, bool writer);

// `from_params` for every walker: optional randomize, the downhill relaxation
// (energy.rs:840-851), then the first bin.
template <class Sys>
__global__ void __launch_bounds__(Sys::BLOCK, Sys::MIN_BLOCKS) init_kernel(const DevParams P, unsigned long long seed0, int init_mode, long long k_base,
                                                         int method_param, double samc_t0, unsigned long long max_relax) {
  extern __shared__ __align__(16) unsigned char smem[];
  const double* zx = stage_zig<Sys>(P, smem);
  const double* zf = zx + SADMC_ZIG_TABLE_LEN;
  constexpr int G = Sys::G;
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t w = tid / G;
  const int lane = (int)(tid % G);
  if (w >= P.n_walkers) return;
  const unsigned gmask = group_mask<G>();
  WalkerRec& wr = P.walkers[w];
  Sys sys(P, w, lane, gmask, smem + zig_smem_bytes<Sys>());
  sys.load(P, w, wr);
  Rng rng;
  seed_from_u64(seed0 + (unsigned long long)w, (uint64_t*)&rng.s0, (uint64_t*)&rng.s1); // energy.rs:835, walker w <-> --seed seed0+w
  if (init_mode == SADMC_INIT_RANDOMIZE) sys.randomize(rng);
  if (P.has_max) {
#pragma unroll 1
    for (unsigned long long it = 0; it < max_relax; it++) {
      double newe;
      if (sys.plan_move(rng, 0.05, zx, zf, newe)) {
        if (newe < sys.energy()) sys.confirm();
        if (sys.energy() < P.max_allowed) break;
      }
    }
  }
  if (P.flags & SADMC_FLAG_BINNING_LINEAR)
    first_bin_linear(P, w, wr, sys.energy(), k_base, method_param, lane == 0);
  else if (P.flags & SADMC_FLAG_BINNING)
    first_bin_binning(P, w, wr, sys.energy(), k_base, method_param, lane == 0);
  else
    first_bin<G>(P, w, wr, sys.energy(), k_base, method_param, lane == 0);
  if (lane == 0) {
    wr.s0 = rng.s0;
    wr.s1 = rng.s1;
    wr.samc_t0 = samc_t0;
  }
  sys.store(P, w, wr, lane == 0);
}

// Systems whose move kernel may run a move's bookkeeping during the NEXT move declare DEFER_BOOK (see move_kernel).
template <class S, class = void>
struct HasDefer {
  static constexpr bool value = false;
};
template <class S>
struct HasDefer<S, decltype((void)S::DEFER_BOOK)> {
  static constexpr bool value = S::DEFER_BOOK;
};

// What `move_once` does after the accept test, for the state the walker is in after move `mv` (energy.rs:934-965):
// histogram and energy moments of its bin, `data_to_collect`, update_weights, round trips.  `i1_ref` = reference index
// of the bin the walker was in before the move.  own_gamma: evaluate gamma(mv) here (deferred bookkeeping: the
// method state has not changed since the accept test of move mv) instead of taking the caller's g.
template <class Sys, class BookT>
__device__ __forceinline__ void bookkeep(Sys& sys, BookT& bk, unsigned long long mv, int i1_ref, double g, bool own_gamma) {
  constexpr int METHOD = BookT::METHOD_KIND;
  const double energy = sys.energy(); // energy.rs:934
  const bool first_visit = bk.c_hist == 0;
  if (first_visit) { // energy.rs:938-940
    bk.c_tfound = mv;
    bk.hi_dirty = true;
    if (METHOD == SADMC_METHOD_SAD && bk.ci >= bk.ilo && bk.ci <= bk.ihi) bk.tfmax = mv;
  }
  bk.c_hist += 1;
  bk.c_etot += energy;
  bk.c_e2 += energy * energy;
  {
    double xv;
    if (sys.extra(mv, xv)) { // energy.rs:944-946, Bins::accumulate_extra 374-386
      bk.c_xcnt += 1;
      bk.c_xtot += xv;
      bk.x_dirty = true;
    }
  }
  if (own_gamma) g = bk.gamma(mv);
  if (METHOD == SADMC_METHOD_SAD) bk.update_weights_sad(energy, mv, g); // energy.rs:948
  if (METHOD == SADMC_METHOD_SAMC) bk.c_lnw += g;
  if (METHOD == SADMC_METHOD_WL) bk.update_weights_wl(energy, mv, first_visit, g);
#ifndef SADMC_ABL_NORT
  bk.round_trips(i1_ref, mv); // energy.rs:950-965
#endif
}

// Systems whose register budget is that of a larger CTA than the one they launch declare LAUNCH_BOUND_THREADS.
template <class S, class = void>
struct LaunchBound {
  static constexpr int threads = S::BLOCK;
};
template <class S>
struct LaunchBound<S, decltype((void)S::LAUNCH_BOUND_THREADS)> {
  static constexpr int threads = S::LAUNCH_BOUND_THREADS > S::BLOCK ? S::LAUNCH_BOUND_THREADS : S::BLOCK; // (a derived system may widen BLOCK)
};

template <class Sys, int METHOD>
__global__ void __launch_bounds__(LaunchBound<Sys>::threads, Sys::MIN_BLOCKS) move_kernel(const DevParams P, unsigned long long moves0, unsigned long long n_moves) {
  extern __shared__ __align__(16) unsigned char smem[];
  const double* zx = stage_zig<Sys>(P, smem);
  const double* zf = zx + SADMC_ZIG_TABLE_LEN;
  constexpr int G = Sys::G;
  constexpr bool HELPERS = HasHelpers<Sys>::value;
  unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  if constexpr (HELPERS) {
    // the upper half of the CTA only ever runs the pair loop for the walkers of the lower half
    if (threadIdx.x >= Sys::WALKERS_PER_BLOCK) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Sys::HELPER_REGS));
      Sys::helper_loop(P, smem + zig_smem_bytes<Sys>(), n_moves);
      return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Sys::MAIN_REGS));
    tid = blockIdx.x * Sys::WALKERS_PER_BLOCK + threadIdx.x;
  }
  const uint32_t w_raw = tid / G;
  const int lane = (int)(tid % G);
  // Systems whose rare paths are warp-cooperative (Sys::COOP) keep every thread of the
  // block alive: threads past the last walker become ghosts that only take part in
  // the cooperative steps.
  const bool ghost = w_raw >= P.n_walkers;
  if (ghost && !Sys::COOP) return;
  const uint32_t w = ghost ? P.n_walkers - 1 : w_raw;
  const unsigned gmask = group_mask<G>();
  WalkerRec& wr = P.walkers[w];
  bool halted = ghost || wr.status != 0; // a walker that left the window stays halted
  if (halted && !Sys::COOP) return;
  Sys sys(P, w, lane, gmask, smem + zig_smem_bytes<Sys>());
  sys.load(P, w, wr);
  sys.set_cooperative(true);
  Book<METHOD, G, Sys::FAST_BOOK> bk(P, w, lane == 0 && !ghost, gmask);
  bk.load(wr);
  Rng rng;
  rng.s0 = wr.s0;
  rng.s1 = wr.s1;
  if (!halted) bk.load_bin(bk.widx(sys.energy()));

  unsigned long long moves = moves0;
  // sqrt(1 / moves) of the NEXT move is evaluated one move ahead, in the shadow of the bin-record load
  // (it depends on nothing but the move counter); same IEEE operations, so the value is unchanged.
  double recent_next = sqrt(1.0 / (double)(moves0 + 1)); // energy.rs:913
  constexpr bool PREDRAW = HasPredraw<Sys>::value;
  PreDraw pre; // the draws of the coming proposal when pre.ok, evaluated during the previous move
  pre.ok = false;
  // energy.rs:907-911: `if moves % (len^2 * 1000) == 0 { system.verify_energy() }`.  The next multiple is kept and
  // recomputed only when the number of bins changes, so a move pays one compare instead of a 64-bit modulo.
  constexpr bool VERIFIES = HasVerify<Sys>::value;
  int v_len = -1;
  unsigned long long v_period = 0, v_at = 0;
  // DEFER (Sys::DEFER_BOOK, SAD, fixed translation scale): the bookkeeping of move m -- histogram and moments, gamma,
  // update_weights, round trips: ~500 dependent scalar instructions that touch no memory -- is run during move m + 1,
  // in the shadow of that move's bin-record load, instead of between the accept test and the next proposal.  Nothing
  // the proposal needs (configuration, energy, generator, translation scale) is touched by it; the accept test of move
  // m + 1 comes after it, so ln w, the SAD range and gamma are what they would have been.  The record requested before
  // the bookkeeping ran is requested again in the one case in which the bookkeeping rewrites bins in HBM (a SAD range
  // extension, energy.rs:544-584); a proposal that needs new bins (prepare_for_state would grow the vectors) waits
  // with its request until the bookkeeping is done, because round_trips works on the old extent.
  constexpr bool DEFER = HasDefer<Sys>::value && METHOD == SADMC_METHOD_SAD && !HasPredraw<Sys>::value && !HELPERS;
  // (the engine only selects a DEFER kernel for runs with a fixed translation scale: with MoveParams::AcceptanceRate the
  // bookkeeping can change the step size the next proposal uses, energy.rs:606-616)
  bool pend = false; // the previous move's bookkeeping is still to do
  int pend_i1 = 0;   // ... with this reference index of the bin the walker was in before that move
#pragma unroll 1
  for (unsigned long long m = 0; m < n_moves + (DEFER ? 1ull : 0ull); m++) {
    const bool epilogue = DEFER && m == n_moves; // one extra pass that only settles the last move's bookkeeping
    if (!epilogue) moves += 1; // energy.rs:905
    if constexpr (VERIFIES) {
      if (!epilogue) {
        if (bk.len != v_len) {
          v_len = bk.len;
          v_period = (unsigned long long)v_len * (unsigned long long)v_len * 1000ull;
          v_at = ((moves - 1) / v_period + 1) * v_period; // smallest multiple >= moves
        }
        if (moves == v_at) {
          v_at += v_period;
          if (!halted && !sys.verify_energy()) { // the reference panics here (lj.rs:259, wca.rs:248, optsquare.rs:200)
            bk.status = SADMC_ERR_VERIFY;
            halted = true;
          }
        }
      }
    }
    const double e1 = sys.energy();
    const int i1 = bk.ci;
    const double recent_scale = recent_next;
    double e2 = 0.0;
    bool accepted = false, proposing = false, need_bins = false;
    int i2 = i1;
    BinLo r2;
    BinHi h2;
    // ln w and histogram of the bin the proposal lands in: the current bin's cached values unless the load below
    // replaces them (no select on the loaded registers right behind the load: it would stall the warp there for the
    // whole DRAM latency instead of at the accept test)
    r2.lnw = bk.c_lnw;
    r2.hist = bk.c_hist;
    r2.etot = 0.0;
    r2.e2tot = 0.0;
    h2.t_found = 0;
    h2.rt_stamp = 0;
    h2.round_trips = 0;
    h2.wl_hist = 0;
    bool some_paired = false;
    if constexpr (HELPERS) some_paired = sys.plan_move_paired(!halted, rng, bk.tscale, zx, zf, e2); // every thread: barriers inside
    if (!halted && !epilogue) {
      bk.acc_rate *= 1.0 - recent_scale;
      bool some;
      if constexpr (HELPERS) {
        some = some_paired;
      } else if constexpr (PREDRAW) {
        if (pre.ok) {
          rng.s0 = pre.s0;
          rng.s1 = pre.s1;
          some = sys.plan_move_drawn((int)pre.which, pre.v0, pre.v1, pre.v2, bk.tscale, e2);
        } else {
          some = sys.plan_move(rng, bk.tscale, zx, zf, e2);
        }
      } else {
        some = sys.plan_move(rng, bk.tscale, zx, zf, e2);
      }
      if (some) { // energy.rs:915
        bool out_of_bounds = false;
        if (P.has_max) out_of_bounds = e2 > P.max_allowed && e2 > e1;
        if (P.has_min) out_of_bounds = out_of_bounds || (e2 < P.min_allowed && e2 < e1);
        if (!out_of_bounds) {
          // with bookkeeping pending, a proposal outside the current extent of the bins waits for it (see above)
          need_bins = DEFER && pend && (e2 < bk.bmin || e2 >= bk.bmin + P.width * (double)bk.len);
          if (!need_bins) {
            if (!bk.prepare_for_state(e2)) { // energy.rs:925
              bk.status = SADMC_ERR_WINDOW;
              halted = true;
            } else {
              i2 = bk.widx(e2);
              proposing = true;
            }
          }
        }
      }
    }
    // The one access of a move that goes to HBM: the record of the proposed bin.  Both sectors are
    // requested together (one DRAM access, no second dependent load if the walker moves there) ...
#ifdef SADMC_ABL_NOLOAD /* ablation experiment only */
    bool other_bin = false;
#else
    bool other_bin = proposing && i2 != i1;
#endif
    double x2tot = 0.0;
    unsigned long long x2cnt = 0;
    if (other_bin) {
      load_rec(bk.rec + i2, r2, h2);
      if constexpr (HasExtra<Sys>::value) {
        if (P.extra_total) bk.load_extra(i2, x2tot, x2cnt);
      }
    }
    // ... and what does not depend on it is computed while it is in flight: next move's sqrt(1/moves),
#ifdef SADMC_EXP_RSQRT /* experiment: tolerance tier only (acceptance_rate is a diagnostic unless the move plan is AcceptanceRate) */
    recent_next = Sys::FAST_BOOK ? rsqrt_newton((double)(moves + 1)) : sqrt(1.0 / (double)(moves + 1));
#else
    recent_next = sqrt(1.0 / (double)(moves + 1));
#endif
    // the previous move's bookkeeping (DEFER),
    if constexpr (DEFER) {
      if (pend) {
        pend = false;
        bk.wrote_bins = false;
        bookkeep(sys, bk, moves - (epilogue ? 0ull : 1ull), pend_i1, 0.0, true);
        if (need_bins) { // now the vectors may grow (energy.rs:925)
          if (!bk.prepare_for_state(e2)) {
            bk.status = SADMC_ERR_WINDOW;
            halted = true;
          } else {
            i2 = bk.widx(e2);
            proposing = true;
#ifndef SADMC_ABL_NOLOAD
            other_bin = i2 != i1;
#endif
            if (other_bin) {
              load_rec(bk.rec + i2, r2, h2);
              if constexpr (HasExtra<Sys>::value) {
                if (P.extra_total) bk.load_extra(i2, x2tot, x2cnt);
              }
            }
          }
        } else if (bk.wrote_bins && other_bin) {
          load_rec(bk.rec + i2, r2, h2); // the range extension may have rewritten ln w of the requested bin
        }
      }
      if (epilogue) break;
    }
    // and this move's gamma (energy.rs:799-824; re-evaluated below in the rare case that reject_move's
    // first-visit hook changes its inputs).
    double gamma_now = bk.gamma(moves);
    // ... and the next proposal's random numbers, for both places the stream can be at after the accept test
    PreDraw pre_b;
    unsigned long long rs0 = 0, rs1 = 0;
    if constexpr (PREDRAW) {
      pre.ok = false;
      pre_b.ok = false;
      if (!halted) {
        predraw_both(rng, sys.predraw_n(), sys.predraw_zone(), zx, zf, pre, pre_b);
        rs0 = rng.s0;
        rs1 = rng.s1;
      }
    }
    if (proposing) {
      // DEFER: the bookkeeping that just ran has updated the current bin's cached record (the load has long landed)
      const double lnw2 = DEFER ? (other_bin ? r2.lnw : bk.c_lnw) : r2.lnw;
      const unsigned long long hist2 = DEFER ? (other_bin ? r2.hist : bk.c_hist) : r2.hist;
      const unsigned long long tL_before = bk.tL;
      if (!bk.reject_move(e1, e2, i2, lnw2, hist2, moves, rng)) { // energy.rs:927-931
        accepted = true;
        bk.accepted += 1;
        bk.acc_rate += recent_scale;
        sys.confirm();
      }
      if (METHOD == SADMC_METHOD_SAD && bk.tL != tL_before) gamma_now = bk.gamma(moves); // energy.rs:466-483 fired
    }
    if constexpr (PREDRAW) {
      if (rng.s0 != rs0 || rng.s1 != rs1) pre = pre_b; // the accept test drew its uniform: the proposal starts one word later
    }
    if (Sys::COOP) sys.finish_move(); // converged point: warp-cooperative energy recomputation
    if (accepted) {
      const double e_now = sys.energy();
      const int inew = e_now == e2 ? i2 : bk.widx(e_now); // set_energy may have recomputed E (lj.rs:117-120)
      if (inew != i1) {
        bk.flush();
        if (inew == i2) {
          if (HasExtra<Sys>::value && P.extra_total)
            bk.adopt_bin(i2, r2, h2, x2tot, x2cnt);
          else
            bk.adopt_bin(i2, r2, h2);
        } else if (inew < bk.lo || inew >= bk.lo + bk.len) {
          bk.status = SADMC_ERR_WINDOW;
          halted = true;
        } else {
          bk.load_bin(inew);
        }
      }
    }
    if (!halted) {
      if constexpr (DEFER) { // settled during the next move (or by the epilogue pass): ONE copy of the bookkeeping code
        pend = true;
        pend_i1 = i1 - bk.lo;
      } else {
        bookkeep(sys, bk, moves, i1 - bk.lo, gamma_now, false);
      }
    }
  }
  if (ghost) return;
  if (wr.status != 0) return; // was halted before this launch: leave its state alone
  if (bk.status != 0 && lane == 0) atomicAdd(&P.halted[bk.status == SADMC_ERR_VERIFY ? 1 : 0], 1u); // surfaced by sadmc_sync
  bk.store(wr);
  sys.store(P, w, wr, lane == 0);
  if (lane == 0) {
    wr.s0 = rng.s0;
    wr.s1 = rng.s1;
  }
}

// ---- trait-shaped single-walker shims (src/system/mod.rs:54-120) ---------------
template <class Sys>
__global__ void __launch_bounds__(Sys::BLOCK) shim_kernel(const DevParams P, uint32_t w, int op, double arg, ShimOut* out, double* pending) {
  extern __shared__ __align__(16) unsigned char smem[];
  const double* zx = stage_zig<Sys>(P, smem);
  const double* zf = zx + SADMC_ZIG_TABLE_LEN;
  constexpr int G = Sys::G;
  const int lane = (int)threadIdx.x;
  if (lane >= G) return;
  const unsigned gmask = group_mask<G>();
  WalkerRec& wr = P.walkers[w];
  Sys sys(P, w, lane, gmask, smem + zig_smem_bytes<Sys>());
  sys.load(P, w, wr);
  Rng rng;
  rng.s0 = wr.s0;
  rng.s1 = wr.s1;
  ShimOut o;
  o.value = 0.0;
  o.some = 1;
  o.ok = 1;
  switch (op) {
    case OP_ENERGY: o.value = sys.energy(); break;
    case OP_COMPUTE_ENERGY: o.value = sys.compute_energy(); break;
    case OP_VERIFY: o.ok = sys.verify_energy() ? 1 : 0; break;
    case OP_PLAN_MOVE: {
      double e2 = 0.0;
      o.some = sys.plan_move(rng, arg, zx, zf, e2) ? 1 : 0;
      o.value = e2;
      sys.get_pending(pending, lane == 0, o.some != 0);
      if (lane == 0) {
        wr.s0 = rng.s0;
        wr.s1 = rng.s1;
      }
      break;
    }
    case OP_RANDOMIZE: // System::randomize (system/mod.rs:59), driven by the walker's own generator
      o.value = sys.randomize(rng);
      sys.store(P, w, wr, lane == 0);
      if (lane == 0) {
        wr.s0 = rng.s0;
        wr.s1 = rng.s1;
        pending[0] = 0.0;
      }
      break;
    case OP_CONFIRM:
      if (sys.set_pending(pending)) sys.confirm();
      sys.store(P, w, wr, lane == 0);
      if (lane == 0) pending[0] = 0.0; // Change::None afterwards (lj.rs:344)
      break;
  }
  if (lane == 0) *out = o;
}

} // namespace sadmc
