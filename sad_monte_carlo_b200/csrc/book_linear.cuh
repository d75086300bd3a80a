// book_linear.cuh -- the `binning` Monte Carlo over `binning::linear::Bins` (src/mc/binning/linear.rs), selected with
// SADMC_FLAG_BINNING | SADMC_FLAG_BINNING_LINEAR: ln w, the (f64) counts and every `extra` accumulator are spread over the
// two neighbouring bin points in proportion to the distance (increment_count 99-134, accumulate_extra 303-345) and read
// back by linear interpolation (interpret_float_index 85-98, get_total / get_count 136-149).  energy_binning.rs is the
// same code for both variants (book_binning.cuh lists what it does); what changes is every access to the bins.
//
// A correctness path, not a fast one: nothing is cached, every access goes to the walker's records in HBM (L1 / L2
// catch the reuse), the minimum of the "hist" counts is re-scanned whenever one of the two touched points held it
// (linear.rs:319-321 does the same), and only one-thread-per-walker systems are built.  No job script of the reference
// passes --linear-bin.
//
// Record layout (the 64-byte BinRec slots, all eight words f64): lnw.total, lnw.count, "energy".total, "energy".count,
// "t_found".total, "t_found".count, (unused), "hist".count.
#pragma once
#include "book_binning.cuh"

namespace sadmc {

enum { L_LNW = 0, L_CNT = 1, L_ETOT = 2, L_ECNT = 3, L_TFT = 4, L_TFC = 5, L_HC = 7 };

template <int METHOD>
struct BookL {
  const DevParams& P;
  const uint32_t w;
  double* const rec; // 8 doubles per bin point
  unsigned long long accepted;
  double acc_rate, tscale;
  int method, status;
  double too_lo, too_hi, latest_parameter, tF;
  unsigned long long tL, num_states;
  double samc_t0, wl_gamma;
  double bmin, min_e, max_e;
  int lo, len;
  double max_count; // lnw.max_count (f64, linear.rs:26)
  double tf_max;    // "t_found".max_total
  double hist_min;  // "hist".min_count: the true minimum over all points at all times
  unsigned long long hist_total;

  __device__ BookL(const DevParams& p, uint32_t walker) : P(p), w(walker), rec(reinterpret_cast<double*>(p.rec + (size_t)walker * p.cap)) {}

  __device__ void load(const WalkerRec& r) {
    accepted = r.accepted;
    acc_rate = r.acc_rate;
    tscale = r.tscale;
    method = r.method;
    status = r.status;
    too_lo = r.too_lo;
    too_hi = r.too_hi;
    latest_parameter = r.latest_parameter;
    tF = r.b_tF;
    tL = r.tL;
    num_states = r.num_states;
    samc_t0 = r.samc_t0;
    wl_gamma = r.wl_gamma;
    bmin = r.bmin;
    min_e = r.b_min_e;
    max_e = r.b_max_e;
    lo = r.lo;
    len = r.len;
    max_count = r.l_max_count;
    tf_max = r.b_tf_max;
    hist_min = r.l_hist_min;
    hist_total = r.b_hist_total;
  }
  __device__ void store(WalkerRec& r) const {
    r.accepted = accepted;
    r.acc_rate = acc_rate;
    r.tscale = tscale;
    r.method = method;
    r.status = status;
    r.too_lo = too_lo;
    r.too_hi = too_hi;
    r.latest_parameter = latest_parameter;
    r.b_tF = tF;
    r.tL = tL;
    r.num_states = num_states;
    r.samc_t0 = samc_t0;
    r.wl_gamma = wl_gamma;
    r.bmin = bmin;
    r.b_min_e = min_e;
    r.b_max_e = max_e;
    r.lo = lo;
    r.len = len;
    r.l_max_count = max_count;
    r.b_tf_max = tf_max;
    r.l_hist_min = hist_min;
    r.b_hist_total = hist_total;
  }

  __device__ __forceinline__ double& at(int i, int field) const { return rec[(size_t)(lo + i) * 8 + field]; } // reference index i
  __device__ __forceinline__ double centre(int i) const { return bmin + ((double)i + 0.5) * P.width; }        // linear.rs:200-202
  __device__ __forceinline__ double fidx(double e) const { return (e - bmin) / P.width; }                    // linear.rs:203-205

  // linear.rs:206-226.  false: the fixed window cannot hold e.
  __device__ __forceinline__ bool prep_for_e(double e) {
    bool grown = false;
    while (e < bmin) {
      if (lo == 0) return false;
      lo -= 1;
      len += 1;
      bmin -= P.width;
      grown = true;
    }
    while (e >= bmin + P.width * ((double)len - 1.0)) {
      if (lo + len >= (int)P.cap) return false;
      len += 1;
      grown = true;
    }
    if (grown && METHOD == SADMC_METHOD_WL) hist_min = 0.0; // insert_zero / push_zero: min_count = 0 (linear.rs:73-84)
    return true;
  }
  // get_total / get_count of one BinCounts field at float index f (linear.rs:85-98, 40-57, 136-149)
  __device__ __forceinline__ double interp(int field, double f) const {
    const double flen = (double)len;
    if (f < -1.0) return 0.0;
    if (f < 0.0) return 0.0 + at(0, field) * (1.0 - (-f));
    if (f < flen - 1.0) {
      const int i = (int)f;
      const double o = f - (double)i;
      double acc = 0.0;
      acc += at(i + 1, field) * o;
      acc += at(i, field) * (1.0 - o);
      return acc;
    }
    if (f < flen) return 0.0 + at(len - 1, field) * (1.0 - (f - (flen - 1.0)));
    return 0.0;
  }
  __device__ __forceinline__ double get_lnw(double e) const { return interp(L_LNW, fidx(e)); }
  __device__ __forceinline__ double get_count(double e) const { return interp(L_CNT, fidx(e)) / P.width; }

  // accumulate_extra for the "energy" / "t_found" accumulators held in the record (linear.rs:303-345); e's points exist
  __device__ __forceinline__ void accumulate(int ftot, int fcnt, double e, double value, double* running_max_total) {
    const double f = fidx(e);
    const int i = (int)f;
    const double off = f - (double)i;
    at(i, fcnt) += 1.0 - off;
    at(i + 1, fcnt) += off;
    const double t0 = at(i, ftot) + value * (1.0 - off);
    const double t1 = at(i + 1, ftot) + value * off;
    at(i, ftot) = t0;
    at(i + 1, ftot) = t1;
    if (running_max_total) {
      if (t0 > *running_max_total) *running_max_total = t0;
      if (t1 > *running_max_total) *running_max_total = t1;
    }
  }
  // the system's own data_to_collect accumulator (side arrays; its counts are f64 here, kept in the u64 array's words)
  __device__ __forceinline__ void accumulate_system_extra(double e, double value) {
    const double f = fidx(e);
    const int i = (int)f;
    const double off = f - (double)i;
    const size_t base = (size_t)w * P.cap + (size_t)lo;
    double* cnt = reinterpret_cast<double*>(P.extra_count);
    cnt[base + i] += 1.0 - off;
    cnt[base + i + 1] += off;
    P.extra_total[base + i] = P.extra_total[base + i] + value * (1.0 - off);
    P.extra_total[base + i + 1] = P.extra_total[base + i + 1] + value * off;
  }

  __device__ __forceinline__ double gamma(unsigned long long moves) const { // energy_binning.rs:507-533
    if (METHOD == SADMC_METHOD_SAD) {
      const double ns = (double)num_states;
      if (latest_parameter * tF * ns == 0.0) return 0.0;
      const double t = (double)moves;
      return (latest_parameter + t / tF) / (latest_parameter + t / ns * (t / tF));
    }
    if (METHOD == SADMC_METHOD_SAMC || method == SADMC_METHOD_SAMC) {
      const double t = (double)moves;
      return t > samc_t0 ? samc_t0 / t : 1.0;
    }
    return wl_gamma;
  }

  template <class RNG>
  __device__ __forceinline__ bool reject_move(double e1, double e2, unsigned long long moves, RNG& rng) { // energy_binning.rs:276-321
    double lnw1 = get_lnw(e1), lnw2 = get_lnw(e2);
    if (METHOD == SADMC_METHOD_SAD) {
      lnw1 = e1 < too_lo ? get_lnw(too_lo) + (e1 - too_lo) / P.min_T : (e1 > too_hi ? get_lnw(too_hi) : lnw1);
      lnw2 = e2 < too_lo ? get_lnw(too_lo) + (e2 - too_lo) / P.min_T : (e2 > too_hi ? get_lnw(too_hi) : lnw2);
    }
    const bool rejected = lnw2 > lnw1 && exp_cmp(rng.gen_f64(), lnw1 - lnw2) > 0;
    if (METHOD == SADMC_METHOD_SAD) {
      if (!rejected && get_count(e2) == 0.0 && e2 < too_hi && e2 > too_lo) tL = moves;
    }
    return rejected;
  }

  // count_states(|e, _| e >= too_lo && e <= too_hi) (energy_binning.rs:368-370): bin centres inside the range
  __device__ __forceinline__ unsigned long long centres_in_range() const {
    int first = (int)fmax(0.0, fmin(fidx(too_lo), (double)(len - 1)));
    int last = (int)fmax(0.0, fmin(fidx(too_hi), (double)(len - 1)));
    while (first > 0 && centre(first - 1) >= too_lo) first--;
    while (first < len && !(centre(first) >= too_lo)) first++;
    while (last < len - 1 && centre(last + 1) <= too_hi) last++;
    while (last >= 0 && !(centre(last) <= too_hi)) last--;
    return last >= first ? (unsigned long long)(last - first + 1) : 0ull;
  }

  __device__ __forceinline__ double hist_rescan() const { // min_of(&data.count), linear.rs:320
    double m = at(0, L_HC);
    for (int i = 1; i < len; i++) m = fmin(m, at(i, L_HC));
    return m;
  }

  // update_weights (energy_binning.rs:323-503); the vectors already hold `energy`'s points
  __device__ __forceinline__ void update_weights(double energy, unsigned long long moves, double g, double old_highest_hist, double old_hist_here) {
    { // bins.increment_count(energy, gamma): linear.rs:241-258, 99-134
      const double f = fidx(energy);
      const int i = (int)f;
      const double off = f - (double)i;
      const double v = g * 1.0 / P.width; // rescaled_gamma
      const double c0 = at(i, L_CNT) + (1.0 - off), c1 = at(i + 1, L_CNT) + off;
      at(i, L_CNT) = c0;
      at(i + 1, L_CNT) = c1;
      at(i, L_LNW) += v * (1.0 - off);
      at(i + 1, L_LNW) += v * off;
      if (c0 > max_count) max_count = c0;
      if (c1 > max_count) max_count = c1;
    }
    if (METHOD == SADMC_METHOD_SAD) {
      const double hist_here = get_count(energy);
      if (old_hist_here == 0.0) accumulate(L_TFT, L_TFC, energy, (double)moves, &tf_max);
      if (hist_here > old_highest_hist) {
        if (energy > too_hi) {
          const double v = get_lnw(too_hi);
          int i0 = (int)fmax(0.0, fmin(fidx(too_hi), (double)(len - 1))) - 1;
          if (i0 < 0) i0 = 0;
          for (int i = i0; i < len; i++) { // set_lnw, linear.rs:259-267
            const double e = centre(i);
            if (e > too_hi && get_count(e) > 0.0) {
              at(i, L_LNW) = v;
              at(i, L_CNT) = 0.0;
            }
          }
          latest_parameter = (energy - too_lo) / P.min_T;
          tL = moves;
          too_hi = energy;
          num_states = centres_in_range();
        } else if (energy < too_lo) {
          const double v = get_lnw(too_lo);
          const double old_lo = too_lo;
          int i1 = (int)fmax(0.0, fmin(fidx(too_lo), (double)(len - 1))) + 1;
          if (i1 > len - 1) i1 = len - 1;
          for (int i = 0; i <= i1; i++) {
            const double e = centre(i);
            if (e < old_lo && get_count(e) > 0.0) {
              at(i, L_LNW) = v + (e - old_lo) / P.min_T;
              at(i, L_CNT) = 0.0;
            }
          }
          latest_parameter = (too_hi - energy) / P.min_T;
          tL = moves;
          too_lo = energy;
          num_states = centres_in_range();
        }
      }
      if (tL == moves) {
        const double old_tF = tF;
        tF = tf_max;
        if (old_tF != tF && P.move_plan == SADMC_MOVE_ACCEPTANCE_RATE) {
          double s = acc_rate / P.move_value;
          s = s < 0.8 ? 0.8 : (s > 1.2 ? 1.2 : s);
          tscale *= s;
        }
      }
    } else if (METHOD == SADMC_METHOD_WL) {
      if (method == SADMC_METHOD_SAMC) return;
      const double old_lowest = hist_min / P.width;
      { // accumulate_extra("hist", energy, 0.0)
        const double f = fidx(energy);
        const int i = (int)f;
        const double off = f - (double)i;
        hist_total += 1;
        const double oc = at(i, L_HC), op = at(i + 1, L_HC);
        at(i, L_HC) = oc + (1.0 - off);
        at(i + 1, L_HC) = op + off;
        if (oc == hist_min || op == hist_min) hist_min = hist_rescan();
      }
      if (P.has_min_gamma && wl_gamma < P.min_gamma) return; // production run
      const double lowest = hist_min / P.width;
      if (lowest > old_lowest && (!P.has_min || get_count(P.min_allowed) > 0.0) && (!P.has_max || get_count(P.max_allowed) > 0.0)) {
        const double mean = (double)hist_total / (P.width * (double)len);
        if ((P.inv_t && lowest > 0.0) || lowest >= 0.8 * mean) {
          wl_gamma *= 0.5;
          for (int i = 0; i < len; i++) at(i, L_HC) = 0.0; // zero_out_extra
          hist_min = 0.0;
          hist_total = 0;
          if (P.has_min_gamma && wl_gamma < P.min_gamma) wl_gamma = 0.0;
        }
        if (P.inv_t && wl_gamma < (double)len / (double)moves) {
          method = SADMC_METHOD_SAMC;
          samc_t0 = (double)len;
        }
      }
    }
  }
};

// Method::new + Bins::new for one walker of a linear engine; as first_bin_binning, the points that move 1's
// `accumulate_extra("energy", e1, e1)` creates are created here (prep_for_e with empty vectors, linear.rs:208-225).
__device__ inline void first_bin_linear(const DevParams& P, uint32_t w, WalkerRec& r, double e0, long long kb_base, int method_param, bool writer) {
  first_bin_binning(P, w, r, e0, kb_base, method_param, writer);
  if (!writer || r.status != 0) return;
  const double k0 = floor(e0 / P.width);
  double bmin = k0 * P.width;
  long long lo = (long long)k0 - kb_base;
  int len = 0;
  while (e0 < bmin) {
    lo -= 1;
    len += 1;
    bmin -= P.width;
  }
  while (e0 >= bmin + P.width * ((double)len - 1.0) && len < 8) len += 1;
  r.bmin = bmin;
  r.len = len;
  r.lo = (int)lo;
  if (lo < 0 || lo + len > (long long)P.cap || len >= 8) {
    r.lo = 0;
    r.len = 2;
    r.status = SADMC_ERR_WINDOW;
    atomicAdd(&P.halted[0], 1u);
    return;
  }
  r.l_max_count = 0.0;
  r.l_hist_min = 0.0;
}

template <class Sys, int METHOD>
__global__ void __launch_bounds__(Sys::BLOCK, Sys::MIN_BLOCKS) move_kernel_linear(const DevParams P, unsigned long long moves0, unsigned long long n_moves) {
  static_assert(Sys::G == 1, "binning::linear is built for one-thread-per-walker systems");
  extern __shared__ __align__(16) unsigned char smem[];
  const double* zx = stage_zig<Sys>(P, smem);
  const double* zf = zx + SADMC_ZIG_TABLE_LEN;
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  const bool ghost = tid >= P.n_walkers;
  if (ghost && !Sys::COOP) return;
  const uint32_t w = ghost ? P.n_walkers - 1 : tid;
  WalkerRec& wr = P.walkers[w];
  bool halted = ghost || wr.status != 0;
  if (halted && !Sys::COOP) return;
  Sys sys(P, w, 0, group_mask<1>(), smem + zig_smem_bytes<Sys>());
  sys.load(P, w, wr);
  sys.set_cooperative(true);
  BookL<METHOD> bk(P, w);
  bk.load(wr);
  Rng rng;
  rng.s0 = wr.s0;
  rng.s1 = wr.s1;
  unsigned long long moves = moves0;
  double hr_min = wr.hr_min;
  int hr_lo = wr.hr_lo, hr_len = wr.hr_len;
  constexpr bool VERIFIES = HasVerify<Sys>::value;
#pragma unroll 1
  for (unsigned long long m = 0; m < n_moves; m++) {
    moves += 1;
    if constexpr (VERIFIES) {
      if (moves % 100000000ull == 0 && !halted && !sys.verify_energy()) {
        bk.status = SADMC_ERR_VERIFY;
        halted = true;
      }
    }
    const double e1 = sys.energy();
    if (!halted) bk.accumulate(L_ETOT, L_ECNT, e1, e1, nullptr); // accumulate_extra("energy", e1, e1); its points exist
    {
      double xv;
      if (sys.extra(moves, xv) && !halted) bk.accumulate_system_extra(e1, xv);
    }
    const double recent_scale = sqrt(1.0 / (double)moves);
    bool accepted = false;
    if (!halted) {
      bk.acc_rate *= 1.0 - recent_scale;
      double e2 = 0.0;
      if (sys.plan_move(rng, bk.tscale, zx, zf, e2)) {
        bool out_of_bounds = false;
        if (P.has_max) out_of_bounds = e2 > P.max_allowed && e2 > e1;
        if (P.has_min) out_of_bounds = out_of_bounds || (e2 < P.min_allowed && e2 < e1);
        if (!out_of_bounds && !bk.reject_move(e1, e2, moves, rng)) {
          accepted = true;
          bk.accepted += 1;
          bk.acc_rate += recent_scale;
          sys.confirm();
        }
      }
    }
    if (Sys::COOP) sys.finish_move();
    if (!halted) {
      const double energy = sys.energy();
      const double g = bk.gamma(moves);
      const double old_highest = bk.max_count / P.width; // bins.max_count()
      const double old_here = bk.get_count(energy);      // before the vectors grow
      if (energy > bk.max_e) bk.max_e = energy;
      if (energy < bk.min_e) bk.min_e = energy;
      if (accepted && !bk.prep_for_e(energy)) {
        bk.status = SADMC_ERR_WINDOW;
        halted = true;
      } else {
        if (P.hr_count && !high_resolution_increment(P, w, true, hr_min, hr_lo, hr_len, energy)) { // energy_binning.rs:328-330
          bk.status = SADMC_ERR_WINDOW;
          halted = true;
        }
        if (!halted) bk.update_weights(energy, moves, g, old_highest, old_here);
      }
    }
  }
  if (ghost) return;
  if (wr.status != 0) return;
  if (bk.status != 0) atomicAdd(&P.halted[bk.status == SADMC_ERR_VERIFY ? 1 : 0], 1u);
  bk.store(wr);
  wr.hr_min = hr_min;
  wr.hr_lo = hr_lo;
  wr.hr_len = hr_len;
  sys.store(P, w, wr, true);
  wr.s0 = rng.s0;
  wr.s1 = rng.s1;
}

} // namespace sadmc
