// book_binning.cuh -- the `binning` binary's bookkeeping on the device: `EnergyMC` of src/mc/energy_binning.rs
// over `binning::histogram::Bins` (src/mc/binning/histogram.rs), selected with SADMC_FLAG_BINNING.
//
//   reject_move     energy_binning.rs:276-321
//   update_weights  energy_binning.rs:323-503
//   gamma           energy_binning.rs:507-533
//   move_once       energy_binning.rs:592-633
//   Bins            histogram.rs:99-372 (prep_for_e 147-167, energy_to_index 135-146, increment_count 181-191,
//                   set_lnw 192-200, count_states 201-210, accumulate_extra 235-261, zero_out_extra 262-276)
//
// What differs from energy.rs (book.cuh), and therefore from `Book`:
//  * the proposal does NOT grow the bins (no prepare_for_state): ln w and the count of an energy outside the
//    vectors read as 0 (histogram.rs:82-95); bins grow where the walker actually goes (increment_count) and for
//    the energy it is at BEFORE the move (`accumulate_extra("energy", e1, e1)`, energy_binning.rs:599-600);
//  * bin edges sit on multiples of the width (min = floor(e / width) width, histogram.rs:149-151), and an energy
//    exactly on the upper edge of the last bin belongs to that bin (histogram.rs:140-142);
//  * every visit adds gamma to ln w -- also outside [too_lo, too_hi] (no undo, cf. energy.rs:535-538);
//  * too_lo / too_hi are raw energies (energy_binning.rs:365-367, 387-389); a range extension rewrites ln w of
//    EVERY visited bin beyond the old end and zeroes its count (set_lnw, 355-361 / 377-383), and num_states counts
//    the bin centres inside the new range whether visited or not (368-370);
//  * t_found and the WL histogram are `extra` accumulators ("t_found", "hist"): tF is the running maximum of the
//    t_found totals (a bin whose count was zeroed is "found" again and ADDS the move number, 345-349), WL flatness
//    compares the smallest "hist" count over ALL bins with 0.8 of the mean (462-472);
//  * no round-trip diagnostics; verify_energy every 1e8 moves (595-597).
//
// Record layout: the same 64-byte BinRec slots as book.cuh, read as
//   lo = { lnw.total f64, lnw.count u64, "energy".total f64, "energy".count u64 }
//   hi = { "t_found".total f64, "t_found".count u64, (unused), "hist".count u64 }     ("hist".total is always 0)
// The system's own data_to_collect accumulator (WCA pressure, two-wells which) stays in the side arrays.
//
// The BinCounts aggregates that the sampler never reads (min_total, max_total of ln w, e_max_*, min_count of ln w,
// max_count / min_total of the extras) are NOT maintained on the device: the reference updates them lazily with
// O(len) rescans at data-dependent moments (histogram.rs:60-80), and the host reports them as what a full rescan
// gives.  The three that the sampler does read are exact: lnw.max_count, "t_found".max_total, "hist".min_count.
#pragma once
#include "move_kernel.cuh"

namespace sadmc {

struct BLo {
  double lnw;
  unsigned long long count;
  double etot;
  unsigned long long ecount;
};
struct BHi {
  double tft;
  unsigned long long tfc, spare, hc;
};

__device__ __forceinline__ void load_brec(const BinRec* p, BLo& l, BHi& h) {
  unsigned long long a0, a1, a2, a3, b0, b1, b2, b3;
  asm volatile("ld.global.v2.u64 {%0, %1}, [%2];" : "=l"(a0), "=l"(a1) : "l"(p));
  asm volatile("ld.global.v2.u64 {%0, %1}, [%2+16];" : "=l"(a2), "=l"(a3) : "l"(p));
  asm volatile("ld.global.v2.u64 {%0, %1}, [%2+32];" : "=l"(b0), "=l"(b1) : "l"(p));
  asm volatile("ld.global.v2.u64 {%0, %1}, [%2+48];" : "=l"(b2), "=l"(b3) : "l"(p));
  l.lnw = __longlong_as_double((long long)a0);
  l.count = a1;
  l.etot = __longlong_as_double((long long)a2);
  l.ecount = a3;
  h.tft = __longlong_as_double((long long)b0);
  h.tfc = b1;
  h.spare = b2;
  h.hc = b3;
}
__device__ __forceinline__ void store_brec(BinRec* p, const BLo& l, const BHi& h) {
  BinLo a;
  a.lnw = l.lnw;
  a.hist = l.count;
  a.etot = l.etot;
  a.e2tot = __longlong_as_double((long long)l.ecount);
  BinHi b;
  b.t_found = (unsigned long long)__double_as_longlong(h.tft);
  b.rt_stamp = h.tfc;
  b.round_trips = h.spare;
  b.wl_hist = h.hc;
  p->lo = a;
  p->hi = b;
}

// `high_resolution.increment_count(energy, 0.)` (energy_binning.rs:328-330; histogram.rs:181-191 with its own min / width):
// counts only, never read by the sampler -- one fire-and-forget reduction per move.  false: its window cannot hold e.
__device__ __forceinline__ bool high_resolution_increment(const DevParams& P, uint32_t w, bool writer, double& hr_min, int& hr_lo, int& hr_len, double e) {
  const double hw = P.hr_width;
  if (hr_len == 0) { // prep_for_e on empty vectors: min = floor(e / width) width (histogram.rs:149-151)
    const double k0 = floor(e / hw);
    hr_min = k0 * hw;
    const long long l0 = (long long)k0 - P.hr_kbase;
    if (l0 < 0 || l0 >= (long long)P.hr_cap) return false;
    hr_lo = (int)l0;
  }
  while (e < hr_min) {
    if (hr_lo == 0) return false;
    hr_lo -= 1;
    hr_len += 1;
    hr_min -= hw;
  }
  while (e >= hr_min + hw * (double)hr_len) {
    if (hr_lo + hr_len >= (int)P.hr_cap) return false;
    hr_len += 1;
  }
  const double fi = (e - hr_min) / hw;
  const int idx = fi == (double)hr_len ? hr_len - 1 : (int)fi;
  if (writer) atomicAdd(P.hr_count + (size_t)w * P.hr_cap + (size_t)(hr_lo + idx), 1ull);
  return true;
}

template <int METHOD, int G>
struct BookB {
  const DevParams& P;
  const uint32_t w;
  const bool writer;
  const unsigned gmask;
  BinRec* const rec;
  // EnergyMC / Method scalars
  unsigned long long accepted;
  double acc_rate, tscale;
  int method, status;
  double too_lo, too_hi, latest_parameter, tF;
  unsigned long long tL, num_states;
  double samc_t0, wl_gamma;
  // Bins
  double bmin, min_e, max_e;
  int lo, len;
  unsigned long long max_count;   // lnw.max_count
  double tf_max;                  // "t_found".max_total
  unsigned long long hist_min;    // "hist".min_count == the true minimum over [0, len) at all times
  long long hist_nmin;            // how many bins hold it
  unsigned long long hist_total;  // "hist".total_count
  double max_S;                   // running maximum of ln w (alignment constant of the reporting fold)
  double hr_min;                  // high-resolution histogram: Bins::min and the extent of its vectors in the window
  int hr_lo, hr_len;
  // SAD: window indices of the bins that hold too_lo / too_hi and their ln w as it sits in HBM
  int ilo, ihi;
  double b_lnw_lo, b_lnw_hi;
  // cached current bin
  int ci;
  BLo c;
  BHi ch;
  double c_xtot;
  unsigned long long c_xcnt;
  bool x_dirty;

  __device__ BookB(const DevParams& p, uint32_t walker, bool is_writer, unsigned mask)
      : P(p), w(walker), writer(is_writer), gmask(mask), rec(p.rec + (size_t)walker * p.cap) {}

  __device__ __forceinline__ void sync() const {
    if (G > 1) __syncwarp(gmask);
  }
  __device__ __forceinline__ size_t side(int i) const { return (size_t)w * P.cap + (size_t)i; }

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
    max_count = r.highest_hist;
    tf_max = r.b_tf_max;
    hist_min = r.b_hist_min;
    hist_nmin = r.b_hist_nmin;
    hist_total = r.b_hist_total;
    max_S = r.max_S;
    hr_min = r.hr_min;
    hr_lo = r.hr_lo;
    hr_len = r.hr_len;
    ilo = r.ilo;
    ihi = r.ihi;
    b_lnw_lo = 0.0;
    b_lnw_hi = 0.0;
    if (METHOD == SADMC_METHOD_SAD && status == 0) {
      b_lnw_lo = rec[ilo].lo.lnw;
      b_lnw_hi = rec[ihi].lo.lnw;
    }
    ci = -1;
    x_dirty = false;
    c_xtot = 0.0;
    c_xcnt = 0;
  }
  __device__ void store(WalkerRec& r) {
    flush();
    if (!writer) return;
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
    r.highest_hist = max_count;
    r.b_tf_max = tf_max;
    r.b_hist_min = hist_min;
    r.b_hist_nmin = hist_nmin;
    r.b_hist_total = hist_total;
    r.max_S = max_S;
    r.hr_min = hr_min;
    r.hr_lo = hr_lo;
    r.hr_len = hr_len;
    r.ilo = ilo;
    r.ihi = ihi;
  }

  // ---- Bins ----------------------------------------------------------------
  __device__ __forceinline__ double centre(int j) const { return bmin + ((double)(j - lo) + 0.5) * P.width; } // histogram.rs:132-134
  // histogram.rs:135-146, shifted into the window; -1 = no such bin (get_total / get_count then read 0, 82-95)
  __device__ __forceinline__ int widx(double e) const {
    if (e < bmin) return -1;
    const double fi = (e - bmin) / P.width;
    if (fi == (double)len) return lo + len - 1;
    if (!(fi < (double)len)) return -1; // beyond the vectors (or NaN -> index 0 in Rust; a NaN energy never gets here)
    return lo + (int)fi;
  }
  // histogram.rs:147-167.  false: the fixed window cannot hold e.
  __device__ __forceinline__ bool prep_for_e(double e) {
    int grown = 0;
    bool front = false;
    while (e < bmin) {
      if (lo == 0) return false;
      lo -= 1;
      len += 1;
      bmin -= P.width;
      grown++;
      front = true;
    }
    while (e >= bmin + P.width * (double)len) {
      if (lo + len >= (int)P.cap) return false;
      len += 1;
      grown++;
    }
    if (grown) {
      if (METHOD == SADMC_METHOD_WL) { // insert_zero / push_zero of the "hist" extra: min_count = 0 (histogram.rs:48-59)
        if (hist_min > 0) {
          hist_min = 0;
          hist_nmin = grown;
        } else {
          hist_nmin += grown;
        }
      }
      // min moved by repeated subtraction: the bins that hold too_lo / too_hi are looked up again with the new
      // arithmetic, as the reference does on every get_lnw (they can only differ for an energy within rounding of an edge)
      if (METHOD == SADMC_METHOD_SAD && front) relocate_range_bins();
    }
    return true;
  }
  __device__ __forceinline__ void relocate_range_bins() {
    const int nlo = widx(too_lo), nhi = widx(too_hi);
    if (nlo >= 0 && nlo != ilo) {
      ilo = nlo;
      b_lnw_lo = rec[ilo].lo.lnw;
    }
    if (nhi >= 0 && nhi != ihi) {
      ihi = nhi;
      b_lnw_hi = rec[ihi].lo.lnw;
    }
  }

  __device__ __forceinline__ void adopt(int i, const BLo& l, const BHi& h, double xt, unsigned long long xc) {
    ci = i;
    c = l;
    ch = h;
    c_xtot = xt;
    c_xcnt = xc;
    x_dirty = false;
  }
  // the system's data_to_collect accumulator of bin i, requested together with its record
  __device__ __forceinline__ void load_extra(int i, double& xt, unsigned long long& xc) const {
    unsigned long long a, b;
    asm volatile("ld.global.u64 %0, [%1];" : "=l"(a) : "l"(P.extra_total + side(i)));
    asm volatile("ld.global.u64 %0, [%1];" : "=l"(b) : "l"(P.extra_count + side(i)));
    xt = __longlong_as_double((long long)a);
    xc = b;
  }
  __device__ __forceinline__ void load_bin(int i) {
    BLo l;
    BHi h;
    load_brec(rec + i, l, h);
    double xt = 0.0;
    unsigned long long xc = 0;
    if (P.extra_total) {
      xt = P.extra_total[side(i)];
      xc = P.extra_count[side(i)];
    }
    adopt(i, l, h, xt, xc);
  }
  __device__ __forceinline__ void flush() {
    if (METHOD == SADMC_METHOD_SAD) {
      if (ci == ilo) b_lnw_lo = c.lnw;
      if (ci == ihi) b_lnw_hi = c.lnw;
    }
    if (ci >= 0 && writer) {
      store_brec(rec + ci, c, ch);
      if (x_dirty) {
        P.extra_total[side(ci)] = c_xtot;
        P.extra_count[side(ci)] = c_xcnt;
      }
    }
    x_dirty = false;
    sync();
  }
  __device__ __forceinline__ double lnw_lo() const { return ci == ilo ? c.lnw : b_lnw_lo; } // get_lnw(too_lo)
  __device__ __forceinline__ double lnw_hi() const { return ci == ihi ? c.lnw : b_lnw_hi; } // get_lnw(too_hi)

  __device__ __forceinline__ bool high_resolution_count(double e) { return high_resolution_increment(P, w, writer, hr_min, hr_lo, hr_len, e); }

  // ---- gamma (energy_binning.rs:507-533) -------------------------------------
  __device__ __forceinline__ double gamma(unsigned long long moves) const {
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

  // ---- reject_move (energy_binning.rs:276-321) --------------------------------
  // lnw_i2 / count_i2: ln w and count of the bin e2 falls into (0 where there is none)
  template <class RNG>
  __device__ __forceinline__ bool reject_move(double e1, double e2, double lnw_i2, unsigned long long count_i2, unsigned long long moves,
                                              RNG& rng) {
    double lnw1 = c.lnw, lnw2 = lnw_i2;
    if (METHOD == SADMC_METHOD_SAD) {
      lnw1 = e1 < too_lo ? lnw_lo() + (e1 - too_lo) / P.min_T : (e1 > too_hi ? lnw_hi() : lnw1);
      lnw2 = e2 < too_lo ? lnw_lo() + (e2 - too_lo) / P.min_T : (e2 > too_hi ? lnw_hi() : lnw2);
    }
    const bool rejected = lnw2 > lnw1 && exp_cmp(rng.gen_f64(), lnw1 - lnw2) > 0;
    if (METHOD == SADMC_METHOD_SAD) {
      // get_count(e2) == 0 compares count / width with 0: true exactly when the integer count is 0
      if (!rejected && count_i2 == 0 && e2 < too_hi && e2 > too_lo) tL = moves;
    }
    return rejected;
  }

  // count_states(|e, _| e >= too_lo && e <= too_hi) (energy_binning.rs:368-370, 390-392): bin centres inside the range
  __device__ __forceinline__ unsigned long long centres_in_range() const {
    int first = ilo, last = ihi;
    // the centre of the bin that holds too_lo may lie on either side of it; neighbours cannot
    while (first > lo && centre(first - 1) >= too_lo) first--;
    while (first < lo + len && !(centre(first) >= too_lo)) first++;
    while (last < lo + len - 1 && centre(last + 1) <= too_hi) last++;
    while (last >= lo && !(centre(last) <= too_hi)) last--;
    return last >= first ? (unsigned long long)(last - first + 1) : 0ull;
  }

  // ---- SAD range extension: set_lnw + the new range (energy_binning.rs:351-393) ------
  __device__ __forceinline__ void sad_extend_range(double energy, unsigned long long moves) {
    flush();
    const int i = ci;
    if (energy > too_hi) {
      const double v = b_lnw_hi; // get_lnw(too_hi); flush() has brought it up to date
      for (int j0 = ihi; j0 < lo + len; j0 += 4) {
        unsigned long long cs[4];
#pragma unroll
        for (int u = 0; u < 4; u++) cs[u] = rec[j0 + u < lo + len ? j0 + u : lo + len - 1].lo.hist;
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int j = j0 + u;
          if (j < lo + len && centre(j) > too_hi && cs[u] > 0 && writer) {
            rec[j].lo.lnw = v;
            rec[j].lo.hist = 0;
          }
        }
      }
      latest_parameter = (energy - too_lo) / P.min_T;
      tL = moves;
      too_hi = energy;
      ihi = i;
    } else {
      const double v = b_lnw_lo;
      for (int j0 = lo; j0 <= ilo; j0 += 4) {
        unsigned long long cs[4];
#pragma unroll
        for (int u = 0; u < 4; u++) cs[u] = rec[j0 + u <= ilo ? j0 + u : ilo].lo.hist;
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int j = j0 + u;
          if (j <= ilo) {
            const double ej = centre(j);
            if (ej < too_lo && cs[u] > 0 && writer) {
              rec[j].lo.lnw = v + (ej - too_lo) / P.min_T;
              rec[j].lo.hist = 0;
            }
          }
        }
      }
      latest_parameter = (too_hi - energy) / P.min_T;
      tL = moves;
      too_lo = energy;
      ilo = i;
    }
    sync();
    num_states = centres_in_range();
    c.lnw = rec[i].lo.lnw; // the walker's own bin may have been rewritten
    c.count = rec[i].lo.hist;
    b_lnw_lo = rec[ilo].lo.lnw;
    b_lnw_hi = rec[ihi].lo.lnw;
  }

  // count of "hist" bins equal to v over [lo, lo + len) (rare: when the minimum rises)
  __device__ __forceinline__ long long hist_count_equal(unsigned long long v) const {
    long long n = 0;
    for (int j = lo; j < lo + len; j++)
      if (rec[j].hi.wl_hist == v) n++;
    return n;
  }
  // lnw.count of the bin that holds e, 0 where there is none (get_count, histogram.rs:218-221)
  __device__ __forceinline__ unsigned long long count_at(double e) const {
    const int j = widx(e);
    if (j < 0) return 0;
    return j == ci ? c.count : rec[j].lo.hist;
  }

  // ---- update_weights after increment_count (energy_binning.rs:331-495) ------------------
  // old_highest: lnw.max_count before this move's increment; old_here: the bin's count before it
  __device__ __forceinline__ void after_increment(double energy, unsigned long long moves, unsigned long long old_highest,
                                                  unsigned long long old_here) {
    if (METHOD == SADMC_METHOD_SAD) {
      if (old_here == 0) { // accumulate_extra("t_found", energy, moves as f64)
        ch.tft += (double)moves;
        ch.tfc += 1;
        if (ch.tft > tf_max) tf_max = ch.tft;
      }
      // hist_here > old_highest_hist: both are counts divided by the same width
      if ((double)c.count / P.width > (double)old_highest / P.width) {
        if (energy > too_hi || energy < too_lo) sad_extend_range(energy, moves);
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
      if (method == SADMC_METHOD_SAMC) return; // 1/t-WL after its switch (energy_binning.rs:496-498)
      const unsigned long long old_lowest = hist_min;
      // accumulate_extra("hist", energy, 0.0)
      hist_total += 1;
      if (ch.hc == hist_min) {
        hist_nmin -= 1;
        ch.hc += 1;
        if (hist_nmin == 0) { // the last bin at the minimum just left it: every bin is now >= hist_min + 1, this one equal
          hist_min += 1;
          flush();
          hist_nmin = hist_count_equal(hist_min);
        }
      } else {
        ch.hc += 1;
      }
      if (P.has_min_gamma && wl_gamma < P.min_gamma) return; // production run
      if ((double)hist_min / P.width > (double)old_lowest / P.width && (!P.has_min || count_at(P.min_allowed) > 0) &&
          (!P.has_max || count_at(P.max_allowed) > 0)) {
        const double lowest = (double)hist_min / P.width;
        const double mean = (double)hist_total / (P.width * (double)len);
        if ((P.inv_t && lowest > 0.0) || lowest >= 0.8 * mean) {
          wl_gamma *= 0.5;
          flush(); // zero_out_extra("hist")
          if (writer)
            for (int j = lo; j < lo + len; j++) rec[j].hi.wl_hist = 0;
          sync();
          ch.hc = 0;
          hist_min = 0;
          hist_nmin = len;
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

// Method::new (energy_binning.rs:150-174) and Bins::new (histogram.rs:170-180) for one walker.  The vectors of the
// reference are empty until the first move's `accumulate_extra("energy", e1, e1)` creates the bin of the starting
// energy (prep_for_e with len == 0: min = floor(e / width) width); the device creates it here -- e1 of move 1 is the
// starting energy -- and the host reports an empty `Bins` while moves == 0.  kb_base: window bin j covers
// [(kb_base + j) width, (kb_base + j + 1) width).
__device__ inline void first_bin_binning(const DevParams& P, uint32_t w, WalkerRec& r, double e0, long long kb_base, int method_param,
                                         bool writer) {
  if (!writer) return;
  r.accepted = 0;
  r.acc_rate = 0.5;
  r.tscale = P.move_plan == SADMC_MOVE_TRANSLATION_SCALE ? P.move_value : 0.05; // energy_binning.rs:572-575
  r.status = 0;
  const double k0 = floor(e0 / P.width);
  double bmin = k0 * P.width;
  long long lo = (long long)k0 - kb_base;
  int len = 0;
  while (e0 < bmin) { // histogram.rs:152-160
    lo -= 1;
    len += 1;
    bmin -= P.width;
  }
  while (e0 >= bmin + P.width * (double)len) { // histogram.rs:161-166
    len += 1;
    if (len > 4) break;
  }
  r.bmin = bmin;
  r.len = len;
  if (lo < 0 || lo + len > (long long)P.cap || !(e0 == e0) || len > 4) {
    r.lo = 0;
    r.len = 1;
    r.status = SADMC_ERR_WINDOW;
    atomicAdd(&P.halted[0], 1u);
    return;
  }
  r.lo = (int)lo;
  r.method = method_param == SADMC_METHOD_INV_T_WL ? SADMC_METHOD_WL : method_param;
  r.too_lo = e0;
  r.too_hi = e0;
  r.latest_parameter = 0.0;
  r.tL = 0;
  r.tF = 0;
  r.b_tF = 0.0;
  r.num_states = 0; // energy_binning.rs:154
  r.highest_hist = 0;
  r.tfmax = 0;
  const double fi = (e0 - bmin) / P.width;
  const int i0 = (int)lo + (fi == (double)len ? len - 1 : (int)fi);
  r.ilo = i0;
  r.ihi = i0;
  r.wl_gamma = 1.0;
  r.b_tf_max = 0.0;
  r.b_min_e = e0;
  r.b_max_e = e0;
  r.b_hist_min = 0;
  r.b_hist_nmin = len;
  r.b_hist_total = 0;
  r.max_S = 0.0;
  r.max_S_index = 0;
  r.verify_fail = 0;
  r.t_range = 0;
  r.hr_min = (round(e0 / P.hr_width) - 0.5) * P.hr_width; // Bins::new (histogram.rs:170-180); replaced by the first count
  r.hr_lo = 0;
  r.hr_len = 0;
}

// n_moves x `move_once` of energy_binning.rs:592-633 for every walker, in one launch.
template <class Sys, int METHOD>
__global__ void __launch_bounds__(Sys::BLOCK, Sys::MIN_BLOCKS) move_kernel_binning(const DevParams P, unsigned long long moves0, unsigned long long n_moves) {
  extern __shared__ __align__(16) unsigned char smem[];
  const double* zx = stage_zig<Sys>(P, smem);
  const double* zf = zx + SADMC_ZIG_TABLE_LEN;
  constexpr int G = Sys::G;
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t w_raw = tid / G;
  const int lane = (int)(tid % G);
  const bool ghost = w_raw >= P.n_walkers;
  if (ghost && !Sys::COOP) return;
  const uint32_t w = ghost ? P.n_walkers - 1 : w_raw;
  const unsigned gmask = group_mask<G>();
  WalkerRec& wr = P.walkers[w];
  bool halted = ghost || wr.status != 0;
  if (halted && !Sys::COOP) return;
  Sys sys(P, w, lane, gmask, smem + zig_smem_bytes<Sys>());
  sys.load(P, w, wr);
  sys.set_cooperative(true);
  BookB<METHOD, G> bk(P, w, lane == 0 && !ghost, gmask);
  bk.load(wr);
  Rng rng;
  rng.s0 = wr.s0;
  rng.s1 = wr.s1;
  if (!halted) {
    const int i0 = bk.widx(sys.energy());
    if (i0 < 0) {
      bk.status = SADMC_ERR_WINDOW;
      halted = true;
    } else {
      bk.load_bin(i0);
    }
  }
  unsigned long long moves = moves0;
  constexpr bool VERIFIES = HasVerify<Sys>::value;
#pragma unroll 1
  for (unsigned long long m = 0; m < n_moves; m++) {
    moves += 1; // energy_binning.rs:593
    if constexpr (VERIFIES) {
      if (moves % 100000000ull == 0 && !halted && !sys.verify_energy()) { // energy_binning.rs:595-597
        bk.status = SADMC_ERR_VERIFY;
        halted = true;
      }
    }
    const double e1 = sys.energy();
    double e2 = 0.0;
    bool accepted = false, proposing = false;
    int i2 = -1;
    BLo r2;
    BHi h2;
    r2.lnw = 0.0;
    r2.count = 0;
    r2.etot = 0.0;
    r2.ecount = 0;
    h2.tft = 0.0;
    h2.tfc = 0;
    h2.spare = 0;
    h2.hc = 0;
    double x2tot = 0.0;
    unsigned long long x2cnt = 0;
    if (!halted) {
      // accumulate_extra("energy", e1, e1) + data_to_collect (energy_binning.rs:598-603): the walker's own bin
      bk.c.etot += e1;
      bk.c.ecount += 1;
    }
    {
      double xv;
      if (sys.extra(moves, xv) && !halted) { // cooperative for the fluids: every lane calls it
        bk.c_xcnt += 1;
        bk.c_xtot += xv;
        bk.x_dirty = true;
      }
    }
    const double recent_scale = sqrt(1.0 / (double)moves);
    if (!halted) {
      bk.acc_rate *= 1.0 - recent_scale;
      if (sys.plan_move(rng, bk.tscale, zx, zf, e2)) {
        bool out_of_bounds = false;
        if (P.has_max) out_of_bounds = e2 > P.max_allowed && e2 > e1;
        if (P.has_min) out_of_bounds = out_of_bounds || (e2 < P.min_allowed && e2 < e1);
        if (!out_of_bounds) {
          proposing = true;
          i2 = bk.widx(e2);
        }
      }
    }
    const bool other_bin = proposing && i2 >= 0 && i2 != bk.ci;
    if (other_bin) { // the one HBM access of a move
      load_brec(bk.rec + i2, r2, h2);
      if constexpr (HasExtra<Sys>::value) {
        if (P.extra_total) bk.load_extra(i2, x2tot, x2cnt);
      }
    }
    const double g = bk.gamma(moves); // "compute gamma out front" (energy_binning.rs:324): reject_move only touches tL
    if (proposing) {
      const double lnw2 = i2 < 0 ? 0.0 : (other_bin ? r2.lnw : bk.c.lnw);
      const unsigned long long cnt2 = i2 < 0 ? 0ull : (other_bin ? r2.count : bk.c.count);
      if (!bk.reject_move(e1, e2, lnw2, cnt2, moves, rng)) {
        accepted = true;
        bk.accepted += 1;
        bk.acc_rate += recent_scale;
        sys.confirm();
      }
    }
    if (Sys::COOP) sys.finish_move();
    if (!halted) {
      const double energy = sys.energy();
      const unsigned long long old_highest = bk.max_count;
      // old_hist_here = get_count(energy) is looked up BEFORE increment_count grows the vectors (energy_binning.rs:326):
      // an energy exactly on the upper edge of the last bin reads that bin's count (histogram.rs:140-142), not the 0 of
      // the bin that is about to be created for it -- it matters for systems with discrete energies
      unsigned long long old_here = bk.c.count;
      if (accepted) {
        const int iold = bk.widx(energy);
        old_here = iold < 0 ? 0ull : (iold == bk.ci ? bk.c.count : (iold == i2 && other_bin ? r2.count : bk.rec[iold].lo.hist));
      }
      // increment_count(energy, gamma): histogram.rs:181-191
      if (energy > bk.max_e) bk.max_e = energy;
      if (energy < bk.min_e) bk.min_e = energy;
      if (accepted) {
        if (!bk.prep_for_e(energy)) {
          bk.status = SADMC_ERR_WINDOW;
          halted = true;
        } else {
          const int inew = bk.widx(energy);
          if (inew != bk.ci) {
            bk.flush();
            if (inew == i2 && other_bin)
              bk.adopt(inew, r2, h2, x2tot, x2cnt);
            else
              bk.load_bin(inew); // a bin that did not exist at the proposal (zeros), or E was re-summed by set_energy
          }
        }
      }
      if (!halted) {
        bk.c.lnw += g;
        bk.c.count += 1;
        if (bk.c.count > bk.max_count) bk.max_count = bk.c.count;
        if (bk.c.lnw > bk.max_S) bk.max_S = bk.c.lnw;
        if (P.hr_count && !bk.high_resolution_count(energy)) { // energy_binning.rs:328-330, before the method's part
          bk.status = SADMC_ERR_WINDOW;
          halted = true;
        }
        bk.after_increment(energy, moves, old_highest, old_here);
      }
    }
  }
  if (ghost) return;
  if (wr.status != 0) return;
  if (bk.status != 0 && lane == 0) atomicAdd(&P.halted[bk.status == SADMC_ERR_VERIFY ? 1 : 0], 1u);
  bk.store(wr);
  sys.store(P, w, wr, lane == 0);
  if (lane == 0) {
    wr.s0 = rng.s0;
    wr.s1 = rng.s1;
  }
}

} // namespace sadmc
