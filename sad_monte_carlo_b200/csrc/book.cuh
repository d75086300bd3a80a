// book.cuh -- flat-histogram bookkeeping of one walker, on the device.
//
// Device form of the reference's `EnergyMC` fields and of
//   prepare_for_state  src/mc/energy.rs:400-434
//   reject_move        src/mc/energy.rs:440-512
//   update_weights     src/mc/energy.rs:514-761
//   gamma              src/mc/energy.rs:799-824 (+ SadVersion::compute_gamma 21-39)
//   move_once          src/mc/energy.rs:904-965 (driver in move_kernel.cuh)
//
// Layout decisions (B200):
//  * The reference grows five parallel Vecs (front inserts included).  Here each
//    walker owns a FIXED window of `cap` bins in HBM; `lo`/`len`/`bmin` track which
//    part of the window the reference's vectors would currently occupy, so every
//    `bins.len()`-dependent rule is reproduced and `bmin` is decremented by
//    `width` per front insert exactly as energy.rs:419 does.
//  * lnw / histogram / energy_total / energy_squared_total of a bin share one
//    32-byte record = one HBM sector; t_found and the round-trip arrays are side
//    arrays that are touched only on first visits / bin changes.
//  * The record of the walker's CURRENT bin is cached in registers and written
//    back only when the walker leaves the bin (most proposals are rejected or
//    stay in the bin), so a rejected move costs one 32-byte read (lnw of the
//    proposed bin) and no write.
//  * All G lanes that cooperate on a walker hold identical copies of the scalars
//    and execute the bookkeeping redundantly; only lane 0 stores to HBM.
#pragma once
#include <stdint.h>

#include "../../include/sadmc_gpu.h"
#include "../../include/sadmc_math.h"
#include "fastmath.cuh"

namespace sadmc {

// One energy bin of one walker: 64 bytes = two 32-byte sectors of the same 64-byte DRAM atom.
// The first sector is what every proposal needs (lnw, histogram) and what every move updates;
// the second holds what is touched on first visits / bin changes only.
struct __align__(32) BinLo {
  double lnw;
  unsigned long long hist;
  double etot;
  double e2tot;
};
struct __align__(32) BinHi {
  unsigned long long t_found;
  unsigned long long rt_stamp;    // move at which have_visited_since_maxentropy[i] was last set individually
  unsigned long long round_trips; // stored minus one (new bins start at 1, energy.rs:418,432)
  unsigned long long wl_hist;     // Method::WL::hist
};
struct __align__(64) BinRec {
  BinLo lo;
  BinHi hi;
};

// Both sectors of a record in one go.  The asm is volatile so that the compiler cannot sink the
// second load to its first use (it did: the `hi` half is only consumed when the walker moves into
// the bin, and a load issued there exposes a second full DRAM latency per accepted move).
__device__ __forceinline__ void load_rec(const BinRec* p, BinLo& l, BinHi& h) {
  unsigned long long a0, a1, a2, a3, b0, b1, b2, b3;
  asm volatile("ld.global.v2.u64 {%0, %1}, [%2];" : "=l"(a0), "=l"(a1) : "l"(p));
  asm volatile("ld.global.v2.u64 {%0, %1}, [%2+16];" : "=l"(a2), "=l"(a3) : "l"(p));
  asm volatile("ld.global.v2.u64 {%0, %1}, [%2+32];" : "=l"(b0), "=l"(b1) : "l"(p));
  asm volatile("ld.global.v2.u64 {%0, %1}, [%2+48];" : "=l"(b2), "=l"(b3) : "l"(p));
  l.lnw = __longlong_as_double((long long)a0);
  l.hist = a1;
  l.etot = __longlong_as_double((long long)a2);
  l.e2tot = __longlong_as_double((long long)a3);
  h.t_found = b0;
  h.rt_stamp = b1;
  h.round_trips = b2;
  h.wl_hist = b3;
}

// Per-walker scalars as they sit in HBM between launches.
struct __align__(16) WalkerRec {
  unsigned long long s0, s1; // rng
  unsigned long long accepted;
  double acc_rate, tscale;
  double E, err; // system energy cache (+ error estimate for LJ/WCA)
  double bmin;
  int lo, len;
  int method, status;
  // Sad
  double too_lo, too_hi, latest_parameter;
  unsigned long long tL, tF, num_states, highest_hist;
  unsigned long long tfmax; // max(t_found[ilo..=ihi]), maintained incrementally
  int ilo, ihi;             // window indices of too_lo / too_hi
  // Samc
  double samc_t0;
  // WL
  double wl_gamma, wl_num_states, wl_min_energy;
  unsigned long long wl_lowest, wl_highest, wl_total;
  long long wl_low_count; // visited bins with hist <= wl_lowest (flatness test in O(1))
  int wl_hist_len, wl_pad;
  // round trips
  double max_S;
  int max_S_index;   // REFERENCE index (not shifted on front inserts, as in the reference)
  int rt_fill_val;   // value of the last bulk fill of have_visited_since_maxentropy
  unsigned long long rt_fill_time;
  int rt_fill_lo, rt_fill_hi; // window extent covered by that fill
  // system extras
  double d_squared; // two-wells
  unsigned long long verify_fail;
  // engine-side diagnostic, not part of the reference's state: the last move at which the BIN INDICES of the SAD range
  // (ilo, ihi) changed.  (`tL` cannot serve: the reference refreshes it whenever the end bin, which is visited twice as
  // often as its neighbours, sets a new histogram record from the half that lies outside the range, energy.rs:540-584.)
  // Written straight to HBM from the rare path; 0 after a resume.
  unsigned long long t_range;
  // SADMC_FLAG_BINNING (book_binning.cuh): Method::Sad::tF is an f64 there; aggregates of the `extra` BinCounts that the
  // sampler reads ("t_found".max_total, "hist".min_count with the number of bins that hold it, "hist".total_count) and
  // Bins::min_e / max_e
  double b_tF, b_tf_max, b_min_e, b_max_e;
  unsigned long long b_hist_min, b_hist_total;
  long long b_hist_nmin;
  // the optional high-resolution histogram (energy_binning.rs:124-125): its Bins::min and the part of its window in use
  double hr_min;
  int hr_lo, hr_len;
  // SADMC_FLAG_BINNING_LINEAR (book_linear.cuh): lnw.max_count and "hist".min_count are f64 there
  double l_max_count, l_hist_min;
};

struct DevParams {
  BinRec* rec;
  double* extra_total;
  unsigned long long* extra_count;
  WalkerRec* walkers;
  double* sys;         // [n_walkers][sys_stride] f64 system image (LJ/WCA/SW/fake/two-wells)
  uint32_t* sys_words; // Ising: [n_walkers][ising_words] packed spins
  const double* zig;   // X[257] then F[257]
  double* zstream;     // LJ move kernels with the z coordinates streamed from L2 (sys_lj_thread.cuh, ZG): 32 doubles per thread, else null
  unsigned int* halted; // [0] walkers that left the bin window, [1] walkers whose verify_energy failed (since creation)
  unsigned long long* hr_count; // [n_walkers][hr_cap] counts of the high-resolution histogram, null when there is none
  double hr_width;
  long long hr_kbase; // its window bin j covers [(hr_kbase + j) hr_width, (hr_kbase + j + 1) hr_width)
  uint32_t hr_cap;
  uint32_t n_walkers, cap, sys_stride, ising_words;
  double width;
  int has_min, has_max;
  double min_allowed, max_allowed;
  double min_T;
  int inv_t, has_min_gamma;
  int method_kind; // sadmc_method_kind of the run (host-side kernel choice; the kernels are compiled per method)
  double min_gamma, canonical_T;
  int move_plan;
  double move_value;
  uint32_t flags;
  // system parameters
  uint32_t N;
  double lj_R2, lj_R;
  double box[3];
  int ncell[3];
  double r_cut2, well2;
  int fake_fn, fake_dim;
  double fake_a, fake_b, fake_e1, fake_e2, fake_sigma;
  double tw_h2h1, tw_r2, tw_rw;
  double erfinv_mean;
  unsigned long long zone_a, zone_b; // precomputed integer-sampling zones
};

// Rust `x as usize` for f64 (saturating, NaN -> 0), clipped to int range.
__device__ __forceinline__ int f64_as_index(double x) {
  // cvt.rzi.s32.f64 saturates and maps NaN to 0 by itself; only negative values need the clamp (branch-free)
  return max(__double2int_rz(x), 0);
}

// FAST (tolerance tier, SADMC_FLAG_FAST_MATH): bin indices by multiplication with 1/width and the SAD
// gamma with cached reciprocals of tF and num_states instead of three IEEE divides per move.  Results
// differ from the reference's arithmetic by a few ulp (gamma) / for energies within an ulp of a bin
// edge (index); the bit-exact tier never uses it.
template <int METHOD, int G, bool FAST = false>
struct Book {
  static constexpr int METHOD_KIND = METHOD;
  const DevParams& P;
  const uint32_t w;
  const bool writer; // lane 0 of the walker's group
  const unsigned gmask;
  BinRec* const rec;
  // persistent
  unsigned long long accepted;
  double acc_rate, tscale, bmin;
  int lo, len, method, status;
  double too_lo, too_hi, latest_parameter;
  unsigned long long tL, tF, num_states, highest_hist, tfmax;
  int ilo, ihi;
  double samc_t0;
  double wl_gamma, wl_num_states, wl_min_energy;
  unsigned long long wl_lowest, wl_highest, wl_total;
  long long wl_low_count;
  int wl_hist_len;
  double max_S;
  int max_S_index, rt_fill_val, rt_fill_lo, rt_fill_hi;
  unsigned long long rt_fill_time;
  // cached current bin
  int ci;
  double c_lnw, c_etot, c_e2;
  unsigned long long c_hist, c_tfound, c_stamp, c_rt, c_wlh;
  bool c_visited, hi_dirty;
  // cached `extra` BinCounts of the current bin (energy.rs:136-142, 374-386)
  double c_xtot;
  unsigned long long c_xcnt;
  bool x_dirty;
  // SAD: ln w of the two boundary bins (ilo, ihi) as they sit in HBM; reject_move needs them for every walker
  // outside [too_lo, too_hi] and a dependent global load there would sit on the critical path of the move
  double b_lnw_lo, b_lnw_hi;
  bool wrote_bins; // set by the rare paths that rewrite other bins' records in HBM (move_kernel's deferred bookkeeping)
  // FAST only
  double inv_width, inv_tF, inv_ns, inv_min_T;
  unsigned long long g_tF, g_ns;

  __device__ Book(const DevParams& p, uint32_t walker, bool is_writer, unsigned mask)
      : P(p), w(walker), writer(is_writer), gmask(mask), rec(p.rec + (size_t)walker * p.cap), inv_width(1.0 / p.width), inv_tF(0.0),
        inv_ns(0.0), inv_min_T(1.0 / p.min_T), g_tF(0), g_ns(0) {
    wrote_bins = false;
  }

  __device__ __forceinline__ void sync() const {
    if (G > 1) __syncwarp(gmask);
  }
  __device__ __forceinline__ size_t side(int i) const { return (size_t)w * P.cap + (size_t)i; }

  __device__ void load(const WalkerRec& r) {
    accepted = r.accepted;
    acc_rate = r.acc_rate;
    tscale = r.tscale;
    bmin = r.bmin;
    lo = r.lo;
    len = r.len;
    method = r.method;
    status = r.status;
    too_lo = r.too_lo;
    too_hi = r.too_hi;
    latest_parameter = r.latest_parameter;
    tL = r.tL;
    tF = r.tF;
    num_states = r.num_states;
    highest_hist = r.highest_hist;
    tfmax = r.tfmax;
    ilo = r.ilo;
    ihi = r.ihi;
    if (METHOD == SADMC_METHOD_SAD) {
      b_lnw_lo = rec[ilo].lo.lnw;
      b_lnw_hi = rec[ihi].lo.lnw;
    }
    samc_t0 = r.samc_t0;
    wl_gamma = r.wl_gamma;
    wl_num_states = r.wl_num_states;
    wl_min_energy = r.wl_min_energy;
    wl_lowest = r.wl_lowest;
    wl_highest = r.wl_highest;
    wl_total = r.wl_total;
    wl_low_count = r.wl_low_count;
    wl_hist_len = r.wl_hist_len;
    max_S = r.max_S;
    max_S_index = r.max_S_index;
    rt_fill_val = r.rt_fill_val;
    rt_fill_time = r.rt_fill_time;
    rt_fill_lo = r.rt_fill_lo;
    rt_fill_hi = r.rt_fill_hi;
    ci = -1;
    hi_dirty = false;
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
    r.bmin = bmin;
    r.lo = lo;
    r.len = len;
    r.method = method;
    r.status = status;
    r.too_lo = too_lo;
    r.too_hi = too_hi;
    r.latest_parameter = latest_parameter;
    r.tL = tL;
    r.tF = tF;
    r.num_states = num_states;
    r.highest_hist = highest_hist;
    r.tfmax = tfmax;
    r.ilo = ilo;
    r.ihi = ihi;
    r.samc_t0 = samc_t0;
    r.wl_gamma = wl_gamma;
    r.wl_num_states = wl_num_states;
    r.wl_min_energy = wl_min_energy;
    r.wl_lowest = wl_lowest;
    r.wl_highest = wl_highest;
    r.wl_total = wl_total;
    r.wl_low_count = wl_low_count;
    r.wl_hist_len = wl_hist_len;
    r.max_S = max_S;
    r.max_S_index = max_S_index;
    r.rt_fill_val = rt_fill_val;
    r.rt_fill_time = rt_fill_time;
    r.rt_fill_lo = rt_fill_lo;
    r.rt_fill_hi = rt_fill_hi;
  }

  // ---- bins ------------------------------------------------------------
  // Bins::state_to_index (energy.rs:371-373) shifted into the window.
  __device__ __forceinline__ int widx(double e) const {
    return lo + f64_as_index(FAST ? (e - bmin) * inv_width : (e - bmin) / P.width);
  }
  // Bins::index_to_state (energy.rs:366-370) for window index j.
  __device__ __forceinline__ double centre(int j) const { return bmin + ((double)(j - lo) + 0.5) * P.width; }

  __device__ __forceinline__ bool visited_flag(int i, unsigned long long stamp) const {
    if (P.flags & SADMC_FLAG_NO_ROUND_TRIPS) return true;
    if (i < rt_fill_lo || i >= rt_fill_hi) return true; // created after the last bulk fill (energy.rs:417,431)
    return stamp > rt_fill_time ? true : (rt_fill_val != 0);
  }
  // the `extra` accumulators of bin i (data_to_collect systems), requested together with its record: volatile like
  // load_rec, so that the loads are not sunk to adopt_bin (a second DRAM latency per accepted move)
  __device__ __forceinline__ void load_extra(int i, double& xt, unsigned long long& xc) const {
    unsigned long long a, b;
    asm volatile("ld.global.u64 %0, [%1];" : "=l"(a) : "l"(P.extra_total + side(i)));
    asm volatile("ld.global.u64 %0, [%1];" : "=l"(b) : "l"(P.extra_count + side(i)));
    xt = __longlong_as_double((long long)a);
    xc = b;
  }
  // make bin i the cached current bin from an already loaded record (+ already loaded extras, if the system has any)
  __device__ __forceinline__ void adopt_bin(int i, const BinLo& l, const BinHi& h, double xt, unsigned long long xc) {
    adopt_bin(i, l, h, false);
    c_xtot = xt;
    c_xcnt = xc;
  }
  __device__ __forceinline__ void adopt_bin(int i, const BinLo& l, const BinHi& h, bool fetch_extra = true) {
    ci = i;
    c_lnw = l.lnw;
    c_hist = l.hist;
    c_etot = l.etot;
    c_e2 = l.e2tot;
    c_tfound = h.t_found;
    c_stamp = h.rt_stamp;
    c_rt = h.round_trips;
    c_wlh = h.wl_hist;
    c_visited = visited_flag(i, h.rt_stamp);
    hi_dirty = false;
    if (fetch_extra && P.extra_total) {
      c_xtot = P.extra_total[side(i)];
      c_xcnt = P.extra_count[side(i)];
    }
    x_dirty = false;
  }
  __device__ __forceinline__ void load_bin(int i) {
    BinLo l;
    BinHi h;
    load_rec(rec + i, l, h);
    adopt_bin(i, l, h);
  }
  __device__ __forceinline__ void flush() {
    if (METHOD == SADMC_METHOD_SAD) {
      if (ci == ilo) b_lnw_lo = c_lnw;
      if (ci == ihi) b_lnw_hi = c_lnw;
    }
    if (ci >= 0 && writer) {
      BinLo l;
      l.lnw = c_lnw;
      l.hist = c_hist;
      l.etot = c_etot;
      l.e2tot = c_e2;
      rec[ci].lo = l;
      if (hi_dirty) {
        BinHi h;
        h.t_found = c_tfound;
        h.rt_stamp = c_stamp;
        h.round_trips = c_rt;
        h.wl_hist = c_wlh;
        rec[ci].hi = h;
      }
      if (x_dirty) {
        P.extra_total[side(ci)] = c_xtot;
        P.extra_count[side(ci)] = c_xcnt;
      }
    }
    hi_dirty = false;
    x_dirty = false;
    sync();
  }
  __device__ __forceinline__ double lnw_lo() const { return ci == ilo ? c_lnw : b_lnw_lo; }
  __device__ __forceinline__ double lnw_hi() const { return ci == ihi ? c_lnw : b_lnw_hi; }
  __device__ __forceinline__ double over_min_T(double x) const { return FAST ? x * inv_min_T : x / P.min_T; }

  // energy.rs:400-434.  Returns false when the fixed window cannot hold e.
  __device__ __forceinline__ bool prepare_for_state(double e) {
    while (e < bmin) {
      if (lo == 0) return false;
      lo -= 1;
      len += 1;
      bmin -= P.width;
    }
    while (e >= bmin + P.width * (double)len) {
      if (lo + len >= (int)P.cap) return false;
      len += 1;
    }
    return true;
  }

  // ---- gamma (energy.rs:799-824) ------------------------------------------
  __device__ __forceinline__ double gamma(unsigned long long moves) {
    if (METHOD == SADMC_METHOD_CANONICAL) return 0.0;
    if (METHOD == SADMC_METHOD_SAD) {
      const double t = (double)moves, tf = (double)tF, ns = (double)num_states;
      if (latest_parameter * tf * ns == 0.0) return 0.0; // energy.rs:23-25
      if (FAST) {
        if (tF != g_tF || num_states != g_ns) { // both change rarely: keep their reciprocals
          g_tF = tF;
          g_ns = num_states;
          inv_tF = 1.0 / tf;
          inv_ns = 1.0 / ns;
        }
        const double ttf = t * inv_tF;
        return (latest_parameter + ttf) * rcp_newton(fma(t * inv_ns, ttf, latest_parameter));
      }
      return (latest_parameter + t / tf) / (latest_parameter + t / ns * (t / tf));
    }
    if (METHOD == SADMC_METHOD_SAMC || method == SADMC_METHOD_SAMC) {
      const double t = (double)moves;
      if (FAST) return t > samc_t0 ? samc_t0 * rcp_newton(t) : 1.0; // tolerance tier: no IEEE divide per move
      return t > samc_t0 ? samc_t0 / t : 1.0;
    }
    return wl_gamma;
  }

  // ---- reject_move (energy.rs:440-512) -------------------------------------
  // i2 = widx(e2); r2 = record of bin i2 (already loaded).  `u01` draws gen::<f64>().
  template <class RNG>
  __device__ __forceinline__ bool reject_move(double e1, double e2, int i2, double lnw_i2, unsigned long long hist_i2,
                                              unsigned long long moves, RNG& rng) {
    if (METHOD == SADMC_METHOD_CANONICAL) {
      if (e1 >= e2) return false;
      return exp_cmp(rng.gen_f64(), (e1 - e2) / P.canonical_T) > 0; // u > exp(...), decided exactly
    }
    double lnw1, lnw2;
    if (METHOD == SADMC_METHOD_SAD) {
      lnw1 = e1 < too_lo ? lnw_lo() + over_min_T(e1 - too_lo) : (e1 > too_hi ? lnw_hi() : c_lnw);
      lnw2 = e2 < too_lo ? lnw_lo() + over_min_T(e2 - too_lo) : (e2 > too_hi ? lnw_hi() : lnw_i2);
    } else {
      lnw1 = c_lnw;
      lnw2 = lnw_i2;
    }
    const bool rejected = lnw2 > lnw1 && exp_cmp(rng.gen_f64(), lnw1 - lnw2) > 0; // u > exp(lnw1 - lnw2), decided exactly
    if (METHOD == SADMC_METHOD_SAD) {
      if (!rejected && hist_i2 == 0 && e2 < too_hi && e2 > too_lo) { // energy.rs:466-483
        latest_parameter = (too_hi - too_lo) / P.min_T;
        num_states += 1;
        tL = moves;
      }
    } else if (METHOD == SADMC_METHOD_WL) {
      if (method != SADMC_METHOD_SAMC && !rejected && hist_i2 == 0 && wl_lowest > 0) wl_num_states += 1.0; // energy.rs:499-501
    }
    return rejected;
  }

  // ---- SAD part of update_weights (energy.rs:523-636) ------------------------
  __device__ __forceinline__ void sad_extend_range(double energy, unsigned long long moves) {
    // Reached when histogram[i] just exceeded highest_hist AND energy lies outside [too_lo, too_hi].
    wrote_bins = true;
    flush();
    const int i = ci;
    if (energy > too_hi) {
      const double lnw_hi = rec[ihi].lo.lnw;
      // energy.rs:544-555 loops over every bin; only bins from ihi up to the walker's bin can satisfy
      // `ej > too_hi && ej <= energy`.  The records are loaded four at a time (independent DRAM accesses).
      for (int j0 = ihi; j0 <= i; j0 += 4) {
        unsigned long long hs[4], tfs[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int j = j0 + u <= i ? j0 + u : i;
          hs[u] = rec[j].lo.hist;
          tfs[u] = rec[j].hi.t_found;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int j = j0 + u;
          if (j > i) break;
          const double ej = centre(j);
          if (ej > too_hi && ej <= energy) {
            if (hs[u] != 0) {
              if (writer) rec[j].lo.lnw = lnw_hi;
              num_states += 1;
            } else if (writer) {
              rec[j].lo.lnw = 0.0;
            }
          }
          if (j > ihi && tfs[u] > tfmax) tfmax = tfs[u];
        }
      }
      latest_parameter = (energy - too_lo) / P.min_T;
      tL = moves;
      too_hi = centre(i);
      if (i != ihi && writer) P.walkers[w].t_range = moves;
      ihi = i;
    } else { // energy < too_lo
      const double lnw_lo = rec[ilo].lo.lnw;
      for (int j0 = i; j0 <= ilo; j0 += 4) {
        unsigned long long hs[4], tfs[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int j = j0 + u <= ilo ? j0 + u : ilo;
          hs[u] = rec[j].lo.hist;
          tfs[u] = rec[j].hi.t_found;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int j = j0 + u;
          if (j > ilo) break;
          const double ej = centre(j);
          if (ej < too_lo && ej >= energy) {
            if (hs[u] != 0) {
              double v = lnw_lo + (ej - too_lo) / P.min_T;
              if (v < 0.0) v = 0.0;
              if (writer) rec[j].lo.lnw = v;
              num_states += 1;
            } else if (writer) {
              rec[j].lo.lnw = 0.0;
            }
          }
          if (j < ilo && tfs[u] > tfmax) tfmax = tfs[u];
        }
      }
      latest_parameter = (too_hi - energy) / P.min_T;
      tL = moves;
      too_lo = centre(i);
      if (i != ilo && writer) P.walkers[w].t_range = moves;
      ilo = i;
    }
    sync();
    c_lnw = rec[i].lo.lnw; // the walker's own bin may have been overwritten
    b_lnw_lo = rec[ilo].lo.lnw;
    b_lnw_hi = rec[ihi].lo.lnw;
  }

  // g = gamma(moves), evaluated by the caller before anything of this move touched its inputs
  __device__ __forceinline__ void update_weights_sad(double energy, unsigned long long moves, double g) {
    const double old_lnw = c_lnw;
    c_lnw += g;
    if (too_lo > too_hi || energy < too_lo || energy > too_hi) c_lnw = old_lnw; // energy.rs:535-538
    if (c_hist > highest_hist) {
      highest_hist = c_hist;
      if (energy > too_hi || energy < too_lo) sad_extend_range(energy, moves);
    }
    if (tL == moves) { // energy.rs:585-635
      const unsigned long long old_tF = tF;
      tF = tfmax;
      if (old_tF != tF && P.move_plan == SADMC_MOVE_ACCEPTANCE_RATE) {
        double s = acc_rate / P.move_value;
        s = s < 0.8 ? 0.8 : (s > 1.2 ? 1.2 : s);
        tscale *= s;
      }
    }
  }

  // ---- WL part of update_weights (energy.rs:638-752) -------------------------
  // Number of visited bins whose WL hist is <= wl_lowest, by a full scan (rare).
  __device__ __forceinline__ long long wl_count_low() {
    long long n = 0;
    for (int j = lo; j < lo + len; j++)
      if (rec[j].lo.hist != 0 && rec[j].hi.wl_hist <= wl_lowest) n++;
    return n;
  }
  __device__ __forceinline__ void wl_regroup(unsigned long long moves) {
    // energy.rs:656-687: hist.len() != lnw.len()
    flush();
    if (wl_hist_len == 0 || (wl_gamma != 1.0 && wl_lowest > 0)) {
      wl_gamma = 1.0;
      wl_lowest = 0;
      wl_highest = 0;
      wl_total = 0;
      if (writer)
        for (int j = lo; j < lo + len; j++) rec[j].hi.wl_hist = 0;
      wl_min_energy = bmin;
    } else {
      // the window keeps zeros where the reference pads; replay the min_energy arithmetic
      while (wl_min_energy > bmin) wl_min_energy -= P.width;
      unsigned long long mn = ~0ull;
      for (int j = lo; j < lo + len; j++) {
        const unsigned long long h = rec[j].hi.wl_hist;
        if (h < mn) mn = h;
      }
      wl_lowest = mn;
    }
    wl_hist_len = len;
    sync();
    c_wlh = rec[ci].hi.wl_hist;
    wl_low_count = wl_count_low();
    (void)moves;
  }
  __device__ __forceinline__ void update_weights_wl(double energy, unsigned long long moves, bool first_visit, double g) {
    (void)energy;
    c_lnw += g;
    if (method == SADMC_METHOD_SAMC) return; // 1/t-WL after its switch (energy.rs:754-756)
    if (P.has_min_gamma && wl_gamma < P.min_gamma) { // production run, energy.rs:649-655
      c_wlh += 1;
      hi_dirty = true;
      return;
    }
    if (wl_hist_len != len) {
      wl_regroup(moves);
    } else if (first_visit && c_wlh <= wl_lowest) {
      wl_low_count += 1; // a bin just became "visited" (histogram != 0) for the flatness filter
    }
    if (c_wlh == wl_lowest) wl_low_count -= 1; // it is about to exceed wl_lowest
    c_wlh += 1;
    hi_dirty = true;
    if (c_wlh > wl_highest) wl_highest = c_wlh;
    wl_total += 1;
    const double max_energy = wl_min_energy + (double)wl_hist_len * P.width;
    // energy.rs:695-708; `min over visited bins == lowest + 1`  <=>  no visited bin is still <= lowest
    if (c_wlh == wl_lowest + 1 && wl_hist_len > 1 && (!P.has_min || P.min_allowed >= wl_min_energy) &&
        (!P.has_max || P.max_allowed <= max_energy) && wl_low_count == 0) {
      wl_lowest = c_wlh;
      bool rescan = true;
      if ((P.inv_t && wl_lowest > 0) || (double)wl_lowest >= 0.8 * (double)wl_total / wl_num_states) {
        wl_gamma *= 0.5;
        flush();
        if (writer)
          for (int j = lo; j < lo + len; j++) rec[j].hi.wl_hist = 0;
        sync();
        c_wlh = 0;
        hi_dirty = true;
        wl_total = 0;
        wl_lowest = 0;
        wl_highest = 0;
        if (P.has_min_gamma && wl_gamma < P.min_gamma) wl_gamma = 0.0;
      }
      if (rescan) {
        flush();
        wl_low_count = wl_count_low();
        hi_dirty = true;
      }
      if (P.inv_t && wl_gamma < wl_num_states / (double)moves) {
        method = SADMC_METHOD_SAMC;
        samc_t0 = wl_num_states;
      }
    }
  }

  // ---- round trips (energy.rs:950-965) ------------------------------------
  __device__ __forceinline__ void round_trips(int i1_ref, unsigned long long moves) {
    const int i_ref = ci - lo;
    if (P.flags & SADMC_FLAG_NO_ROUND_TRIPS) {
      // max_S / max_S_index are still kept: the largest ln w of the walker is what the merge for reporting aligns
      // walkers by (fold_kernels.cuh), and it costs two instructions here instead of a pass over the window there
      if (c_lnw > max_S) {
        max_S = c_lnw;
        max_S_index = i_ref;
      }
      return;
    }
    if (c_lnw > max_S) {
      max_S = c_lnw;
      max_S_index = i_ref;
      rt_fill_val = 1;
      rt_fill_time = moves;
      rt_fill_lo = lo;
      rt_fill_hi = lo + len;
      c_visited = true;
    } else if (i_ref == max_S_index) {
      if (i1_ref != i_ref) {
        rt_fill_val = 0;
        rt_fill_time = moves;
        rt_fill_lo = lo;
        rt_fill_hi = lo + len;
        c_visited = false;
      }
    } else if (!c_visited) {
      c_visited = true;
      c_stamp = moves;
      c_rt += 1;
      hi_dirty = true;
    }
  }
};

} // namespace sadmc
