// replicas_round.cuh -- what `MC::run_once` does after the moves (src/mc/energy_replicas.rs:527-600): neighbour swaps,
// the median estimator below the lowest cutoff, splitting off a new replica.  Included by engine.cu only; one CTA per
// simulation, thread 0 runs the (serial, few-dozen-step) logic, the block copies / swaps system images and selects the
// median.  The per-round moves are replicas.cuh.
#pragma once
#include "book.cuh"
#include "kernel_set.cuh"
#include "replicas.cuh"
#include "rng.cuh"

namespace sadmc {

constexpr int REPLICA_ESTIMATOR_SIZE = 4096; // energy_replicas.rs:49

__device__ __forceinline__ void replica_swap_rows(const DevParams& P, uint32_t a, uint32_t b, int t, int nt) {
  if (P.sys_stride) {
    double* x = P.sys + (size_t)a * P.sys_stride;
    double* y = P.sys + (size_t)b * P.sys_stride;
    for (uint32_t k = t; k < P.sys_stride; k += nt) {
      const double v = x[k];
      x[k] = y[k];
      y[k] = v;
    }
  }
  if (P.ising_words) {
    uint32_t* x = P.sys_words + (size_t)a * P.ising_words;
    uint32_t* y = P.sys_words + (size_t)b * P.ising_words;
    for (uint32_t k = t; k < P.ising_words; k += nt) {
      const uint32_t v = x[k];
      x[k] = y[k];
      y[k] = v;
    }
  }
}

__global__ void __launch_bounds__(128) replica_round_kernel(const DevParams P, ReplicaRec* reps, ReplicaSim* sims, double* slot_energy, double* median_buf,
                                                           uint32_t r_max, unsigned long long steps, double dimensionality) {
  __shared__ int sh_swap[128]; // pairs to swap this round (slot index of the upper partner), -1 terminated
  __shared__ int sh_nswap, sh_new, sh_len;
  __shared__ double sh_mid, sh_next, sh_prev;
  __shared__ int sh_has_next, sh_has_prev;
  const uint32_t sim = blockIdx.x;
  const int t = threadIdx.x, nt = blockDim.x;
  ReplicaSim& S = sims[sim];
  ReplicaRec* R = reps + (size_t)sim * r_max;
  double* E = slot_energy + (size_t)sim * r_max;
  double* med = median_buf + (size_t)sim * REPLICA_ESTIMATOR_SIZE;
  const uint32_t base = sim * r_max;
  if (t == 0) {
    Rng g;
    g.s0 = S.s0;
    g.s1 = S.s1;
    const int first = (g.next() >> 63) != 0 ? 0 : 1; // gen::<bool>(): chunks of replicas[..] or of replicas[1..] (527-533)
    int ns = 0;
    for (int i = first; i + 1 < S.n_rep; i += 2) {
      ReplicaRec& r0 = R[i];
      ReplicaRec& r1 = R[i + 1];
      if (E[i] < r1.max_energy) { // 538-559
        if (ns < 128) sh_swap[ns++] = i;
        const double e = E[i];
        E[i] = E[i + 1];
        E[i + 1] = e;
        const double l = r0.lowest_max;
        r0.lowest_max = r1.lowest_max;
        r1.lowest_max = l;
        r0.collecting = 1;
        r1.collecting = 1;
        if (r1.lowest_max > r1.max_energy) {
          r1.unique_visitors += 1;
          r1.lowest_max = r1.max_energy;
        }
      }
    }
    sh_nswap = ns;
    // the median estimator of the energies below the lowest cutoff (562-567, MedianEstimator::add_energy 60-69)
    const int last = S.n_rep - 1;
    const double last_energy = E[last];
    if (last_energy < R[last].cutoff) {
      if (S.median_len < REPLICA_ESTIMATOR_SIZE) {
        med[S.median_len++] = last_energy;
      } else if (g.gen_f64() < 1.0 / ((double)S.median_len + 1.0)) {
        const uint32_t i = g.below((uint32_t)S.median_len, zone_single((uint64_t)S.median_len)); // gen_range(0, len)
        med[i] = last_energy;
      }
    }
    S.s0 = g.s0;
    S.s1 = g.s1;
    // a new replica? (568-590)
    const ReplicaRec& r = R[last];
    int want = 0;
    if (r.unique_visitors >= S.indep && r.lowest_max == r.max_energy) {
      const double mean_below = r.below_total / (double)r.below_count;
      if (mean_below + S.min_T < r.cutoff && last_energy < r.cutoff) want = 1;
    }
    if (want && S.n_rep >= (int)r_max) {
      S.overflow = 1; // no free slot: the simulation carries on without splitting (reported to the host)
      want = 0;
    }
    sh_new = want;
    sh_len = S.median_len;
  }
  __syncthreads();
  // the swaps: std::mem::swap(&mut r0.system, &mut r1.system) (540)
  for (int k = 0; k < sh_nswap; k++) {
    const uint32_t a = base + sh_swap[k], b = a + 1;
    replica_swap_rows(P, a, b, t, nt);
    if (t == 0) {
      WalkerRec& w0 = P.walkers[a];
      WalkerRec& w1 = P.walkers[b];
      double v = w0.E;
      w0.E = w1.E;
      w1.E = v;
      v = w0.err;
      w0.err = w1.err;
      w1.err = v;
      v = w0.d_squared;
      w0.d_squared = w1.d_squared;
      w1.d_squared = v;
    }
  }
  __syncthreads();
  if (sh_new) {
    // MedianEstimator::median (70-98) without the sort (the vector is reset right afterwards): the element of rank len / 2,
    // then the smallest value above it, else the largest below it
    const int len = sh_len, middle = len / 2;
    if (t == 0) {
      sh_has_next = 0;
      sh_has_prev = 0;
    }
    __syncthreads();
    for (int i = t; i < len; i += nt) {
      const double v = med[i];
      int less = 0, eq = 0;
      for (int j = 0; j < len; j++) {
        const double u = med[j];
        less += u < v ? 1 : 0;
        eq += u == v ? 1 : 0;
      }
      if (less <= middle && middle < less + eq) sh_mid = v; // every candidate that qualifies holds the same value
    }
    __syncthreads();
    const double mid = sh_mid;
    // block-wide min over values > mid and max over values < mid, through shared atomics on the ordered bit patterns
    __shared__ unsigned long long sh_next_bits, sh_prev_bits;
    if (t == 0) {
      sh_next_bits = ~0ull;
      sh_prev_bits = 0ull;
    }
    __syncthreads();
    auto ordered = [](double x) { // monotone map of doubles onto unsigned integers
      const unsigned long long b = (unsigned long long)__double_as_longlong(x);
      return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    };
    for (int i = t; i < len; i += nt) {
      const double v = med[i];
      if (v > mid) atomicMin(&sh_next_bits, ordered(v));
      if (v < mid) atomicMax(&sh_prev_bits, ordered(v));
    }
    __syncthreads();
    if (t == 0) {
      auto unordered = [](unsigned long long o) {
        const unsigned long long b = (o >> 63) ? (o & 0x7fffffffffffffffull) : ~o;
        return __longlong_as_double((long long)b);
      };
      double median_below = mid;
      if (sh_next_bits != ~0ull)
        median_below = 0.5 * (mid + unordered(sh_next_bits));
      else if (sh_prev_bits != 0ull)
        median_below = 0.5 * (mid + unordered(sh_prev_bits));
      med[0] = median_below; // MedianEstimator::reset (55-58)
      S.median_len = 1;
      const int last = S.n_rep - 1;
      const ReplicaRec r = R[last];
      ReplicaRec nr = r; // r.clone()
      nr.max_energy = r.cutoff;
      nr.cutoff = median_below;
      // decimate (171-196)
      nr.upwelling = 0;
      if (nr.above_count > 1) {
        nr.above_total /= (double)nr.above_count;
        nr.above_sq /= (double)nr.above_count;
        nr.above_count = 1;
      }
      if (nr.below_count > 1) {
        nr.below_total /= (double)nr.below_count;
        nr.below_sq /= (double)nr.below_count;
        nr.below_count = 1;
      }
      if (nr.xcnt > 1) {
        nr.xtot /= (double)nr.xcnt;
        nr.xcnt = 1;
      }
      nr.accepted = 1;
      nr.rejected = 1;
      nr.unique_visitors = 1;
      nr.lowest_max = -INFINITY; // 580
      nr.tscale = r.tscale * pow(0.5, 1.0 / dimensionality);
      R[last + 1] = nr;
      E[last + 1] = E[last];
      WalkerRec& src = P.walkers[base + last];
      WalkerRec& dst = P.walkers[base + last + 1];
      dst.E = src.E;
      dst.err = src.err;
      dst.d_squared = src.d_squared;
      dst.status = 0;
      Rng g;
      g.s0 = src.s0;
      g.s1 = src.s1;
      xoroshiro_jump_device(g); // newr.rng.jump() (588)
      dst.s0 = g.s0;
      dst.s1 = g.s1;
    }
    // newr.system = r.system.clone()
    {
      const uint32_t a = base + S.n_rep - 1, b = a + 1;
      if (P.sys_stride)
        for (uint32_t k = t; k < P.sys_stride; k += nt) P.sys[(size_t)b * P.sys_stride + k] = P.sys[(size_t)a * P.sys_stride + k];
      if (P.ising_words)
        for (uint32_t k = t; k < P.ising_words; k += nt) P.sys_words[(size_t)b * P.ising_words + k] = P.sys_words[(size_t)a * P.ising_words + k];
    }
    __syncthreads();
    if (t == 0) S.n_rep += 1;
  }
  if (t == 0) S.moves += steps * (unsigned long long)(sh_new ? S.n_rep - 1 : S.n_rep); // these_moves: steps per replica that ran (509-516, 595-597)
}

} // namespace sadmc
