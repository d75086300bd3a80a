// fold_kernels.cuh -- merge of the local walkers' bins for reporting (included by engine.cu only).
#pragma once
#include "book.cuh"

namespace sadmc {

// ---- merge for reporting ---------------------------------------------------
// One thread per window bin and walker chunk (each warp reads 32 consecutive bins of
// one walker: the 32-byte `lo` sectors of 32 adjacent records).  lnw is aligned per walker by
// subtracting that walker's maximum lnw (plotting/parse-binning.py:169) before
// it is summed; bins a walker never visited do not contribute to the lnw sums.
//
// Selection (sadmc_fold_select / sadmc_fold_select_ex): walkers first, first + stride, ... (at most `count` of them,
// 0 = to the end) take part: interleaved groups give ensemble error bars, contiguous blocks are the shards of a
// multi-GPU run.  sad_range_only: 1 = a SAD walker contributes ln w only for the bins inside its own [too_lo, too_hi] --
// the part of ln w that SAD defines (plotting/parse-binning.py:150-164 replaces the rest); 2 = only the bins strictly
// inside: too_lo / too_hi are bin centres and `update_weights` reverts the increment for energies outside the range
// (energy.rs:535-538), so the two end bins receive the increments of only half of their visits.
//
// Alignment constant of a walker: its largest ln w over the bins that count.  Without a range restriction that is
// the walker's running maximum max_S (energy.rs:950-953, kept by the move kernel also when the round-trip diagnostics
// are switched off): ln w of a bin never decreases and a SAD range extension only writes values that exist already,
// so the merge is ONE pass over the records.  With a range restriction the maximum over the restricted set is taken
// by walker_max_lnw_kernel first.
//
// tl_max (sadmc_fold_settled; 0 = off): a SAD walker contributes ln w only if the bins of its range [too_lo, too_hi] have
// not changed since move tl_max (WalkerRec::t_range).  A bin that has just joined a walker's range starts from a copied
// or zero ln w (energy.rs:544-584), the former end bin from a ln w that received only half of its increments, and at
// gamma ~ 1/t such a bin needs about as long again to settle; a minority of such walkers otherwise dominates the error
// of an ensemble mean.  Histograms and energy moments are not filtered.
//
// SADMC_FLAG_BINNING engines (binning != 0; record layout of book_binning.cuh): the histogram that is merged is the count of
// the "energy" accumulator -- every visit, never zeroed, whereas lnw.count is reset by SAD range extensions
// (energy_binning.rs:355-361) -- energy_squared_total is not collected by the reference (0), and the alignment constant
// is always taken by walker_max_lnw_kernel (set_lnw can lower a walker's largest ln w).
struct FoldSel {
  uint32_t first, stride, count;
  int sad_range_only;
  unsigned long long tl_max;
  int binning;
};
__device__ __forceinline__ unsigned long long fold_visits(const BinLo& b, const FoldSel s) {
  return s.binning ? (unsigned long long)__double_as_longlong(b.e2tot) : b.hist;
}
__device__ __forceinline__ void fold_range(const WalkerRec& r, const FoldSel s, int& ilo, int& ihi) {
  const bool sad = r.method == SADMC_METHOD_SAD;
  const bool ranged = s.sad_range_only != 0 && sad;
  const int shrink = s.sad_range_only == 2 ? 1 : 0;
  ilo = ranged ? r.ilo + shrink : r.lo;
  ihi = ranged ? r.ihi - shrink : r.lo + r.len - 1;
  if (sad && s.tl_max != 0 && r.t_range > s.tl_max) ihi = ilo - 1; // not settled: no bin counts
}
__device__ __forceinline__ bool fold_lnw_counts(const WalkerRec& r, const BinLo& b, int j, const FoldSel s) {
  if (fold_visits(b, s) == 0) return false;
  int ilo, ihi;
  fold_range(r, s, ilo, ihi);
  return j >= ilo && j <= ihi;
}
__global__ void __launch_bounds__(256) walker_max_lnw_kernel(const DevParams P, double* wmax, const FoldSel sel) {
  const uint32_t w = sel.first + blockIdx.x * sel.stride;
  const WalkerRec& r = P.walkers[w];
  double m = -1e300;
  for (int j = r.lo + (int)threadIdx.x; j < r.lo + r.len; j += blockDim.x) {
    const BinLo b = P.rec[(size_t)w * P.cap + j].lo;
    if (fold_lnw_counts(r, b, j, sel) && b.lnw > m) m = b.lnw;
  }
  __shared__ double sm[256];
  sm[threadIdx.x] = m;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s && sm[threadIdx.x + s] > sm[threadIdx.x]) sm[threadIdx.x] = sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) wmax[w] = sm[0];
}

// Stage 1: grid (bin blocks, walker chunks).  A block walks its chunk of the selected walkers in tiles of
// 256 whose window extents are staged in shared memory, every thread owns one bin and keeps eight record
// loads in flight (the loop is bound by HBM latency otherwise: one 32-byte sector per walker and thread).
// Stage 2 adds the chunk partials in chunk order, so the result does not depend on scheduling.
struct FoldMeta {
  int lo, end, ilo, ihi; // window extent [lo, end) and the extent in which ln w counts
  double wmax;
};
constexpr int FOLD_FIELDS = 6;
__global__ void __launch_bounds__(256) fold_partial_kernel(const DevParams P, const double* wmax, double* partial, const FoldSel sel,
                                                          uint32_t n_sel, uint32_t per_chunk) {
  __shared__ FoldMeta meta[256];
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t c = blockIdx.y;
  const uint32_t s_begin = c * per_chunk;
  const uint32_t s_end = s_begin + per_chunk < n_sel ? s_begin + per_chunk : n_sel;
  unsigned long long h = 0, cnt = 0;
  double et = 0.0, e2 = 0.0, ls = 0.0, lq = 0.0;
  for (uint32_t s0 = s_begin; s0 < s_end; s0 += 256) {
    __syncthreads();
    const uint32_t s = s0 + threadIdx.x;
    if (s < s_end) {
      const uint32_t w = sel.first + s * sel.stride;
      const WalkerRec& r = P.walkers[w];
      FoldMeta m;
      m.lo = r.lo;
      m.end = r.lo + r.len;
      fold_range(r, sel, m.ilo, m.ihi);
      m.wmax = wmax ? wmax[w] : r.max_S; // one pass: the walker's running maximum (see above)
      meta[threadIdx.x] = m;
    }
    __syncthreads();
    const uint32_t n_tile = s_end - s0 < 256 ? s_end - s0 : 256;
    if (j >= P.cap) continue;
    for (uint32_t k0 = 0; k0 < n_tile; k0 += 8) {
      BinLo b[8];
      bool in[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const uint32_t k = k0 + u;
        in[u] = k < n_tile && (int)j >= meta[k < n_tile ? k : 0].lo && (int)j < meta[k < n_tile ? k : 0].end;
        if (in[u]) b[u] = P.rec[(size_t)(sel.first + (s0 + k) * sel.stride) * P.cap + j].lo;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        if (!in[u]) continue;
        const FoldMeta m = meta[k0 + u];
        const unsigned long long visits = fold_visits(b[u], sel);
        h += visits;
        et += b[u].etot;
        if (!sel.binning) e2 += b[u].e2tot;
        if (visits != 0 && (int)j >= m.ilo && (int)j <= m.ihi) {
          const double a = b[u].lnw - m.wmax;
          ls += a;
          lq += a * a;
          cnt += 1;
        }
      }
    }
  }
  if (j >= P.cap) return;
  double* p = partial + ((size_t)c * FOLD_FIELDS) * P.cap + j;
  p[0] = __longlong_as_double((long long)h);
  p[(size_t)1 * P.cap] = et;
  p[(size_t)2 * P.cap] = e2;
  p[(size_t)3 * P.cap] = ls;
  p[(size_t)4 * P.cap] = lq;
  p[(size_t)5 * P.cap] = __longlong_as_double((long long)cnt);
}

// `packed` != nullptr: everything as ONE f64 buffer [7][cap] for a single collective -- histogram as two exact halves
// (h >> 32, h & 0xffffffff: their sums over ranks stay below 2^53), lnw_count, energy_total, energy_squared_total,
// lnw_sum, lnw_sq_sum (parallel.py recombines the histogram in integer arithmetic).
constexpr int FOLD_PACKED_FIELDS = 7;
__global__ void __launch_bounds__(256) fold_final_kernel(const DevParams P, const double* partial, uint32_t n_chunks, unsigned long long* histogram,
                                                        double* energy_total, double* energy_squared_total, double* lnw_sum, double* lnw_sq_sum,
                                                        unsigned long long* lnw_count, double* packed) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= P.cap) return;
  unsigned long long h = 0, cnt = 0;
  double et = 0.0, e2 = 0.0, ls = 0.0, lq = 0.0;
  for (uint32_t c = 0; c < n_chunks; c++) {
    const double* p = partial + ((size_t)c * FOLD_FIELDS) * P.cap + j;
    h += (unsigned long long)__double_as_longlong(p[0]);
    et += p[(size_t)1 * P.cap];
    e2 += p[(size_t)2 * P.cap];
    ls += p[(size_t)3 * P.cap];
    lq += p[(size_t)4 * P.cap];
    cnt += (unsigned long long)__double_as_longlong(p[(size_t)5 * P.cap]);
  }
  if (histogram) histogram[j] = h;
  if (energy_total) energy_total[j] = et;
  if (energy_squared_total) energy_squared_total[j] = e2;
  if (lnw_sum) lnw_sum[j] = ls;
  if (lnw_sq_sum) lnw_sq_sum[j] = lq;
  if (lnw_count) lnw_count[j] = cnt;
  if (packed) {
    const size_t n = P.cap;
    packed[j] = (double)(h >> 32);
    packed[n + j] = (double)(h & 0xffffffffull);
    packed[2 * n + j] = (double)cnt;
    packed[3 * n + j] = et;
    packed[4 * n + j] = e2;
    packed[5 * n + j] = ls;
    packed[6 * n + j] = lq;
  }
}

} // namespace sadmc
