// fold_kernels.cuh -- merge of the local walkers' bins for reporting (included by engine.cu only).
#pragma once
#include "book.cuh"

namespace sadmc {

// ---- merge for reporting ---------------------------------------------------
// One thread per window bin; loops over the local walkers (each warp reads 32
// consecutive bins of one walker: coalesced 1 KB).  lnw is aligned per walker by
// subtracting that walker's maximum lnw (plotting/parse-binning.py:169) before
// it is summed; bins a walker never visited do not contribute to the lnw sums.
//
// Selection (sadmc_fold_select): walkers first, first + stride, ... take part (interleaved groups give
// ensemble error bars); with sad_range_only a SAD walker contributes ln w only for the bins inside its own
// [too_lo, too_hi] -- the part of ln w that SAD defines (plotting/parse-binning.py:150-164 replaces the rest).
struct FoldSel {
  uint32_t first, stride;
  int sad_range_only;
};
__device__ __forceinline__ bool fold_lnw_counts(const WalkerRec& r, const BinLo& b, int j, const FoldSel s) {
  if (b.hist == 0) return false;
  if (s.sad_range_only && r.method == SADMC_METHOD_SAD) return j >= r.ilo && j <= r.ihi;
  return true;
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

__global__ void __launch_bounds__(256) fold_kernel(const DevParams P, const double* wmax, unsigned long long* histogram, double* energy_total,
                                                  double* energy_squared_total, double* lnw_sum, double* lnw_sq_sum,
                                                  unsigned long long* lnw_count, const FoldSel sel) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= P.cap) return;
  unsigned long long h = 0, cnt = 0;
  double et = 0.0, e2 = 0.0, ls = 0.0, lq = 0.0;
  for (uint32_t w = sel.first; w < P.n_walkers; w += sel.stride) {
    const WalkerRec& r = P.walkers[w];
    if ((int)j < r.lo || (int)j >= r.lo + r.len) continue;
    const BinLo b = P.rec[(size_t)w * P.cap + j].lo;
    h += b.hist;
    et += b.etot;
    e2 += b.e2tot;
    if (fold_lnw_counts(r, b, (int)j, sel)) {
      const double a = b.lnw - wmax[w];
      ls += a;
      lq += a * a;
      cnt += 1;
    }
  }
  if (histogram) histogram[j] = h;
  if (energy_total) energy_total[j] = et;
  if (energy_squared_total) energy_squared_total[j] = e2;
  if (lnw_sum) lnw_sum[j] = ls;
  if (lnw_sq_sum) lnw_sq_sum[j] = lq;
  if (lnw_count) lnw_count[j] = cnt;
}

} // namespace sadmc
