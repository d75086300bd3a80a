// kernels_lj_thread_fast.cu -- LJ clusters, one thread per walker, tolerance tier (SADMC_FLAG_FAST_MATH).
#include "make_set.cuh"
#include "sys_lj_thread.cuh"
namespace sadmc {
bool kernels_lj_thread_fast(int N, int G, const DevParams& P, KernelSet* out) {
  if (N > 64 || G != 1) return false;
  if (N == 31)
    *out = make_set<LjThreadSys<true, 31, 1>, true>(P);
  else if (N == 38)
    *out = make_set<LjThreadSys<true, 38, 1>, true>(P);
  else
    *out = make_set<LjThreadSys<true, 0, 1>, true>(P);
  return true;
}
} // namespace sadmc
