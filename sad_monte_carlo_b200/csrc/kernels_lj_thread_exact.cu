// kernels_lj_thread_exact.cu -- LJ clusters, one thread per walker, reference operation order (bit-exact tier).
#include "make_set.cuh"
#include "sys_lj_thread.cuh"
namespace sadmc {
bool kernels_lj_thread_exact(int N, const DevParams& P, KernelSet* out) {
  if (N > 64) return false;
  if (N == 31)
    *out = make_set<LjThreadSys<false, 31, 1>, true>(P);
  else if (N == 38)
    *out = make_set<LjThreadSys<false, 38, 1>, true>(P);
  else
    *out = make_set<LjThreadSys<false, 0, 1>, true>(P);
  return true;
}
} // namespace sadmc
