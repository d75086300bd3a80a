// kernels_lj_thread_fast_multi.cu -- LJ clusters in shared memory, 2 or 4 threads per walker (tolerance tier).
#include "make_set.cuh"
#include "sys_lj_thread.cuh"
namespace sadmc {
bool kernels_lj_thread_fast_multi(int N, int G, const DevParams& P, KernelSet* out) {
  if (N > 64 || (G != 2 && G != 4)) return false;
#define CASE(nt)                                                                       \
  if (N == nt || nt == 0) {                                                            \
    *out = G == 2 ? make_set<LjThreadSys<true, nt, 2>>(P) : make_set<LjThreadSys<true, nt, 4>>(P); \
    return true;                                                                       \
  }
  CASE(31) CASE(38) CASE(0)
#undef CASE
  return false;
}
} // namespace sadmc
