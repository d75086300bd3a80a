// kernels_lj_warp_small.cu -- LJ clusters with the atoms in registers, 4 or 8 lanes per walker.
#include "make_set.cuh"
#include "sys_lj.cuh"
namespace sadmc {
bool kernels_lj_warp_small(int G, int A, const DevParams& P, KernelSet* out) {
#define CASE(g, a)                  \
  if (G == g && A == a) {           \
    *out = make_set<LjSys<g, a>>(P); \
    return true;                    \
  }
  CASE(8, 1) CASE(8, 2) CASE(8, 4) CASE(8, 5) CASE(4, 4) CASE(4, 8)
#undef CASE
  return false;
}
} // namespace sadmc
