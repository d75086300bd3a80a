// kernels_lj_warp.cu -- LJ clusters with the atoms in registers, 16 or 32 lanes per walker.
#include "make_set.cuh"
#include "sys_lj.cuh"
namespace sadmc {
bool kernels_lj_warp(int G, int A, const DevParams& P, KernelSet* out) {
#define CASE(g, a)                  \
  if (G == g && A == a) {           \
    *out = make_set<LjSys<g, a>>(P); \
    return true;                    \
  }
  CASE(32, 1) CASE(32, 2) CASE(16, 2) CASE(16, 3)
#undef CASE
  return false;
}
} // namespace sadmc
