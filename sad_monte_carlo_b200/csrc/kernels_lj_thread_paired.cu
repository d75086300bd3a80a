// kernels_lj_thread_paired.cu -- EXPERIMENT: LJ31 / LJ38 thread-per-walker kernels with helper warps (sys_lj_paired.cuh).
#include "make_set.cuh"
#include "sys_lj_paired.cuh"
namespace sadmc {
template <int NT>
static KernelSet paired_set(const DevParams& P) {
  KernelSet k = make_set<LjThreadSys<true, NT, 1>>(P); // init and shims: the one-lane kernels (same layout)
  typedef LjPairedSys<NT> S;
  k.move[SADMC_METHOD_SAD] = move_kernel<S, SADMC_METHOD_SAD>;
  k.move[SADMC_METHOD_SAMC] = move_kernel<S, SADMC_METHOD_SAMC>;
  k.move[SADMC_METHOD_WL] = move_kernel<S, SADMC_METHOD_WL>;
  k.move[SADMC_METHOD_INV_T_WL] = move_kernel<S, SADMC_METHOD_WL>;
  k.move[SADMC_METHOD_CANONICAL] = move_kernel<S, SADMC_METHOD_CANONICAL>;
  k.move_block = S::BLOCK;
  k.move_threads_per_walker = S::BLOCK / S::WALKERS_PER_BLOCK;
  k.move_smem = ZIG_SMEM_BYTES + S::smem_bytes(P, S::BLOCK);
  return k;
}
bool kernels_lj_thread_paired(int N, const DevParams& P, KernelSet* out) {
  if (N == 31)
    *out = paired_set<31>(P);
  else if (N == 38)
    *out = paired_set<38>(P);
  else
    return false;
  return true;
}
} // namespace sadmc
