// kernels_lj_thread_fast.cu -- LJ clusters, one thread per walker, tolerance tier (SADMC_FLAG_FAST_MATH).
#include "make_set.cuh"
#include "sys_lj_thread.cuh"
namespace sadmc {
bool kernels_lj_thread_fast(int N, int G, const DevParams& P, KernelSet* out) {
  if (N > 64 || G != 1) return false;
#ifdef SADMC_EXP_NT /* occupancy probe (tools/exp_build.sh): a compile-time atom count other than 31 / 38 */
  if (N == SADMC_EXP_NT) {
    *out = make_set<LjThreadSys<true, SADMC_EXP_NT, 1>, true>(P);
    return true;
  }
#endif
  if (N == 31) {
    *out = make_set<LjThreadSys<true, 31, 1>, true>(P);
#ifndef SADMC_LJ31_SMEM_Z /* (defined: only the all-shared-memory move kernels are built) */
    // Which layout needs less time for this many walkers: a wave of 384 walkers per SM (stream) takes 1.457 x as long as
    // a wave of 256 (shared memory) -- 9.05e9 against 8.79e9 moves/s with whole waves of either, profiles/r02_zg_ab.log.
    bool stream = (P.flags & SADMC_FLAG_LJ_STREAM_Z) != 0;
    if (!(P.flags & (SADMC_FLAG_LJ_STREAM_Z | SADMC_FLAG_LJ_SMEM_Z))) {
      int dev = 0, sms = 148;
      if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      const long long w = P.n_walkers, zg_wave = 384ll * sms, sm_wave = 256ll * sms;
      stream = (double)((w + zg_wave - 1) / zg_wave) * 1.457 < (double)((w + sm_wave - 1) / sm_wave);
    }
    if (stream) {
    // histogram-method move kernels: z streamed from L2, three CTAs per SM (sys_lj_thread.cuh, ZG); init, shims, binning,
    // tempering and replicas keep the shared-memory layout
    typedef LjThreadSys<true, 31, 1, 0, true> S;
    out->move[SADMC_METHOD_SAD] = move_kernel<S, SADMC_METHOD_SAD>;
    out->move[SADMC_METHOD_SAMC] = move_kernel<S, SADMC_METHOD_SAMC>;
    out->move[SADMC_METHOD_WL] = move_kernel<S, SADMC_METHOD_WL>;
    out->move[SADMC_METHOD_INV_T_WL] = move_kernel<S, SADMC_METHOD_WL>;
    out->move[SADMC_METHOD_CANONICAL] = move_kernel<S, SADMC_METHOD_CANONICAL>;
    out->move_block = S::BLOCK;
    out->move_threads_per_walker = 1;
    out->move_smem = zig_smem_bytes<S>() + S::smem_bytes(P, S::BLOCK);
    out->zstream_per_thread = S::ZSTREAM_PER_WALKER;
    }
#endif
  }
  else if (N == 38)
    *out = make_set<LjThreadSys<true, 38, 1>, true>(P);
  else
    *out = make_set<LjThreadSys<true, 0, 1>, true>(P);
  return true;
}
} // namespace sadmc
