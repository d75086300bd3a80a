// kernels_lj_thread_fast.cu -- LJ clusters, one thread per walker, tolerance tier (SADMC_FLAG_FAST_MATH).
#include "make_set.cuh"
#include "sys_lj_thread.cuh"
namespace sadmc {
template <class S>
static void use_stream(const DevParams& P, KernelSet* out) {
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
bool kernels_lj_thread_fast(int N, int G, const DevParams& P, KernelSet* out) {
  if (N > 64 || G != 1) return false;
#ifdef SADMC_EXP_NT /* occupancy probe (tools/exp_build.sh): a compile-time atom count other than 31 / 38 */
  if (N == SADMC_EXP_NT) {
    *out = make_set<LjThreadSys<true, SADMC_EXP_NT, 1>, true>(P);
    return true;
  }
#endif
  if (N == 31 || N == 38) {
    if (N == 31)
      *out = make_set<LjThreadSys<true, 31, 1>, true>(P);
    else
      *out = make_set<LjThreadSys<true, 38, 1>, true>(P);
#ifndef SADMC_LJ_SMEM_Z_ONLY /* (defined: only the all-shared-memory move kernels are built) */
    // Histogram-method move kernels with z streamed from L2 (sys_lj_thread.cuh, ZG): LJ31 three 128-thread CTAs per SM instead of
    // two, LJ38 one 320-thread CTA instead of one of 224.  Init, shims, binning, tempering and replicas keep the shared-memory
    // layout.  Which layout needs less time for this many walkers, from the rates measured with whole waves of either:
    //   LJ31 SAD    a wave of 384 walkers per SM takes 1.44 x as long as a wave of 256 (9.16e9 / 8.79e9 moves/s, profiles/r02_zg_ab.log);
    //   LJ31 WL     1.19 x (1/t-WL 5.52e9 / 4.39e9, profiles/r02_lj31_wl_zg.log);
    //   LJ38 WL     a wave of 320 takes 1.16 x as long as a wave of 224 (1/t-WL 4.05e9 / 3.30e9, profiles/r02_lj38_zg.log);
    //   LJ38 other  1.40 x (SAD 6.06e9 / 5.92e9).
    const long long zg_per_sm = N == 31 ? 384 : 320, sm_per_sm = N == 31 ? 256 : 224;
    const bool wl = P.method_kind == SADMC_METHOD_WL || P.method_kind == SADMC_METHOD_INV_T_WL;
    const double wave_ratio = N == 31 ? (wl ? 1.19 : 1.44) : (wl ? 1.16 : 1.40);
    bool stream = (P.flags & SADMC_FLAG_LJ_STREAM_Z) != 0;
    if (!(P.flags & (SADMC_FLAG_LJ_STREAM_Z | SADMC_FLAG_LJ_SMEM_Z))) {
      int dev = 0, sms = 148;
      if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      const long long w = P.n_walkers, zg_wave = zg_per_sm * sms, sm_wave = sm_per_sm * sms;
      stream = (double)((w + zg_wave - 1) / zg_wave) * wave_ratio < (double)((w + sm_wave - 1) / sm_wave);
    }
    if (stream) {
      if (N == 31)
        use_stream<LjThreadSys<true, 31, 1, 0, true>>(P, out);
      else
        use_stream<LjThreadSys<true, 38, 1, 0, true>>(P, out);
    }
#endif
  }
  else
    *out = make_set<LjThreadSys<true, 0, 1>, true>(P);
  return true;
}
} // namespace sadmc
