// make_set.cuh -- instantiates every kernel of one system type (included by the kernels_*.cu units).
#pragma once
#include "kernel_set.cuh"
#include "book_binning.cuh"
#include "book_linear.cuh"
#include "move_kernel.cuh"
#include "replicas.cuh"
#include "tempering.cuh"

namespace sadmc {

// BINNING: also build the energy_binning.rs move kernels (the `binning` binary's bookkeeping) and the replica-exchange
// move kernel (the `tempering` binary) for this system
template <class Sys, bool BINNING = false>
static KernelSet make_set(const DevParams& P) {
  KernelSet k;
  memset(&k, 0, sizeof k);
  k.move[SADMC_METHOD_SAD] = move_kernel<Sys, SADMC_METHOD_SAD>;
  k.move[SADMC_METHOD_SAMC] = move_kernel<Sys, SADMC_METHOD_SAMC>;
  k.move[SADMC_METHOD_WL] = move_kernel<Sys, SADMC_METHOD_WL>;
  k.move[SADMC_METHOD_INV_T_WL] = move_kernel<Sys, SADMC_METHOD_WL>;
  k.move[SADMC_METHOD_CANONICAL] = move_kernel<Sys, SADMC_METHOD_CANONICAL>;
  if constexpr (BINNING) {
    k.move_binning[SADMC_METHOD_SAD] = move_kernel_binning<Sys, SADMC_METHOD_SAD>;
    k.move_binning[SADMC_METHOD_SAMC] = move_kernel_binning<Sys, SADMC_METHOD_SAMC>;
    k.move_binning[SADMC_METHOD_WL] = move_kernel_binning<Sys, SADMC_METHOD_WL>;
    k.move_binning[SADMC_METHOD_INV_T_WL] = move_kernel_binning<Sys, SADMC_METHOD_WL>;
    k.temper = temper_move_kernel<Sys>;
    k.replica_init = replica_init_kernel<Sys>;
    k.replica_move = replica_move_kernel<Sys>;
    if constexpr (Sys::G == 1) {
      k.move_linear[SADMC_METHOD_SAD] = move_kernel_linear<Sys, SADMC_METHOD_SAD>;
      k.move_linear[SADMC_METHOD_SAMC] = move_kernel_linear<Sys, SADMC_METHOD_SAMC>;
      k.move_linear[SADMC_METHOD_WL] = move_kernel_linear<Sys, SADMC_METHOD_WL>;
      k.move_linear[SADMC_METHOD_INV_T_WL] = move_kernel_linear<Sys, SADMC_METHOD_WL>;
    }
  }
  k.init = init_kernel<Sys>;
  k.shim = shim_kernel<Sys>;
  k.G = Sys::G;
  k.block = Sys::BLOCK;
  k.smem = zig_smem_bytes<Sys>() + Sys::smem_bytes(P, Sys::BLOCK);
  return k;
}

} // namespace sadmc
