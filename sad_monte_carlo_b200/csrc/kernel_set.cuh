// kernel_set.cuh -- the kernels of one system type, as function pointers the host code launches.
//
// Each system family is instantiated in its own translation unit (kernels_*.cu) so that the library
// builds in parallel; engine.cu only sees these factory functions.
#pragma once
#include <cstring>

#include "book.cuh"

namespace sadmc {

// trait-shaped single-walker shims (src/system/mod.rs:54-120)
enum SysOp { OP_ENERGY = 0, OP_COMPUTE_ENERGY = 1, OP_PLAN_MOVE = 2, OP_CONFIRM = 3, OP_VERIFY = 4, OP_RANDOMIZE = 5 };

struct ShimOut {
  double value;
  int some;
  int ok;
};

// one replica slot of a tempering simulation (tempering.cuh; `Replica`, src/mc/tempering.rs:46-73, minus system and generator)
struct TemperRec {
  double T;
  unsigned long long rejected, accepted, rejected_swap, accepted_swap, ignored;
  double total_energy, total_energy_squared;
  double tscale; // Replica::translation_scale: 1.0 from the constructor (tempering.rs:88), a serialised field otherwise
};

// one replica slot of a `replicas` simulation (replicas.cuh; `Replica`, src/mc/energy_replicas.rs:103-145, minus system and generator)
struct ReplicaRec {
  double max_energy, cutoff, lowest_max, tscale;
  unsigned long long rejected, accepted, above_count, below_count, upwelling, unique_visitors;
  double above_total, below_total, above_sq, below_sq;
  double xtot; // above_extra of the system's one data_to_collect key: total, count
  unsigned long long xcnt;
  int collecting, pad;
};
// one simulation (`MC`, energy_replicas.rs:307-333, minus the replicas)
struct ReplicaSim {
  unsigned long long s0, s1; // MC::rng
  unsigned long long moves, indep;
  double min_T;
  int n_rep, median_len, overflow, pad;
};

typedef void (*temper_fn)(const DevParams, TemperRec*, unsigned long long);
typedef void (*replica_init_fn)(const DevParams, unsigned long long, uint32_t, uint32_t, double*, uint32_t, unsigned long long*);
typedef void (*replica_move_fn)(const DevParams, ReplicaRec*, const ReplicaSim*, uint32_t, unsigned long long, double*);
typedef void (*move_fn)(const DevParams, unsigned long long, unsigned long long);
typedef void (*init_fn)(const DevParams, unsigned long long, int, long long, int, double, unsigned long long);
typedef void (*shim_fn)(const DevParams, uint32_t, int, double, ShimOut*, double*);

struct KernelSet {
  move_fn move[6]; // indexed by sadmc_method_kind (WL and INV_T_WL share)
  move_fn move_binning[6]; // SADMC_FLAG_BINNING: energy_binning.rs bookkeeping (book_binning.cuh); null where not built
  move_fn move_linear[6];  // ... | SADMC_FLAG_BINNING_LINEAR: the same over binning::linear (book_linear.cuh)
  temper_fn temper; // Replica::run_once x steps (tempering.cuh); null where not built
  replica_init_fn replica_init; // energy_replicas.rs (replicas.cuh); null where not built
  replica_move_fn replica_move;
  init_fn init;
  shim_fn shim;
  int G, block;
  size_t smem;
  // when the move kernels are launched differently from init / shim (helper warps): 0 = as above
  int move_block, move_threads_per_walker;
  size_t move_smem;
  int zstream_per_thread; // doubles of DevParams::zstream per move-kernel thread (0: none)
};

// factories, one per translation unit; the bool ones return false when no instance was built
KernelSet kernels_ising(const DevParams& P);
KernelSet kernels_fake(const DevParams& P);
KernelSet kernels_two_wells(const DevParams& P);
KernelSet kernels_erfinv(const DevParams& P);
KernelSet kernels_cell_fluid(bool square_well, const DevParams& P);
bool kernels_wca_group(int G, bool fast, const DevParams& P, KernelSet* out);
bool kernels_lj_thread_exact(int N, const DevParams& P, KernelSet* out);
bool kernels_lj_thread_fast(int N, int G, const DevParams& P, KernelSet* out);
bool kernels_lj_thread_paired(int N, const DevParams& P, KernelSet* out);
bool kernels_lj_thread_fast_multi(int N, int G, const DevParams& P, KernelSet* out);
bool kernels_lj_warp(int G, int A, const DevParams& P, KernelSet* out);
bool kernels_lj_warp_small(int G, int A, const DevParams& P, KernelSet* out);

} // namespace sadmc
