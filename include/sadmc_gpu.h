/* sadmc_gpu.h -- C ABI of the B200 walker engine (libsadmc_gpu.so).
 *
 * This is the drop-in boundary for ONE path of droundy/sad-monte-carlo: the
 * propose / dE / accept loop of `EnergyMC::move_once` (src/mc/energy.rs:904-974)
 * over the `System`/`ConfirmSystem`/`MovableSystem` traits
 * (src/system/mod.rs:54-120), for thousands of independent walkers at once.
 *
 * The reference has no FFI for this path (it is Rust generics,
 * `EnergyMC<S: MovableSystem>` energy.rs:827).  A per-move FFI call would be
 * slower than the CPU move itself, so the boundary sits one level up: the host
 * asks the engine to advance ALL walkers by `n` moves, where `n` is what the
 * reference's `PluginManager::run` computes as the next `period`
 * (src/mc/plugin.rs:93-144).  Report/Save/Movie stay on the host.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function
 * returns 0 on success or a negative sadmc_status; the message for the last
 * failure of the calling thread is sadmc_last_error().  Nothing unwinds across
 * this boundary (the reference panics instead: ising.rs:39, wca.rs:186-191,
 * two_wells.rs:240-245, lj.rs:259).  A handle is single-owner and not
 * thread-safe (the reference's MC is `&mut self` everywhere).  `Option<f64>`
 * fields of the reference are doubles where NaN means `None`.
 *
 * There is NO CPU fallback: sadmc_create fails with SADMC_ERR_CUDA when no
 * device is usable.
 */
#ifndef SADMC_GPU_H
#define SADMC_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SADMC_ABI_VERSION 1

typedef enum {
  SADMC_OK = 0,
  SADMC_ERR_INVALID = -1,     /* bad config / argument (reference: panic!/assert!) */
  SADMC_ERR_CUDA = -2,        /* CUDA runtime failure, or no device              */
  SADMC_ERR_WINDOW = -3,      /* a walker's energy left the device bin window     */
  SADMC_ERR_UNSUPPORTED = -4, /* valid in the reference, not built here           */
  SADMC_ERR_VERIFY = -5       /* verify_energy failed (lj.rs:249-261 etc.)        */
} sadmc_status;

/* AnyParams variants, src/system/any.rs:10-27 */
typedef enum {
  SADMC_SYS_FAKE = 1,        /* fake.rs            */
  SADMC_SYS_FAKE_ERFINV = 2, /* erfinv.rs          */
  SADMC_SYS_WCA = 3,         /* wca.rs + optcell.rs */
  SADMC_SYS_LJ = 4,          /* lj.rs              */
  SADMC_SYS_ISING = 6,       /* ising.rs           */
  SADMC_SYS_SW = 7,          /* optsquare.rs       */
  SADMC_SYS_TWO_WELLS = 8    /* two_wells.rs       */
} sadmc_system_kind;

/* fake::Function, src/system/fake.rs:12-36 */
typedef enum {
  SADMC_FAKE_LINEAR = 0,
  SADMC_FAKE_QUADRATIC = 1,
  SADMC_FAKE_PIECES = 2,
  SADMC_FAKE_GAUSSIAN = 3
} sadmc_fake_function;

/* MethodParams, src/mc/energy.rs:44-69 (Method at run time: 213-244) */
typedef enum {
  SADMC_METHOD_SAD = 1,
  SADMC_METHOD_SAMC = 2,
  SADMC_METHOD_WL = 3,
  SADMC_METHOD_INV_T_WL = 4,
  SADMC_METHOD_CANONICAL = 5
} sadmc_method_kind;

/* MoveParams, src/mc/energy.rs:73-78 */
typedef enum { SADMC_MOVE_TRANSLATION_SCALE = 0, SADMC_MOVE_ACCEPTANCE_RATE = 1 } sadmc_move_plan;

/* How walker w gets its first configuration. */
typedef enum {
  /* The reference constructor (`Any::from(AnyParams)`, any.rs:80-93): every
   * walker starts from the SAME configuration; walker w then equals a
   * reference process run with `--seed (seed+w)`. */
  SADMC_INIT_REFERENCE = 0,
  /* `System::randomize` (e.g. lj.rs:262-279) driven by the walker's own MC
   * stream, then the reference's downhill relaxation (energy.rs:840-851).
   * Used by the throughput configs ("synthetic random-start walkers"). */
  SADMC_INIT_RANDOMIZE = 1,
  /* Leave systems unset; the caller supplies them with sadmc_set_system and
   * then calls sadmc_start (resume path, mc/mod.rs:70-84). */
  SADMC_INIT_EXTERNAL = 2
} sadmc_init_mode;

#define SADMC_FLAG_NO_ROUND_TRIPS 1u /* skip energy.rs:950-965 diagnostics (never read by the sampler) */
#define SADMC_FLAG_SUM_TREE 2u       /* reserved (oracle only): sum LJ pair terms in the kernel's lane order */
/* LJ, lanes_per_walker = 1 only: FMA-contracted pair arithmetic with a Newton reciprocal and a
 * warp-cooperative energy recomputation instead of the reference's exact operation order.
 * Per-move energies then agree with the reference to <= 1e-12 relative instead of bit for bit.
 * WCA (lanes_per_walker 4/8/16): the same arithmetic shortcuts, and the whole-system energy is re-summed every
 * 65 536 accepted moves instead of every ~10 (wca.rs:164-177 with a 6 553.6 times larger error allowance). */
#define SADMC_FLAG_FAST_MATH 4u
/* EXPERIMENT, with SADMC_FLAG_FAST_MATH and lanes_per_walker = 1, LJ31 / LJ38 only: a helper warp per bookkeeping warp
 * sums half of the pair loop (csrc/sys_lj_paired.cuh).  Same tolerance tier as SADMC_FLAG_FAST_MATH. */
#define SADMC_FLAG_HELPER_WARPS 8u
/* The bookkeeping of the reference's `binning` binary instead of `histogram`'s: `EnergyMC` of src/mc/energy_binning.rs
 * (reject_move 276-321, update_weights 323-503, gamma 507-533, move_once 592-633) over `binning::histogram::Bins`
 * (src/mc/binning/histogram.rs; `--histogram-bin`, what fake/run-fake.py:25-26 runs).  `energy_bin` is the histogram bin.
 * Methods: SAD, SAMC, WL, 1/t-WL.  State out: sadmc_get_binning_walker / sadmc_get_binning_bins (sadmc_get_bins,
 * sadmc_set_walker_bins, sadmc_resume and sadmc_set_lnw are for the energy.rs layout and refuse such an engine).
 * Built for the one-thread-per-walker systems (Ising, fake, two-wells, erfinv, LJ with lanes_per_walker = 1) and the
 * warp-per-walker fluids (square well, WCA with lanes_per_walker = 32).  Not built: `binning::linear`. */
#define SADMC_FLAG_BINNING 16u
/* With SADMC_FLAG_BINNING: `BinningParams::Linear { bin }` (binning.rs:57-60, src/mc/binning/linear.rs) instead of the histogram:
 * ln w, the counts (f64 there) and every accumulator are spread over the two neighbouring bin points in proportion to the
 * distance and read back by linear interpolation.  One-thread-per-walker systems only; state out: sadmc_get_binning_walker
 * (lnw_max_count_f64 / hist_min_count_f64) and sadmc_get_binning_bins_f64.  A correctness path: every access goes to HBM
 * uncached (no job script of the reference uses --linear-bin). */
#define SADMC_FLAG_BINNING_LINEAR 32u
/* LJ31 and LJ38, SADMC_FLAG_FAST_MATH, one lane per walker: the histogram move kernels exist in two layouts with the same results
 * bit for bit (tests/test_gpu_lj.py) -- all three coordinates in shared memory (two 128-thread CTAs per SM, 249 registers),
 * or x and y in shared memory and z streamed from an L2-resident array through a cp.async ring (three CTAs per SM, 168
 * registers; LJ31 ~4 % faster when the walkers fill whole waves of 384 per SM; LJ38: one 320-thread CTA instead of one of 224,
 * 1/t-WL + 23 %).  By default the engine takes whichever needs less time for the walker count and method; these flags force one. */
#define SADMC_FLAG_LJ_SMEM_Z 64u
#define SADMC_FLAG_LJ_STREAM_Z 128u

typedef struct sadmc_config {
  uint32_t abi_version; /* = SADMC_ABI_VERSION */
  int32_t system;       /* sadmc_system_kind */

  /* --- system parameters (the `--<sys>-*` flags of the reference CLI) --- */
  uint32_t N;              /* ising N (ising.rs:14) | lj N (lj.rs:16) | wca N (wca.rs:375) | sw N |
                              two-wells N (two_wells.rs:15) | erfinv N | fake quadratic `dimensions` */
  double lj_radius;        /* lj.rs:18 */
  double reduced_density;  /* wca: CellDimensionsGivenNumber::ReducedDensity (wca.rs:365) */
  double filling_fraction; /* sw: FillingFraction (optsquare.rs:337) */
  double cell_width[3];    /* wca/sw: CellWidth; used when cell_width[0] > 0 */
  double sw_well_width;    /* optsquare.rs:345 */
  int32_t fake_function;   /* sadmc_fake_function */
  int32_t _pad0;
  double fake_a, fake_b, fake_e1, fake_e2; /* Function::Pieces, fake.rs:21-30 */
  double fake_sigma;                       /* Function::Gaussian, fake.rs:32-35 */
  double tw_h2_to_h1, tw_barrier_over_h1, tw_r2; /* two_wells.rs:13-22 */
  double erfinv_mean_energy;                     /* erfinv.rs:12-15 */

  /* --- EnergyMCParams, energy.rs:82-97 --- */
  int32_t method; /* sadmc_method_kind */
  int32_t move_plan;
  double sad_min_T;          /* MethodParams::Sad     */
  double samc_t0;            /* MethodParams::Samc    */
  double wl_min_gamma;       /* MethodParams::WL, NaN = None */
  double canonical_T;        /* MethodParams::_Canonical */
  uint64_t seed;             /* walker w (global index) is seeded `seed + w`; None == 0 (energy.rs:835) */
  double energy_bin;         /* NaN = None -> delta_energy() or 1.0 (energy.rs:831-833) */
  double min_allowed_energy; /* NaN = None */
  double max_allowed_energy; /* NaN = None */
  double move_value;         /* TranslationScale(sigma) or AcceptanceRate(r) */

  /* --- engine --- */
  uint32_t n_walkers;     /* walkers held by THIS engine (this GPU) */
  uint32_t walker_offset; /* global index of local walker 0 (rank * n_walkers when sharded) */
  int32_t device;         /* CUDA ordinal */
  int32_t init_mode;      /* sadmc_init_mode */
  /* Energy window [lo, hi) that the per-walker device bin arrays must cover.
   * The reference grows its vectors without bound (energy.rs:400-434); the
   * engine pre-allocates and a walker that leaves the window stops with
   * SADMC_ERR_WINDOW.  NaN = derive from min/max_allowed_energy and the
   * system's lowest_possible_energy(). */
  double bin_window_lo, bin_window_hi;
  int32_t lanes_per_walker; /* LJ: 0 = auto; 32/16/8/4 = that many lanes of a warp cooperate on one walker
                               (registers + shuffles); 1 = one thread per walker, cluster in shared memory.
                               WCA: 0 = auto (8 with SADMC_FLAG_FAST_MATH, else 32); 4/8/16 = lanes sharing a walker's cell-list
                               lookups; 32 = a warp per walker */
  uint32_t flags;
  /* SADMC_FLAG_BINNING only: `high_resolution_de` (energy_binning.rs:62-63, 124-125, 328-330): a second, finer histogram of
   * the visited energies that rides along with the weights' bins (counts only).  NaN or <= 0 = None. */
  double high_resolution_de;
} sadmc_config;

/* Per-walker scalars: the non-vector fields of `EnergyMC` (energy.rs:167-210)
 * and of `Method` (213-244), as a checkpoint would hold them. */
typedef struct sadmc_walker_state {
  uint64_t moves, accepted_moves;
  double acceptance_rate, translation_scale;
  uint64_t rng_s0, rng_s1; /* rand_xoshiro serde fields s0,s1 */
  double energy;           /* system.energy() */
  double bins_min, bins_width;
  uint32_t bins_len;     /* bins.lnw.len() */
  uint32_t window_first; /* device window index of reference bin 0 */
  int32_t method;        /* current sadmc_method_kind (1/t-WL may have become SAMC, energy.rs:754-756) */
  int32_t status;        /* sadmc_status of this walker */
  /* Sad */
  double too_lo, too_hi, latest_parameter;
  uint64_t tL, tF, num_states, highest_hist;
  /* Samc */
  double samc_t0;
  /* WL */
  double wl_gamma, wl_num_states, wl_min_energy;
  uint64_t wl_lowest_hist, wl_highest_hist, wl_total_hist;
  uint32_t wl_hist_len;
  int32_t wl_inv_t;
  /* round trips */
  double max_S;
  uint32_t max_S_index; /* reference index (relative to bins_min) */
  uint32_t _pad;
} sadmc_walker_state;

/* SADMC_FLAG_BINNING: the non-vector fields of energy_binning.rs's `EnergyMC` (92-126), `Method` (128-148) and
 * `histogram::Bins` (histogram.rs:99-111), plus the `BinCounts` aggregates (histogram.rs:12-32) that the sampler reads. */
typedef struct sadmc_binning_state {
  uint64_t moves, accepted_moves;
  double acceptance_rate, translation_scale;
  uint64_t rng_s0, rng_s1;
  double energy;
  double bins_min, bins_width; /* Bins::min, Bins::width */
  double bins_min_e, bins_max_e; /* lowest / highest energy counted so far (histogram.rs:182-187) */
  uint32_t bins_len;     /* lnw.total.len(): 0 until the first move */
  uint32_t window_first; /* device window index of bin 0 */
  int32_t method;        /* current sadmc_method_kind (1/t-WL may have become SAMC, energy_binning.rs:496-498) */
  int32_t status;
  /* Sad (energy_binning.rs:131-139): too_lo / too_hi are raw energies, tF an f64 */
  double too_lo, too_hi, latest_parameter, tF;
  uint64_t tL, num_states;
  double samc_t0;  /* Samc */
  double wl_gamma; /* WL */
  int32_t wl_inv_t, _pad;
  uint64_t lnw_max_count, lnw_total_count;  /* bins.lnw.max_count (SAD's old_highest_hist), bins.lnw.total_count */
  double t_found_max_total;                 /* bins.extra["t_found"].max_total (SAD's tF source) */
  uint64_t hist_min_count, hist_total_count; /* bins.extra["hist"] (WL flatness) */
  /* the same two aggregates as f64: what binning::linear keeps (its counts are f64, linear.rs:22-28); for the histogram
   * variant the integer values converted */
  double lnw_max_count_f64, hist_min_count_f64;
} sadmc_binning_state;

typedef struct sadmc_engine sadmc_engine;

/* ---- lifecycle --------------------------------------------------------- */
/* `EnergyMC::from_params` for every walker (energy.rs:830-898) + the system
 * constructor selected by cfg->init_mode. */
int sadmc_create(const sadmc_config* cfg, sadmc_engine** out);
void sadmc_destroy(sadmc_engine* e);
const char* sadmc_last_error(void);
int sadmc_abi_version(void);

/* The reference's system constructor `Any::from(AnyParams)` (any.rs:80-93: lj.rs:126-205,
 * ising.rs:31-53, optsquare.rs:357-436, wca.rs:392-498 with fcc = false, fake.rs:85-93,
 * two_wells.rs:248-250, erfinv.rs:50-58) run on the HOST: one system image in the layout of
 * sadmc_get_system.  It is what SADMC_INIT_REFERENCE replicates to every walker, exported so a
 * host can write the `system` of a checkpoint before any move.  Needs no device.  `*needed`
 * (optional) receives the image length in doubles; buf == NULL only queries it. */
int sadmc_reference_system(const sadmc_config* cfg, double* buf, size_t n, size_t* needed);

/* With SADMC_INIT_EXTERNAL: finish from_params (relaxation + first bin) after
 * the systems were supplied. */
int sadmc_start(sadmc_engine* e);

/* Launch on this CUDA stream (a cudaStream_t) instead of the engine's own. */
int sadmc_set_stream(sadmc_engine* e, void* cuda_stream);
void* sadmc_get_stream(sadmc_engine* e);

/* ---- the hot path ------------------------------------------------------ */
/* n_moves x `move_once` (energy.rs:904-974) for every walker.  Blocking.
 * Returns SADMC_ERR_WINDOW when walkers left the device bin window during this call (the reference would have grown
 * its vectors, energy.rs:400-434; here such a walker freezes with its status set and the caller is told at once),
 * SADMC_ERR_VERIFY when the in-loop `verify_energy` of energy.rs:907-911 failed for a walker (the reference
 * panics: lj.rs:259, wca.rs:248, optsquare.rs:200).  Each halted walker is reported by one call only; the other
 * walkers have completed their n_moves and the engine stays usable. */
int sadmc_run(sadmc_engine* e, uint64_t n_moves);
/* Same, but only enqueues on the engine's stream. */
int sadmc_run_async(sadmc_engine* e, uint64_t n_moves);
/* Waits for the stream; reports newly halted walkers like sadmc_run. */
int sadmc_sync(sadmc_engine* e);
/* Walkers halted since creation: left the bin window / failed verify_energy.  Either pointer may be NULL. */
int sadmc_num_halted(sadmc_engine* e, uint64_t* left_window, uint64_t* failed_verify);
/* Device time of the move kernel(s) of the last sadmc_run, CUDA events. */
int sadmc_last_run_ms(sadmc_engine* e, float* ms);
/* How many kernels of this library have been launched by this engine. */
int sadmc_launch_count(sadmc_engine* e, uint64_t* n);
/* How the move kernel of this engine is launched: threads per CTA, threads per walker, dynamic shared memory per CTA, and
 * the bytes of walker state streamed from L2 per walker (0 unless the LJ z stream is in use, see SADMC_FLAG_LJ_STREAM_Z).
 * Any pointer may be null.  (GPU-side diagnostic: the reference has no counterpart.) */
int sadmc_move_launch_shape(sadmc_engine* e, uint32_t* block, uint32_t* threads_per_walker, uint64_t* shared_bytes, uint32_t* stream_bytes_per_walker);

/* ---- state out (what Report/Save/Movie and the parity tests read) ------ */
int sadmc_num_moves(sadmc_engine* e, uint64_t* moves);                  /* MonteCarlo::num_moves, energy.rs:981 */
int sadmc_num_accepted_moves(sadmc_engine* e, uint64_t* accepted_sum);  /* energy.rs:984, summed over walkers */
/* Smallest and largest per-walker accepted-move count (what one reference process would report lies in between). */
int sadmc_accepted_moves_range(sadmc_engine* e, uint64_t* min_accepted, uint64_t* max_accepted);
int sadmc_get_walker(sadmc_engine* e, uint32_t w, sadmc_walker_state* out);
int sadmc_get_energies(sadmc_engine* e, double* energies /* [n_walkers] */);
/* `Bins` (energy.rs:146-163) + round-trip vectors (203-205) of walker w, in
 * reference index order (element 0 = bin at bins_min).  Any pointer may be
 * NULL.  `cap` is the capacity of every non-NULL array; fails if < bins_len. */
int sadmc_get_bins(sadmc_engine* e, uint32_t w, uint32_t cap, uint64_t* histogram, uint64_t* t_found,
                   double* lnw, double* energy_total, double* energy_squared_total,
                   uint64_t* round_trips, uint8_t* have_visited_since_maxentropy, uint64_t* wl_hist,
                   double* extra_total, uint64_t* extra_count);
/* SADMC_FLAG_BINNING engines: scalars and per-bin vectors of walker w in reference index order (element 0 = bin at
 * bins_min): bins.lnw.{total,count}, bins.extra["energy"|"t_found"].{total,count}, bins.extra["hist"].count (its total
 * is always 0), and the system's data_to_collect accumulator ("pressure" / "which").  Any pointer may be NULL. */
int sadmc_get_binning_walker(sadmc_engine* e, uint32_t w, sadmc_binning_state* out);
int sadmc_get_binning_bins(sadmc_engine* e, uint32_t w, uint32_t cap, double* lnw_total, uint64_t* lnw_count,
                           double* energy_total, uint64_t* energy_count, double* t_found_total, uint64_t* t_found_count,
                           uint64_t* hist_count, double* extra_total, uint64_t* extra_count);
/* The same with every count as f64: required for SADMC_FLAG_BINNING_LINEAR engines (their counts are fractional), exact
 * for histogram engines below 2^53. */
int sadmc_get_binning_bins_f64(sadmc_engine* e, uint32_t w, uint32_t cap, double* lnw_total, double* lnw_count,
                               double* energy_total, double* energy_count, double* t_found_total, double* t_found_count,
                               double* hist_count, double* extra_total, double* extra_count);
/* The `high_resolution` histogram (cfg->high_resolution_de) of walker w: `histogram::Bins::min`, the counts in reference
 * index order (the `lnw.count` of that Bins; its totals are all 0), their number in *len.  cap = capacity of `count`. */
int sadmc_get_high_resolution(sadmc_engine* e, uint32_t w, uint32_t cap, double* bins_min, uint32_t* len, uint64_t* count);
int sadmc_set_high_resolution(sadmc_engine* e, uint32_t w, double bins_min, uint32_t len, const uint64_t* count); /* resume */
/* Resume of a SADMC_FLAG_BINNING engine (created with SADMC_INIT_EXTERNAL): sadmc_set_system(s), then this for every
 * walker -- the inverse of the two getters above, same field meanings (t_found_*, hist_count, extra_* may be NULL) -- then
 * sadmc_resume(e, moves).  The resumed engine continues bit for bit like the one that was checkpointed. */
int sadmc_set_binning_walker(sadmc_engine* e, uint32_t w, const sadmc_binning_state* s, const double* lnw_total,
                             const uint64_t* lnw_count, const double* energy_total, const uint64_t* energy_count,
                             const double* t_found_total, const uint64_t* t_found_count, const uint64_t* hist_count,
                             const double* extra_total, const uint64_t* extra_count);
/* System configuration of walker w as f64s.  Layout: LJ/WCA/SW: x0,y0,z0,x1,..
 * (3N), then E, then error (WCA/LJ; 0 for SW).  Fake/ErfInv/TwoWells:
 * position[dim] (+ d_squared for two-wells).  Ising: N*N spins as +-1.0, then E. */
int sadmc_system_len(sadmc_engine* e, size_t* n_doubles);
int sadmc_get_system(sadmc_engine* e, uint32_t w, double* buf, size_t n);
int sadmc_set_system(sadmc_engine* e, uint32_t w, const double* buf, size_t n);
/* Bulk variants: [n_walkers][system_len] row-major host buffers. */
int sadmc_get_systems(sadmc_engine* e, double* buf, size_t n);
int sadmc_set_systems(sadmc_engine* e, const double* buf, size_t n);
int sadmc_get_rngs(sadmc_engine* e, uint64_t* s /* [n_walkers][2] */);
int sadmc_set_rngs(sadmc_engine* e, const uint64_t* s /* [n_walkers][2] */);

/* ---- resume (mc/mod.rs:70-84: the reference deserialises the whole EnergyMC and calls update_caches) ----
 * On an engine created with SADMC_INIT_EXTERNAL: for every walker sadmc_set_system(s) and sadmc_set_walker_bins
 * (the inverse of sadmc_get_walker + sadmc_get_bins: same field meanings, reference index order, element 0 = bin at
 * s->bins_min; t_found, round_trips, have_visited, wl_hist, extra_* may be NULL), then sadmc_resume(e, moves)
 * instead of sadmc_start.  A resumed engine continues bit for bit like the one that was checkpointed
 * (tests/resume-sad.rs). */
int sadmc_set_walker_bins(sadmc_engine* e, uint32_t w, const sadmc_walker_state* s, const uint64_t* histogram,
                          const uint64_t* t_found, const double* lnw, const double* energy_total,
                          const double* energy_squared_total, const uint64_t* round_trips,
                          const uint8_t* have_visited_since_maxentropy, const uint64_t* wl_hist,
                          const double* extra_total, const uint64_t* extra_count);
int sadmc_resume(sadmc_engine* e, uint64_t moves);

/* ---- fixed weights: a production run -------------------------------------------------
 * Every walker's ln w := lnw_window[j] for window bin j (n must equal sadmc_window's nbins).  With
 * SADMC_METHOD_SAMC and samc_t0 = 0 (gamma = t0 / t = 0, energy.rs:816-820) the weights then never change: all walkers
 * sample the same multicanonical ensemble, their histograms add up, and S(E) = ln w(E) + ln H(E) + const holds without
 * any bias from the weight-learning phase -- the many-walker counterpart of the reference's WL production mode
 * (energy.rs:649-655).  tools/lj31_production.py uses it for the heat capacity of LJ31. */
int sadmc_set_lnw(sadmc_engine* e, const double* lnw_window, uint32_t n);

/* ---- window geometry ---------------------------------------------------- */
/* Device window: bin j of every walker covers [lo + j*width, lo + (j+1)*width). */
int sadmc_window(sadmc_engine* e, double* lo, double* width, uint32_t* nbins);
/* `Cell::box_diagonal` and `Cell::r_cutoff` (optcell.rs:27-33) of the periodic fluids, as the engine derived them
 * from CellWidth / CellVolume / ReducedDensity / FillingFraction (wca.rs:396-401, optsquare.rs:360-368, optcell.rs:44-50):
 * what a checkpoint's `cell` records.  SADMC_ERR_INVALID for the other systems. */
int sadmc_cell_box(sadmc_engine* e, double box_diagonal[3], double* r_cutoff);

/* ---- merge for reporting ------------------------------------------------ */
/* Fold the local walkers' bins into window-aligned sums, written to DEVICE
 * buffers of length nbins (caller-owned, e.g. torch tensors; a following
 * NCCL all-reduce(sum) merges GPUs): histogram (u64), energy_total,
 * energy_squared_total (f64), lnw_sum / lnw_sq_sum / lnw_count: sum, sum of
 * squares and number of walkers contributing their max-aligned lnw
 * (plotting/parse-binning.py:169 alignment).  Any pointer may be NULL. */
/* Which walkers the following folds merge: first_walker, first_walker + walker_stride, ... (default 0, 1 =
 * all; interleaved groups give ensemble error bars).  sad_range_only != 0: a SAD walker contributes its
 * ln w only for bins inside its own [too_lo, too_hi], the range in which SAD defines ln w
 * (plotting/parse-binning.py:150-164 reconstructs the rest from the histogram). */
int sadmc_fold_select(sadmc_engine* e, uint32_t first_walker, uint32_t walker_stride, int sad_range_only);
/* Same with at most `walker_count` walkers (0 = to the end): with stride 1 a contiguous block, i.e. the shard a rank of
 * a multi-GPU run would hold.  sad_range_only = 2 counts only the bins STRICTLY inside (too_lo, too_hi): the two end
 * bins are centred on too_lo / too_hi and `update_weights` reverts the increment for the half of their visits that
 * lies outside the range (energy.rs:535-538), so their ln w is not an estimate of the entropy. */
int sadmc_fold_select_ex(sadmc_engine* e, uint32_t first_walker, uint32_t walker_stride, uint32_t walker_count,
                         int sad_range_only);
/* tl_max != 0: a SAD walker contributes its ln w to the following folds only if the bins of its range
 * [too_lo, too_hi] have not changed since move tl_max: a bin that has just joined a walker's range starts from a copied
 * or zero ln w (energy.rs:544-584) and settles slowly at gamma ~ 1/t.  (An engine-side record of the last range
 * change is used, not `Sad::tL`: the reference refreshes tL far more often.)  0 (default) = every walker. */
int sadmc_fold_settled(sadmc_engine* e, uint64_t tl_max);
int sadmc_fold_device(sadmc_engine* e, void* d_histogram, void* d_energy_total, void* d_energy_squared_total,
                      void* d_lnw_sum, void* d_lnw_sq_sum, void* d_lnw_count);
/* The same fold as ONE device buffer of 7 x nbins doubles, for a single collective: histogram >> 32,
 * histogram & 0xffffffff (both exact in f64, also after summing over ranks), lnw_count, energy_total,
 * energy_squared_total, lnw_sum, lnw_sq_sum. */
int sadmc_fold_packed_device(sadmc_engine* e, void* d_packed);
/* Same into HOST buffers (single-GPU convenience). */
int sadmc_fold(sadmc_engine* e, uint64_t* histogram, double* energy_total, double* energy_squared_total,
               double* lnw_sum, double* lnw_sq_sum, uint64_t* lnw_count);

/* ---- trait-shaped single-walker shims (src/system/mod.rs:54-120) -------- */
/* For parity tests and for a Rust `impl MovableSystem for GpuSystem`.  Each is
 * one tiny kernel launch; never use them in a loop that matters. */
int sadmc_sys_energy(sadmc_engine* e, uint32_t w, double* energy);            /* System::energy          */
int sadmc_sys_compute_energy(sadmc_engine* e, uint32_t w, double* energy);    /* System::compute_energy  */
/* MovableSystem::plan_move: draws from walker w's RNG; *some = 0 is `None`. */
int sadmc_sys_plan_move(sadmc_engine* e, uint32_t w, double mean_distance, int* some, double* e_new);
int sadmc_sys_confirm(sadmc_engine* e, uint32_t w);                           /* ConfirmSystem::confirm  */
/* System::randomize (system/mod.rs:59; lj.rs:262-279, wca.rs:252-270, fake.rs:102-111): draws from walker w's RNG. */
int sadmc_sys_randomize(sadmc_engine* e, uint32_t w, double* energy);
int sadmc_sys_verify_energy(sadmc_engine* e, uint32_t w);                     /* System::verify_energy   */

/* ---- replica exchange: the `tempering` binary (src/mc/tempering.rs) ------------------------------------------
 * `n_sim` independent tempering simulations x `n_T` temperatures on one GPU; simulation k is the reference process run
 * with `--seed (cfg->seed + cfg->walker_offset + k)` (MC::from_params, tempering.rs:152-175: every replica starts from
 * the SAME system and a CLONE of the same generator; the simulation's own generator is that generator after jump()).
 * cfg: the system parameters, seed, walker_offset, device, init_mode (SADMC_INIT_REFERENCE, or SADMC_INIT_EXTERNAL +
 * sadmc_tempering_set_system) and n_walkers = n_sim; method / bins are not used.  Same systems as SADMC_FLAG_BINNING. */
typedef struct sadmc_tempering sadmc_tempering;
/* `Replica` (tempering.rs:46-73) without its system */
typedef struct sadmc_replica_state {
  double T;
  uint64_t rejected_count, accepted_count, rejected_swap_count, accepted_swap_count, ignored_count;
  double total_energy, total_energy_squared;
  double translation_scale; /* 1.0 from the constructor (tempering.rs:88) unless sadmc_tempering_set_translation_scales */
  uint64_t rng_s0, rng_s1;
  double energy; /* system.energy() */
} sadmc_replica_state;
int sadmc_tempering_create(const sadmc_config* cfg, const double* T, uint32_t n_T, uint64_t canonical_steps, sadmc_tempering** out);
void sadmc_tempering_destroy(sadmc_tempering* t);
/* n_rounds x `MC::run_once` (tempering.rs:272-342): min_moves_to_randomize() * canonical_steps moves for every replica
 * (one launch), then one swap attempt per neighbouring pair (one launch).  Blocking. */
int sadmc_tempering_run(sadmc_tempering* t, uint64_t n_rounds);
int sadmc_tempering_num_moves(sadmc_tempering* t, uint64_t* moves);        /* MC::moves of each simulation */
int sadmc_tempering_steps_per_round(sadmc_tempering* t, uint64_t* steps); /* moves per replica and round (274) */
int sadmc_tempering_get_replicas(sadmc_tempering* t, uint32_t sim, sadmc_replica_state* out /* [n_T] */);
int sadmc_tempering_get_rng(sadmc_tempering* t, uint32_t sim, uint64_t s[2]); /* MC::rng */
/* Resume (tempering.rs:196-213 deserialises the whole MC): counters, moments, translation scales and generators of the
 * replicas of one simulation, the simulation's own generator, and MC::moves (a multiple of steps x replicas); systems go
 * back with sadmc_tempering_set_system. */
int sadmc_tempering_set_replicas(sadmc_tempering* t, uint32_t sim, const sadmc_replica_state* in /* [n_T] */);
int sadmc_tempering_set_rng(sadmc_tempering* t, uint32_t sim, const uint64_t s[2]);
int sadmc_tempering_set_num_moves(sadmc_tempering* t, uint64_t moves);
/* `Replica::translation_scale` of replica r of every simulation := scale[r] (n_T values).  The reference's constructor
 * fixes 1.0 and has no flag for it, but the field is part of the serialised state (tempering.rs:71-72), so a resumed
 * reference run uses whatever the checkpoint holds; clusters and fluids need a step that shrinks with temperature. */
int sadmc_tempering_set_translation_scales(sadmc_tempering* t, const double* scale /* [n_T] */);
int sadmc_tempering_system_len(sadmc_tempering* t, size_t* n_doubles);
int sadmc_tempering_get_system(sadmc_tempering* t, uint32_t sim, uint32_t replica, double* buf, size_t n);
int sadmc_tempering_set_system(sadmc_tempering* t, uint32_t sim, uint32_t replica, const double* buf, size_t n);
/* sadmc_cell_box for the periodic fluids (what a checkpoint's `cell` records). */
int sadmc_tempering_cell_box(sadmc_tempering* t, double box_diagonal[3], double* r_cutoff);
/* Device time of the last sadmc_tempering_run (move and swap kernels), CUDA events. */
int sadmc_tempering_last_run_ms(sadmc_tempering* t, float* ms);

/* ---- energy-ceiling replicas: the `replicas` binary (src/mc/energy_replicas.rs) ----------------------------------
 * `n_sim` independent simulations on one GPU (simulation k = the reference process `--seed cfg->seed + cfg->walker_offset + k`),
 * each with up to `max_replicas` replica slots: a replica accepts every proposal below its max_energy, neighbours swap
 * systems when the upper one has come below the lower one's ceiling, and a new, lower replica is split off at the median of
 * the energies seen below the lowest cutoff once `independent_systems_before_new_bin` independent systems have visited it
 * (MC::run_once, energy_replicas.rs:504-600).  cfg: system parameters, seed, walker_offset, device, n_walkers = n_sim; the
 * systems need `System::randomize` (Ising, LJ, WCA, fake, erfinv; the reference leaves the square well's as todo!(), the
 * two-wells sampler is not restated).  max_init: MAX_INIT of from_params (346-368), 0 = the reference's 1 << 15. */
typedef struct sadmc_replicas sadmc_replicas;
/* `Replica` (energy_replicas.rs:103-145) without its system; above_extra holds the system's one data_to_collect key */
typedef struct sadmc_zeno_replica_state {
  double max_energy, cutoff_energy, lowest_max_energy, translation_scale;
  uint64_t rejected_count, accepted_count, above_count, below_count, upwelling_count, unique_visitors;
  double above_total, below_total, above_total_squared, below_total_squared;
  double above_extra_total;
  uint64_t above_extra_count;
  int32_t collecting_data, _pad;
  uint64_t rng_s0, rng_s1;
  double energy;
} sadmc_zeno_replica_state;
int sadmc_replicas_create(const sadmc_config* cfg, double min_T, uint64_t independent_systems_before_new_bin, uint32_t max_replicas,
                          uint32_t max_init, sadmc_replicas** out);
void sadmc_replicas_destroy(sadmc_replicas* z);
/* n_rounds x MC::run_once: two launches per round (moves; swaps + median + split).  Blocking.  SADMC_ERR_WINDOW when a
 * simulation wanted to split off a replica with all max_replicas slots in use (it carries on without). */
int sadmc_replicas_run(sadmc_replicas* z, uint64_t n_rounds);
int sadmc_replicas_num_moves(sadmc_replicas* z, uint32_t sim, uint64_t* moves);            /* MC::moves */
int sadmc_replicas_num_replicas(sadmc_replicas* z, uint32_t sim, uint32_t* n);            /* replicas.len() */
int sadmc_replicas_get_replicas(sadmc_replicas* z, uint32_t sim, uint32_t cap, sadmc_zeno_replica_state* out);
int sadmc_replicas_get_rng(sadmc_replicas* z, uint32_t sim, uint64_t s[2]);                /* MC::rng */
int sadmc_replicas_get_median(sadmc_replicas* z, uint32_t sim, uint32_t cap, double* energies, uint32_t* len); /* MedianEstimator */
int sadmc_replicas_system_len(sadmc_replicas* z, size_t* n_doubles);
int sadmc_replicas_get_system(sadmc_replicas* z, uint32_t sim, uint32_t replica, double* buf, size_t n);
int sadmc_replicas_last_run_ms(sadmc_replicas* z, float* ms);

/* ---- measurement utility (bench.py) --------------------------------------- */
/* Achievable FP64 FMA throughput of `device` in TFLOP/s (independent DFMA chains,
 * best of `reps`): the denominator of the FP64 roofline fraction. */
int sadmc_measure_fp64_peak(int device, int reps, double* tflops);

/* Self-test of the exact exp-comparison filter used by the accept test (csrc/fastmath.cuh): evaluates n
 * random and adversarial (v, d) pairs on `device`; *mismatches counts decisions that differ from the plain
 * `v <=> sadmc_exp(d)` comparison or estimates outside the proven bound (must be 0), *exact_evaluations
 * how many pairs needed the full exp. */
int sadmc_selftest_exp_cmp(int device, uint64_t seed, uint64_t n, uint64_t* mismatches, uint64_t* exact_evaluations);

#ifdef __cplusplus
}
#endif
#endif /* SADMC_GPU_H */
