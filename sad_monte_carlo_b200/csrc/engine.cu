// engine.cu -- the C ABI of include/sadmc_gpu.h: device memory, launches, state I/O.
//
// Replaces, for many walkers at once, what `EnergyMC::from_params`
// (src/mc/energy.rs:830-898) and the `loop { mc.move_once() }` of
// src/bin/histogram.rs:6-11 do for one.  No CPU fallback: every entry point that
// computes launches a kernel of this library.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sadmc_gpu.h"
#include "book.cuh"
#include "fold_kernels.cuh"
#include "host_ctor.hpp"
#include "kernel_set.cuh"
#include "rng.cuh"
#include "sys_limits.hpp"
#include "replicas_round.cuh"
#include "tempering_swap.cuh"

using namespace sadmc;

static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t _e = (call);                                                                             \
    if (_e != cudaSuccess) return fail(SADMC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

struct sadmc_engine {
  sadmc_config cfg;
  DevParams P;
  KernelSet ks;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  unsigned long long moves = 0;
  unsigned long long launches = 0;
  long long k_base = 0;
  size_t sys_len = 0; // doubles per walker in the ABI image
  bool started = false;
  bool host_only = false; // sadmc_reference_system: system parameters only, no bin window, no device
  bool has_extra = false; // the system reports data_to_collect values (two-wells `which`, WCA `pressure`)
  double* d_zig = nullptr;
  ShimOut* d_shim = nullptr;
  double* d_pending = nullptr;
  double* d_wmax = nullptr;
  void* d_fold = nullptr;
  double* d_fold_part = nullptr; // per-chunk partial sums of the two-stage fold
  size_t fold_part_bytes = 0;
  float last_ms = 0.f;
  FoldSel fold_sel = {0u, 1u, 0u, 0, 0ull, 0};
  unsigned int* h_halted = nullptr;     // pinned copy of P.halted, refreshed behind every launch
  unsigned int halted_seen[2] = {0, 0}; // what sadmc_sync has reported already
  std::vector<void*> allocs;
};

static int dev_alloc(sadmc_engine* e, void** p, size_t bytes, bool zero) {
  if (bytes == 0) bytes = 8;
  cudaError_t er = cudaMalloc(p, bytes);
  if (er != cudaSuccess) return fail(SADMC_ERR_CUDA, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(er));
  e->allocs.push_back(*p);
  if (zero) {
    er = cudaMemsetAsync(*p, 0, bytes, e->stream);
    if (er != cudaSuccess) return fail(SADMC_ERR_CUDA, "cudaMemset failed: %s", cudaGetErrorString(er));
  }
  return 0;
}

static int pick_kernels(sadmc_engine* e) {
  const sadmc_config& c = e->cfg;
  DevParams& P = e->P;
  switch (c.system) {
    case SADMC_SYS_ISING: e->ks = kernels_ising(P); return 0;
    case SADMC_SYS_FAKE: e->ks = kernels_fake(P); return 0;
    case SADMC_SYS_WCA: {
      // lanes_per_walker: 0 = auto (below), 4 / 8 / 16 = that many lanes share a walker (sys_wca_group.cuh),
      // 32 = one warp per walker (sys_cell_fluid.cuh, the kernel the square well uses)
      const bool fast = (c.flags & SADMC_FLAG_FAST_MATH) != 0;
      // auto: the fast tier runs 8 lanes per walker (0.99e9 moves/s at N = 256); with the reference's re-summation of the
      // whole energy every ~10 accepted moves that sum is the kernel, and a warp per walker does it fastest (1.3e8 vs 0.8e8)
      const int G = c.lanes_per_walker == 0 ? (fast ? 8 : 32) : c.lanes_per_walker;
      if (G == 32 && !fast) {
        e->ks = kernels_cell_fluid(false, P);
        return 0;
      }
      if (kernels_wca_group(G, fast, P, &e->ks)) return 0;
      return fail(SADMC_ERR_UNSUPPORTED, "wca: lanes_per_walker must be 0, 4, 8, 16 or 32 (32 without SADMC_FLAG_FAST_MATH), not %d", G);
    }
    case SADMC_SYS_SW: e->ks = kernels_cell_fluid(true, P); return 0;
    case SADMC_SYS_TWO_WELLS: e->ks = kernels_two_wells(P); return 0;
    case SADMC_SYS_FAKE_ERFINV: e->ks = kernels_erfinv(P); return 0;
    case SADMC_SYS_LJ: {
      int G = c.lanes_per_walker;
      const bool fast = (c.flags & SADMC_FLAG_FAST_MATH) != 0;
      // many walkers: one thread per walker, cluster in shared memory (the fastest measured); few: a warp or part of one
      if (G == 0 && (c.flags & SADMC_FLAG_BINNING)) G = 1; // the energy_binning.rs kernels exist for one thread per walker
      if (G == 0) G = c.n_walkers >= 16384 ? 1 : (c.n_walkers >= 4096 ? 8 : 32);
      if (G == 1 || (fast && (G == 2 || G == 4))) { // configuration in shared memory (sys_lj_thread.cuh)
        if (c.N > 64) return fail(SADMC_ERR_UNSUPPORTED, "lj: shared-memory kernels hold N <= 64 atoms (N=%u)", c.N);
        if (fast && G == 1 && (c.flags & SADMC_FLAG_HELPER_WARPS)) { // experiment: helper warps for the pair loop
          if (kernels_lj_thread_paired((int)c.N, P, &e->ks)) return 0;
          return fail(SADMC_ERR_UNSUPPORTED, "lj: helper-warp kernels exist for N = 31 and 38 only (N=%u)", c.N);
        }
        const bool ok = !fast ? kernels_lj_thread_exact((int)c.N, P, &e->ks)
                              : (G == 1 ? kernels_lj_thread_fast((int)c.N, G, P, &e->ks) : kernels_lj_thread_fast_multi((int)c.N, G, P, &e->ks));
        if (ok) return 0;
        return fail(SADMC_ERR_UNSUPPORTED, "lj: no shared-memory kernel instance for N=%u, lanes_per_walker=%d", c.N, G);
      }
      if (G == 2) return fail(SADMC_ERR_UNSUPPORTED, "lj: lanes_per_walker = 2 needs SADMC_FLAG_FAST_MATH");
      const int A = ((int)c.N + G - 1) / G;
      if (kernels_lj_warp(G, A, P, &e->ks) || kernels_lj_warp_small(G, A, P, &e->ks)) return 0;
      return fail(SADMC_ERR_UNSUPPORTED, "lj: no kernel instance for N=%u with lanes_per_walker=%d (atoms per lane %d)", c.N, G, A);
    }
    default: return fail(SADMC_ERR_UNSUPPORTED, "system kind %d has no kernel yet", c.system);
  }
}

static bool is_none(double x) { return std::isnan(x); }

static int setup_params(sadmc_engine* e) {
  const sadmc_config& c = e->cfg;
  DevParams& P = e->P;
  memset(&P, 0, sizeof P);
  if (c.abi_version != SADMC_ABI_VERSION) return fail(SADMC_ERR_INVALID, "abi_version %u != %d", c.abi_version, SADMC_ABI_VERSION);
  if (c.n_walkers == 0) return fail(SADMC_ERR_INVALID, "n_walkers must be > 0");
  if (c.method < SADMC_METHOD_SAD || c.method > SADMC_METHOD_CANONICAL) return fail(SADMC_ERR_INVALID, "unknown method %d", c.method);
  P.n_walkers = c.n_walkers;
  P.N = c.N;
  P.flags = c.flags;
  P.has_min = !is_none(c.min_allowed_energy);
  P.has_max = !is_none(c.max_allowed_energy);
  P.min_allowed = c.min_allowed_energy;
  P.max_allowed = c.max_allowed_energy;
  P.min_T = c.sad_min_T;
  P.inv_t = c.method == SADMC_METHOD_INV_T_WL;
  P.method_kind = c.method;
  P.has_min_gamma = c.method == SADMC_METHOD_WL && !is_none(c.wl_min_gamma);
  P.min_gamma = c.wl_min_gamma;
  P.canonical_T = c.canonical_T;
  P.move_plan = c.move_plan;
  P.move_value = c.move_value;
  if (c.method == SADMC_METHOD_SAD && !(c.sad_min_T > 0)) return fail(SADMC_ERR_INVALID, "sad_min_T must be > 0");

  double native_de = NAN, lowest = NAN, greatest = NAN;
  switch (c.system) {
    case SADMC_SYS_ISING:
      if (!(c.N > 1)) return fail(SADMC_ERR_INVALID, "ising N must be > 1 (ising.rs:39)");
      if (c.N > 256) return fail(SADMC_ERR_UNSUPPORTED, "ising N > 256 does not fit the shared-memory lattice");
      native_de = 4.0; // ising.rs:76-78
      lowest = -2.0 * c.N * c.N;
      greatest = 2.0 * c.N * c.N;
      P.ising_words = (c.N * c.N + 31) / 32;
      P.zone_a = zone_single(c.N);
      e->sys_len = (size_t)c.N * c.N + 1;
      break;
    case SADMC_SYS_LJ:
      if (c.N < 1) return fail(SADMC_ERR_INVALID, "lj N must be >= 1");
      if (!(c.lj_radius > 0)) return fail(SADMC_ERR_INVALID, "lj radius must be > 0");
      P.lj_R = c.lj_radius;
      P.lj_R2 = c.lj_radius * c.lj_radius;
      P.zone_b = zone_uniform(c.N);
      lowest = std::fmax(-0.5 * c.N * (c.N - 1.0), -8.7 * c.N); // lj.rs:245-248, tightened by the bulk fcc cohesive energy
      e->sys_len = 3 * (size_t)c.N + 2;
      P.sys_stride = (uint32_t)e->sys_len;
      break;
    case SADMC_SYS_WCA:
    case SADMC_SYS_SW: {
      const bool sw = c.system == SADMC_SYS_SW;
      if (c.N < 1 || c.N > 4096) return fail(SADMC_ERR_UNSUPPORTED, "cell fluids hold 1..4096 atoms per walker (N=%u)", c.N);
      double box[3];
      if (c.cell_width[0] > 0) { // CellDimensions::CellWidth, optcell.rs:46
        for (int k = 0; k < 3; k++) box[k] = std::fabs(c.cell_width[k]);
      } else {
        // ReducedDensity -> CellVolume(N / rho) (wca.rs:399-401); FillingFraction -> CellVolume(N pi/6 / eta) (optsquare.rs:365-367)
        const double vol = sw ? (double)c.N * (M_PI * 1.0 * 1.0 * 1.0 / 6.0) / c.filling_fraction : (double)c.N / c.reduced_density;
        if (!(vol > 0)) return fail(SADMC_ERR_INVALID, "cell volume must be positive");
        box[0] = box[1] = box[2] = std::cbrt(vol); // optcell.rs:47-50
      }
      const double r_cut = sw ? c.sw_well_width * 1.0 : std::pow(2.0, 1.0 / 6.0); // optsquare.rs:155, wca.rs:61-63
      for (int k = 0; k < 3; k++) {
        if (r_cut > box[k]) return fail(SADMC_ERR_INVALID, "The cell is not large enough for the well width, sorry! (wca.rs:186-191)");
        P.box[k] = box[k];
        P.ncell[k] = (int)std::floor(box[k] / r_cut); // optcell.rs:64-66
        if (P.ncell[k] < 3)
          return fail(SADMC_ERR_UNSUPPORTED, "box of %.3f holds %d subcells along axis %d; the device cell list needs >= 3", box[k], P.ncell[k], k);
      }
      if ((long long)P.ncell[0] * P.ncell[1] * P.ncell[2] > 32000) return fail(SADMC_ERR_UNSUPPORTED, "too many subcells for the 16-bit cell list");
      P.r_cut2 = r_cut * r_cut;
      P.well2 = r_cut * r_cut;
      P.zone_b = zone_uniform(c.N);
      if (sw) {
        native_de = 1.0; // optsquare.rs:190-192
        lowest = -(double)c.N * (double)hostctor::max_balls_within(r_cut); // optsquare.rs:196-198
        greatest = 0.0;
      } else {
        lowest = 0.0; // wca.rs:234-236
        e->has_extra = true;
      }
      e->sys_len = 3 * (size_t)c.N + 2;
      P.sys_stride = (uint32_t)e->sys_len;
      break;
    }
    case SADMC_SYS_FAKE: {
      P.fake_fn = c.fake_function;
      P.fake_dim = c.fake_function == SADMC_FAKE_LINEAR ? 1 : (c.fake_function == SADMC_FAKE_QUADRATIC ? (int)c.N : 3); // fake.rs:39-46
      if (P.fake_dim < 1 || P.fake_dim > FAKE_MAX_DIM) return fail(SADMC_ERR_UNSUPPORTED, "fake: dimensions must be 1..%d", FAKE_MAX_DIM);
      P.fake_a = c.fake_a;
      P.fake_b = c.fake_b;
      P.fake_e1 = c.fake_e1;
      P.fake_e2 = c.fake_e2;
      P.fake_sigma = c.fake_sigma;
      P.zone_a = zone_single(P.fake_dim);
      if (c.fake_function == SADMC_FAKE_LINEAR || c.fake_function == SADMC_FAKE_QUADRATIC) {
        lowest = 0.0;
        greatest = 1.0;
      } else if (c.fake_function == SADMC_FAKE_GAUSSIAN) {
        lowest = -1.0;
        greatest = 0.0;
      } else {
        const double f1 = ((1.0 - c.fake_b) / (c.fake_b - c.fake_a)) * ((1.0 - c.fake_b) / (c.fake_b - c.fake_a)) * c.fake_e2 - c.fake_e2;
        lowest = std::fmin(-c.fake_e1, -c.fake_e2);
        greatest = std::fmax(0.0, f1);
      }
      e->sys_len = P.fake_dim;
      P.sys_stride = (uint32_t)e->sys_len;
      break;
    }
    case SADMC_SYS_TWO_WELLS:
      if (c.N % 3 != 0 || c.N == 0) return fail(SADMC_ERR_INVALID, "The number of dimensions %u is not divisible by three! (two_wells.rs:240-245)", c.N);
      if (c.N > TW_MAX_DIM) return fail(SADMC_ERR_UNSUPPORTED, "two-wells: N <= %d", TW_MAX_DIM);
      P.tw_h2h1 = c.tw_h2_to_h1;
      P.tw_r2 = c.tw_r2;
      P.tw_rw = std::sqrt(c.tw_barrier_over_h1) * 1.0 + c.tw_r2 * std::sqrt(1.0 + c.tw_barrier_over_h1 - 1.0 / c.tw_h2_to_h1); // two_wells.rs:246-247
      P.zone_a = zone_single(c.N / 3);
      lowest = std::fmin(-1.0, -c.tw_h2_to_h1);
      greatest = 0.0;
      e->sys_len = (size_t)c.N + 1;
      P.sys_stride = (uint32_t)e->sys_len;
      e->has_extra = true;
      break;
    case SADMC_SYS_FAKE_ERFINV:
      if (c.N < 1 || c.N > ERFINV_MAX_DIM) return fail(SADMC_ERR_UNSUPPORTED, "erfinv: N must be 1..%d", ERFINV_MAX_DIM);
      P.erfinv_mean = c.erfinv_mean_energy;
      P.zone_a = zone_single(c.N);
      e->sys_len = c.N;
      P.sys_stride = (uint32_t)e->sys_len;
      break;
    default: return fail(SADMC_ERR_UNSUPPORTED, "system kind %d has no kernel yet", c.system);
  }
  P.width = !is_none(c.energy_bin) ? c.energy_bin : (!is_none(native_de) ? native_de : 1.0); // energy.rs:831-833
  if (!(P.width > 0)) return fail(SADMC_ERR_INVALID, "energy_bin must be > 0 (energy.rs:402)");

  if (e->host_only) return 0;
  double wlo = c.bin_window_lo, whi = c.bin_window_hi;
  if (is_none(wlo)) wlo = P.has_min ? c.min_allowed_energy - 2 * P.width : lowest;
  if (is_none(whi)) whi = P.has_max ? c.max_allowed_energy + 2 * P.width : greatest;
  if (is_none(wlo) || is_none(whi))
    return fail(SADMC_ERR_INVALID, "cannot derive the bin window: give bin_window_lo/hi or min/max_allowed_energy");
  if (!(whi > wlo)) return fail(SADMC_ERR_INVALID, "empty bin window [%g, %g)", wlo, whi);
  const bool binning = (c.flags & SADMC_FLAG_BINNING) != 0;
  if (binning && c.method == SADMC_METHOD_CANONICAL) return fail(SADMC_ERR_INVALID, "energy_binning.rs has no canonical method (MethodParams, energy_binning.rs:22-40)");
  // energy.rs centres bins on multiples of the width (energy.rs:852), histogram.rs puts their edges there (histogram.rs:149-151)
  e->k_base = binning ? (long long)std::floor(wlo / P.width) - 1 : (long long)std::floor(wlo / P.width + 0.5) - 1;
  const long long k_top = (long long)std::ceil(whi / P.width + 0.5) + 1;
  const long long cap = k_top - e->k_base + 1;
  if (cap > (1ll << 28)) return fail(SADMC_ERR_INVALID, "bin window needs %lld bins per walker", cap);
  P.cap = (uint32_t)cap;
  P.hr_count = nullptr;
  P.hr_cap = 0;
  P.hr_width = 1.0;
  P.hr_kbase = 0;
  if (!is_none(c.high_resolution_de) && c.high_resolution_de > 0) { // energy_binning.rs:62-63
    if (!binning) return fail(SADMC_ERR_INVALID, "high_resolution_de belongs to the `binning` Monte Carlo: set SADMC_FLAG_BINNING");
    P.hr_width = c.high_resolution_de;
    P.hr_kbase = (long long)std::floor(wlo / P.hr_width) - 1;
    const long long hcap = (long long)std::ceil(whi / P.hr_width) + 2 - P.hr_kbase;
    if (hcap > (1ll << 28)) return fail(SADMC_ERR_INVALID, "the high-resolution histogram needs %lld bins per walker", hcap);
    P.hr_cap = (uint32_t)hcap;
  }
  return 0;
}

static int refuse_binning(const sadmc_engine* e, const char* what) {
  if (e && (e->cfg.flags & SADMC_FLAG_BINNING)) return fail(SADMC_ERR_INVALID, "%s works on the energy.rs bin layout; this engine runs SADMC_FLAG_BINNING", what);
  return 0;
}

static int launch_cfg(const sadmc_engine* e, int* grid) {
  const long long threads = (long long)e->cfg.n_walkers * e->ks.G;
  *grid = (int)((threads + e->ks.block - 1) / e->ks.block);
  return 0;
}

// `Any::from(AnyParams)` (any.rs:80-93) on the host: one system image in sadmc_get_system's layout.
static int reference_image(sadmc_engine* e, std::vector<double>& img) {
  const sadmc_config& c = e->cfg;
  switch (c.system) {
    case SADMC_SYS_ISING: img = hostctor::ising_image(c.N); break;
    case SADMC_SYS_LJ: img = hostctor::lj_image(c.N, c.lj_radius); break;
    case SADMC_SYS_FAKE: img.assign(e->sys_len, 0.0); break;       // fake.rs:85-93: the origin
    case SADMC_SYS_SW: {
      std::string why;
      img = hostctor::sw_image(c.N, e->P.box, why);
      if (img.empty()) return fail(SADMC_ERR_INVALID, "sw: %s", why.c_str());
      break;
    }
    case SADMC_SYS_WCA: // From<WcaNParams>, fcc = false (wca.rs:448-496); the fcc start needs rand's choose_multiple
      img = hostctor::wca_image(c.N, e->P.box, 0, getenv("SADMC_HOST_THREADS") ? (unsigned)atoi(getenv("SADMC_HOST_THREADS")) : 0u);
      if (img.empty()) return fail(SADMC_ERR_INVALID, "wca: no random placement of %u atoms below 1e80 epsilon", c.N);
      break;
    case SADMC_SYS_FAKE_ERFINV: img.assign(e->sys_len, 0.5); break; // erfinv.rs:50-58
    case SADMC_SYS_TWO_WELLS: {                                   // two_wells.rs:248-250
      img.assign(e->sys_len, 0.0);
      img[0] = -0.99;
      double d2 = 0.0;
      for (uint32_t k = 0; k < c.N; k++) d2 += img[k] * img[k];
      img[c.N] = d2;
      break;
    }
    default: return fail(SADMC_ERR_UNSUPPORTED, "no reference constructor for system %d yet", c.system);
  }
  return 0;
}

static int upload_initial_systems(sadmc_engine* e) {
  const sadmc_config& c = e->cfg;
  if (c.init_mode != SADMC_INIT_REFERENCE) return 0;
  std::vector<double> img;
  const int rc = reference_image(e, img);
  if (rc) return rc;
  std::vector<double> all((size_t)c.n_walkers * e->sys_len);
  for (uint32_t w = 0; w < c.n_walkers; w++) memcpy(&all[(size_t)w * e->sys_len], img.data(), e->sys_len * sizeof(double));
  return sadmc_set_systems(e, all.data(), all.size());
}

// ---- measurement utility: the chip's FP64 FMA peak, for the roofline denominator ----
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      x0 = fma(x0, a, b);
      x1 = fma(x1, a, b);
      x2 = fma(x2, a, b);
      x3 = fma(x3, a, b);
      x4 = fma(x4, a, b);
      x5 = fma(x5, a, b);
      x6 = fma(x6, a, b);
      x7 = fma(x7, a, b);
    }
  }
  const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 12345.678) out[0] = s;
}

// ---- self-test of exp_cmp (fastmath.cuh): its decision must equal the plain comparison with sadmc_exp ----
__global__ void __launch_bounds__(256) exp_cmp_test_kernel(unsigned long long seed, unsigned long long n, unsigned long long* mismatches,
                                                         unsigned long long* filtered) {
  const unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  Rng r;
  seed_from_u64(seed + tid, (uint64_t*)&r.s0, (uint64_t*)&r.s1);
  unsigned long long bad = 0, slow = 0;
  for (unsigned long long k = tid; k < n; k += (unsigned long long)gridDim.x * blockDim.x) {
    // d spread over [-90, 0.5]; v uniform, or within a few ulp / a few 1e-5 of e^d (the adversarial cases)
    const double d = (k % 7 == 0) ? -r.gen_f64() * 1e-3 : (0.5 - 90.5 * r.gen_f64());
    const double ex = sadmc_exp(d);
    double v;
    switch (k % 5) {
      case 0: v = r.gen_f64(); break;
      case 1: v = ex; break;
      case 2: v = sadmc_bits_f64(sadmc_f64_bits(ex) + (r.next() % 5) - 2); break;
      case 3: v = ex * (1.0 + (r.gen_f64() - 0.5) * 4e-4); break;
      default: v = ex * r.gen_f64() * 2.0; break;
    }
    const int want = v > ex ? 1 : (v < ex ? -1 : 0);
    if (exp_cmp(v, d) != want) bad++;
    if (d <= 0.0 && d > -80.0) {
      const float t = __double2float_rn(d) * 1.4426950408889634f;
      float a;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(t));
      if (!(v > (double)a * 1.0001) && !(v < (double)a * 0.9999)) slow++;
      // the bound itself: the float estimate must sit within 1e-4 of sadmc_exp
      if (fabs((double)a - ex) > 0.99e-4 * ex) bad++;
    }
  }
  atomicAdd(mismatches, bad);
  atomicAdd(filtered, slow);
}

extern "C" {

int sadmc_selftest_exp_cmp(int device, uint64_t seed, uint64_t n, uint64_t* mismatches, uint64_t* exact_evaluations) {
  if (!mismatches) return fail(SADMC_ERR_INVALID, "null argument");
  CK(cudaSetDevice(device));
  unsigned long long* d = nullptr;
  CK(cudaMalloc(&d, 16));
  CK(cudaMemset(d, 0, 16));
  exp_cmp_test_kernel<<<1024, 256>>>(seed, n, d, d + 1);
  CK(cudaGetLastError());
  unsigned long long h[2];
  CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
  cudaFree(d);
  *mismatches = h[0];
  if (exact_evaluations) *exact_evaluations = h[1];
  return 0;
}

// TFLOP/s of dependent-chain-free DFMA (2 flops each) on `device`; best of `reps`.
int sadmc_measure_fp64_peak(int device, int reps, double* tflops) {
  if (!tflops) return fail(SADMC_ERR_INVALID, "null argument");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  double* d = nullptr;
  CK(cudaMalloc(&d, 8));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  double best = 0.0;
  for (int r = 0; r < reps + 1; r++) {
    CK(cudaEventRecord(a));
    fp64_peak_kernel<<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    const double fl = 2.0 * 64.0 * (double)iters * blocks * threads;
    if (r > 0 && fl / (ms * 1e-3) / 1e12 > best) best = fl / (ms * 1e-3) / 1e12;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d);
  *tflops = best;
  return 0;
}

const char* sadmc_last_error(void) { return g_err.c_str(); }
int sadmc_abi_version(void) { return SADMC_ABI_VERSION; }
size_t sadmc_sizeof_config(void) { return sizeof(sadmc_config); }
size_t sadmc_sizeof_walker_state(void) { return sizeof(sadmc_walker_state); }
size_t sadmc_sizeof_binning_state(void) { return sizeof(sadmc_binning_state); }
size_t sadmc_sizeof_replica_state(void) { return sizeof(sadmc_replica_state); }
size_t sadmc_sizeof_zeno_replica_state(void) { return sizeof(sadmc_zeno_replica_state); }

int sadmc_reference_system(const sadmc_config* cfg, double* buf, size_t n, size_t* needed) {
  if (!cfg) return fail(SADMC_ERR_INVALID, "null argument");
  sadmc_engine e; // host fields only: no device, no stream
  e.cfg = *cfg;
  e.host_only = true;
  int rc = setup_params(&e);
  if (rc) return rc;
  if (needed) *needed = e.sys_len;
  if (!buf) return 0;
  if (n < e.sys_len) return fail(SADMC_ERR_INVALID, "system image needs %zu doubles, buffer holds %zu", e.sys_len, n);
  std::vector<double> img;
  try {
    rc = reference_image(&e, img);
  } catch (const std::exception& ex) {
    return fail(SADMC_ERR_INVALID, "%s", ex.what());
  }
  if (rc) return rc;
  memcpy(buf, img.data(), e.sys_len * sizeof(double));
  return 0;
}

void sadmc_destroy(sadmc_engine* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  for (void* p : e->allocs) cudaFree(p);
  if (e->h_halted) cudaFreeHost(e->h_halted);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

int sadmc_create(const sadmc_config* cfg, sadmc_engine** out) {
  if (!cfg || !out) return fail(SADMC_ERR_INVALID, "null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(SADMC_ERR_CUDA, "no CUDA device: the walker engine has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(SADMC_ERR_INVALID, "device %d out of range (%d devices)", cfg->device, ndev);
  sadmc_engine* e = new sadmc_engine;
  e->cfg = *cfg;
  int rc = setup_params(e);
  if (rc) {
    delete e;
    return rc;
  }
  rc = pick_kernels(e);
  if (!rc && (e->cfg.flags & SADMC_FLAG_BINNING_LINEAR) && !(e->cfg.flags & SADMC_FLAG_BINNING))
    rc = fail(SADMC_ERR_INVALID, "SADMC_FLAG_BINNING_LINEAR selects the bins of the `binning` Monte Carlo: set SADMC_FLAG_BINNING as well");
  if (!rc && (e->cfg.flags & SADMC_FLAG_BINNING_LINEAR) && !e->ks.move_linear[e->cfg.method])
    rc = fail(SADMC_ERR_UNSUPPORTED, "SADMC_FLAG_BINNING_LINEAR: binning::linear is built for the one-thread-per-walker systems only (method %d)", e->cfg.method);
  if (!rc && (e->cfg.flags & SADMC_FLAG_BINNING) && !e->ks.move_binning[e->cfg.method])
    rc = fail(SADMC_ERR_UNSUPPORTED, "SADMC_FLAG_BINNING: no energy_binning.rs kernel for this system / lanes_per_walker / method %d", e->cfg.method);
  if (rc) {
    delete e;
    return rc;
  }
#define BAIL(expr)          \
  do {                      \
    int _rc = (expr);       \
    if (_rc) {              \
      sadmc_destroy(e);     \
      return _rc;           \
    }                       \
  } while (0)
#define CKB(call)                                                                                   \
  do {                                                                                              \
    cudaError_t _e = (call);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      fail(SADMC_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(_e));                         \
      sadmc_destroy(e);                                                                             \
      return SADMC_ERR_CUDA;                                                                        \
    }                                                                                               \
  } while (0)
  CKB(cudaSetDevice(cfg->device));
  CKB(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  e->own_stream = true;
  CKB(cudaEventCreate(&e->ev0));
  CKB(cudaEventCreate(&e->ev1));
  DevParams& P = e->P;
  const size_t nb = (size_t)P.n_walkers * P.cap;
  size_t need = (size_t)P.n_walkers * P.hr_cap * 8 + nb * (sizeof(BinRec) + (e->has_extra ? 16 : 0)) + (size_t)P.n_walkers * (sizeof(WalkerRec) + e->sys_len * 8 + P.ising_words * 4);
  size_t free_b = 0, total_b = 0;
  CKB(cudaMemGetInfo(&free_b, &total_b));
  if (need > free_b) {
    fail(SADMC_ERR_INVALID, "%u walkers x %u bins need %.1f GB of HBM, %.1f GB free: shrink the bin window or the walker count",
         P.n_walkers, P.cap, need / 1e9, free_b / 1e9);
    sadmc_destroy(e);
    return SADMC_ERR_INVALID;
  }
  BAIL(dev_alloc(e, (void**)&P.rec, nb * sizeof(BinRec), true));
  if (e->has_extra) {
    BAIL(dev_alloc(e, (void**)&P.extra_total, nb * 8, true));
    BAIL(dev_alloc(e, (void**)&P.extra_count, nb * 8, true));
  }
  if (P.hr_cap) BAIL(dev_alloc(e, (void**)&P.hr_count, (size_t)P.n_walkers * P.hr_cap * 8, true));
  BAIL(dev_alloc(e, (void**)&P.walkers, (size_t)P.n_walkers * sizeof(WalkerRec), true));
  BAIL(dev_alloc(e, (void**)&P.sys, (size_t)P.n_walkers * (P.sys_stride ? P.sys_stride : 1) * 8, true));
  if (e->ks.zstream_per_thread) { // scratch of the move kernels that stream part of the configuration from L2: laid out by thread
    const size_t blk = e->ks.move_block ? e->ks.move_block : e->ks.block;
    const size_t tpw = e->ks.move_block ? e->ks.move_threads_per_walker : e->ks.G;
    const size_t threads = ((size_t)P.n_walkers * tpw + blk - 1) / blk * blk;
    BAIL(dev_alloc(e, (void**)&P.zstream, threads * e->ks.zstream_per_thread * 8, true));
  }
  BAIL(dev_alloc(e, (void**)&P.sys_words, (size_t)P.n_walkers * (P.ising_words ? P.ising_words : 1) * 4, true));
  BAIL(dev_alloc(e, (void**)&e->d_zig, 2 * SADMC_ZIG_TABLE_LEN * 8, false));
  BAIL(dev_alloc(e, (void**)&e->d_shim, sizeof(ShimOut), true));
  BAIL(dev_alloc(e, (void**)&e->d_pending, 8 * 8, true));
  BAIL(dev_alloc(e, (void**)&P.halted, 2 * sizeof(unsigned int), true));
  CKB(cudaMallocHost((void**)&e->h_halted, 2 * sizeof(unsigned int)));
  e->h_halted[0] = e->h_halted[1] = 0;
  {
    double z[2 * SADMC_ZIG_TABLE_LEN];
    memcpy(z, hostctor::H_ZX, sizeof hostctor::H_ZX);
    memcpy(z + SADMC_ZIG_TABLE_LEN, hostctor::H_ZF, sizeof hostctor::H_ZF);
    CKB(cudaMemcpyAsync(e->d_zig, z, sizeof z, cudaMemcpyHostToDevice, e->stream));
    CKB(cudaStreamSynchronize(e->stream));
  }
  P.zig = e->d_zig;
  if (e->ks.smem > 48 * 1024 || e->ks.move_smem > 48 * 1024) {
    for (int m = 1; m <= 5; m++)
      if (e->ks.move[m])
        CKB(cudaFuncSetAttribute((const void*)e->ks.move[m], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(e->ks.move_smem ? e->ks.move_smem : e->ks.smem)));
    for (int m = 1; m <= 5; m++)
      if (e->ks.move_binning[m])
        CKB(cudaFuncSetAttribute((const void*)e->ks.move_binning[m], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->ks.smem));
    for (int m = 1; m <= 5; m++)
      if (e->ks.move_linear[m])
        CKB(cudaFuncSetAttribute((const void*)e->ks.move_linear[m], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->ks.smem));
    CKB(cudaFuncSetAttribute((const void*)e->ks.init, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->ks.smem));
    CKB(cudaFuncSetAttribute((const void*)e->ks.shim, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->ks.smem));
  }
  BAIL(upload_initial_systems(e));
  if (cfg->init_mode != SADMC_INIT_EXTERNAL) BAIL(sadmc_start(e));
  *out = e;
  return 0;
#undef BAIL
#undef CKB
}

int sadmc_start(sadmc_engine* e) {
  if (!e) return fail(SADMC_ERR_INVALID, "null engine");
  if (e->started) return fail(SADMC_ERR_INVALID, "engine already started");
  CK(cudaSetDevice(e->cfg.device));
  int grid;
  launch_cfg(e, &grid);
  e->ks.init<<<grid, e->ks.block, e->ks.smem, e->stream>>>(e->P, e->cfg.seed + e->cfg.walker_offset, e->cfg.init_mode, e->k_base,
                                                         e->cfg.method, e->cfg.samc_t0, 100000000ull);
  e->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(e->h_halted, e->P.halted, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  e->started = true;
  e->moves = 0;
  return 0;
}

int sadmc_set_stream(sadmc_engine* e, void* s) {
  if (!e) return fail(SADMC_ERR_INVALID, "null engine");
  CK(cudaStreamSynchronize(e->stream));
  if (e->own_stream) cudaStreamDestroy(e->stream);
  e->stream = (cudaStream_t)s;
  e->own_stream = false;
  return 0;
}
void* sadmc_get_stream(sadmc_engine* e) { return e ? (void*)e->stream : nullptr; }

int sadmc_run_async(sadmc_engine* e, uint64_t n_moves) {
  if (!e) return fail(SADMC_ERR_INVALID, "null engine");
  if (!e->started) return fail(SADMC_ERR_INVALID, "engine not started (call sadmc_start)");
  if (n_moves == 0) return 0;
  CK(cudaSetDevice(e->cfg.device));
  int grid;
  launch_cfg(e, &grid);
  int block = e->ks.block;
  size_t smem = e->ks.smem;
  if (e->ks.move_block) { // move kernels with their own launch shape (helper warps)
    block = e->ks.move_block;
    smem = e->ks.move_smem;
    const long long threads = (long long)e->cfg.n_walkers * e->ks.move_threads_per_walker;
    grid = (int)((threads + block - 1) / block);
  }
  move_fn f = (e->cfg.flags & SADMC_FLAG_BINNING_LINEAR) ? e->ks.move_linear[e->cfg.method]
              : (e->cfg.flags & SADMC_FLAG_BINNING)      ? e->ks.move_binning[e->cfg.method]
                                                         : e->ks.move[e->cfg.method];
  if (!f) return fail(SADMC_ERR_UNSUPPORTED, "no move kernel for method %d with these flags", e->cfg.method);
  CK(cudaEventRecord(e->ev0, e->stream));
  f<<<grid, block, smem, e->stream>>>(e->P, e->moves, n_moves);
  CK(cudaGetLastError());
  CK(cudaEventRecord(e->ev1, e->stream));
  CK(cudaMemcpyAsync(e->h_halted, e->P.halted, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, e->stream));
  e->launches++;
  e->moves += n_moves;
  return 0;
}
// A walker that leaves its bin window (or fails verify_energy, where the reference panics) freezes with its status
// set; the run reports it at the next synchronisation instead of carrying on silently.  Each halted walker is
// reported once; sadmc_num_halted gives the totals at any time.
int sadmc_sync(sadmc_engine* e) {
  if (!e) return fail(SADMC_ERR_INVALID, "null engine");
  CK(cudaStreamSynchronize(e->stream));
  if (e->h_halted) {
    const unsigned int win = e->h_halted[0], ver = e->h_halted[1];
    const unsigned int new_win = win - e->halted_seen[0], new_ver = ver - e->halted_seen[1];
    e->halted_seen[0] = win;
    e->halted_seen[1] = ver;
    if (new_ver) return fail(SADMC_ERR_VERIFY, "verify_energy failed for %u walker(s) (%u in total); they are halted (sadmc_get_walker(...).status)", new_ver, ver);
    if (new_win)
      return fail(SADMC_ERR_WINDOW, "%u walker(s) left the device bin window [bin_window_lo, bin_window_hi) and are halted (%u in total): widen the window",
                  new_win, win);
  }
  return 0;
}
int sadmc_num_halted(sadmc_engine* e, uint64_t* left_window, uint64_t* failed_verify) {
  if (!e) return fail(SADMC_ERR_INVALID, "null engine");
  CK(cudaStreamSynchronize(e->stream));
  if (left_window) *left_window = e->h_halted ? e->h_halted[0] : 0;
  if (failed_verify) *failed_verify = e->h_halted ? e->h_halted[1] : 0;
  return 0;
}
int sadmc_run(sadmc_engine* e, uint64_t n_moves) {
  int rc = sadmc_run_async(e, n_moves);
  if (rc) return rc;
  return sadmc_sync(e);
}
int sadmc_last_run_ms(sadmc_engine* e, float* ms) {
  if (!e || !ms) return fail(SADMC_ERR_INVALID, "null argument");
  CK(cudaEventSynchronize(e->ev1));
  CK(cudaEventElapsedTime(ms, e->ev0, e->ev1));
  return 0;
}
int sadmc_move_launch_shape(sadmc_engine* e, uint32_t* block, uint32_t* threads_per_walker, uint64_t* shared_bytes, uint32_t* stream_bytes_per_walker) {
  if (!e) return fail(SADMC_ERR_INVALID, "null engine");
  const bool own = e->ks.move_block != 0; // move kernels with their own launch shape
  if (block) *block = (uint32_t)(own ? e->ks.move_block : e->ks.block);
  if (threads_per_walker) *threads_per_walker = (uint32_t)(own ? e->ks.move_threads_per_walker : e->ks.G);
  if (shared_bytes) *shared_bytes = own ? e->ks.move_smem : e->ks.smem;
  if (stream_bytes_per_walker) *stream_bytes_per_walker = (uint32_t)e->ks.zstream_per_thread * 8u * (uint32_t)(own ? e->ks.move_threads_per_walker : e->ks.G);
  return 0;
}
int sadmc_launch_count(sadmc_engine* e, uint64_t* n) {
  if (!e || !n) return fail(SADMC_ERR_INVALID, "null argument");
  *n = e->launches;
  return 0;
}
int sadmc_num_moves(sadmc_engine* e, uint64_t* moves) {
  if (!e || !moves) return fail(SADMC_ERR_INVALID, "null argument");
  *moves = e->moves;
  return 0;
}

static int fetch_walkers(sadmc_engine* e, std::vector<WalkerRec>& v) {
  v.resize(e->P.n_walkers);
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaMemcpyAsync(v.data(), e->P.walkers, v.size() * sizeof(WalkerRec), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}
static int fetch_walker(sadmc_engine* e, uint32_t w, WalkerRec* r) {
  if (w >= e->P.n_walkers) return fail(SADMC_ERR_INVALID, "walker %u out of range", w);
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaMemcpyAsync(r, e->P.walkers + w, sizeof(WalkerRec), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}

int sadmc_num_accepted_moves(sadmc_engine* e, uint64_t* accepted_sum) {
  if (!e || !accepted_sum) return fail(SADMC_ERR_INVALID, "null argument");
  std::vector<WalkerRec> v;
  int rc = fetch_walkers(e, v);
  if (rc) return rc;
  uint64_t s = 0;
  for (auto& r : v) s += r.accepted;
  *accepted_sum = s;
  return 0;
}

int sadmc_accepted_moves_range(sadmc_engine* e, uint64_t* min_accepted, uint64_t* max_accepted) {
  if (!e) return fail(SADMC_ERR_INVALID, "null engine");
  std::vector<WalkerRec> v;
  int rc = fetch_walkers(e, v);
  if (rc) return rc;
  uint64_t lo = ~0ull, hi = 0;
  for (auto& r : v) {
    if (r.accepted < lo) lo = r.accepted;
    if (r.accepted > hi) hi = r.accepted;
  }
  if (min_accepted) *min_accepted = lo;
  if (max_accepted) *max_accepted = hi;
  return 0;
}

int sadmc_get_walker(sadmc_engine* e, uint32_t w, sadmc_walker_state* s) {
  if (!e || !s) return fail(SADMC_ERR_INVALID, "null argument");
  WalkerRec r;
  int rc = fetch_walker(e, w, &r);
  if (rc) return rc;
  memset(s, 0, sizeof *s);
  s->moves = e->moves;
  s->accepted_moves = r.accepted;
  s->acceptance_rate = r.acc_rate;
  s->translation_scale = r.tscale;
  s->rng_s0 = r.s0;
  s->rng_s1 = r.s1;
  s->energy = r.E;
  s->bins_min = r.bmin;
  s->bins_width = e->P.width;
  s->bins_len = (uint32_t)r.len;
  s->window_first = (uint32_t)r.lo;
  s->method = r.method == SADMC_METHOD_WL && e->P.inv_t ? SADMC_METHOD_INV_T_WL : r.method;
  s->status = r.status;
  s->too_lo = r.too_lo;
  s->too_hi = r.too_hi;
  s->latest_parameter = r.latest_parameter;
  s->tL = r.tL;
  s->tF = r.tF;
  s->num_states = r.num_states;
  s->highest_hist = r.highest_hist;
  s->samc_t0 = r.samc_t0;
  s->wl_gamma = r.wl_gamma;
  s->wl_num_states = r.wl_num_states;
  s->wl_min_energy = r.wl_min_energy;
  s->wl_lowest_hist = r.wl_lowest;
  s->wl_highest_hist = r.wl_highest;
  s->wl_total_hist = r.wl_total;
  s->wl_hist_len = (uint32_t)r.wl_hist_len;
  s->wl_inv_t = e->P.inv_t;
  s->max_S = r.max_S;
  s->max_S_index = (uint32_t)r.max_S_index;
  return 0;
}

int sadmc_get_energies(sadmc_engine* e, double* energies) {
  if (!e || !energies) return fail(SADMC_ERR_INVALID, "null argument");
  std::vector<WalkerRec> v;
  int rc = fetch_walkers(e, v);
  if (rc) return rc;
  for (size_t i = 0; i < v.size(); i++) energies[i] = v[i].E;
  return 0;
}

int sadmc_get_bins(sadmc_engine* e, uint32_t w, uint32_t cap, uint64_t* histogram, uint64_t* t_found, double* lnw, double* energy_total,
                   double* energy_squared_total, uint64_t* round_trips, uint8_t* have_visited, uint64_t* wl_hist, double* extra_total,
                   uint64_t* extra_count) {
  if (!e) return fail(SADMC_ERR_INVALID, "null engine");
  if (refuse_binning(e, "sadmc_get_bins")) return SADMC_ERR_INVALID;
  WalkerRec r;
  int rc = fetch_walker(e, w, &r);
  if (rc) return rc;
  const size_t n = (size_t)r.len;
  if (cap < n) return fail(SADMC_ERR_INVALID, "capacity %u < bins_len %zu", cap, n);
  const size_t base = (size_t)w * e->P.cap + (size_t)r.lo;
  std::vector<BinRec> recs(n);
  CK(cudaMemcpyAsync(recs.data(), e->P.rec + base, n * sizeof(BinRec), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  const bool rt = !(e->P.flags & SADMC_FLAG_NO_ROUND_TRIPS);
  for (size_t i = 0; i < n; i++) {
    const BinRec& b = recs[i];
    if (histogram) histogram[i] = b.lo.hist;
    if (lnw) lnw[i] = b.lo.lnw;
    if (energy_total) energy_total[i] = b.lo.etot;
    if (energy_squared_total) energy_squared_total[i] = b.lo.e2tot;
    if (t_found) t_found[i] = b.hi.t_found;
    if (round_trips) round_trips[i] = rt ? b.hi.round_trips + 1 : 1; // stored minus one
    if (wl_hist) wl_hist[i] = r.wl_hist_len > 0 ? b.hi.wl_hist : 0;
    if (have_visited) {
      const int j = r.lo + (int)i;
      bool v = true;
      if (rt && j >= r.rt_fill_lo && j < r.rt_fill_hi) v = b.hi.rt_stamp > r.rt_fill_time ? true : (r.rt_fill_val != 0);
      have_visited[i] = v ? 1 : 0;
    }
  }
  if (extra_total) {
    if (e->P.extra_total) {
      CK(cudaMemcpyAsync(extra_total, e->P.extra_total + base, n * 8, cudaMemcpyDeviceToHost, e->stream));
      CK(cudaStreamSynchronize(e->stream));
    } else
      for (size_t i = 0; i < n; i++) extra_total[i] = 0;
  }
  if (extra_count) {
    if (e->P.extra_count) {
      CK(cudaMemcpyAsync(extra_count, e->P.extra_count + base, n * 8, cudaMemcpyDeviceToHost, e->stream));
      CK(cudaStreamSynchronize(e->stream));
    } else
      for (size_t i = 0; i < n; i++) extra_count[i] = 0;
  }
  return 0;
}

// ---- SADMC_FLAG_BINNING: the state of energy_binning.rs's EnergyMC over binning::histogram::Bins ----
static int need_binning(sadmc_engine* e) {
  if (!e) return fail(SADMC_ERR_INVALID, "null engine");
  if (!(e->cfg.flags & SADMC_FLAG_BINNING)) return fail(SADMC_ERR_INVALID, "engine was not created with SADMC_FLAG_BINNING");
  return 0;
}
int sadmc_get_binning_walker(sadmc_engine* e, uint32_t w, sadmc_binning_state* s) {
  int rc = need_binning(e);
  if (rc) return rc;
  if (!s) return fail(SADMC_ERR_INVALID, "null argument");
  WalkerRec r;
  rc = fetch_walker(e, w, &r);
  if (rc) return rc;
  memset(s, 0, sizeof *s);
  s->moves = e->moves;
  s->accepted_moves = r.accepted;
  s->acceptance_rate = r.acc_rate;
  s->translation_scale = r.tscale;
  s->rng_s0 = r.s0;
  s->rng_s1 = r.s1;
  s->energy = r.E;
  s->bins_width = e->P.width;
  if (e->moves == 0) { // nothing has called prep_for_e yet: Bins::new (histogram.rs:170-180) with empty vectors
    s->bins_min = (std::round(r.E / e->P.width) - 0.5) * e->P.width;
    s->bins_len = 0;
  } else {
    s->bins_min = r.bmin;
    s->bins_len = (uint32_t)r.len;
  }
  s->bins_min_e = r.b_min_e;
  s->bins_max_e = r.b_max_e;
  s->window_first = (uint32_t)r.lo;
  s->method = r.method == SADMC_METHOD_WL && e->P.inv_t ? SADMC_METHOD_INV_T_WL : r.method;
  s->status = r.status;
  s->too_lo = r.too_lo;
  s->too_hi = r.too_hi;
  s->latest_parameter = r.latest_parameter;
  s->tF = r.b_tF;
  s->tL = r.tL;
  s->num_states = r.num_states;
  s->samc_t0 = r.samc_t0;
  s->wl_gamma = r.wl_gamma;
  s->wl_inv_t = e->P.inv_t;
  const bool linear = (e->cfg.flags & SADMC_FLAG_BINNING_LINEAR) != 0;
  s->lnw_max_count = linear ? (uint64_t)r.l_max_count : r.highest_hist;
  s->lnw_max_count_f64 = linear ? r.l_max_count : (double)r.highest_hist;
  s->lnw_total_count = e->moves;
  s->t_found_max_total = r.b_tf_max;
  s->hist_min_count = linear ? (uint64_t)r.l_hist_min : r.b_hist_min;
  s->hist_min_count_f64 = linear ? r.l_hist_min : (double)r.b_hist_min;
  s->hist_total_count = r.b_hist_total;
  return 0;
}
int sadmc_get_binning_bins(sadmc_engine* e, uint32_t w, uint32_t cap, double* lnw_total, uint64_t* lnw_count, double* energy_total,
                           uint64_t* energy_count, double* t_found_total, uint64_t* t_found_count, uint64_t* hist_count, double* extra_total,
                           uint64_t* extra_count) {
  int rc = need_binning(e);
  if (rc) return rc;
  if (e->cfg.flags & SADMC_FLAG_BINNING_LINEAR) return fail(SADMC_ERR_INVALID, "binning::linear keeps f64 counts: use sadmc_get_binning_bins_f64");
  WalkerRec r;
  rc = fetch_walker(e, w, &r);
  if (rc) return rc;
  const size_t n = e->moves == 0 ? 0 : (size_t)r.len;
  if (cap < n) return fail(SADMC_ERR_INVALID, "capacity %u < bins_len %zu", cap, n);
  if (n == 0) return 0;
  const size_t base = (size_t)w * e->P.cap + (size_t)r.lo;
  std::vector<BinRec> recs(n);
  CK(cudaMemcpyAsync(recs.data(), e->P.rec + base, n * sizeof(BinRec), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  for (size_t i = 0; i < n; i++) { // record layout: book_binning.cuh
    const BinRec& b = recs[i];
    if (lnw_total) lnw_total[i] = b.lo.lnw;
    if (lnw_count) lnw_count[i] = b.lo.hist;
    if (energy_total) energy_total[i] = b.lo.etot;
    if (energy_count) memcpy(&energy_count[i], &b.lo.e2tot, 8);
    if (t_found_total) memcpy(&t_found_total[i], &b.hi.t_found, 8);
    if (t_found_count) t_found_count[i] = b.hi.rt_stamp;
    if (hist_count) hist_count[i] = b.hi.wl_hist;
  }
  if (extra_total) {
    if (e->P.extra_total) {
      CK(cudaMemcpyAsync(extra_total, e->P.extra_total + base, n * 8, cudaMemcpyDeviceToHost, e->stream));
      CK(cudaStreamSynchronize(e->stream));
    } else
      for (size_t i = 0; i < n; i++) extra_total[i] = 0;
  }
  if (extra_count) {
    if (e->P.extra_count) {
      CK(cudaMemcpyAsync(extra_count, e->P.extra_count + base, n * 8, cudaMemcpyDeviceToHost, e->stream));
      CK(cudaStreamSynchronize(e->stream));
    } else
      for (size_t i = 0; i < n; i++) extra_count[i] = 0;
  }
  return 0;
}

int sadmc_get_binning_bins_f64(sadmc_engine* e, uint32_t w, uint32_t cap, double* lnw_total, double* lnw_count, double* energy_total,
                               double* energy_count, double* t_found_total, double* t_found_count, double* hist_count, double* extra_total,
                               double* extra_count) {
  int rc = need_binning(e);
  if (rc) return rc;
  const bool linear = (e->cfg.flags & SADMC_FLAG_BINNING_LINEAR) != 0;
  WalkerRec r;
  rc = fetch_walker(e, w, &r);
  if (rc) return rc;
  const size_t n = e->moves == 0 ? 0 : (size_t)r.len;
  if (cap < n) return fail(SADMC_ERR_INVALID, "capacity %u < bins_len %zu", cap, n);
  if (n == 0) return 0;
  const size_t base = (size_t)w * e->P.cap + (size_t)r.lo;
  std::vector<BinRec> recs(n);
  CK(cudaMemcpyAsync(recs.data(), e->P.rec + base, n * sizeof(BinRec), cudaMemcpyDeviceToHost, e->stream));
  std::vector<uint64_t> xc(n, 0);
  if (e->P.extra_count) CK(cudaMemcpyAsync(xc.data(), e->P.extra_count + base, n * 8, cudaMemcpyDeviceToHost, e->stream));
  if (extra_total) {
    if (e->P.extra_total)
      CK(cudaMemcpyAsync(extra_total, e->P.extra_total + base, n * 8, cudaMemcpyDeviceToHost, e->stream));
    else
      for (size_t i = 0; i < n; i++) extra_total[i] = 0;
  }
  CK(cudaStreamSynchronize(e->stream));
  auto word = [&](const void* p) { // a count word of the record: f64 for linear engines (book_linear.cuh), u64 otherwise
    double d;
    uint64_t u;
    memcpy(&d, p, 8);
    memcpy(&u, p, 8);
    return linear ? d : (double)u;
  };
  for (size_t i = 0; i < n; i++) {
    const BinRec& b = recs[i];
    if (lnw_total) lnw_total[i] = b.lo.lnw;
    if (lnw_count) lnw_count[i] = word(&b.lo.hist);
    if (energy_total) energy_total[i] = b.lo.etot;
    if (energy_count) energy_count[i] = word(&b.lo.e2tot);
    if (t_found_total) memcpy(&t_found_total[i], &b.hi.t_found, 8);
    if (t_found_count) t_found_count[i] = word(&b.hi.rt_stamp);
    if (hist_count) hist_count[i] = word(&b.hi.wl_hist);
    if (extra_count) extra_count[i] = word(&xc[i]);
  }
  return 0;
}
int sadmc_get_high_resolution(sadmc_engine* e, uint32_t w, uint32_t cap, double* bins_min, uint32_t* len, uint64_t* count) {
  int rc = need_binning(e);
  if (rc) return rc;
  if (!e->P.hr_count) return fail(SADMC_ERR_INVALID, "engine was created without high_resolution_de");
  WalkerRec r;
  rc = fetch_walker(e, w, &r);
  if (rc) return rc;
  if (bins_min) *bins_min = r.hr_min;
  if (len) *len = (uint32_t)r.hr_len;
  if (count) {
    if (cap < (uint32_t)r.hr_len) return fail(SADMC_ERR_INVALID, "capacity %u < %d high-resolution bins", cap, r.hr_len);
    if (r.hr_len) {
      CK(cudaMemcpyAsync(count, e->P.hr_count + (size_t)w * e->P.hr_cap + (size_t)r.hr_lo, (size_t)r.hr_len * 8, cudaMemcpyDeviceToHost, e->stream));
      CK(cudaStreamSynchronize(e->stream));
    }
  }
  return 0;
}
int sadmc_set_high_resolution(sadmc_engine* e, uint32_t w, double bins_min, uint32_t len, const uint64_t* count) {
  int rc = need_binning(e);
  if (rc) return rc;
  if (!e->P.hr_count) return fail(SADMC_ERR_INVALID, "engine was created without high_resolution_de");
  if (w >= e->P.n_walkers) return fail(SADMC_ERR_INVALID, "walker %u out of range", w);
  if (len && !count) return fail(SADMC_ERR_INVALID, "null argument");
  const DevParams& P = e->P;
  const long long lo = len ? (long long)std::floor(bins_min / P.hr_width + 0.5) - P.hr_kbase : 0;
  if (lo < 0 || lo + (long long)len > (long long)P.hr_cap) return fail(SADMC_ERR_WINDOW, "the checkpointed high-resolution histogram does not fit its device window");
  WalkerRec r;
  rc = fetch_walker(e, w, &r);
  if (rc) return rc;
  r.hr_min = bins_min;
  r.hr_lo = (int)lo;
  r.hr_len = (int)len;
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaMemsetAsync(P.hr_count + (size_t)w * P.hr_cap, 0, (size_t)P.hr_cap * 8, e->stream));
  if (len) CK(cudaMemcpyAsync(P.hr_count + (size_t)w * P.hr_cap + lo, count, (size_t)len * 8, cudaMemcpyHostToDevice, e->stream));
  CK(cudaMemcpyAsync(P.walkers + w, &r, sizeof r, cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}

// resume for SADMC_FLAG_BINNING engines: the inverse of sadmc_get_binning_walker / sadmc_get_binning_bins
int sadmc_set_binning_walker(sadmc_engine* e, uint32_t w, const sadmc_binning_state* s, const double* lnw_total, const uint64_t* lnw_count,
                             const double* energy_total, const uint64_t* energy_count, const double* t_found_total, const uint64_t* t_found_count,
                             const uint64_t* hist_count, const double* extra_total, const uint64_t* extra_count) {
  int rc = need_binning(e);
  if (rc) return rc;
  if (!s || !lnw_total || !lnw_count || !energy_total || !energy_count)
    return fail(SADMC_ERR_INVALID, "null argument (lnw_total, lnw_count, energy_total, energy_count are required)");
  if (e->cfg.flags & SADMC_FLAG_BINNING_LINEAR) return fail(SADMC_ERR_UNSUPPORTED, "resuming a binning::linear engine is not built");
  if (w >= e->P.n_walkers) return fail(SADMC_ERR_INVALID, "walker %u out of range", w);
  if (e->cfg.init_mode != SADMC_INIT_EXTERNAL) return fail(SADMC_ERR_INVALID, "resume needs an engine created with SADMC_INIT_EXTERNAL");
  const DevParams& P = e->P;
  if (s->bins_width != P.width) return fail(SADMC_ERR_INVALID, "checkpoint bin width %g differs from the engine's %g", s->bins_width, P.width);
  const size_t n = s->bins_len;
  if (n == 0) return fail(SADMC_ERR_INVALID, "a checkpoint without bins (moves == 0) is a fresh start: create the engine without SADMC_INIT_EXTERNAL");
  // window bin j covers [(k_base + j) width, (k_base + j + 1) width); bins_min sits on such an edge up to the rounding of
  // the repeated `min -= width` (histogram.rs:159)
  const long long lo = (long long)std::floor(s->bins_min / P.width + 0.5) - e->k_base;
  if (lo < 0 || lo + (long long)n > (long long)P.cap)
    return fail(SADMC_ERR_WINDOW, "checkpointed bins [%g, %g) do not fit the device window", s->bins_min, s->bins_min + n * P.width);
  WalkerRec old;
  rc = fetch_walker(e, w, &old);
  if (rc) return rc;
  WalkerRec r;
  memset(&r, 0, sizeof r);
  r.err = old.err; // the system-side fields that sadmc_set_system(s) already placed in the record
  r.d_squared = old.d_squared;
  r.hr_min = old.hr_min; // ... and sadmc_set_high_resolution
  r.hr_lo = old.hr_lo;
  r.hr_len = old.hr_len;
  r.s0 = s->rng_s0;
  r.s1 = s->rng_s1;
  r.accepted = s->accepted_moves;
  r.acc_rate = s->acceptance_rate;
  r.tscale = s->translation_scale;
  r.E = s->energy;
  r.bmin = s->bins_min;
  r.lo = (int)lo;
  r.len = (int)n;
  r.method = s->method == SADMC_METHOD_INV_T_WL ? SADMC_METHOD_WL : s->method;
  r.too_lo = s->too_lo;
  r.too_hi = s->too_hi;
  r.latest_parameter = s->latest_parameter;
  r.b_tF = s->tF;
  r.tL = s->tL;
  r.num_states = s->num_states;
  r.samc_t0 = s->samc_t0;
  r.wl_gamma = s->wl_gamma;
  r.highest_hist = s->lnw_max_count;
  r.b_tf_max = s->t_found_max_total;
  r.b_min_e = s->bins_min_e;
  r.b_max_e = s->bins_max_e;
  r.b_hist_total = s->hist_total_count;
  // derived device fields
  auto widx = [&](double energy) { // histogram.rs:135-146 shifted into the window (book_binning.cuh widx)
    if (energy < s->bins_min) return (int)lo;
    const double fi = (energy - s->bins_min) / P.width;
    if (fi == (double)n || !(fi < (double)n)) return (int)(lo + (long long)n - 1);
    return (int)(lo + (long long)fi);
  };
  r.ilo = widx(s->too_lo);
  r.ihi = widx(s->too_hi);
  unsigned long long hmin = ~0ull;
  long long nmin = 0;
  double mx = 0.0;
  for (size_t j = 0; j < n; j++) {
    const unsigned long long h = hist_count ? hist_count[j] : 0;
    if (h < hmin) {
      hmin = h;
      nmin = 1;
    } else if (h == hmin) {
      nmin++;
    }
    if (lnw_total[j] > mx) mx = lnw_total[j];
  }
  r.b_hist_min = hmin; // == s->hist_min_count: the reference's min_count is the true minimum at all times (histogram.rs:241-245)
  r.b_hist_nmin = nmin;
  r.max_S = mx;
  std::vector<BinRec> recs(n);
  for (size_t j = 0; j < n; j++) { // record layout: book_binning.cuh
    BinRec& b = recs[j];
    b.lo.lnw = lnw_total[j];
    b.lo.hist = lnw_count[j];
    b.lo.etot = energy_total[j];
    memcpy(&b.lo.e2tot, &energy_count[j], 8);
    const double tft = t_found_total ? t_found_total[j] : 0.0;
    memcpy(&b.hi.t_found, &tft, 8);
    b.hi.rt_stamp = t_found_count ? t_found_count[j] : 0;
    b.hi.round_trips = 0;
    b.hi.wl_hist = hist_count ? hist_count[j] : 0;
  }
  CK(cudaSetDevice(e->cfg.device));
  const size_t base = (size_t)w * P.cap;
  CK(cudaMemsetAsync(P.rec + base, 0, (size_t)P.cap * sizeof(BinRec), e->stream));
  CK(cudaMemcpyAsync(P.rec + base + lo, recs.data(), n * sizeof(BinRec), cudaMemcpyHostToDevice, e->stream));
  if (P.extra_total) {
    CK(cudaMemsetAsync(P.extra_total + base, 0, (size_t)P.cap * 8, e->stream));
    CK(cudaMemsetAsync(P.extra_count + base, 0, (size_t)P.cap * 8, e->stream));
    if (extra_total) CK(cudaMemcpyAsync(P.extra_total + base + lo, extra_total, n * 8, cudaMemcpyHostToDevice, e->stream));
    if (extra_count) CK(cudaMemcpyAsync(P.extra_count + base + lo, extra_count, n * 8, cudaMemcpyHostToDevice, e->stream));
  }
  CK(cudaMemcpyAsync(P.walkers + w, &r, sizeof r, cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}

// ---- resume: the inverse of sadmc_get_walker / sadmc_get_bins (mc/mod.rs:70-84 deserialises a whole EnergyMC) ----
int sadmc_set_walker_bins(sadmc_engine* e, uint32_t w, const sadmc_walker_state* s, const uint64_t* histogram, const uint64_t* t_found,
                          const double* lnw, const double* energy_total, const double* energy_squared_total, const uint64_t* round_trips,
                          const uint8_t* have_visited, const uint64_t* wl_hist, const double* extra_total, const uint64_t* extra_count) {
  if (!e || !s || !histogram || !lnw || !energy_total || !energy_squared_total)
    return fail(SADMC_ERR_INVALID, "null argument (histogram, lnw, energy_total, energy_squared_total are required)");
  if (refuse_binning(e, "sadmc_set_walker_bins")) return SADMC_ERR_INVALID;
  if (w >= e->P.n_walkers) return fail(SADMC_ERR_INVALID, "walker %u out of range", w);
  if (e->started && e->cfg.init_mode != SADMC_INIT_EXTERNAL) return fail(SADMC_ERR_INVALID, "resume needs an engine created with SADMC_INIT_EXTERNAL");
  const DevParams& P = e->P;
  if (s->bins_width != P.width) return fail(SADMC_ERR_INVALID, "checkpoint bin width %g differs from the engine's %g", s->bins_width, P.width);
  const size_t n = s->bins_len;
  // window index of reference bin 0: bin j of the window covers [(k_base + j - 0.5) w, (k_base + j + 0.5) w)
  const long long lo = (long long)std::floor(s->bins_min / P.width + 0.5 + 0.5) - e->k_base;
  if (n == 0 || lo < 0 || lo + (long long)n > (long long)P.cap)
    return fail(SADMC_ERR_WINDOW, "checkpointed bins [%g, %g) do not fit the device window", s->bins_min, s->bins_min + n * P.width);
  WalkerRec r;
  memset(&r, 0, sizeof r);
  r.s0 = s->rng_s0;
  r.s1 = s->rng_s1;
  r.accepted = s->accepted_moves;
  r.acc_rate = s->acceptance_rate;
  r.tscale = s->translation_scale;
  r.E = s->energy;
  r.bmin = s->bins_min;
  r.lo = (int)lo;
  r.len = (int)n;
  r.method = s->method == SADMC_METHOD_INV_T_WL ? SADMC_METHOD_WL : s->method;
  r.status = 0;
  r.too_lo = s->too_lo;
  r.too_hi = s->too_hi;
  r.latest_parameter = s->latest_parameter;
  r.tL = s->tL;
  r.tF = s->tF;
  r.num_states = s->num_states;
  r.highest_hist = s->highest_hist;
  r.samc_t0 = s->samc_t0;
  r.wl_gamma = s->wl_gamma;
  r.wl_num_states = s->wl_num_states;
  r.wl_min_energy = s->wl_min_energy;
  r.wl_lowest = s->wl_lowest_hist;
  r.wl_highest = s->wl_highest_hist;
  r.wl_total = s->wl_total_hist;
  r.wl_hist_len = (int)s->wl_hist_len;
  r.max_S = s->max_S;
  r.max_S_index = (int)s->max_S_index;
  // derived device fields
  auto ref_index = [&](double energy) { // Bins::state_to_index, energy.rs:371-373 (saturating cast)
    const double x = (energy - s->bins_min) / P.width;
    long long i = !(x > 0.0) ? 0 : (x >= 2147483647.0 ? 2147483647ll : (long long)x);
    if (i >= (long long)n) i = (long long)n - 1;
    return (int)i;
  };
  r.ilo = r.lo + ref_index(s->too_lo);
  r.ihi = r.lo + ref_index(s->too_hi);
  unsigned long long tfmax = 0;
  if (t_found)
    for (int j = r.ilo - r.lo; j <= r.ihi - r.lo; j++)
      if (j >= 0 && j < (int)n && t_found[j] > tfmax) tfmax = t_found[j];
  r.tfmax = tfmax;
  long long low = 0;
  if (wl_hist && r.wl_hist_len > 0)
    for (size_t j = 0; j < n; j++)
      if (histogram[j] != 0 && wl_hist[j] <= r.wl_lowest) low++;
  r.wl_low_count = low;
  // have_visited_since_maxentropy as "all false since move 0" plus individual stamps at move 1
  r.rt_fill_val = 0;
  r.rt_fill_time = 0;
  r.rt_fill_lo = r.lo;
  r.rt_fill_hi = r.lo + r.len;
  // keep the system-side fields that sadmc_set_system(s) already placed in the record
  WalkerRec old;
  int rc = fetch_walker(e, w, &old);
  if (rc) return rc;
  r.err = old.err;
  r.d_squared = old.d_squared;
  std::vector<BinRec> recs(n);
  for (size_t j = 0; j < n; j++) {
    BinRec& b = recs[j];
    b.lo.lnw = lnw[j];
    b.lo.hist = histogram[j];
    b.lo.etot = energy_total[j];
    b.lo.e2tot = energy_squared_total[j];
    b.hi.t_found = t_found ? t_found[j] : 0;
    b.hi.rt_stamp = (have_visited && have_visited[j]) ? 1 : 0;
    b.hi.round_trips = round_trips ? (round_trips[j] > 0 ? round_trips[j] - 1 : 0) : 0; // stored minus one
    b.hi.wl_hist = (wl_hist && r.wl_hist_len > 0) ? wl_hist[j] : 0;
  }
  CK(cudaSetDevice(e->cfg.device));
  const size_t base = (size_t)w * P.cap;
  CK(cudaMemsetAsync(P.rec + base, 0, (size_t)P.cap * sizeof(BinRec), e->stream));
  CK(cudaMemcpyAsync(P.rec + base + lo, recs.data(), n * sizeof(BinRec), cudaMemcpyHostToDevice, e->stream));
  if (P.extra_total) {
    CK(cudaMemsetAsync(P.extra_total + base, 0, (size_t)P.cap * 8, e->stream));
    CK(cudaMemsetAsync(P.extra_count + base, 0, (size_t)P.cap * 8, e->stream));
    if (extra_total) CK(cudaMemcpyAsync(P.extra_total + base + lo, extra_total, n * 8, cudaMemcpyHostToDevice, e->stream));
    if (extra_count) CK(cudaMemcpyAsync(P.extra_count + base + lo, extra_count, n * 8, cudaMemcpyHostToDevice, e->stream));
  }
  CK(cudaMemcpyAsync(P.walkers + w, &r, sizeof r, cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}

int sadmc_resume(sadmc_engine* e, uint64_t moves) {
  if (!e) return fail(SADMC_ERR_INVALID, "null engine");
  if (e->started) return fail(SADMC_ERR_INVALID, "engine already started");
  if (e->cfg.init_mode != SADMC_INIT_EXTERNAL) return fail(SADMC_ERR_INVALID, "resume needs an engine created with SADMC_INIT_EXTERNAL");
  e->started = true;
  e->moves = moves;
  return 0;
}

int sadmc_system_len(sadmc_engine* e, size_t* n) {
  if (!e || !n) return fail(SADMC_ERR_INVALID, "null argument");
  *n = e->sys_len;
  return 0;
}

// ---- system images: ABI f64 layout <-> device-native layout -------------------
static int systems_to_host(sadmc_engine* e, uint32_t w0, uint32_t nw, double* buf) {
  const DevParams& P = e->P;
  std::vector<WalkerRec> recs(nw);
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaMemcpyAsync(recs.data(), P.walkers + w0, nw * sizeof(WalkerRec), cudaMemcpyDeviceToHost, e->stream));
  if (e->cfg.system == SADMC_SYS_ISING) {
    std::vector<uint32_t> words((size_t)nw * P.ising_words);
    CK(cudaMemcpyAsync(words.data(), P.sys_words + (size_t)w0 * P.ising_words, words.size() * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    const size_t sites = (size_t)P.N * P.N;
    for (uint32_t w = 0; w < nw; w++) {
      double* o = buf + (size_t)w * e->sys_len;
      const uint32_t* wd = &words[(size_t)w * P.ising_words];
      for (size_t s = 0; s < sites; s++) o[s] = ((wd[s >> 5] >> (s & 31)) & 1u) ? 1.0 : -1.0;
      o[sites] = recs[w].E;
    }
    return 0;
  }
  CK(cudaMemcpyAsync(buf, P.sys + (size_t)w0 * P.sys_stride, (size_t)nw * P.sys_stride * 8, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}
static int systems_from_host(sadmc_engine* e, uint32_t w0, uint32_t nw, const double* buf) {
  const DevParams& P = e->P;
  CK(cudaSetDevice(e->cfg.device));
  std::vector<WalkerRec> recs(nw);
  CK(cudaMemcpyAsync(recs.data(), P.walkers + w0, nw * sizeof(WalkerRec), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  if (e->cfg.system == SADMC_SYS_ISING) {
    std::vector<uint32_t> words((size_t)nw * P.ising_words, 0u);
    const size_t sites = (size_t)P.N * P.N;
    for (uint32_t w = 0; w < nw; w++) {
      const double* in = buf + (size_t)w * e->sys_len;
      uint32_t* wd = &words[(size_t)w * P.ising_words];
      for (size_t s = 0; s < sites; s++)
        if (in[s] > 0) wd[s >> 5] |= 1u << (s & 31);
      recs[w].E = in[sites];
      recs[w].err = 0;
    }
    CK(cudaMemcpyAsync(P.sys_words + (size_t)w0 * P.ising_words, words.data(), words.size() * 4, cudaMemcpyHostToDevice, e->stream));
  } else {
    CK(cudaMemcpyAsync(P.sys + (size_t)w0 * P.sys_stride, buf, (size_t)nw * P.sys_stride * 8, cudaMemcpyHostToDevice, e->stream));
    for (uint32_t w = 0; w < nw; w++) {
      const double* in = buf + (size_t)w * e->sys_len;
      if (e->cfg.system == SADMC_SYS_LJ || e->cfg.system == SADMC_SYS_WCA || e->cfg.system == SADMC_SYS_SW) {
        recs[w].E = in[3 * (size_t)P.N];
        recs[w].err = in[3 * (size_t)P.N + 1];
      } else if (e->cfg.system == SADMC_SYS_TWO_WELLS) {
        recs[w].d_squared = in[P.N];
      }
    }
  }
  CK(cudaMemcpyAsync(P.walkers + w0, recs.data(), nw * sizeof(WalkerRec), cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}

int sadmc_get_system(sadmc_engine* e, uint32_t w, double* buf, size_t n) {
  if (!e || !buf) return fail(SADMC_ERR_INVALID, "null argument");
  if (w >= e->P.n_walkers) return fail(SADMC_ERR_INVALID, "walker %u out of range", w);
  if (n < e->sys_len) return fail(SADMC_ERR_INVALID, "buffer of %zu doubles < system_len %zu", n, e->sys_len);
  return systems_to_host(e, w, 1, buf);
}
int sadmc_set_system(sadmc_engine* e, uint32_t w, const double* buf, size_t n) {
  if (!e || !buf) return fail(SADMC_ERR_INVALID, "null argument");
  if (w >= e->P.n_walkers) return fail(SADMC_ERR_INVALID, "walker %u out of range", w);
  if (n < e->sys_len) return fail(SADMC_ERR_INVALID, "buffer of %zu doubles < system_len %zu", n, e->sys_len);
  return systems_from_host(e, w, 1, buf);
}
int sadmc_get_systems(sadmc_engine* e, double* buf, size_t n) {
  if (!e || !buf) return fail(SADMC_ERR_INVALID, "null argument");
  if (n < e->sys_len * e->P.n_walkers) return fail(SADMC_ERR_INVALID, "buffer too small");
  return systems_to_host(e, 0, e->P.n_walkers, buf);
}
int sadmc_set_systems(sadmc_engine* e, const double* buf, size_t n) {
  if (!e || !buf) return fail(SADMC_ERR_INVALID, "null argument");
  if (n < e->sys_len * e->P.n_walkers) return fail(SADMC_ERR_INVALID, "buffer too small");
  return systems_from_host(e, 0, e->P.n_walkers, buf);
}
int sadmc_get_rngs(sadmc_engine* e, uint64_t* s) {
  if (!e || !s) return fail(SADMC_ERR_INVALID, "null argument");
  std::vector<WalkerRec> v;
  int rc = fetch_walkers(e, v);
  if (rc) return rc;
  for (size_t i = 0; i < v.size(); i++) {
    s[2 * i] = v[i].s0;
    s[2 * i + 1] = v[i].s1;
  }
  return 0;
}
int sadmc_set_rngs(sadmc_engine* e, const uint64_t* s) {
  if (!e || !s) return fail(SADMC_ERR_INVALID, "null argument");
  std::vector<WalkerRec> v;
  int rc = fetch_walkers(e, v);
  if (rc) return rc;
  for (size_t i = 0; i < v.size(); i++) {
    v[i].s0 = s[2 * i];
    v[i].s1 = s[2 * i + 1];
  }
  CK(cudaMemcpyAsync(e->P.walkers, v.data(), v.size() * sizeof(WalkerRec), cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}

int sadmc_window(sadmc_engine* e, double* lo, double* width, uint32_t* nbins) {
  if (!e) return fail(SADMC_ERR_INVALID, "null engine");
  if (lo) *lo = ((double)e->k_base - ((e->cfg.flags & SADMC_FLAG_BINNING) ? 0.0 : 0.5)) * e->P.width;
  if (width) *width = e->P.width;
  if (nbins) *nbins = e->P.cap;
  return 0;
}

int sadmc_cell_box(sadmc_engine* e, double box_diagonal[3], double* r_cutoff) {
  if (!e) return fail(SADMC_ERR_INVALID, "null engine");
  if (e->cfg.system != SADMC_SYS_WCA && e->cfg.system != SADMC_SYS_SW) return fail(SADMC_ERR_INVALID, "system %d has no periodic cell", e->cfg.system);
  if (box_diagonal)
    for (int k = 0; k < 3; k++) box_diagonal[k] = e->P.box[k];
  if (r_cutoff) *r_cutoff = e->cfg.system == SADMC_SYS_SW ? e->cfg.sw_well_width * 1.0 : std::pow(2.0, 1.0 / 6.0);
  return 0;
}

// ---- fixed weights for a production run -------------------------------------------
__global__ void __launch_bounds__(256) set_lnw_kernel(const DevParams P, const double* lnw) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t w = blockIdx.y;
  if (j < P.cap) P.rec[(size_t)w * P.cap + j].lo.lnw = lnw[j];
}

// ---- merge for reporting -------------------------------------------------------
int sadmc_fold_select_ex(sadmc_engine* e, uint32_t first_walker, uint32_t walker_stride, uint32_t walker_count, int sad_range_only) {
  if (!e) return fail(SADMC_ERR_INVALID, "null engine");
  if (walker_stride == 0 || first_walker >= e->P.n_walkers) return fail(SADMC_ERR_INVALID, "fold selection (%u, %u) holds no walker", first_walker, walker_stride);
  if (sad_range_only < 0 || sad_range_only > 2) return fail(SADMC_ERR_INVALID, "sad_range_only must be 0, 1 or 2");
  e->fold_sel.first = first_walker;
  e->fold_sel.stride = walker_stride;
  e->fold_sel.count = walker_count;
  e->fold_sel.sad_range_only = sad_range_only;
  return 0;
}
int sadmc_set_lnw(sadmc_engine* e, const double* lnw_window, uint32_t n) {
  if (!e || !lnw_window) return fail(SADMC_ERR_INVALID, "null argument");
  if (refuse_binning(e, "sadmc_set_lnw")) return SADMC_ERR_INVALID;
  if (n != e->P.cap) return fail(SADMC_ERR_INVALID, "ln w array holds %u bins, the device window %u", n, e->P.cap);
  CK(cudaSetDevice(e->cfg.device));
  double* d = nullptr;
  CK(cudaMalloc(&d, (size_t)n * 8));
  cudaError_t er = cudaMemcpyAsync(d, lnw_window, (size_t)n * 8, cudaMemcpyHostToDevice, e->stream);
  if (er == cudaSuccess) {
    const uint32_t wmax = 65535; // grid.y limit
    for (uint32_t w0 = 0; w0 < e->P.n_walkers && er == cudaSuccess; w0 += wmax) {
      DevParams Q = e->P;
      Q.rec = e->P.rec + (size_t)w0 * e->P.cap;
      const uint32_t nw = e->P.n_walkers - w0 < wmax ? e->P.n_walkers - w0 : wmax;
      set_lnw_kernel<<<dim3((n + 255) / 256, nw), 256, 0, e->stream>>>(Q, d);
      er = cudaGetLastError();
      e->launches++;
    }
  }
  if (er == cudaSuccess) er = cudaStreamSynchronize(e->stream);
  cudaFree(d);
  if (er != cudaSuccess) return fail(SADMC_ERR_CUDA, "sadmc_set_lnw failed: %s", cudaGetErrorString(er));
  return 0;
}
int sadmc_fold_settled(sadmc_engine* e, uint64_t tl_max) {
  if (!e) return fail(SADMC_ERR_INVALID, "null engine");
  e->fold_sel.tl_max = tl_max;
  return 0;
}
int sadmc_fold_select(sadmc_engine* e, uint32_t first_walker, uint32_t walker_stride, int sad_range_only) {
  return sadmc_fold_select_ex(e, first_walker, walker_stride, 0u, sad_range_only);
}
static int fold_launch(sadmc_engine* e, void* d_histogram, void* d_energy_total, void* d_energy_squared_total, void* d_lnw_sum,
                       void* d_lnw_sq_sum, void* d_lnw_count, double* d_packed) {
  if (!e) return fail(SADMC_ERR_INVALID, "null engine");
  CK(cudaSetDevice(e->cfg.device));
  if (e->cfg.flags & SADMC_FLAG_BINNING_LINEAR) return fail(SADMC_ERR_UNSUPPORTED, "the reporting fold is not built for binning::linear engines (f64 counts)");
  FoldSel sel = e->fold_sel;
  sel.binning = (e->cfg.flags & SADMC_FLAG_BINNING) ? 1 : 0;
  uint32_t n_sel = sel.first < e->P.n_walkers ? (e->P.n_walkers - sel.first + sel.stride - 1) / sel.stride : 0;
  if (sel.count && sel.count < n_sel) n_sel = sel.count;
  if (n_sel == 0) return fail(SADMC_ERR_INVALID, "fold selection holds no walker");
  const double* wmax = nullptr; // one pass: the walkers' running maxima (fold_kernels.cuh)
  if (sel.sad_range_only != 0 || sel.binning) {
    if (!e->d_wmax) {
      int rc = dev_alloc(e, (void**)&e->d_wmax, (size_t)e->P.n_walkers * 8, false);
      if (rc) return rc;
    }
    walker_max_lnw_kernel<<<n_sel, 256, 0, e->stream>>>(e->P, e->d_wmax, sel);
    CK(cudaGetLastError());
    e->launches++;
    wmax = e->d_wmax;
  }
  // chunks of walkers so that the grid covers the chip several times over (148 SMs x 8 blocks)
  const uint32_t nbx = (e->P.cap + 255) / 256;
  uint32_t n_chunks = (1184 + nbx - 1) / nbx;
  const uint32_t max_chunks = (n_sel + 63) / 64;
  if (n_chunks > max_chunks) n_chunks = max_chunks;
  if (n_chunks > 65535) n_chunks = 65535;
  const uint32_t per_chunk = (n_sel + n_chunks - 1) / n_chunks;
  n_chunks = (n_sel + per_chunk - 1) / per_chunk;
  const size_t part_bytes = (size_t)n_chunks * FOLD_FIELDS * e->P.cap * 8;
  if (part_bytes > e->fold_part_bytes) {
    int rc = dev_alloc(e, (void**)&e->d_fold_part, part_bytes, false);
    if (rc) return rc;
    e->fold_part_bytes = part_bytes;
  }
  fold_partial_kernel<<<dim3(nbx, n_chunks), 256, 0, e->stream>>>(e->P, wmax, e->d_fold_part, sel, n_sel, per_chunk);
  CK(cudaGetLastError());
  fold_final_kernel<<<nbx, 256, 0, e->stream>>>(e->P, e->d_fold_part, n_chunks, (unsigned long long*)d_histogram, (double*)d_energy_total,
                                               (double*)d_energy_squared_total, (double*)d_lnw_sum, (double*)d_lnw_sq_sum,
                                               (unsigned long long*)d_lnw_count, d_packed);
  CK(cudaGetLastError());
  e->launches += 2;
  return 0;
}
int sadmc_fold_device(sadmc_engine* e, void* d_histogram, void* d_energy_total, void* d_energy_squared_total, void* d_lnw_sum,
                      void* d_lnw_sq_sum, void* d_lnw_count) {
  return fold_launch(e, d_histogram, d_energy_total, d_energy_squared_total, d_lnw_sum, d_lnw_sq_sum, d_lnw_count, nullptr);
}
int sadmc_fold_packed_device(sadmc_engine* e, void* d_packed) {
  if (!d_packed) return fail(SADMC_ERR_INVALID, "null argument");
  return fold_launch(e, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, (double*)d_packed);
}
int sadmc_fold(sadmc_engine* e, uint64_t* histogram, double* energy_total, double* energy_squared_total, double* lnw_sum,
               double* lnw_sq_sum, uint64_t* lnw_count) {
  if (!e) return fail(SADMC_ERR_INVALID, "null engine");
  const size_t n = e->P.cap;
  if (!e->d_fold) {
    int rc = dev_alloc(e, (void**)&e->d_fold, 6 * n * 8, false);
    if (rc) return rc;
  }
  unsigned long long* d = (unsigned long long*)e->d_fold;
  int rc = sadmc_fold_device(e, d, d + n, d + 2 * n, d + 3 * n, d + 4 * n, d + 5 * n);
  if (rc) return rc;
  void* outs[6] = {histogram, energy_total, energy_squared_total, lnw_sum, lnw_sq_sum, lnw_count};
  for (int k = 0; k < 6; k++)
    if (outs[k]) CK(cudaMemcpyAsync(outs[k], d + k * n, n * 8, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}

// ---- replica exchange: the `tempering` binary (src/mc/tempering.rs; kernels in tempering.cuh) ----
struct sadmc_tempering {
  sadmc_engine* e = nullptr; // systems, generators, kernels (its bins are a dummy window; the engine is never started)
  uint32_t n_sim = 0, n_T = 0;
  unsigned long long steps = 0, rounds = 0;
  TemperRec* d_reps = nullptr;
  unsigned long long* d_mc_rng = nullptr;
  float last_ms = 0.f;
  bool settled = false; // every replica's cached energy is what its system says (see temper_settle)
};
static unsigned long long min_moves_to_randomize(const sadmc_config& c) {
  switch (c.system) {
    case SADMC_SYS_ISING: return (unsigned long long)c.N * c.N;                                              // ising.rs:86-88
    case SADMC_SYS_FAKE: return c.fake_function == SADMC_FAKE_LINEAR ? 1 : (c.fake_function == SADMC_FAKE_QUADRATIC ? c.N : 3); // fake.rs:113-115
    default: return c.N; // lj.rs:280-282, optsquare.rs:205-207, wca.rs, erfinv.rs:86-88, two_wells.rs (position.len())
  }
}
static void xoroshiro_jump(Rng& g) { // rand_xoshiro 0.4 Xoroshiro128Plus::jump: 2^64 calls of next_u64
  static const uint64_t JUMP[2] = {0xdf900294d8f554a5ull, 0x170865df4b3201fcull};
  uint64_t s0 = 0, s1 = 0;
  for (int i = 0; i < 2; i++)
    for (int b = 0; b < 64; b++) {
      if (JUMP[i] & (1ull << b)) {
        s0 ^= g.s0;
        s1 ^= g.s1;
      }
      g.next();
    }
  g.s0 = s0;
  g.s1 = s1;
}
// A launch with zero moves: load + store of every system.  Host-side constructors that do not know the energy hand over
// NaN (the fluids sum it when they load, sys_cell_fluid.cuh), and the analytic systems derive theirs from the positions;
// afterwards images and cached energies are what the reference holds after `system.clone()`.
static int temper_settle(sadmc_tempering* t) {
  if (t->settled) return 0;
  sadmc_engine* e = t->e;
  CK(cudaSetDevice(e->cfg.device));
  int grid;
  launch_cfg(e, &grid);
  e->ks.temper<<<grid, e->ks.block, e->ks.smem, e->stream>>>(e->P, t->d_reps, 0ull);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(e->stream));
  e->launches++;
  t->settled = true;
  return 0;
}
int sadmc_tempering_create(const sadmc_config* cfg, const double* T, uint32_t n_T, uint64_t canonical_steps, sadmc_tempering** out) {
  if (!cfg || !T || !out) return fail(SADMC_ERR_INVALID, "null argument");
  *out = nullptr;
  if (n_T < 1 || cfg->n_walkers < 1) return fail(SADMC_ERR_INVALID, "tempering needs at least one temperature and one simulation");
  if ((unsigned long long)n_T * cfg->n_walkers > 0x7fffffffull) return fail(SADMC_ERR_INVALID, "too many replicas");
  if (cfg->init_mode != SADMC_INIT_REFERENCE && cfg->init_mode != SADMC_INIT_EXTERNAL)
    return fail(SADMC_ERR_INVALID, "tempering starts every replica from the reference constructor's system (tempering.rs:161) or from supplied systems");
  for (uint32_t r = 0; r < n_T; r++)
    if (!(T[r] > 0)) return fail(SADMC_ERR_INVALID, "temperature %u is not positive", r);
  sadmc_config c = *cfg;
  const uint32_t n_sim = cfg->n_walkers;
  c.n_walkers = n_sim * n_T;
  c.method = SADMC_METHOD_CANONICAL; // unused: the bins below are a dummy window
  c.canonical_T = 1.0;
  c.energy_bin = 1.0;
  c.min_allowed_energy = c.max_allowed_energy = NAN;
  c.bin_window_lo = 0.0;
  c.bin_window_hi = 1.0;
  c.flags &= ~(uint32_t)(SADMC_FLAG_BINNING | SADMC_FLAG_BINNING_LINEAR);
  c.high_resolution_de = NAN;
  if (c.system == SADMC_SYS_LJ && c.lanes_per_walker == 0) c.lanes_per_walker = 1;
  if (c.system == SADMC_SYS_WCA && c.lanes_per_walker == 0) c.lanes_per_walker = 32;
  c.init_mode = SADMC_INIT_EXTERNAL;
  sadmc_engine* e = nullptr;
  int rc = sadmc_create(&c, &e);
  if (rc) return rc;
  if (!e->ks.temper) {
    sadmc_destroy(e);
    return fail(SADMC_ERR_UNSUPPORTED, "no tempering kernel for this system / lanes_per_walker");
  }
  sadmc_tempering* t = new sadmc_tempering;
  t->e = e;
  t->n_sim = n_sim;
  t->n_T = n_T;
  t->steps = min_moves_to_randomize(*cfg) * canonical_steps; // tempering.rs:274
#define TBAIL(expr)             \
  do {                          \
    int _rc = (expr);           \
    if (_rc) {                  \
      sadmc_tempering_destroy(t); \
      return _rc;               \
    }                           \
  } while (0)
  if (cfg->init_mode == SADMC_INIT_REFERENCE) { // system.clone() for every replica (tempering.rs:161)
    std::vector<double> img;
    e->cfg.init_mode = SADMC_INIT_REFERENCE;
    rc = reference_image(e, img);
    e->cfg.init_mode = SADMC_INIT_EXTERNAL;
    TBAIL(rc);
    std::vector<double> all((size_t)c.n_walkers * e->sys_len);
    for (uint32_t w = 0; w < c.n_walkers; w++) memcpy(&all[(size_t)w * e->sys_len], img.data(), e->sys_len * sizeof(double));
    TBAIL(sadmc_set_systems(e, all.data(), all.size()));
  }
  {
    std::vector<uint64_t> rngs((size_t)c.n_walkers * 2), mc((size_t)n_sim * 2);
    std::vector<TemperRec> reps(c.n_walkers);
    for (uint32_t k = 0; k < n_sim; k++) {
      Rng g;
      seed_from_u64(cfg->seed + cfg->walker_offset + k, &g.s0, &g.s1); // tempering.rs:153
      for (uint32_t r = 0; r < n_T; r++) {                             // rng.clone() (161)
        rngs[2 * ((size_t)k * n_T + r)] = g.s0;
        rngs[2 * ((size_t)k * n_T + r) + 1] = g.s1;
        TemperRec& q = reps[(size_t)k * n_T + r];
        memset(&q, 0, sizeof q);
        q.T = T[r];
        q.tscale = 1.0; // Length::new(1.0), tempering.rs:88
      }
      xoroshiro_jump(g); // tempering.rs:164
      mc[2 * k] = g.s0;
      mc[2 * k + 1] = g.s1;
    }
    TBAIL(sadmc_set_rngs(e, rngs.data()));
    TBAIL(dev_alloc(e, (void**)&t->d_reps, reps.size() * sizeof(TemperRec), false));
    TBAIL(dev_alloc(e, (void**)&t->d_mc_rng, mc.size() * 8, false));
    cudaError_t er = cudaMemcpyAsync(t->d_reps, reps.data(), reps.size() * sizeof(TemperRec), cudaMemcpyHostToDevice, e->stream);
    if (er == cudaSuccess) er = cudaMemcpyAsync(t->d_mc_rng, mc.data(), mc.size() * 8, cudaMemcpyHostToDevice, e->stream);
    if (er == cudaSuccess) er = cudaStreamSynchronize(e->stream);
    if (er != cudaSuccess) {
      sadmc_tempering_destroy(t);
      return fail(SADMC_ERR_CUDA, "tempering upload failed: %s", cudaGetErrorString(er));
    }
  }
  if (e->ks.smem > 48 * 1024) {
    cudaError_t er = cudaFuncSetAttribute((const void*)e->ks.temper, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->ks.smem);
    if (er != cudaSuccess) {
      sadmc_tempering_destroy(t);
      return fail(SADMC_ERR_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(er));
    }
  }
#undef TBAIL
  *out = t;
  return 0;
}
void sadmc_tempering_destroy(sadmc_tempering* t) {
  if (!t) return;
  if (t->e) sadmc_destroy(t->e); // d_reps / d_mc_rng are among the engine's allocations
  delete t;
}
int sadmc_tempering_run(sadmc_tempering* t, uint64_t n_rounds) {
  if (!t) return fail(SADMC_ERR_INVALID, "null argument");
  sadmc_engine* e = t->e;
  int src = temper_settle(t);
  if (src) return src;
  CK(cudaSetDevice(e->cfg.device));
  int grid;
  launch_cfg(e, &grid);
  CK(cudaEventRecord(e->ev0, e->stream));
  for (uint64_t r = 0; r < n_rounds; r++) {
    e->ks.temper<<<grid, e->ks.block, e->ks.smem, e->stream>>>(e->P, t->d_reps, t->steps);
    CK(cudaGetLastError());
    temper_swap_kernel<<<t->n_sim, 128, 0, e->stream>>>(e->P, t->d_reps, t->d_mc_rng, t->n_T);
    CK(cudaGetLastError());
    e->launches += 2;
  }
  CK(cudaEventRecord(e->ev1, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaEventElapsedTime(&t->last_ms, e->ev0, e->ev1));
  t->rounds += n_rounds;
  return 0;
}
int sadmc_tempering_cell_box(sadmc_tempering* t, double box_diagonal[3], double* r_cutoff) {
  if (!t) return fail(SADMC_ERR_INVALID, "null argument");
  return sadmc_cell_box(t->e, box_diagonal, r_cutoff);
}
int sadmc_tempering_last_run_ms(sadmc_tempering* t, float* ms) {
  if (!t || !ms) return fail(SADMC_ERR_INVALID, "null argument");
  *ms = t->last_ms;
  return 0;
}
int sadmc_tempering_num_moves(sadmc_tempering* t, uint64_t* moves) {
  if (!t || !moves) return fail(SADMC_ERR_INVALID, "null argument");
  *moves = t->rounds * t->steps * t->n_T; // these_moves = steps per replica, summed (tempering.rs:277-284, 321-323)
  return 0;
}
int sadmc_tempering_steps_per_round(sadmc_tempering* t, uint64_t* steps) {
  if (!t || !steps) return fail(SADMC_ERR_INVALID, "null argument");
  *steps = t->steps;
  return 0;
}
int sadmc_tempering_get_replicas(sadmc_tempering* t, uint32_t sim, sadmc_replica_state* out) {
  if (!t || !out) return fail(SADMC_ERR_INVALID, "null argument");
  if (sim >= t->n_sim) return fail(SADMC_ERR_INVALID, "simulation %u out of range", sim);
  sadmc_engine* e = t->e;
  int src = temper_settle(t);
  if (src) return src;
  std::vector<TemperRec> reps(t->n_T);
  std::vector<WalkerRec> recs(t->n_T);
  CK(cudaMemcpyAsync(reps.data(), t->d_reps + (size_t)sim * t->n_T, t->n_T * sizeof(TemperRec), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaMemcpyAsync(recs.data(), e->P.walkers + (size_t)sim * t->n_T, t->n_T * sizeof(WalkerRec), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  for (uint32_t r = 0; r < t->n_T; r++) {
    sadmc_replica_state& o = out[r];
    o.T = reps[r].T;
    o.rejected_count = reps[r].rejected;
    o.accepted_count = reps[r].accepted;
    o.rejected_swap_count = reps[r].rejected_swap;
    o.accepted_swap_count = reps[r].accepted_swap;
    o.ignored_count = reps[r].ignored;
    o.total_energy = reps[r].total_energy;
    o.total_energy_squared = reps[r].total_energy_squared;
    o.translation_scale = reps[r].tscale;
    o.rng_s0 = recs[r].s0;
    o.rng_s1 = recs[r].s1;
    o.energy = recs[r].E;
  }
  return 0;
}
int sadmc_tempering_set_translation_scales(sadmc_tempering* t, const double* scale) {
  if (!t || !scale) return fail(SADMC_ERR_INVALID, "null argument");
  sadmc_engine* e = t->e;
  const size_t n = (size_t)t->n_sim * t->n_T;
  std::vector<TemperRec> reps(n);
  CK(cudaMemcpyAsync(reps.data(), t->d_reps, n * sizeof(TemperRec), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  for (uint32_t r = 0; r < t->n_T; r++)
    if (!(scale[r] > 0)) return fail(SADMC_ERR_INVALID, "translation scale %u is not positive", r);
  for (size_t k = 0; k < n; k++) reps[k].tscale = scale[k % t->n_T];
  CK(cudaMemcpyAsync(t->d_reps, reps.data(), n * sizeof(TemperRec), cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}
// resume (tempering.rs:196-213: the whole MC is deserialised): the inverse of sadmc_tempering_get_replicas / _get_rng / _num_moves;
// systems go back with sadmc_tempering_set_system
int sadmc_tempering_set_replicas(sadmc_tempering* t, uint32_t sim, const sadmc_replica_state* in) {
  if (!t || !in) return fail(SADMC_ERR_INVALID, "null argument");
  if (sim >= t->n_sim) return fail(SADMC_ERR_INVALID, "simulation %u out of range", sim);
  sadmc_engine* e = t->e;
  std::vector<TemperRec> reps(t->n_T);
  std::vector<WalkerRec> recs(t->n_T);
  CK(cudaMemcpyAsync(recs.data(), e->P.walkers + (size_t)sim * t->n_T, t->n_T * sizeof(WalkerRec), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  for (uint32_t r = 0; r < t->n_T; r++) {
    if (!(in[r].T > 0) || !(in[r].translation_scale > 0)) return fail(SADMC_ERR_INVALID, "replica %u: temperature and translation scale must be positive", r);
    TemperRec& q = reps[r];
    q.T = in[r].T;
    q.rejected = in[r].rejected_count;
    q.accepted = in[r].accepted_count;
    q.rejected_swap = in[r].rejected_swap_count;
    q.accepted_swap = in[r].accepted_swap_count;
    q.ignored = in[r].ignored_count;
    q.total_energy = in[r].total_energy;
    q.total_energy_squared = in[r].total_energy_squared;
    q.tscale = in[r].translation_scale;
    recs[r].s0 = in[r].rng_s0;
    recs[r].s1 = in[r].rng_s1;
  }
  CK(cudaMemcpyAsync(t->d_reps + (size_t)sim * t->n_T, reps.data(), t->n_T * sizeof(TemperRec), cudaMemcpyHostToDevice, e->stream));
  CK(cudaMemcpyAsync(e->P.walkers + (size_t)sim * t->n_T, recs.data(), t->n_T * sizeof(WalkerRec), cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}
int sadmc_tempering_set_rng(sadmc_tempering* t, uint32_t sim, const uint64_t s[2]) {
  if (!t || !s) return fail(SADMC_ERR_INVALID, "null argument");
  if (sim >= t->n_sim) return fail(SADMC_ERR_INVALID, "simulation %u out of range", sim);
  CK(cudaMemcpyAsync(t->d_mc_rng + 2 * (size_t)sim, s, 16, cudaMemcpyHostToDevice, t->e->stream));
  CK(cudaStreamSynchronize(t->e->stream));
  return 0;
}
int sadmc_tempering_set_num_moves(sadmc_tempering* t, uint64_t moves) {
  if (!t) return fail(SADMC_ERR_INVALID, "null argument");
  const unsigned long long per_round = t->steps * t->n_T;
  if (per_round == 0 || moves % per_round) return fail(SADMC_ERR_INVALID, "MC::moves of a checkpoint is a multiple of %llu (steps x replicas)", per_round);
  t->rounds = moves / per_round;
  return 0;
}
int sadmc_tempering_get_rng(sadmc_tempering* t, uint32_t sim, uint64_t s[2]) {
  if (!t || !s) return fail(SADMC_ERR_INVALID, "null argument");
  if (sim >= t->n_sim) return fail(SADMC_ERR_INVALID, "simulation %u out of range", sim);
  CK(cudaMemcpyAsync(s, t->d_mc_rng + 2 * (size_t)sim, 16, cudaMemcpyDeviceToHost, t->e->stream));
  CK(cudaStreamSynchronize(t->e->stream));
  return 0;
}
int sadmc_tempering_system_len(sadmc_tempering* t, size_t* n) {
  if (!t) return fail(SADMC_ERR_INVALID, "null argument");
  return sadmc_system_len(t->e, n);
}
int sadmc_tempering_get_system(sadmc_tempering* t, uint32_t sim, uint32_t replica, double* buf, size_t n) {
  if (!t) return fail(SADMC_ERR_INVALID, "null argument");
  if (sim >= t->n_sim || replica >= t->n_T) return fail(SADMC_ERR_INVALID, "replica (%u, %u) out of range", sim, replica);
  int src = temper_settle(t);
  if (src) return src;
  return sadmc_get_system(t->e, sim * t->n_T + replica, buf, n);
}
int sadmc_tempering_set_system(sadmc_tempering* t, uint32_t sim, uint32_t replica, const double* buf, size_t n) {
  if (!t) return fail(SADMC_ERR_INVALID, "null argument");
  if (sim >= t->n_sim || replica >= t->n_T) return fail(SADMC_ERR_INVALID, "replica (%u, %u) out of range", sim, replica);
  t->settled = false;
  return sadmc_set_system(t->e, sim * t->n_T + replica, buf, n);
}

// ---- energy-ceiling replicas: the `replicas` binary (src/mc/energy_replicas.rs; kernels in replicas.cuh / replicas_round.cuh) ----
struct sadmc_replicas {
  sadmc_engine* e = nullptr; // systems, generators, kernels (dummy bins; never started)
  uint32_t n_sim = 0, r_max = 0;
  unsigned long long steps = 0;
  double dimensionality = 1.0;
  ReplicaRec* d_reps = nullptr;
  ReplicaSim* d_sims = nullptr;
  double* d_energy = nullptr;
  double* d_median = nullptr;
  float last_ms = 0.f;
};
static double system_max_size(const sadmc_engine* e) { // MovableSystem::max_size
  const sadmc_config& c = e->cfg;
  switch (c.system) {
    case SADMC_SYS_LJ: return c.lj_radius;                                                                // lj.rs:375-377
    case SADMC_SYS_WCA:
    case SADMC_SYS_SW: return std::sqrt(e->P.box[0] * e->P.box[0] + e->P.box[1] * e->P.box[1] + e->P.box[2] * e->P.box[2]); // wca.rs:354-356
    case SADMC_SYS_TWO_WELLS: return 2.0;                                                                 // two_wells.rs:465-467
    default: return 0.5;                                                                                  // fake.rs:145, erfinv.rs:111, ising.rs:120
  }
}
static double system_dimensionality(const sadmc_config& c) { // System::dimensionality
  switch (c.system) {
    case SADMC_SYS_ISING: return (double)c.N * c.N;                                   // ising.rs:89-91
    case SADMC_SYS_FAKE: return (double)min_moves_to_randomize(c);                    // fake.rs:116-118
    case SADMC_SYS_TWO_WELLS: return (double)c.N;                                     // two_wells.rs:405-407
    default: return 3.0 * (double)c.N;                                                // lj.rs:283-285, wca.rs:271-273, erfinv.rs:89-91
  }
}
int sadmc_replicas_create(const sadmc_config* cfg, double min_T, uint64_t indep, uint32_t max_replicas, uint32_t max_init, sadmc_replicas** out) {
  if (!cfg || !out) return fail(SADMC_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->n_walkers < 1) return fail(SADMC_ERR_INVALID, "replicas needs at least one simulation");
  if (max_replicas < 2 || max_replicas > 256) return fail(SADMC_ERR_INVALID, "max_replicas must be 2..256");
  if (cfg->system == SADMC_SYS_SW || cfg->system == SADMC_SYS_TWO_WELLS)
    return fail(SADMC_ERR_UNSUPPORTED, "replicas needs System::randomize: todo!() for the square well in the reference (optsquare.rs:202-204), not restated for two-wells");
  if ((unsigned long long)max_replicas * cfg->n_walkers > 0x7fffffffull) return fail(SADMC_ERR_INVALID, "too many replica slots");
  if (max_init == 0) max_init = 1u << 15; // MAX_INIT, energy_replicas.rs:350
  sadmc_config c = *cfg;
  const uint32_t n_sim = cfg->n_walkers;
  c.n_walkers = n_sim * max_replicas;
  c.method = SADMC_METHOD_CANONICAL; // unused: dummy bins
  c.canonical_T = 1.0;
  c.energy_bin = 1.0;
  c.min_allowed_energy = c.max_allowed_energy = NAN;
  c.bin_window_lo = 0.0;
  c.bin_window_hi = 1.0;
  c.flags &= ~(uint32_t)(SADMC_FLAG_BINNING | SADMC_FLAG_BINNING_LINEAR);
  c.high_resolution_de = NAN;
  if (c.system == SADMC_SYS_LJ && c.lanes_per_walker == 0) c.lanes_per_walker = 1;
  if (c.system == SADMC_SYS_WCA && c.lanes_per_walker == 0) c.lanes_per_walker = 32;
  c.init_mode = SADMC_INIT_EXTERNAL;
  sadmc_engine* e = nullptr;
  int rc = sadmc_create(&c, &e);
  if (rc) return rc;
  if (!e->ks.replica_move || !e->ks.replica_init) {
    sadmc_destroy(e);
    return fail(SADMC_ERR_UNSUPPORTED, "no replicas kernel for this system / lanes_per_walker");
  }
  sadmc_replicas* z = new sadmc_replicas;
  z->e = e;
  z->n_sim = n_sim;
  z->r_max = max_replicas;
  z->steps = min_moves_to_randomize(*cfg); // energy_replicas.rs:507
  z->dimensionality = system_dimensionality(*cfg);
#define ZBAIL(expr)              \
  do {                           \
    int _rc = (expr);            \
    if (_rc) {                   \
      sadmc_replicas_destroy(z); \
      return _rc;                \
    }                            \
  } while (0)
#define ZCK(call)                                                                      \
  do {                                                                                 \
    cudaError_t _e = (call);                                                           \
    if (_e != cudaSuccess) {                                                           \
      sadmc_replicas_destroy(z);                                                       \
      return fail(SADMC_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(_e));     \
    }                                                                                  \
  } while (0)
  { // every slot starts as the constructor's system (only slots 0 and 1 of each simulation are in use at first)
    std::vector<double> img;
    e->cfg.init_mode = SADMC_INIT_REFERENCE;
    rc = reference_image(e, img);
    e->cfg.init_mode = SADMC_INIT_EXTERNAL;
    ZBAIL(rc);
    std::vector<double> all((size_t)c.n_walkers * e->sys_len);
    for (uint32_t w = 0; w < c.n_walkers; w++) memcpy(&all[(size_t)w * e->sys_len], img.data(), e->sys_len * sizeof(double));
    ZBAIL(sadmc_set_systems(e, all.data(), all.size()));
  }
  const size_t n_slots = c.n_walkers;
  double* d_init = nullptr;
  unsigned long long* d_rng = nullptr;
  ZBAIL(dev_alloc(e, (void**)&z->d_reps, n_slots * sizeof(ReplicaRec), true));
  ZBAIL(dev_alloc(e, (void**)&z->d_sims, (size_t)n_sim * sizeof(ReplicaSim), true));
  ZBAIL(dev_alloc(e, (void**)&z->d_energy, n_slots * 8, true));
  ZBAIL(dev_alloc(e, (void**)&z->d_median, (size_t)n_sim * REPLICA_ESTIMATOR_SIZE * 8, true));
  ZBAIL(dev_alloc(e, (void**)&d_init, (size_t)n_sim * max_init * 8, false));
  ZBAIL(dev_alloc(e, (void**)&d_rng, (size_t)n_sim * 16, false));
  if (e->ks.smem > 48 * 1024) {
    ZCK(cudaFuncSetAttribute((const void*)e->ks.replica_init, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->ks.smem));
    ZCK(cudaFuncSetAttribute((const void*)e->ks.replica_move, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->ks.smem));
  }
  {
    const long long threads = (long long)n_sim * e->ks.G;
    const int grid = (int)((threads + e->ks.block - 1) / e->ks.block);
    e->ks.replica_init<<<grid, e->ks.block, e->ks.smem, e->stream>>>(e->P, cfg->seed + cfg->walker_offset, n_sim, max_replicas, d_init, max_init, d_rng);
    ZCK(cudaGetLastError());
    e->launches++;
  }
  std::vector<double> en((size_t)n_sim * max_init);
  std::vector<unsigned long long> mc((size_t)n_sim * 2);
  ZCK(cudaMemcpyAsync(en.data(), d_init, en.size() * 8, cudaMemcpyDeviceToHost, e->stream));
  ZCK(cudaMemcpyAsync(mc.data(), d_rng, mc.size() * 8, cudaMemcpyDeviceToHost, e->stream));
  ZCK(cudaStreamSynchronize(e->stream));
  std::vector<ReplicaRec> reps(n_slots);
  std::vector<ReplicaSim> sims(n_sim);
  std::vector<double> med((size_t)n_sim * REPLICA_ESTIMATOR_SIZE, 0.0);
  const double max_size = system_max_size(e);
  for (uint32_t k = 0; k < n_sim; k++) { // energy_replicas.rs:369-398
    double* a = &en[(size_t)k * max_init];
    std::sort(a, a + max_init);
    const double half = a[max_init / 2], quarter = a[max_init / 4];
    memset(&reps[(size_t)k * max_replicas], 0, sizeof(ReplicaRec) * max_replicas);
    ReplicaRec& r0 = reps[(size_t)k * max_replicas];
    ReplicaRec& r1 = reps[(size_t)k * max_replicas + 1];
    r0.max_energy = INFINITY;
    r0.cutoff = half;
    r1.max_energy = half;
    r1.cutoff = quarter;
    for (ReplicaRec* r : {&r0, &r1}) { // Replica::new, 147-170
      r->lowest_max = r->max_energy;
      r->tscale = max_size;
      r->unique_visitors = 1;
      r->collecting = 1;
    }
    ReplicaSim& S = sims[k];
    memset(&S, 0, sizeof S);
    S.s0 = mc[2 * k];
    S.s1 = mc[2 * k + 1];
    S.indep = indep;
    S.min_T = min_T;
    S.n_rep = 2;
    S.median_len = 1;
    med[(size_t)k * REPLICA_ESTIMATOR_SIZE] = quarter; // MedianEstimator::new(energies[len / 4])
  }
  ZCK(cudaMemcpyAsync(z->d_reps, reps.data(), reps.size() * sizeof(ReplicaRec), cudaMemcpyHostToDevice, e->stream));
  ZCK(cudaMemcpyAsync(z->d_sims, sims.data(), sims.size() * sizeof(ReplicaSim), cudaMemcpyHostToDevice, e->stream));
  ZCK(cudaMemcpyAsync(z->d_median, med.data(), med.size() * 8, cudaMemcpyHostToDevice, e->stream));
  ZCK(cudaStreamSynchronize(e->stream));
#undef ZBAIL
#undef ZCK
  *out = z;
  return 0;
}
void sadmc_replicas_destroy(sadmc_replicas* z) {
  if (!z) return;
  if (z->e) sadmc_destroy(z->e);
  delete z;
}
int sadmc_replicas_run(sadmc_replicas* z, uint64_t n_rounds) {
  if (!z) return fail(SADMC_ERR_INVALID, "null argument");
  sadmc_engine* e = z->e;
  CK(cudaSetDevice(e->cfg.device));
  int grid;
  launch_cfg(e, &grid);
  CK(cudaEventRecord(e->ev0, e->stream));
  for (uint64_t r = 0; r < n_rounds; r++) {
    e->ks.replica_move<<<grid, e->ks.block, e->ks.smem, e->stream>>>(e->P, z->d_reps, z->d_sims, z->r_max, z->steps, z->d_energy);
    CK(cudaGetLastError());
    replica_round_kernel<<<z->n_sim, 128, 0, e->stream>>>(e->P, z->d_reps, z->d_sims, z->d_energy, z->d_median, z->r_max, z->steps, z->dimensionality);
    CK(cudaGetLastError());
    e->launches += 2;
  }
  CK(cudaEventRecord(e->ev1, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaEventElapsedTime(&z->last_ms, e->ev0, e->ev1));
  std::vector<ReplicaSim> sims(z->n_sim);
  CK(cudaMemcpy(sims.data(), z->d_sims, sims.size() * sizeof(ReplicaSim), cudaMemcpyDeviceToHost));
  uint32_t over = 0;
  for (const ReplicaSim& S : sims) over += S.overflow ? 1u : 0u;
  if (over) return fail(SADMC_ERR_WINDOW, "%u simulation(s) wanted to split off a replica with all %u slots in use: raise max_replicas", over, z->r_max);
  return 0;
}
int sadmc_replicas_last_run_ms(sadmc_replicas* z, float* ms) {
  if (!z || !ms) return fail(SADMC_ERR_INVALID, "null argument");
  *ms = z->last_ms;
  return 0;
}
static int fetch_sim(sadmc_replicas* z, uint32_t sim, ReplicaSim* S) {
  if (!z) return fail(SADMC_ERR_INVALID, "null argument");
  if (sim >= z->n_sim) return fail(SADMC_ERR_INVALID, "simulation %u out of range", sim);
  CK(cudaMemcpyAsync(S, z->d_sims + sim, sizeof(ReplicaSim), cudaMemcpyDeviceToHost, z->e->stream));
  CK(cudaStreamSynchronize(z->e->stream));
  return 0;
}
int sadmc_replicas_num_moves(sadmc_replicas* z, uint32_t sim, uint64_t* moves) {
  ReplicaSim S;
  int rc = fetch_sim(z, sim, &S);
  if (!rc && moves) *moves = S.moves;
  return rc;
}
int sadmc_replicas_num_replicas(sadmc_replicas* z, uint32_t sim, uint32_t* n) {
  ReplicaSim S;
  int rc = fetch_sim(z, sim, &S);
  if (!rc && n) *n = (uint32_t)S.n_rep;
  return rc;
}
int sadmc_replicas_get_rng(sadmc_replicas* z, uint32_t sim, uint64_t s[2]) {
  ReplicaSim S;
  int rc = fetch_sim(z, sim, &S);
  if (!rc && s) {
    s[0] = S.s0;
    s[1] = S.s1;
  }
  return rc;
}
int sadmc_replicas_get_median(sadmc_replicas* z, uint32_t sim, uint32_t cap, double* energies, uint32_t* len) {
  ReplicaSim S;
  int rc = fetch_sim(z, sim, &S);
  if (rc) return rc;
  if (len) *len = (uint32_t)S.median_len;
  if (energies) {
    if (cap < (uint32_t)S.median_len) return fail(SADMC_ERR_INVALID, "capacity %u < %d energies", cap, S.median_len);
    CK(cudaMemcpyAsync(energies, z->d_median + (size_t)sim * REPLICA_ESTIMATOR_SIZE, (size_t)S.median_len * 8, cudaMemcpyDeviceToHost, z->e->stream));
    CK(cudaStreamSynchronize(z->e->stream));
  }
  return 0;
}
int sadmc_replicas_get_replicas(sadmc_replicas* z, uint32_t sim, uint32_t cap, sadmc_zeno_replica_state* out) {
  ReplicaSim S;
  int rc = fetch_sim(z, sim, &S);
  if (rc) return rc;
  if (!out) return fail(SADMC_ERR_INVALID, "null argument");
  if (cap < (uint32_t)S.n_rep) return fail(SADMC_ERR_INVALID, "capacity %u < %d replicas", cap, S.n_rep);
  sadmc_engine* e = z->e;
  const size_t base = (size_t)sim * z->r_max;
  std::vector<ReplicaRec> reps(S.n_rep);
  std::vector<WalkerRec> recs(S.n_rep);
  CK(cudaMemcpyAsync(reps.data(), z->d_reps + base, reps.size() * sizeof(ReplicaRec), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaMemcpyAsync(recs.data(), e->P.walkers + base, recs.size() * sizeof(WalkerRec), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  for (int r = 0; r < S.n_rep; r++) {
    sadmc_zeno_replica_state& o = out[r];
    memset(&o, 0, sizeof o);
    const ReplicaRec& q = reps[r];
    o.max_energy = q.max_energy;
    o.cutoff_energy = q.cutoff;
    o.lowest_max_energy = q.lowest_max;
    o.translation_scale = q.tscale;
    o.rejected_count = q.rejected;
    o.accepted_count = q.accepted;
    o.above_count = q.above_count;
    o.below_count = q.below_count;
    o.upwelling_count = q.upwelling;
    o.unique_visitors = q.unique_visitors;
    o.above_total = q.above_total;
    o.below_total = q.below_total;
    o.above_total_squared = q.above_sq;
    o.below_total_squared = q.below_sq;
    o.above_extra_total = q.xtot;
    o.above_extra_count = q.xcnt;
    o.collecting_data = q.collecting;
    o.rng_s0 = recs[r].s0;
    o.rng_s1 = recs[r].s1;
    o.energy = recs[r].E;
  }
  return 0;
}
int sadmc_replicas_system_len(sadmc_replicas* z, size_t* n) {
  if (!z) return fail(SADMC_ERR_INVALID, "null argument");
  return sadmc_system_len(z->e, n);
}
int sadmc_replicas_get_system(sadmc_replicas* z, uint32_t sim, uint32_t replica, double* buf, size_t n) {
  if (!z) return fail(SADMC_ERR_INVALID, "null argument");
  if (sim >= z->n_sim || replica >= z->r_max) return fail(SADMC_ERR_INVALID, "replica (%u, %u) out of range", sim, replica);
  return sadmc_get_system(z->e, sim * z->r_max + replica, buf, n);
}

// ---- trait shims -------------------------------------------------------------
static int run_shim(sadmc_engine* e, uint32_t w, int op, double arg, ShimOut* o) {
  if (!e) return fail(SADMC_ERR_INVALID, "null engine");
  if (w >= e->P.n_walkers) return fail(SADMC_ERR_INVALID, "walker %u out of range", w);
  CK(cudaSetDevice(e->cfg.device));
  e->ks.shim<<<1, e->ks.block, e->ks.smem, e->stream>>>(e->P, w, op, arg, e->d_shim, e->d_pending);
  e->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(o, e->d_shim, sizeof(ShimOut), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}
int sadmc_sys_energy(sadmc_engine* e, uint32_t w, double* energy) {
  ShimOut o;
  int rc = run_shim(e, w, OP_ENERGY, 0, &o);
  if (!rc) *energy = o.value;
  return rc;
}
int sadmc_sys_compute_energy(sadmc_engine* e, uint32_t w, double* energy) {
  ShimOut o;
  int rc = run_shim(e, w, OP_COMPUTE_ENERGY, 0, &o);
  if (!rc) *energy = o.value;
  return rc;
}
int sadmc_sys_plan_move(sadmc_engine* e, uint32_t w, double mean_distance, int* some, double* e_new) {
  ShimOut o;
  int rc = run_shim(e, w, OP_PLAN_MOVE, mean_distance, &o);
  if (!rc) {
    *some = o.some;
    *e_new = o.value;
  }
  return rc;
}
int sadmc_sys_randomize(sadmc_engine* e, uint32_t w, double* energy) {
  ShimOut o;
  int rc = run_shim(e, w, OP_RANDOMIZE, 0, &o);
  if (!rc && energy) *energy = o.value;
  return rc;
}
int sadmc_sys_confirm(sadmc_engine* e, uint32_t w) {
  ShimOut o;
  return run_shim(e, w, OP_CONFIRM, 0, &o);
}
int sadmc_sys_verify_energy(sadmc_engine* e, uint32_t w) {
  ShimOut o;
  int rc = run_shim(e, w, OP_VERIFY, 0, &o);
  if (rc) return rc;
  return o.ok ? 0 : fail(SADMC_ERR_VERIFY, "verify_energy failed for walker %u", w);
}

} // extern "C"
