// sys_ising.cuh -- 2-D periodic Ising lattice, one thread per walker.
//
// Device form of `Ising` (src/system/ising.rs): plan_move 104-119, confirm 95-100,
// compute_energy 59-75, randomize 79-85.  The reference keeps `Vec<i8>` spins and
// an f64 energy; every value involved is an integer, so the lattice is bit-packed
// (bit = 1 <=> spin +1; site index i + j*N as in the reference) and dE is integer
// arithmetic -- results are identical, not merely close.  The packed lattice of
// each walker lives in shared memory, word-interleaved across the block's threads
// so that 32 walkers reading 32 different random words never bank-conflict.
#pragma once
#include "book.cuh"
#include "rng.cuh"

// resident CTAs per SM the one-thread-per-walker kernels of the small systems are compiled for (register cap = 65536 / (128 * this))
#ifndef SADMC_SMALL_MIN_BLOCKS
#define SADMC_SMALL_MIN_BLOCKS 4
#endif

namespace sadmc {

struct IsingSys {
  static constexpr int G = 1;
  static constexpr bool FAST_BOOK = false;
  static constexpr int BLOCK = 128;
  static constexpr int MIN_BLOCKS = SADMC_SMALL_MIN_BLOCKS;
  static constexpr bool COOP = false;
  __device__ __forceinline__ void set_cooperative(bool) {}
  __device__ __forceinline__ void finish_move() {}
  uint32_t* sp; // this thread's words: sp[k * stride]
  int stride, N, words;
  int E, ch_site, ch_e;
  unsigned long long zone;

  static __host__ __device__ size_t smem_bytes(const DevParams& P, int block) { return (size_t)P.ising_words * block * sizeof(uint32_t); }

  __device__ IsingSys(const DevParams& P, uint32_t, int, unsigned, unsigned char* smem)
      : sp(reinterpret_cast<uint32_t*>(smem) + threadIdx.x), stride(blockDim.x), N((int)P.N), words((int)P.ising_words), zone(P.zone_a) {}

  __device__ void load(const DevParams& P, uint32_t w, const WalkerRec& r) {
    const uint32_t* g = P.sys_words + (size_t)w * words;
    for (int k = 0; k < words; k++) sp[k * stride] = g[k];
    E = (int)r.E;
    ch_site = 0;
    ch_e = E;
  }
  __device__ void store(const DevParams& P, uint32_t w, WalkerRec& r, bool) {
    uint32_t* g = P.sys_words + (size_t)w * words;
    for (int k = 0; k < words; k++) g[k] = sp[k * stride];
    r.E = (double)E;
    r.err = 0.0;
  }
  __device__ __forceinline__ int spin(int site) const { return (int)((sp[(site >> 5) * stride] >> (site & 31)) & 1u) * 2 - 1; }
  __device__ __forceinline__ double energy() const { return (double)E; }

  __device__ __forceinline__ bool plan_move(Rng& rng, double, const double*, const double*, double& e2) {
    const int i = (int)rng.below((uint32_t)N, zone); // gen_range(0, N), ising.rs:105
    const int j = (int)rng.below((uint32_t)N, zone); // ising.rs:106
    const int jp = j + 1 == N ? 0 : j + 1, jm = j == 0 ? N - 1 : j - 1;
    const int ip = i + 1 == N ? 0 : i + 1, im = i == 0 ? N - 1 : i - 1;
    const int nt = spin(i + jp * N) + spin(i + jm * N) + spin(ip + j * N) + spin(im + j * N);
    ch_site = i + j * N;
    ch_e = E - nt * spin(ch_site) * 2; // ising.rs:116
    e2 = (double)ch_e;
    return true;
  }
  __device__ __forceinline__ void confirm() {
    sp[(ch_site >> 5) * stride] ^= 1u << (ch_site & 31);
    E = ch_e;
  }
  __device__ double compute_energy() const { // ising.rs:59-75
    int e = 0;
    for (int i1 = 0; i1 < N; i1++)
      for (int j1 = 0; j1 < N; j1++) {
        const int j2 = j1 + 1 == N ? 0 : j1 + 1;
        const int i2 = i1 + 1 == N ? 0 : i1 + 1;
        e += (spin(i1 + j2 * N) + spin(i2 + j1 * N)) * spin(i1 + j1 * N);
      }
    return (double)e;
  }
  __device__ double randomize(Rng& rng) { // ising.rs:79-85: bit 0 of successive next_u64
    for (int k = 0; k < words; k++) sp[k * stride] = 0;
    for (int s = 0; s < N * N; s++)
      if (rng.next() & 1ull) sp[(s >> 5) * stride] |= 1u << (s & 31);
    E = (int)compute_energy();
    return (double)E;
  }
  __device__ bool verify_energy() const { return true; } // trait default (mod.rs:85)
  __device__ __forceinline__ bool extra(unsigned long long, double&) const { return false; }
  // pending change across the trait shims (possible_change, ising.rs:28)
  __device__ void get_pending(double* p, bool writer, bool some) const {
    if (!writer || !some) return;
    p[0] = 1.0;
    p[1] = (double)ch_site;
    p[5] = (double)ch_e;
  }
  __device__ bool set_pending(const double* p) {
    if (p[0] == 0.0) return false;
    ch_site = (int)p[1];
    ch_e = (int)p[5];
    return true;
  }
};

} // namespace sadmc
