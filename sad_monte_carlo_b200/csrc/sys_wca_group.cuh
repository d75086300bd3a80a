// sys_wca_group.cuh -- periodic WCA fluid on a cell list, G = 4, 8 or 16 lanes per walker.
//
// Device form of `Wca` (src/system/wca.rs): move_atom 119-140, potential 66-76, potential_pressure 79-92,
// plan_move 342-353, confirm 277-289, set_energy 164-177, expected_accuracy 178-180, compute_energy 222-230,
// verify_energy 237-251, data_to_collect (pressure) 202-218, randomize 252-270; and of `Cell`
// (src/system/optcell.rs): get_subcell 112-124, add_to_subcells 131-160 (image offsets), put_in_cell 277-309.
//
// Why not a warp per walker (sys_cell_fluid.cuh, still used for the square well): a move touches ~30 candidate
// atoms in 27 subcells -- one or two distance tests per lane -- and then runs a scalar tail (xoroshiro, three
// ziggurat draws, bin record, accept test, bookkeeping) that a warp executes once per WALKER.  With G lanes per walker
// a warp carries 32 / G walkers through that tail together, and the candidate loop still fills the lanes (lane l walks
// subcells l, l + G, ...).  What bounds the number of resident walkers is shared memory (positions, 24 N bytes + the
// linked cell lists), so the ziggurat tables are read from global memory (ZIG_GLOBAL) and the `cell_of` array of the
// warp kernel is gone (an atom's subcell is recomputed from its position when it moves).
//
// Each atom sits in ONE list (head[cell], next[atom], 16-bit) and a lookup visits the 27 neighbouring cells with the
// image offset the reference would have stored (optcell.rs:135-157): -L when the neighbour wrapped below 0, +L when it
// wrapped past the last subcell.  When the old and the new position share a subcell (the usual case: steps are a few
// percent of a subcell) ONE walk over the 27 lists yields both sums of wca.rs:124-133.
//
// Arithmetic tiers (both are tolerance tier, <= 1e-12 relative per move: lane order instead of list order):
//   default   the reference's operations (IEEE divide, no FMA) and its `set_energy` error budget, which recomputes the
//             whole energy every ~10 accepted moves (wca.rs:164-177: 1e-14 |E| N^2 per move against 1e-13 |E| N^2).
//             That sum runs over half of the neighbour shell (each pair once), lane-parallel over atoms.
//   FAST      (SADMC_FLAG_FAST_MATH) FMA + Newton reciprocal, subcell indices by multiplication, and the same error
//             budget with a 2^16 / 10 times larger allowance: the accumulated energy is re-summed every 65 536 accepted
//             moves instead of every 10 (rounding drift in between <= 2^8 ulp, i.e. < 1e-13 relative; tested).
#pragma once
#include "book.cuh"
#include "rng.cuh"

namespace sadmc {

template <int G_, bool FAST>
struct WcaGroupSys {
  static_assert(G_ == 4 || G_ == 8 || G_ == 16, "lanes per walker");
  static constexpr int G = G_;
  static constexpr bool FAST_BOOK = FAST;
#ifndef SADMC_WCA_WALKERS_PER_BLOCK
#define SADMC_WCA_WALKERS_PER_BLOCK 8
#endif
  static constexpr int WALKERS_PER_BLOCK = SADMC_WCA_WALKERS_PER_BLOCK;
  static constexpr int BLOCK = WALKERS_PER_BLOCK * G_;
#ifndef SADMC_WCA_MIN_BLOCKS
#define SADMC_WCA_MIN_BLOCKS (G_ == 16 ? 3 : 4)
#endif
  static constexpr int MIN_BLOCKS = SADMC_WCA_MIN_BLOCKS;
  static constexpr bool COOP = false;
  static constexpr bool VERIFIES = true; // wca.rs:237-251
  static constexpr bool ZIG_GLOBAL = true;
  static constexpr bool HAS_EXTRA = true; // the pressure, wca.rs:202-218
  // how much larger than the reference's the error allowance is before the energy is re-summed (FAST)
  static constexpr double RELAX = FAST ? 6553.6 : 1.0;

  __device__ __forceinline__ void set_cooperative(bool) {}
  __device__ __forceinline__ void finish_move() {}

  double *px, *py, *pz;
  short *next, *head;
  int N, lig, ncx, ncy, ncz, my_ncells;
  unsigned long long my_cells; // this lane's neighbours lig, lig + G, ...: (dx+1) | (dy+1) << 2 | (dz+1) << 4, 6 bits each (up to 7)
  unsigned gmask;
  double Lx, Ly, Lz, sx, sy, sz, rc2; // s* = n* / L* (FAST subcell index)
  double E, err;
  unsigned long long zone, x_at;
  // pending change
  int ch_which, ch_cold, ch_cnew;
  double tx, ty, tz, ch_e, ch_dabse;

  static __host__ __device__ size_t walker_bytes(uint32_t N, int ncells) {
    size_t b = (size_t)3 * N * sizeof(double) + (size_t)N * sizeof(short) + (size_t)ncells * sizeof(short);
    return (b + 15) & ~(size_t)15;
  }
  static __host__ __device__ size_t smem_bytes(const DevParams& P, int block) {
    return walker_bytes(P.N, P.ncell[0] * P.ncell[1] * P.ncell[2]) * (size_t)(block / G_);
  }

  __device__ WcaGroupSys(const DevParams& P, uint32_t, int lane_in_group, unsigned group_mask_, unsigned char* smem)
      : N((int)P.N), lig(lane_in_group), ncx(P.ncell[0]), ncy(P.ncell[1]), ncz(P.ncell[2]), gmask(group_mask_), Lx(P.box[0]), Ly(P.box[1]),
        Lz(P.box[2]), rc2(P.r_cut2), zone(P.zone_b), x_at(0), ch_which(-1) {
    sx = (double)ncx / Lx;
    sy = (double)ncy / Ly;
    sz = (double)ncz / Lz;
    unsigned char* base = smem + walker_bytes(P.N, ncx * ncy * ncz) * (threadIdx.x / G_);
    px = reinterpret_cast<double*>(base);
    py = px + N;
    pz = py + N;
    next = reinterpret_cast<short*>(pz + N);
    head = next + N;
    my_cells = 0;
    my_ncells = 0;
    for (int k = lig; k < 27; k += G_) {
      my_cells |= (unsigned long long)((k / 9) | (((k / 3) % 3) << 2) | ((k % 3) << 4)) << (6 * my_ncells);
      my_ncells++;
    }
  }
  __device__ __forceinline__ void gsync() const { __syncwarp(gmask); }
  __device__ __forceinline__ int ncells() const { return ncx * ncy * ncz; }

  // optcell.rs:112-124.  The index only has to be the SAME function wherever it is used (linking and lookup): a
  // position within rounding of a subcell face may land on either side, and a pair closer than the cutoff is still
  // found because subcells are at least one cutoff wide.
  static __device__ __forceinline__ int cell1(double x, double L, double s, int n) {
    int c = (int)floor(FAST ? x * s : x / L * (double)n);
    return c >= n ? n - 1 : (c < 0 ? 0 : c);
  }
  __device__ __forceinline__ void subcell(double x, double y, double z, int& cx, int& cy, int& cz) const {
    cx = cell1(x, Lx, sx, ncx);
    cy = cell1(y, Ly, sy, ncy);
    cz = cell1(z, Lz, sz, ncz);
  }
  __device__ __forceinline__ int flat(int cx, int cy, int cz) const { return (cx * ncy + cy) * ncz + cz; }
  static __device__ __forceinline__ double wrap1(double v, double L) { // optcell.rs:277-309
    if (v < 0.0) {
      do {
        v += L;
      } while (v < 0.0);
    } else {
      while (v >= L) v -= L;
    }
    return v;
  }
  __device__ __forceinline__ double potential(double r2) const { // wca.rs:66-76
    if (FAST) {
      // no branch: most candidates of a lookup lie beyond the cutoff, but with four walkers in a warp some lane is
      // nearly always inside, so the branch only added divergence.  The argument is clamped so that two overlapping
      // atoms of a random start (r^2 ~ 1e-6) still give a finite number where the reference does.
      const double s = rcp_newton(r2 > 1e-300 ? r2 : 1e-300);
      const double s3 = s * s * s;
      const double u = fma(4.0, fma(s3, s3, -s3), 1.0);
      return r2 < rc2 ? u : 0.0;
    }
    if (r2 < rc2) {
      const double s = 1.0 / r2;
      const double s3 = s * s * s;
      return 4.0 * (s3 * s3 - s3) + 1.0;
    }
    return 0.0;
  }
  __device__ __forceinline__ double pressure(double r2) const { // wca.rs:79-92
    if (r2 < rc2) {
      const double s = FAST ? rcp_newton(r2) : 1.0 / r2;
      const double s3 = s * s * s;
      return 4.0 * 3.0 * (2.0 * (s3 * s3) - s3);
    }
    return 0.0;
  }
  __device__ __forceinline__ double group_sum(double v) const {
#pragma unroll
    for (int off = G / 2; off >= 1; off >>= 1) v += __shfl_xor_sync(gmask, v, off, G);
    return v;
  }

  // neighbour subcell number k (0..26) of (cx, cy, cz): its list and the image shift of its atoms (optcell.rs:135-157:
  // -L when the neighbour wrapped below 0, +L when it wrapped past the last subcell).  Branch-free.
  __device__ __forceinline__ int neighbour_d(int dx, int dy, int dz, int cx, int cy, int cz, double& shx, double& shy, double& shz) const {
    int qx = cx + dx, qy = cy + dy, qz = cz + dz;
    const bool xl = qx < 0, xh = qx >= ncx, yl = qy < 0, yh = qy >= ncy, zl = qz < 0, zh = qz >= ncz;
    qx += xl ? ncx : (xh ? -ncx : 0);
    qy += yl ? ncy : (yh ? -ncy : 0);
    qz += zl ? ncz : (zh ? -ncz : 0);
    shx = xl ? -Lx : (xh ? Lx : 0.0);
    shy = yl ? -Ly : (yh ? Ly : 0.0);
    shz = zl ? -Lz : (zh ? Lz : 0.0);
    return flat(qx, qy, qz);
  }
  __device__ __forceinline__ int neighbour(int k, int cx, int cy, int cz, double& shx, double& shy, double& shz) const {
    return neighbour_d(k / 9 - 1, (k / 3) % 3 - 1, k % 3 - 1, cx, cy, cz, shx, shy, shz);
  }
  // Every candidate the reference's `maybe_interacting_atoms_excluding(r, exclude)` returns for a position in subcell
  // (cx, cy, cz) (optcell.rs:93-110): f(j, image position).  The group's lanes share the 27 lists: lane l walks
  // neighbours l, l + G, ... whose offsets it keeps packed in `my_cells` (2 bits per axis, 6 bits per neighbour).
  // Two of the lane's lists are walked side by side: the two pointer chases (head -> next -> next, ~30 cycles of
  // shared-memory latency per hop) overlap, and the loop runs max(len_a, len_b) times instead of len_a + len_b.  A
  // list that has run out keeps feeding atom 0 with skip = true.
  template <class F>
  __device__ __forceinline__ void visit(int cx, int cy, int cz, int exclude, F&& f) const {
    unsigned long long cells = my_cells;
#pragma unroll 1
    for (int r = 0; r < my_ncells; r += 2, cells >>= 12) {
      double ax, ay, az, bx, by, bz;
      const int qa = neighbour_d((int)(cells & 3ull) - 1, (int)((cells >> 2) & 3ull) - 1, (int)((cells >> 4) & 3ull) - 1, cx, cy, cz, ax, ay, az);
      const int qb = neighbour_d((int)((cells >> 6) & 3ull) - 1, (int)((cells >> 8) & 3ull) - 1, (int)((cells >> 10) & 3ull) - 1, cx, cy, cz, bx, by, bz);
      int ja = head[qa];
      int jb = r + 1 < my_ncells ? head[qb] : -1; // (bits past the lane's last neighbour decode to offset -1: a valid cell, not used)
      while ((ja & jb) >= 0 || ja >= 0 || jb >= 0) {
        const int ia = ja >= 0 ? ja : 0, ib = jb >= 0 ? jb : 0;
        const double xa = px[ia] + ax, ya = py[ia] + ay, za = pz[ia] + az;
        const double xb = px[ib] + bx, yb = py[ib] + by, zb = pz[ib] + bz;
        const int na = next[ia], nb = next[ib];
        f(ja < 0 || ja == exclude, xa, ya, za);
        f(jb < 0 || jb == exclude, xb, yb, zb);
        ja = ja >= 0 ? na : -1;
        jb = jb >= 0 ? nb : -1;
      }
    }
  }

  __device__ void link_all(int natoms) { // optcell.rs:63-73 (update_caches); one lane, once per launch
    gsync();
    for (int c = lig; c < ncells(); c += G) head[c] = -1;
    gsync();
    if (lig == 0) {
      for (int a = natoms - 1; a >= 0; a--) {
        int cx, cy, cz;
        subcell(px[a], py[a], pz[a], cx, cy, cz);
        const int c = flat(cx, cy, cz);
        next[a] = head[c];
        head[c] = (short)a;
      }
    }
    gsync();
  }
  __device__ void load(const DevParams& P, uint32_t w, const WalkerRec& r) {
    const double* g = P.sys + (size_t)w * P.sys_stride;
    for (int a = lig; a < N; a += G) {
      px[a] = g[3 * a];
      py[a] = g[3 * a + 1];
      pz[a] = g[3 * a + 2];
    }
    E = r.E;
    err = r.err;
    link_all(N);
    if (E != E) { // NaN in the image: "compute it" (host-side constructors that do not know the energy)
      E = compute_energy();
      err = 0.0;
    }
  }
  __device__ void store(const DevParams& P, uint32_t w, WalkerRec& r, bool writer) {
    gsync();
    double* g = P.sys + (size_t)w * P.sys_stride;
    for (int a = lig; a < N; a += G) {
      g[3 * a] = px[a];
      g[3 * a + 1] = py[a];
      g[3 * a + 2] = pz[a];
    }
    if (writer) {
      g[3 * N] = E;
      g[3 * N + 1] = err;
      r.E = E;
      r.err = err;
    }
  }
  __device__ __forceinline__ double energy() const { return E; }

  __device__ __forceinline__ bool plan_move(Rng& rng, double scale, const double* zx, const double* zf, double& e2) {
    const int which = (int)rng.below((uint32_t)N, zone); // Uniform::new(0, N), wca.rs:345
    double vx, vy, vz;
    rng.normal3(zx, zf, vx, vy, vz); // rng.rs:111-117: three StandardNormal draws
    const double fx = px[which], fy = py[which], fz = pz[which];
    tx = wrap1(fx + vx * scale, Lx); // put_in_cell(pos + vector * mean_distance), wca.rs:346-351
    ty = wrap1(fy + vy * scale, Ly);
    tz = wrap1(fz + vz * scale, Lz);
    int cx, cy, cz, ox, oy, oz;
    subcell(tx, ty, tz, cx, cy, cz);
    subcell(fx, fy, fz, ox, oy, oz);
    ch_which = which;
    ch_cnew = flat(cx, cy, cz);
    ch_cold = flat(ox, oy, oz);
    double snew = 0.0, sold = 0.0;
    const double ttx = tx, tty = ty, ttz = tz;
    if (ch_cnew == ch_cold) { // one walk over the 27 lists serves both sums
      visit(cx, cy, cz, which, [&](bool skip, double ix, double iy, double iz) {
        const double ax = ix - ttx, ay = iy - tty, az = iz - ttz;
        const double bx = ix - fx, by = iy - fy, bz = iz - fz;
        const double rn = FAST ? fma(az, az, fma(ay, ay, ax * ax)) : ax * ax + ay * ay + az * az;
        const double ro = FAST ? fma(bz, bz, fma(by, by, bx * bx)) : bx * bx + by * by + bz * bz;
        snew += potential(skip ? 1e300 : rn); // the moved atom itself is not a neighbour (`_excluding`)
        sold += potential(skip ? 1e300 : ro);
      });
    } else {
      visit(cx, cy, cz, which, [&](bool skip, double ix, double iy, double iz) {
        const double ax = ix - ttx, ay = iy - tty, az = iz - ttz;
        snew += potential(skip ? 1e300 : (FAST ? fma(az, az, fma(ay, ay, ax * ax)) : ax * ax + ay * ay + az * az));
      });
      visit(ox, oy, oz, which, [&](bool skip, double ix, double iy, double iz) {
        const double bx = ix - fx, by = iy - fy, bz = iz - fz;
        sold += potential(skip ? 1e300 : (FAST ? fma(bz, bz, fma(by, by, bx * bx)) : bx * bx + by * by + bz * bz));
      });
    }
    snew = group_sum(snew);
    sold = group_sum(sold);
    ch_e = E + snew - sold; // wca.rs:124-133
    ch_dabse = snew + sold;
    e2 = ch_e;
    return true;
  }

  // Whole-system sum over half of the neighbour shell: the pair (i, j) is met once, from the atom whose subcell
  // precedes the other's in (dx, dy, dz) order (own subcell: j > i).  Lane-parallel over atoms.  f(r^2) per pair.
  template <class F>
  __device__ __forceinline__ void for_half_pairs(int natoms, F&& f) const {
    for (int i = lig; i < natoms; i += G) {
      const double rx = px[i], ry = py[i], rz = pz[i];
      int cx, cy, cz;
      subcell(rx, ry, rz, cx, cy, cz);
      for (int k = 13; k < 27; k++) { // k = 13 is the atom's own subcell, 14..26 the "later" half of the shell
        double shx, shy, shz;
        const int q = neighbour(k, cx, cy, cz, shx, shy, shz);
        for (int j = head[q]; j >= 0; j = next[j]) {
          if (k == 13 && j <= i) continue;
          const double dx = px[j] + shx - rx, dy = py[j] + shy - ry, dz = pz[j] + shz - rz;
          f(FAST ? fma(dz, dz, fma(dy, dy, dx * dx)) : dx * dx + dy * dy + dz * dz);
        }
      }
    }
  }
  // wca.rs:222-230 over the atoms currently linked (randomize adds them one by one).  With fewer than three subcells
  // along an axis two different k would name the same list; the engine refuses such boxes.
  __device__ double compute_energy_first(int natoms) const {
    double acc = 0.0;
    for_half_pairs(natoms, [&](double r2) { acc += potential(r2); });
    return group_sum(acc);
  }
  __device__ double compute_energy() const { return compute_energy_first(N); }
  __device__ __forceinline__ double expected_accuracy(double newe) const { return fabs(newe) * 1e-13 * (double)N * (double)N; } // wca.rs:178-180

  // wca.rs:164-177 with `natoms` atoms in the cell
  __device__ __forceinline__ void set_energy(double new_e, double dabse, int natoms) {
    const double n = (double)natoms;
    const double single_error = dabse > fabs(new_e) ? 1e-14 * dabse * n : 1e-14 * fabs(new_e) * n;
    err += single_error * n;
    if (err > fabs(new_e) * (1e-13 * RELAX) * n * n) {
      E = compute_energy_first(natoms);
      err = 1e-15 * E * n;
    } else {
      E = new_e;
    }
  }
  __device__ __forceinline__ void unlink(int which, int c) {
    if (head[c] == which) {
      head[c] = next[which];
    } else {
      int p = head[c];
      while (next[p] != which) p = next[p];
      next[p] = next[which];
    }
  }
  __device__ __forceinline__ void confirm() { // wca.rs:277-289 + optcell.rs:162-175
    if (ch_which < 0) return;
    gsync();
    if (lig == 0) {
      px[ch_which] = tx;
      py[ch_which] = ty;
      pz[ch_which] = tz;
      if (ch_cnew != ch_cold) {
        unlink(ch_which, ch_cold);
        next[ch_which] = head[ch_cnew];
        head[ch_cnew] = (short)ch_which;
      }
    }
    gsync();
    set_energy(ch_e, ch_dabse, N);
    ch_which = -1;
  }

  __device__ double randomize(Rng& rng) { // wca.rs:252-270
    // remove every atom, then add_atom_at + confirm one by one with E and error left stale, as the reference does
    gsync();
    for (int c = lig; c < ncells(); c += G) head[c] = -1;
    gsync();
    for (int a = 0; a < N; a++) {
      const double x = wrap1(rng.uniform_f64(0.0, Lx), Lx);
      const double y = wrap1(rng.uniform_f64(0.0, Ly), Ly);
      const double z = wrap1(rng.uniform_f64(0.0, Lz), Lz);
      int cx, cy, cz;
      subcell(x, y, z, cx, cy, cz);
      double dabse = 0.0;
      visit(cx, cy, cz, -1, [&](bool, double ix, double iy, double iz) { // Wca::add_atom_at, wca.rs:104-116
        const double ax = ix - x, ay = iy - y, az = iz - z;
        dabse += potential(FAST ? fma(az, az, fma(ay, ay, ax * ax)) : ax * ax + ay * ay + az * az);
      });
      dabse = group_sum(dabse);
      const double e = E + dabse;
      gsync();
      if (lig == 0) { // Cell::add_atom_at, optcell.rs:126-130
        px[a] = x;
        py[a] = y;
        pz[a] = z;
        const int c = flat(cx, cy, cz);
        next[a] = head[c];
        head[c] = (short)a;
      }
      gsync();
      // set_energy with num_atoms() == a + 1 and the reference's allowance (the configuration is still being built)
      const double n = (double)(a + 1);
      const double single_error = dabse > fabs(e) ? 1e-14 * dabse * n : 1e-14 * fabs(e) * n;
      err += single_error * n;
      if (err > fabs(e) * 1e-13 * n * n) {
        E = compute_energy_first(a + 1);
        err = 1e-15 * E * n;
      } else {
        E = e;
      }
    }
    E = compute_energy();
    return E;
  }
  __device__ bool verify_energy() const { // wca.rs:237-251
    const double egood = compute_energy();
    if (fabs(egood - E) > expected_accuracy(E)) return egood == E; // the reference's tolerance in both tiers
    return true;
  }
  // System::data_to_collect: the pressure every N^2 moves (wca.rs:202-218): sum over ORDERED pairs (no factor 1/2:
  // it is inside potential_pressure), i.e. twice the half-shell sum.  The next due move is kept instead of taking
  // `moves % N^2` every move.
  __device__ __forceinline__ bool extra(unsigned long long moves, double& v) {
    const unsigned long long period = (unsigned long long)N * (unsigned long long)N;
    if (x_at == 0) x_at = ((moves - 1) / period + 1) * period; // smallest multiple >= moves
    if (moves != x_at) return false;
    x_at += period;
    double p = 0.0;
    for_half_pairs(N, [&](double r2) { p += pressure(r2); });
    p = 2.0 * group_sum(p);
    v = p / (3.0 * (Lx * Ly * Lz));
    return true;
  }
  __device__ void get_pending(double* p, bool writer, bool some) const {
    if (!writer) return;
    if (!some) {
      p[0] = 0.0;
      return;
    }
    p[0] = 1.0;
    p[1] = (double)ch_which;
    p[2] = tx;
    p[3] = ty;
    p[4] = tz;
    p[5] = ch_e;
    p[6] = ch_dabse;
  }
  __device__ bool set_pending(const double* p) {
    if (p[0] == 0.0) return false;
    ch_which = (int)p[1];
    tx = p[2];
    ty = p[3];
    tz = p[4];
    ch_e = p[5];
    ch_dabse = p[6];
    int cx, cy, cz;
    subcell(tx, ty, tz, cx, cy, cz);
    ch_cnew = flat(cx, cy, cz);
    subcell(px[ch_which], py[ch_which], pz[ch_which], cx, cy, cz);
    ch_cold = flat(cx, cy, cz);
    return true;
  }
};

} // namespace sadmc
