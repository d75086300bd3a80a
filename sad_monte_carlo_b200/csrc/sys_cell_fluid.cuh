// sys_cell_fluid.cuh -- periodic fluids on a cell list: WCA and square-well / hard spheres,
// one warp per walker, configuration + cell list in shared memory.
//
// Device forms of
//   `Wca`         src/system/wca.rs      move_atom 119-140, potential 66-76, plan_move 342-353,
//                                        confirm 277-289, set_energy 164-177, compute_energy 222-230,
//                                        data_to_collect (pressure) 202-218, randomize 252-270
//   `SquareWell`  src/system/optsquare.rs move_atom 74-95, plan_move 273-284, confirm 213-221,
//                                        compute_energy 176-186, compute_energy_slowly 108-152
//   `Cell`        src/system/optcell.rs  get_subcell 112-124, add_to_subcells 131-160 (image offsets),
//                                        put_in_cell 277-309, NEIGHBORS 373-405
//
// The reference stores every atom in the lists of all 27 subcells around it,
// together with the periodic-image offset, so that one list lookup returns
// already-imaged candidates.  Here each atom sits in ONE list (linked lists:
// head[cell], next[atom], 16-bit) and a lookup visits the 27 neighbouring cells:
// lane l < 27 of the warp walks the list of neighbour cell l, applying the image
// offset the reference would have stored (-1 when the neighbour wrapped below 0,
// +1 when it wrapped past the last subcell).  The candidate SET and the imaged
// coordinates are identical to the reference's, which keeps every square-well
// decision (`r^2 < 1`, `r^2 < w^2`: integer energies) bit-exact; WCA sums the
// same terms in lane order instead of list order (tolerance tier, <= 1e-12).
#pragma once
#include "book.cuh"
#include "rng.cuh"

namespace sadmc {

template <bool SW>
struct CellFluidSys {
  static constexpr int G = 32;
  static constexpr bool FAST_BOOK = false;
  static constexpr int BLOCK = 128;
#ifndef SADMC_FLUID_MIN_BLOCKS
#define SADMC_FLUID_MIN_BLOCKS 5
#endif
  static constexpr int MIN_BLOCKS = SADMC_FLUID_MIN_BLOCKS;
  static constexpr bool COOP = false;
  static constexpr bool HAS_EXTRA = !SW; // WCA reports the pressure (wca.rs:202-218)
  static constexpr bool VERIFIES = true; // overrides System::verify_energy: run at the cadence of energy.rs:907-911
  __device__ __forceinline__ void set_cooperative(bool) {}
  __device__ __forceinline__ void finish_move() {}

  // shared-memory image of this walker
  double *px, *py, *pz;
  short *cell_of, *next, *head;
  int N, lane, ncx, ncy, ncz, ncells;
  double Lx, Ly, Lz, rc2, wsqr;
  double E, err;
  unsigned long long zone;
  // pending change
  int ch_which;
  double tx, ty, tz, ch_e, ch_dabse;
  // lane l < 27: offset of its neighbour cell
  int ndx, ndy, ndz;

  static __host__ __device__ size_t walker_bytes(uint32_t N, int ncells) {
    size_t b = (size_t)3 * N * sizeof(double) + (size_t)2 * N * sizeof(short) + (size_t)ncells * sizeof(short);
    return (b + 15) & ~(size_t)15;
  }
  static __host__ __device__ size_t smem_bytes(const DevParams& P, int block) {
    return walker_bytes(P.N, P.ncell[0] * P.ncell[1] * P.ncell[2]) * (size_t)(block / 32);
  }

  __device__ CellFluidSys(const DevParams& P, uint32_t, int lane_in_group, unsigned, unsigned char* smem)
      : N((int)P.N), lane(lane_in_group), ncx(P.ncell[0]), ncy(P.ncell[1]), ncz(P.ncell[2]), Lx(P.box[0]), Ly(P.box[1]), Lz(P.box[2]),
        rc2(P.r_cut2), wsqr(P.well2), zone(P.zone_b), ch_which(-1) {
    ncells = ncx * ncy * ncz;
    unsigned char* base = smem + walker_bytes(P.N, ncells) * (threadIdx.x / 32);
    px = reinterpret_cast<double*>(base);
    py = px + N;
    pz = py + N;
    cell_of = reinterpret_cast<short*>(pz + N);
    next = cell_of + N;
    head = next + N;
    const int l = lane_in_group < 27 ? lane_in_group : 0;
    ndx = l / 9 - 1;
    ndy = (l / 3) % 3 - 1;
    ndz = l % 3 - 1;
  }

  // optcell.rs:112-124
  __device__ __forceinline__ void subcell(double x, double y, double z, int& cx, int& cy, int& cz) const {
    cx = (int)floor(x / Lx * (double)ncx);
    cy = (int)floor(y / Ly * (double)ncy);
    cz = (int)floor(z / Lz * (double)ncz);
    // x / L * n can round up to n for x one ulp below L (the reference then wraps through `modulus`);
    // keep the index inside the grid
    cx = cx >= ncx ? ncx - 1 : (cx < 0 ? 0 : cx);
    cy = cy >= ncy ? ncy - 1 : (cy < 0 ? 0 : cy);
    cz = cz >= ncz ? ncz - 1 : (cz < 0 ? 0 : cz);
  }
  __device__ __forceinline__ int flat(int cx, int cy, int cz) const { return (cx * ncy + cy) * ncz + cz; } // optcell.rs:349-359
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
  // WCA pair energy, wca.rs:66-76
  __device__ __forceinline__ double wca_potential(double r2) const {
    if (r2 < rc2) {
      const double s = 1.0 / r2;
      const double s3 = s * s * s;
      return 4.0 * (s3 * s3 - s3) + 1.0;
    }
    return 0.0;
  }
  __device__ __forceinline__ double wca_pressure(double r2) const { // wca.rs:79-92
    if (r2 < rc2) {
      const double s = 1.0 / r2;
      const double s3 = s * s * s;
      return 4.0 * 3.0 * (2.0 * (s3 * s3) - s3);
    }
    return 0.0;
  }

  __device__ void rebuild_lists() { // optcell.rs:63-73 (update_caches)
    __syncwarp();
    for (int c = lane; c < ncells; c += 32) head[c] = -1;
    __syncwarp();
    if (lane == 0) {
      for (int a = N - 1; a >= 0; a--) {
        int cx, cy, cz;
        subcell(px[a], py[a], pz[a], cx, cy, cz);
        const int c = flat(cx, cy, cz);
        cell_of[a] = (short)c;
        next[a] = head[c];
        head[c] = (short)a;
      }
    }
    __syncwarp();
  }
  __device__ void load(const DevParams& P, uint32_t w, const WalkerRec& r) {
    const double* g = P.sys + (size_t)w * P.sys_stride;
    for (int a = lane; a < N; a += 32) {
      px[a] = g[3 * a];
      py[a] = g[3 * a + 1];
      pz[a] = g[3 * a + 2];
    }
    E = r.E;
    err = r.err;
    rebuild_lists();
    if (E != E) { // NaN in the image: "compute it" (host-side constructors that do not know the energy)
      E = compute_energy();
      err = 0.0;
    }
  }
  __device__ void store(const DevParams& P, uint32_t w, WalkerRec& r, bool writer) {
    __syncwarp();
    double* g = P.sys + (size_t)w * P.sys_stride;
    for (int a = lane; a < N; a += 32) {
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

  // Visit every candidate the reference's `maybe_interacting_atoms_excluding(r, exclude)` would return
  // (optcell.rs:93-110): f(d2) is called with |image(pos_j) - r|^2 in the reference's arithmetic.
  template <class F>
  __device__ __forceinline__ void for_neighbours(double rx, double ry, double rz, int exclude, F&& f) const {
    if (lane < 27) {
      int cx, cy, cz;
      subcell(rx, ry, rz, cx, cy, cz);
      int qx = cx + ndx, qy = cy + ndy, qz = cz + ndz;
      // the atom's own (sc + n) left the grid on the other side: optcell.rs:135-157
      double ox = 0.0, oy = 0.0, oz = 0.0;
      if (qx < 0) {
        qx += ncx;
        ox = 1.0;
      } else if (qx >= ncx) {
        qx -= ncx;
        ox = -1.0;
      }
      if (qy < 0) {
        qy += ncy;
        oy = 1.0;
      } else if (qy >= ncy) {
        qy -= ncy;
        oy = -1.0;
      }
      if (qz < 0) {
        qz += ncz;
        oz = 1.0;
      } else if (qz >= ncz) {
        qz -= ncz;
        oz = -1.0;
      }
      for (int j = head[flat(qx, qy, qz)]; j >= 0; j = next[j]) {
        if (j == exclude) continue;
        // image = pos - offset * box (optcell.rs:83-88, 103-108); then (image - r).norm2()
        const double ix = px[j] - ox * Lx, iy = py[j] - oy * Ly, iz = pz[j] - oz * Lz;
        const double dx = ix - rx, dy = iy - ry, dz = iz - rz;
        f(dx * dx + dy * dy + dz * dz);
      }
    }
  }
  __device__ __forceinline__ double warp_sum(double v) const {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
  }
  __device__ __forceinline__ int warp_sum_int(int v) const { return __reduce_add_sync(0xffffffffu, v); }

  __device__ __forceinline__ bool plan_move(Rng& rng, double scale, const double* zx, const double* zf, double& e2) {
    const int which = (int)rng.below((uint32_t)N, zone); // Uniform::new(0, N): wca.rs:345, optsquare.rs:276
    const double vx = rng.normal(zx, zf);
    const double vy = rng.normal(zx, zf);
    const double vz = rng.normal(zx, zf);
    const double fx = px[which], fy = py[which], fz = pz[which];
    tx = wrap1(fx + vx * scale, Lx); // put_in_cell(pos + vector * mean_distance)
    ty = wrap1(fy + vy * scale, Ly);
    tz = wrap1(fz + vz * scale, Lz);
    ch_which = which;
    if (SW) { // optsquare.rs:74-95
      int overlap = 0, cnt = 0;
      const double w2 = wsqr;
      for_neighbours(tx, ty, tz, which, [&](double d2) {
        if (d2 < 1.0) overlap = 1;
        if (d2 < w2) cnt -= 1;
      });
      if (__any_sync(0xffffffffu, overlap)) {
        ch_which = -1; // possible_change = Change::None
        return false;
      }
      for_neighbours(fx, fy, fz, which, [&](double d2) {
        if (d2 < w2) cnt += 1;
      });
      ch_e = E + (double)warp_sum_int(cnt);
      e2 = ch_e;
      return true;
    } else { // wca.rs:119-140
      double snew = 0.0, sold = 0.0;
      for_neighbours(tx, ty, tz, which, [&](double d2) { snew += wca_potential(d2); });
      for_neighbours(fx, fy, fz, which, [&](double d2) { sold += wca_potential(d2); });
      snew = warp_sum(snew);
      sold = warp_sum(sold);
      ch_e = E + snew - sold;
      ch_dabse = snew + sold;
      e2 = ch_e;
      return true;
    }
  }

  // Sum over ATOMS in parallel: lane l takes atoms l, l + 32, ... and visits, for each, the 27 neighbouring cells
  // one after the other (same candidate set and imaged coordinates as for_neighbours).  The per-move lookups
  // spread one atom's 27 cells over the lanes; a whole-system sum done that way runs N serial rounds per warp
  // (subcell index with three f64 divides per round), and `set_energy` asks for one every ~10 accepted moves
  // (wca.rs:164-177) -- it was most of the WCA kernel's time.
  template <class F>
  __device__ __forceinline__ void for_all_pairs(int natoms, F&& f) const {
    for (int i = lane; i < natoms; i += 32) {
      const double rx = px[i], ry = py[i], rz = pz[i];
      int cx, cy, cz;
      subcell(rx, ry, rz, cx, cy, cz);
      for (int dx = -1; dx <= 1; dx++) {
        int qx = cx + dx;
        double ox = 0.0;
        if (qx < 0) {
          qx += ncx;
          ox = 1.0;
        } else if (qx >= ncx) {
          qx -= ncx;
          ox = -1.0;
        }
        for (int dy = -1; dy <= 1; dy++) {
          int qy = cy + dy;
          double oy = 0.0;
          if (qy < 0) {
            qy += ncy;
            oy = 1.0;
          } else if (qy >= ncy) {
            qy -= ncy;
            oy = -1.0;
          }
          for (int dz = -1; dz <= 1; dz++) {
            int qz = cz + dz;
            double oz = 0.0;
            if (qz < 0) {
              qz += ncz;
              oz = 1.0;
            } else if (qz >= ncz) {
              qz -= ncz;
              oz = -1.0;
            }
            for (int j = head[flat(qx, qy, qz)]; j >= 0; j = next[j]) {
              if (j == i) continue;
              const double ix = px[j] - ox * Lx, iy = py[j] - oy * Ly, iz = pz[j] - oz * Lz;
              const double ex = ix - rx, ey = iy - ry, ez = iz - rz;
              f(ex * ex + ey * ey + ez * ez);
            }
          }
        }
      }
    }
  }
  __device__ double compute_energy() const { return compute_energy_first(N); } // wca.rs:222-230 / optsquare.rs:176-186
  // optsquare.rs:108-152: all pairs, all 27 images, no cell list (SquareWell::verify_energy)
  __device__ double compute_energy_slowly() const {
    int cnt = 0;
    for (int i = 0; i < N; i++)
      for (int j = lane; j < N; j += 32) {
        double dx = px[i] - px[j], dy = py[i] - py[j], dz = pz[i] - pz[j];
        while (dx > Lx / 2.0) dx -= Lx;
        while (dy > Ly / 2.0) dy -= Ly;
        while (dz > Lz / 2.0) dz -= Lz;
        while (dx < -Lx / 2.0) dx += Lx;
        while (dy < -Ly / 2.0) dy += Ly;
        while (dz < -Lz / 2.0) dz += Lz;
        for (int a = -1; a < 2; a++)
          for (int b = -1; b < 2; b++)
            for (int c = -1; c < 2; c++) {
              const double x = dx + Lx * (double)a, y = dy + Ly * (double)b, z = dz + Lz * (double)c;
              const double d2 = x * x + y * y + z * z;
              if (d2 < wsqr && d2 > 0.0) cnt -= 1;
            }
      }
    return (double)warp_sum_int(cnt) * 0.5;
  }
  __device__ __forceinline__ double expected_accuracy(double newe) const { return fabs(newe) * 1e-13 * (double)N * (double)N; } // wca.rs:178-180
  // wca.rs:164-177 with `natoms` atoms currently in the cell
  __device__ __forceinline__ void set_energy(double new_e, double dabse, int natoms) {
    const double n = (double)natoms;
    const double single_error = dabse > fabs(new_e) ? 1e-14 * dabse * n : 1e-14 * fabs(new_e) * n;
    err += single_error * n;
    if (err > fabs(new_e) * 1e-13 * n * n) {
      E = compute_energy();
      err = 1e-15 * E * n;
    } else {
      E = new_e;
    }
  }
  // optcell.rs:162-175: new position; relink when the subcell changed
  __device__ __forceinline__ void cell_move(int which, double x, double y, double z) {
    __syncwarp();
    if (lane == 0) {
      px[which] = x;
      py[which] = y;
      pz[which] = z;
      int cx, cy, cz;
      subcell(x, y, z, cx, cy, cz);
      const int c = flat(cx, cy, cz), oldc = cell_of[which];
      if (c != oldc) {
        if (head[oldc] == which) {
          head[oldc] = next[which];
        } else {
          int p = head[oldc];
          while (next[p] != which) p = next[p];
          next[p] = next[which];
        }
        next[which] = head[c];
        head[c] = (short)which;
        cell_of[which] = (short)c;
      }
    }
    __syncwarp();
  }
  __device__ __forceinline__ void confirm() { // wca.rs:277-289 / optsquare.rs:213-221
    if (ch_which < 0) return;
    cell_move(ch_which, tx, ty, tz);
    if (SW)
      E = ch_e;
    else
      set_energy(ch_e, ch_dabse, N);
    ch_which = -1;
  }

  __device__ double randomize(Rng& rng) { // wca.rs:252-270 (SquareWell::randomize is todo!() in the reference)
    // remove every atom, then add_atom_at + confirm one by one with E and error left stale, as the reference does
    __syncwarp();
    for (int c = lane; c < ncells; c += 32) head[c] = -1;
    for (int a = lane; a < N; a += 32) {
      px[a] = 1e300; // not yet present
      cell_of[a] = -1;
    }
    __syncwarp();
    for (int a = 0; a < N; a++) {
      const double x = wrap1(rng.uniform_f64(0.0, Lx), Lx);
      const double y = wrap1(rng.uniform_f64(0.0, Ly), Ly);
      const double z = wrap1(rng.uniform_f64(0.0, Lz), Lz);
      double dabse = 0.0;
      for_neighbours(x, y, z, -1, [&](double d2) { dabse += wca_potential(d2); }); // Wca::add_atom_at, wca.rs:104-116
      dabse = warp_sum(dabse);
      const double e = E + dabse;
      __syncwarp();
      if (lane == 0) { // Cell::add_atom_at, optcell.rs:126-130
        px[a] = x;
        py[a] = y;
        pz[a] = z;
        int cx, cy, cz;
        subcell(x, y, z, cx, cy, cz);
        const int c = flat(cx, cy, cz);
        cell_of[a] = (short)c;
        next[a] = head[c];
        head[c] = (short)a;
      }
      __syncwarp();
      // set_energy with num_atoms() == a + 1; compute_energy() must only see the atoms added so far
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
  // the same sum over the first `natoms` atoms (randomize adds them one by one)
  __device__ double compute_energy_first(int natoms) const {
    double acc = 0.0;
    int cnt = 0;
    if (SW) {
      const double w2 = wsqr;
      for_all_pairs(natoms, [&](double d2) {
        if (d2 < w2) cnt -= 1;
      });
      return (double)warp_sum_int(cnt) * 0.5;
    }
    for_all_pairs(natoms, [&](double d2) { acc += wca_potential(d2); });
    return warp_sum(acc) * 0.5;
  }
  __device__ bool verify_energy() const {
    if (SW) return E == compute_energy_slowly(); // optsquare.rs:199-201
    const double egood = compute_energy();       // wca.rs:237-251
    if (fabs(egood - E) > expected_accuracy(E)) return egood == E;
    return true;
  }
  // System::data_to_collect: WCA pressure every N^2 moves (wca.rs:202-218)
  __device__ __forceinline__ bool extra(unsigned long long moves, double& v) const {
    if (SW) return false;
    if (moves % ((unsigned long long)N * (unsigned long long)N) != 0) return false;
    double p = 0.0;
    for_all_pairs(N, [&](double d2) { p += wca_pressure(d2); });
    p = warp_sum(p);
    v = p / (3.0 * (Lx * Ly * Lz));
    return true;
  }
  __device__ void get_pending(double* p, bool writer, bool some) const {
    if (!writer) return;
    if (!some) {
      p[0] = 0.0; // optsquare.rs:80-83: an overlap clears possible_change
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
    return true;
  }
};

} // namespace sadmc
