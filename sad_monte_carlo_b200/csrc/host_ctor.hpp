// host_ctor.hpp -- the reference's system CONSTRUCTORS, run once on the host.
//
// With SADMC_INIT_REFERENCE every walker starts from the configuration the
// reference builds in `Any::from(AnyParams)` (src/system/any.rs:80-93).  Those
// constructors are sequential, RNG-driven set-up code (seconds, once per run),
// not the hot path; they produce one f64 system image (layout of
// sadmc_get_system) that the engine replicates to all walkers.  Arithmetic is in
// the reference's order (no FMA) because the constructors' accept/stop decisions
// depend on it.
#pragma once
#include <cmath>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sadmc_gpu.h"
#include "rng.cuh"

namespace sadmc {
namespace hostctor {

static const double H_ZX[SADMC_ZIG_TABLE_LEN] = SADMC_ZIG_NORM_X_INIT;
static const double H_ZF[SADMC_ZIG_TABLE_LEN] = SADMC_ZIG_NORM_F_INIT;

struct V3 {
  double x, y, z;
};
inline double norm2(const V3& a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline V3 sub(const V3& a, const V3& b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }

inline double lj_potential(double r2) { // lj.rs:78-81
  const double s = 1.0 / r2;
  const double s3 = s * s * s;
  return 4.0 * (s3 * s3 - s3);
}
inline double lj_compute_energy(const std::vector<V3>& p) { // lj.rs:236-244
  double e = 0.0;
  for (size_t which = 0; which < p.size(); which++)
    for (size_t k = 0; k < which; k++) e += lj_potential(norm2(sub(p[which], p[k])));
  return e;
}

// From<LjParams> for Lj, lj.rs:126-205
inline std::vector<double> lj_image(uint32_t n, double radius) {
  Rng rng;
  seed_from_u64(0, &rng.s0, &rng.s1);
  double best_energy = 1e80;
  std::vector<V3> best, pos(n);
  double E = 0.0, error = 0.0;
  bool done = false;
  for (uint64_t attempt = 0; attempt < 10000000ull && !done; attempt++) {
    for (uint32_t k = 0; k < n; k++) {
      V3 r;
      for (;;) {
        r.x = rng.uniform_f64(-1.0, 2.0);
        r.y = rng.uniform_f64(-1.0, 2.0);
        r.z = rng.uniform_f64(-1.0, 2.0);
        if (norm2(r) < 1.0) break;
      }
      pos[k] = V3{r.x * radius, r.y * radius, r.z * radius};
    }
    V3 cm = pos[0];
    for (uint32_t k = 1; k < n; k++) cm = V3{cm.x + pos[k].x, cm.y + pos[k].y, cm.z + pos[k].z};
    cm = V3{cm.x / (double)n, cm.y / (double)n, cm.z / (double)n};
    for (auto& x : pos) x = sub(x, cm);
    bool outside = false;
    for (auto& x : pos)
      if (norm2(x) > radius * radius) {
        outside = true;
        break;
      }
    if (outside) continue;
    E = lj_compute_energy(pos);
    if (E < best_energy) {
      best_energy = E;
      best = pos;
    }
    if (E < 0.0) done = true;
  }
  if (!done) {
    // downhill-only relaxation of the best attempt, lj.rs:183-203
    pos = best;
    E = best_energy;
    error = 0.0;
    const uint64_t zone = zone_uniform(n);
    const double R2 = radius * radius;
    for (uint64_t attempt = 0; attempt < 100000000ull; attempt++) {
      const uint32_t which = rng.below(n, zone);
      const double vx = rng.normal(H_ZX, H_ZF), vy = rng.normal(H_ZX, H_ZF), vz = rng.normal(H_ZX, H_ZF);
      const V3 from = pos[which];
      const V3 to{from.x + vx * 0.03, from.y + vy * 0.03, from.z + vz * 0.03};
      if (norm2(to) > R2 && norm2(to) > norm2(from)) continue; // None
      double e = E;
      for (uint32_t k = 0; k < n; k++) {
        if (k == which) continue;
        e += lj_potential(norm2(sub(pos[k], to))) - lj_potential(norm2(sub(pos[k], from)));
      }
      if (e < E) { // confirm + set_energy, lj.rs:110-123
        pos[which] = to;
        const double new_error = std::fabs(e) > std::fabs(E) ? std::fabs(e) * 1e-15 * (double)n : std::fabs(E) * 1e-15 * (double)n;
        error = new_error + error;
        if (error > std::fabs(e) * 1e-14 * (double)n * (double)n) {
          error *= 0.0;
          E = lj_compute_energy(pos);
        } else {
          E = e;
        }
      }
      if (E < 0.0) break;
    }
  }
  std::vector<double> img;
  for (auto& p : pos) {
    img.push_back(p.x);
    img.push_back(p.y);
    img.push_back(p.z);
  }
  img.push_back(E);
  img.push_back(error);
  return img;
}

// From<IsingParams>, ising.rs:31-53: spins from seed 10137, +-1.0 per site, then E
inline std::vector<double> ising_image(uint32_t N) {
  Rng rng;
  seed_from_u64(10137, &rng.s0, &rng.s1);
  std::vector<double> img((size_t)N * N + 1);
  for (size_t k = 0; k < (size_t)N * N; k++) img[k] = (rng.next() & 1ull) ? 1.0 : -1.0;
  double e = 0.0;
  for (uint32_t i1 = 0; i1 < N; i1++)
    for (uint32_t j1 = 0; j1 < N; j1++) {
      const uint32_t j2 = (j1 + 1) % N, i2 = (i1 + 1) % N;
      e += (img[i1 + (size_t)j2 * N] + img[i2 + (size_t)j1 * N]) * img[i1 + (size_t)j1 * N];
    }
  img[(size_t)N * N] = e;
  return img;
}

// optsquare.rs:290-322
inline uint64_t max_balls_within(double distance) {
  distance += 1e-10;
  const double a = std::sqrt(2.0);
  const long c = (long)std::ceil(distance / a) + 1;
  long num = -1;
  const double d2 = distance * distance;
  for (long n = -c; n < c + 1; n++)
    for (long m = -c; m < c + 1; m++)
      for (long l = -c; l < c + 1; l++) {
        const double x0 = (double)(m + l) * a, y0 = (double)(n + l) * a, z0 = (double)(m + n) * a;
        if (x0 * x0 + y0 * y0 + z0 * z0 <= d2) num++;
        if ((x0 + 0.5 * a) * (x0 + 0.5 * a) + (y0 + 0.5 * a) * (y0 + 0.5 * a) + z0 * z0 <= d2) num++;
        if ((x0 + 0.5 * a) * (x0 + 0.5 * a) + y0 * y0 + (z0 + 0.5 * a) * (z0 + 0.5 * a) <= d2) num++;
        if (x0 * x0 + (y0 + 0.5 * a) * (y0 + 0.5 * a) + (z0 + 0.5 * a) * (z0 + 0.5 * a) <= d2) num++;
      }
  return (uint64_t)num;
}

// From<SquareWellNParams>, optsquare.rs:357-436: N distinct random sites of a stretched FCC grid.
// The energy is left as NaN: the device counts the well overlaps when it loads the image.
inline std::vector<double> sw_image(uint32_t n, const double box[3], std::string& why) {
  const double min_cell_width = 2.0 * std::sqrt(2.0) * 0.5; // units::R = sigma / 2
  size_t cells[3];
  double cw[3];
  for (int k = 0; k < 3; k++) {
    cells[k] = (size_t)(box[k] / min_cell_width);
    if (cells[k] == 0) {
      why = "box too small for an fcc cell";
      return {};
    }
    cw[k] = box[k] / (double)cells[k];
    if (!(cw[k] >= min_cell_width)) {
      why = "fcc cell narrower than 2 sqrt(2) R";
      return {};
    }
  }
  const double off[4][3] = {{0, 0, 0}, {0.0, cw[1] / 2.0, cw[2] / 2.0}, {cw[0] / 2.0, 0.0, cw[2] / 2.0}, {cw[0] / 2.0, cw[1] / 2.0, 0.0}};
  const size_t total_spots = 4 * cells[0] * cells[1] * cells[1]; // sic (optsquare.rs:402)
  if (total_spots < n || 4 * cells[0] * cells[1] * cells[2] < n) {
    why = "not enough fcc spots for N atoms";
    return {};
  }
  std::vector<char> reserved(cells[0] * cells[1] * cells[2] * 4, 0);
  Rng rng;
  seed_from_u64(0, &rng.s0, &rng.s1);
  std::vector<double> img;
  for (uint32_t a = 0; a < n; a++) {
    for (;;) {
      const size_t i = rng.below((uint32_t)cells[0], zone_uniform(cells[0]));
      const size_t j = rng.below((uint32_t)cells[1], zone_uniform(cells[1]));
      const size_t k = rng.below((uint32_t)cells[2], zone_uniform(cells[2]));
      const size_t l = rng.below(4, zone_uniform(4));
      char& spot = reserved[((i * cells[1] + j) * cells[2] + k) * 4 + l];
      if (!spot) {
        spot = 1;
        img.push_back((double)i * cw[0] + off[l][0]);
        img.push_back((double)j * cw[1] + off[l][1]);
        img.push_back((double)k * cw[2] + off[l][2]);
        break;
      }
    }
  }
  img.push_back(std::nan(""));
  img.push_back(0.0);
  return img;
}

// From<WcaNParams> with fcc = false, wca.rs:448-496: N*N times, drop N atoms uniformly in the box
// one after another (add_atom_at + confirm, so E is the running sum the reference keeps, with
// set_energy's error budget, wca.rs:164-177), keep the attempt with the lowest E (strict <, first
// wins).  The RNG stream is one sequential generator seeded 0 and every attempt uses 3N draws, so
// the attempts are independent given the generator state at their start: the states are recorded
// in one cheap sequential pass and the attempts are then scored on all host threads.
struct WcaScratch {
  // the 27-fold subcell lists of optcell.rs:131-160, flattened: list of (atom, image offset) per subcell
  struct Entry {
    uint32_t atom;
    int8_t o[3];
  };
  long nc[3];
  double box[3], rc2;
  std::vector<std::vector<Entry>> lists;
  std::vector<V3> pos;
  double E = 0.0, error = 0.0;

  WcaScratch(const double b[3], double r_cut) {
    for (int k = 0; k < 3; k++) {
      box[k] = b[k];
      nc[k] = (long)std::floor(b[k] / r_cut); // optcell.rs:64-66
    }
    rc2 = r_cut * r_cut;
    lists.resize((size_t)(nc[0] * nc[1] * nc[2]));
  }
  void clear() {
    for (auto& l : lists) l.clear();
    pos.clear();
    E = error = 0.0;
  }
  void subcell(const V3& r, long sc[3]) const { // optcell.rs:112-124
    sc[0] = (long)std::floor(r.x / box[0] * (double)nc[0]);
    sc[1] = (long)std::floor(r.y / box[1] * (double)nc[1]);
    sc[2] = (long)std::floor(r.z / box[2] * (double)nc[2]);
  }
  size_t flat(const long q[3]) const { // optcell.rs:343-359
    return (size_t)(((q[0] + nc[0]) % nc[0]) * nc[1] * nc[2] + ((q[1] + nc[1]) % nc[1]) * nc[2] + ((q[2] + nc[2]) % nc[2]));
  }
  double potential(double r2) const { // wca.rs:66-76
    if (r2 < rc2) {
      const double s = 1.0 / r2;
      const double s3 = s * s * s;
      return 4.0 * (s3 * s3 - s3) + 1.0;
    }
    return 0.0;
  }
  template <class F>
  void neighbours(const V3& r, long exclude, F&& f) const { // optcell.rs:75-110
    long sc[3];
    subcell(r, sc);
    for (const Entry& en : lists[flat(sc)]) {
      if ((long)en.atom == exclude) continue;
      const V3& p = pos[en.atom];
      const V3 img{p.x - (double)en.o[0] * box[0], p.y - (double)en.o[1] * box[1], p.z - (double)en.o[2] * box[2]};
      f(norm2(sub(img, r)));
    }
  }
  double compute_energy() const { // wca.rs:222-230
    double e = 0.0;
    for (size_t a = 0; a < pos.size(); a++) neighbours(pos[a], (long)a, [&](double r2) { e += potential(r2); });
    return e * 0.5;
  }
  void add_and_confirm(const V3& r) { // wca.rs:104-116 + 290-297 + 164-177
    static const int8_t NB[27][3] = {{0, 0, 0},   {1, 0, 0},   {-1, 0, 0},  {0, 1, 0},  {0, -1, 0},  {0, 0, 1},   {0, 0, -1},
                                     {0, 1, 1},   {0, 1, -1},  {0, -1, 1},  {0, -1, -1}, {1, 0, 1},  {1, 0, -1},  {-1, 0, 1},
                                     {-1, 0, -1}, {1, 1, 0},   {1, -1, 0},  {-1, 1, 0}, {-1, -1, 0}, {1, 1, 1},   {-1, 1, 1},
                                     {1, -1, 1},  {1, 1, -1},  {1, -1, -1}, {-1, 1, -1}, {-1, -1, 1}, {-1, -1, -1}}; // optcell.rs:373-405
    double dabse = 0.0;
    neighbours(r, -1, [&](double r2) { dabse += potential(r2); });
    const double new_e = E + dabse;
    const uint32_t index = (uint32_t)pos.size();
    pos.push_back(r);
    long sc[3];
    subcell(r, sc);
    for (int n = 0; n < 27; n++) {
      long q[3];
      Entry en;
      en.atom = index;
      for (int k = 0; k < 3; k++) {
        q[k] = sc[k] + NB[n][k];
        en.o[k] = (int8_t)(q[k] < 0 ? -1 : (q[k] == nc[k] ? 1 : 0));
      }
      lists[flat(q)].push_back(en);
    }
    const double n_at = (double)pos.size();
    const double single = dabse > std::fabs(new_e) ? 1e-14 * dabse * n_at : 1e-14 * std::fabs(new_e) * n_at;
    error += single * n_at;
    if (error > std::fabs(new_e) * 1e-13 * n_at * n_at) {
      E = compute_energy();
      error = 1e-15 * E * n_at;
    } else {
      E = new_e;
    }
  }
};

inline std::vector<double> wca_image(uint32_t n, const double box[3], uint64_t attempts, unsigned n_threads = 0) {
  const double r_cut = std::pow(2.0, 1.0 / 6.0); // wca.rs:61-63
  if (attempts == 0) attempts = (uint64_t)n * n;
  std::vector<Rng> start(attempts);
  {
    Rng rng;
    seed_from_u64(0, &rng.s0, &rng.s1);
    for (uint64_t a = 0; a < attempts; a++) {
      start[a] = rng;
      for (uint32_t k = 0; k < 3 * n; k++) rng.next(); // Uniform<f64>::sample draws exactly one u64
    }
  }
  auto place = [&](Rng rng, std::vector<V3>& out) {
    out.resize(n);
    for (uint32_t k = 0; k < n; k++) {
      out[k].x = rng.uniform_f64(0.0, box[0]);
      out[k].y = rng.uniform_f64(0.0, box[1]);
      out[k].z = rng.uniform_f64(0.0, box[2]);
    }
  };
  if (n_threads == 0) n_threads = std::thread::hardware_concurrency();
  if (n_threads == 0) n_threads = 1;
  if (n_threads > attempts) n_threads = (unsigned)attempts;
  std::vector<double> best_e(n_threads, 1e80);
  std::vector<uint64_t> best_a(n_threads, ~0ull);
  auto worker = [&](unsigned t) {
    WcaScratch w(box, r_cut);
    std::vector<V3> p;
    const uint64_t lo = attempts * t / n_threads, hi = attempts * (t + 1) / n_threads;
    for (uint64_t a = lo; a < hi; a++) {
      place(start[a], p);
      w.clear();
      for (const V3& r : p) w.add_and_confirm(r);
      if (w.E < best_e[t]) {
        best_e[t] = w.E;
        best_a[t] = a;
      }
    }
  };
  std::vector<std::thread> pool;
  for (unsigned t = 1; t < n_threads; t++) pool.emplace_back(worker, t);
  worker(0);
  for (auto& th : pool) th.join();
  // contiguous attempt ranges in order + strict '<' == the sequential scan's "first lowest wins"
  double be = 1e80;
  uint64_t ba = ~0ull;
  for (unsigned t = 0; t < n_threads; t++)
    if (best_e[t] < be) {
      be = best_e[t];
      ba = best_a[t];
    }
  std::vector<V3> p;
  if (ba != ~0ull) place(start[ba], p); // (every attempt at 1e80 or above: the reference keeps no atoms)
  if (p.size() != n) return {};
  // wca.rs:488-495: the kept positions are added once more, then E = compute_energy()
  WcaScratch w(box, r_cut);
  for (const V3& r : p) w.add_and_confirm(r);
  std::vector<double> img;
  for (auto& r : p) {
    img.push_back(r.x);
    img.push_back(r.y);
    img.push_back(r.z);
  }
  img.push_back(w.compute_energy());
  img.push_back(w.error);
  return img;
}

} // namespace hostctor
} // namespace sadmc
