// oracle_systems.hpp -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// CPU restatement of the reference's `System` / `ConfirmSystem` /
// `MovableSystem` implementations (src/system/mod.rs:54-120) that sit on the
// hot path: Ising, Lj, Wca (+ optcell::Cell), optsquare::SquareWell, Fake,
// TwoWells, ErfInv.  One class per reference struct; each method cites the
// lines it follows.  Evaluation order of floating-point expressions follows the
// Rust source (left to right, no FMA: build with -ffp-contract=off).
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "oracle_rng.hpp"

namespace oracle {

struct Vec3 {
  double x = 0, y = 0, z = 0;
  Vec3() {}
  Vec3(double a, double b, double c) : x(a), y(b), z(c) {}
  Vec3 operator+(const Vec3& o) const { return Vec3(x + o.x, y + o.y, z + o.z); }
  Vec3 operator-(const Vec3& o) const { return Vec3(x - o.x, y - o.y, z - o.z); }
  Vec3 operator*(double s) const { return Vec3(x * s, y * s, z * s); }
  Vec3 operator/(double s) const { return Vec3(x / s, y / s, z / s); }
  double norm2() const { return x * x + y * y + z * z; } // vector3d 0.2: x*x + y*y + z*z
};

// src/rng.rs:111-117
inline Vec3 random_vector(Rng& rng) {
  const double a = rng.standard_normal();
  const double b = rng.standard_normal();
  const double c = rng.standard_normal();
  return Vec3(a, b, c);
}

// The trait surface EnergyMC uses (src/system/mod.rs:54-120).
struct System {
  virtual ~System() {}
  virtual double energy() const = 0;                                  // System::energy
  virtual double compute_energy() const = 0;                          // System::compute_energy
  virtual double randomize(Rng& rng) = 0;                             // System::randomize
  virtual bool has_delta_energy(double*) const { return false; }      // System::delta_energy
  virtual bool lowest_possible_energy(double*) const { return false; }
  virtual bool verify_energy() const { return true; }                 // false == the reference would panic
  // System::data_to_collect -> at most one (key, value) in the systems on this path
  virtual bool data_to_collect(uint64_t /*iter*/, std::string* /*key*/, double* /*value*/) const { return false; }
  virtual bool plan_move(Rng& rng, double mean_distance, double* e) = 0; // MovableSystem::plan_move; false == None
  virtual void confirm() = 0;                                            // ConfirmSystem::confirm
  // flat f64 image of the configuration, layout of sadmc_get_system (include/sadmc_gpu.h)
  virtual std::vector<double> get_state() const = 0;
  virtual void set_state(const std::vector<double>& s) = 0;
};

// --------------------------------------------------------------------------
// Ising, src/system/ising.rs
// --------------------------------------------------------------------------
struct Ising : System {
  double E = 0;
  size_t N;
  std::vector<int8_t> S;
  bool has_change = false;
  size_t change_site = 0;
  double change_e = 0;

  explicit Ising(size_t n) : N(n), S(n * n, 1) { // ising.rs:31-53
    if (!(N > 1)) throw std::invalid_argument("ising N must be > 1");
    Rng rng = Rng::seed_from_u64(10137);
    for (auto& s : S) s = (int8_t)(((int8_t)rng.next_u64() & 1) * 2 - 1);
    E = compute_energy();
  }
  double energy() const override { return E; }
  double compute_energy() const override { // ising.rs:59-75
    double e = 0.0;
    for (size_t i1 = 0; i1 < N; i1++)
      for (size_t j1 = 0; j1 < N; j1++) {
        const size_t j2 = (j1 + 1) % N;
        int8_t nt = S[i1 + j2 * N];
        const size_t i2 = (i1 + 1) % N;
        nt = (int8_t)(nt + S[i2 + j1 * N]);
        e += (double)nt * (double)S[i1 + j1 * N];
      }
    return e;
  }
  bool has_delta_energy(double* d) const override { // ising.rs:76-78
    *d = 4.0;
    return true;
  }
  double randomize(Rng& rng) override { // ising.rs:79-85
    for (auto& s : S) s = (int8_t)(((int8_t)rng.next_u64() & 1) * 2 - 1);
    E = compute_energy();
    return E;
  }
  bool plan_move(Rng& rng, double, double* e_out) override { // ising.rs:104-119
    const size_t i = rng.gen_range_usize(0, N);
    const size_t j = rng.gen_range_usize(0, N);
    size_t j2 = (j + 1) % N;
    int8_t nt = S[i + j2 * N];
    j2 = (j + N - 1) % N;
    nt = (int8_t)(nt + S[i + j2 * N]);
    size_t i2 = (i + 1) % N;
    nt = (int8_t)(nt + S[i2 + j * N]);
    i2 = (i + N - 1) % N;
    nt = (int8_t)(nt + S[i2 + j * N]);
    const double e = E - (double)nt * (double)S[i + j * N] * 2.0;
    has_change = true;
    change_site = i + j * N;
    change_e = e;
    *e_out = e;
    return true;
  }
  void confirm() override { // ising.rs:95-100 (the change is NOT cleared)
    if (has_change) {
      S[change_site] = (int8_t)(-S[change_site]);
      E = change_e;
    }
  }
  bool verify_energy() const override { return true; } // trait default: no-op
  std::vector<double> get_state() const override {
    std::vector<double> s(S.begin(), S.end());
    s.push_back(E);
    return s;
  }
  void set_state(const std::vector<double>& s) override {
    for (size_t k = 0; k < N * N; k++) S[k] = (int8_t)s[k];
    E = s[N * N];
    has_change = false;
  }
};

// --------------------------------------------------------------------------
// Lj, src/system/lj.rs
// --------------------------------------------------------------------------
struct Lj : System {
  double E = 0, error = 0;
  std::vector<Vec3> positions;
  double max_radius_squared, max_radius;
  bool has_change = false;
  size_t ch_which = 0;
  Vec3 ch_to;
  double ch_e = 0;
  // Test-only: sum the pair terms of move_atom in the order the kernel's
  // lanes_per_walker = G butterfly does, instead of lj.rs:93-102's sequential
  // order (0 = reference order).  See tests/test_lj_parity.py.
  int tree_lanes = 0;

  static double potential(double r2) { // lj.rs:78-81; powi(6), powi(3) as repeated products
    const double s = 1.0 / r2;
    const double s3 = s * s * s;
    const double s6 = s3 * s3;
    return 4.0 * (s6 - s3);
  }
  struct Empty {};
  Lj(size_t n, double radius, Empty) : positions(n), max_radius_squared(radius * radius), max_radius(radius) {}
  // From<LjParams>, lj.rs:126-205.  `max_attempts`/`max_relax` are the reference's
  // 10^7 / 10^8; tests may shrink them (then the result is NOT the reference's).
  Lj(size_t n, double radius, uint64_t max_attempts = 10000000ull, uint64_t max_relax = 100000000ull)
      : max_radius_squared(radius * radius), max_radius(radius) {
    Rng rng = Rng::seed_from_u64(0);
    double best_energy = 1e80;
    std::vector<Vec3> best_positions;
    for (uint64_t attempt = 0; attempt < max_attempts; attempt++) {
      std::vector<Vec3> pos;
      for (size_t k = 0; k < n; k++) {
        Vec3 r;
        for (;;) {
          const double a = rng.uniform_f64(-1.0, 1.0);
          const double b = rng.uniform_f64(-1.0, 1.0);
          const double c = rng.uniform_f64(-1.0, 1.0);
          r = Vec3(a, b, c);
          if (r.norm2() < 1.0) break;
        }
        pos.push_back(r * radius);
      }
      Vec3 cm = pos[0];
      for (size_t k = 1; k < n; k++) cm = cm + pos[k];
      cm = cm / (double)n;
      for (auto& x : pos) x = x - cm;
      bool outside = false;
      for (auto& x : pos)
        if (x.norm2() > radius * radius) {
          outside = true;
          break;
        }
      if (outside) continue;
      positions = pos;
      E = compute_energy();
      if (E < best_energy) {
        best_energy = E;
        best_positions = positions;
      }
      if (E < 0.0) return;
    }
    E = best_energy;
    error = 0;
    positions = best_positions;
    for (uint64_t attempt = 0; attempt < max_relax; attempt++) {
      double newe;
      if (plan_move(rng, 0.03, &newe)) {
        if (newe < E) confirm();
        if (E < 0.0) return;
      }
    }
  }
  double energy() const override { return E; }
  double compute_energy_tree() const; // oracle_capi.cpp (test-only variant)
  double compute_energy() const override { // lj.rs:236-244
    if (tree_lanes) return compute_energy_tree();
    double e = 0.0;
    for (size_t which = 0; which < positions.size(); which++)
      for (size_t k = 0; k < which; k++) e += potential((positions[which] - positions[k]).norm2());
    return e;
  }
  bool lowest_possible_energy(double* e) const override { // lj.rs:245-248
    const double n = (double)positions.size();
    *e = -0.5 * n * (n - 1.0);
    return true;
  }
  double expected_accuracy(double newe) const { // lj.rs:106-108
    return std::fabs(newe) * 1e-14 * (double)positions.size() * (double)positions.size();
  }
  bool verify_energy() const override { // lj.rs:249-261
    const double egood = compute_energy();
    const double expected = expected_accuracy(E);
    if (std::fabs(egood - E) > expected) return egood == E;
    return true;
  }
  double randomize(Rng& rng) override { // lj.rs:262-279
    for (auto& x : positions) {
      Vec3 r;
      for (;;) {
        const double a = rng.uniform_f64(-1.0, 1.0);
        const double b = rng.uniform_f64(-1.0, 1.0);
        const double c = rng.uniform_f64(-1.0, 1.0);
        r = Vec3(a, b, c);
        if (r.norm2() < 1.0) break;
      }
      x = r * max_radius;
    }
    E = compute_energy();
    return E;
  }
  bool move_atom(size_t which, Vec3 r, double* e_out) { // lj.rs:86-105
    const double previous_rsqr = positions[which].norm2();
    if (r.norm2() > max_radius_squared && r.norm2() > previous_rsqr) return false;
    double e = E;
    const Vec3 from = positions[which];
    if (tree_lanes == 0) {
      for (size_t k = 0; k < positions.size(); k++) {
        if (k == which) continue;
        const Vec3 r1 = positions[k];
        e += potential((r1 - r).norm2()) - potential((r1 - from).norm2());
      }
    } else {
      e = move_atom_tree(which, r);
    }
    has_change = true;
    ch_which = which;
    ch_to = r;
    ch_e = e;
    *e_out = e;
    return true;
  }
  double move_atom_tree(size_t which, Vec3 r) const; // defined in oracle_capi.cpp (test-only variant)
  void set_energy(double new_e) { // lj.rs:110-123
    const double n = (double)positions.size();
    const double new_error = std::fabs(new_e) > std::fabs(E) ? std::fabs(new_e) * 1e-15 * n : std::fabs(E) * 1e-15 * n;
    error = new_error + error;
    if (error > expected_accuracy(new_e)) {
      error *= 0.0;
      E = compute_energy();
    } else {
      E = new_e;
    }
  }
  bool plan_move(Rng& rng, double mean_distance, double* e) override { // lj.rs:365-374
    if (positions.empty()) return false;
    const size_t which = (size_t)rng.uniform_usize(0, positions.size());
    const Vec3 to = positions[which] + random_vector(rng) * mean_distance;
    return move_atom(which, to, e);
  }
  void confirm() override { // lj.rs:339-346
    if (!has_change) return;
    positions[ch_which] = ch_to;
    has_change = false;
    set_energy(ch_e);
  }
  std::vector<double> get_state() const override {
    std::vector<double> s;
    for (auto& p : positions) {
      s.push_back(p.x);
      s.push_back(p.y);
      s.push_back(p.z);
    }
    s.push_back(E);
    s.push_back(error);
    return s;
  }
  void set_state(const std::vector<double>& s) override {
    const size_t n = positions.size();
    for (size_t k = 0; k < n; k++) positions[k] = Vec3(s[3 * k], s[3 * k + 1], s[3 * k + 2]);
    E = s[3 * n];
    error = s[3 * n + 1];
    has_change = false;
  }
};

// --------------------------------------------------------------------------
// optcell::Cell, src/system/optcell.rs
// --------------------------------------------------------------------------
struct Cell {
  struct Neighbor {
    uint32_t index;
    int8_t ox, oy, oz;
  };
  Vec3 box_diagonal;
  double r_cutoff;
  std::vector<Vec3> positions;
  long ncx = 0, ncy = 0, ncz = 0;
  std::vector<std::vector<Neighbor>> subcells;

  static const int (*neighbors())[3] { // optcell.rs:373-405, same order
    static const int NB[27][3] = {{0, 0, 0},   {1, 0, 0},   {-1, 0, 0},  {0, 1, 0},  {0, -1, 0}, {0, 0, 1},   {0, 0, -1},
                                  {0, 1, 1},   {0, 1, -1},  {0, -1, 1},  {0, -1, -1}, {1, 0, 1},  {1, 0, -1},  {-1, 0, 1},
                                  {-1, 0, -1}, {1, 1, 0},   {1, -1, 0},  {-1, 1, 0}, {-1, -1, 0}, {1, 1, 1},   {-1, 1, 1},
                                  {1, -1, 1},  {1, 1, -1},  {1, -1, -1}, {-1, 1, -1}, {-1, -1, 1}, {-1, -1, -1}};
    return NB;
  }
  Cell() : r_cutoff(0) {}
  Cell(Vec3 box, double interaction_length) : box_diagonal(box), r_cutoff(interaction_length) { update_caches(); } // optcell.rs:44-61
  static Cell from_volume(double v, double interaction_length) {
    const double w = std::cbrt(v);
    return Cell(Vec3(w, w, w), interaction_length);
  }
  void update_caches() { // optcell.rs:63-73
    ncx = (long)std::floor(box_diagonal.x / r_cutoff);
    ncy = (long)std::floor(box_diagonal.y / r_cutoff);
    ncz = (long)std::floor(box_diagonal.z / r_cutoff);
    subcells.assign((size_t)(ncx * ncy * ncz), std::vector<Neighbor>());
    for (size_t i = 0; i < positions.size(); i++) add_to_subcells((uint32_t)i, positions[i]);
  }
  void get_subcell(const Vec3& r, long sc[3]) const { // optcell.rs:112-124
    sc[0] = (long)std::floor(r.x / box_diagonal.x * (double)ncx);
    sc[1] = (long)std::floor(r.y / box_diagonal.y * (double)ncy);
    sc[2] = (long)std::floor(r.z / box_diagonal.z * (double)ncz);
  }
  static size_t modulus(long i, long sz) { return (size_t)((i + sz) % sz); } // optcell.rs:343-346
  size_t flat(long ix, long iy, long iz) const { // optcell.rs:349-359
    return modulus(ix, ncx) * (size_t)(ncy * ncz) + modulus(iy, ncy) * (size_t)ncz + modulus(iz, ncz);
  }
  void add_to_subcells(uint32_t index, const Vec3& r) { // optcell.rs:131-160
    long sc[3];
    get_subcell(r, sc);
    const int(*NB)[3] = neighbors();
    for (int n = 0; n < 27; n++) {
      const long px = sc[0] + NB[n][0], py = sc[1] + NB[n][1], pz = sc[2] + NB[n][2];
      Neighbor nb;
      nb.index = index;
      nb.ox = (int8_t)(px < 0 ? -1 : (px == ncx ? 1 : 0));
      nb.oy = (int8_t)(py < 0 ? -1 : (py == ncy ? 1 : 0));
      nb.oz = (int8_t)(pz < 0 ? -1 : (pz == ncz ? 1 : 0));
      subcells[flat(px, py, pz)].push_back(nb);
    }
  }
  // optcell.rs:75-110: candidates come back already imaged: pos - offset * box
  template <class F>
  void for_maybe_interacting(const Vec3& r, long exclude, F&& f) const {
    long sc[3];
    get_subcell(r, sc);
    for (const Neighbor& nb : subcells[flat(sc[0], sc[1], sc[2])]) {
      if (exclude >= 0 && nb.index == (uint32_t)exclude) continue;
      const Vec3& p = positions[nb.index];
      const Vec3 img(p.x - (double)nb.ox * box_diagonal.x, p.y - (double)nb.oy * box_diagonal.y,
                     p.z - (double)nb.oz * box_diagonal.z);
      if (!f(img)) return;
    }
  }
  void add_atom_at(const Vec3& r) { // optcell.rs:126-130
    const uint32_t index = (uint32_t)positions.size();
    positions.push_back(r);
    add_to_subcells(index, r);
  }
  static void remove_if_index(std::vector<Neighbor>& v, uint32_t index) { // optcell.rs:407-418
    for (size_t i = 0; i < v.size(); i++)
      if (v[i].index == index) {
        v[i] = v.back(); // swap_remove
        v.pop_back();
        return;
      }
  }
  Vec3 move_atom(size_t which, const Vec3& r) { // optcell.rs:162-175
    const Vec3 old = positions[which];
    positions[which] = r;
    long sc[3], oldsc[3];
    get_subcell(r, sc);
    get_subcell(old, oldsc);
    if (sc[0] != oldsc[0] || sc[1] != oldsc[1] || sc[2] != oldsc[2]) {
      const int(*NB)[3] = neighbors();
      for (int n = 0; n < 27; n++)
        remove_if_index(subcells[flat(oldsc[0] + NB[n][0], oldsc[1] + NB[n][1], oldsc[2] + NB[n][2])], (uint32_t)which);
      add_to_subcells((uint32_t)which, r);
    }
    return old;
  }
  double volume() const { return box_diagonal.x * box_diagonal.y * box_diagonal.z; } // optcell.rs:195-197
  static double wrap1(double v, double L) { // optcell.rs:277-309, one axis
    if (v < 0.0) {
      do {
        v += L;
      } while (v < 0.0);
    } else {
      while (v >= L) v -= L;
    }
    return v;
  }
  Vec3 put_in_cell(Vec3 r) const { return Vec3(wrap1(r.x, box_diagonal.x), wrap1(r.y, box_diagonal.y), wrap1(r.z, box_diagonal.z)); }
};

// --------------------------------------------------------------------------
// Wca, src/system/wca.rs
// --------------------------------------------------------------------------
struct Wca : System {
  double E = 0, error = 0;
  Cell cell;
  enum { NONE, MOVE, ADD } ch_kind = NONE;
  size_t ch_which = 0;
  Vec3 ch_to;
  double ch_e = 0, ch_dabse = 0;

  static double r_cutoff() { return std::pow(2.0, 1.0 / 6.0); } // wca.rs:61-63
  static double potential(double r2) {                          // wca.rs:66-76
    const double rc = r_cutoff();
    const double rc2 = rc * rc;
    if (r2 < rc2) {
      const double s = 1.0 / r2;
      const double s3 = s * s * s;
      const double s6 = s3 * s3;
      return 4.0 * (s6 - s3) + 1.0;
    }
    return 0.0;
  }
  static double potential_pressure(double r2) { // wca.rs:79-92
    const double rc = r_cutoff();
    const double rc2 = rc * rc;
    if (r2 < rc2) {
      const double s = 1.0 / r2;
      const double s3 = s * s * s;
      const double s6 = s3 * s3;
      return 4.0 * 3.0 * (2.0 * s6 - s3);
    }
    return 0.0;
  }
  explicit Wca(const Cell& c) : cell(c) { // From<WcaParams>, wca.rs:183-199
    if (cell.r_cutoff > cell.box_diagonal.x || cell.r_cutoff > cell.box_diagonal.y || cell.r_cutoff > cell.box_diagonal.z)
      throw std::invalid_argument("The cell is not large enough for the well width, sorry!");
  }
  size_t num_atoms() const { return cell.positions.size(); }
  // From<WcaNParams> with fcc = false, wca.rs:392-498 (n*n random attempts, keep the lowest energy)
  static Wca from_n(size_t n, Vec3 box, uint64_t attempts_override = ~0ull) {
    const Cell proto(box, r_cutoff());
    Rng rng = Rng::seed_from_u64(0);
    Wca probe(proto);
    double best_energy = 1e80;
    std::vector<Vec3> best_positions;
    const uint64_t attempts = attempts_override == ~0ull ? (uint64_t)n * n : attempts_override;
    for (uint64_t attempt = 0; attempt < attempts; attempt++) {
      std::vector<Vec3> pos;
      for (size_t k = 0; k < n; k++) {
        const double a = rng.uniform_f64(0.0, box.x);
        const double b = rng.uniform_f64(0.0, box.y);
        const double c = rng.uniform_f64(0.0, box.z);
        pos.push_back(Vec3(a, b, c));
      }
      Wca w(proto);
      for (auto& r : pos) {
        w.add_atom_at(r);
        w.confirm();
      }
      if (w.E < best_energy) {
        best_energy = w.E;
        best_positions = pos;
      }
    }
    Wca w(proto);
    for (auto& r : best_positions) {
      w.add_atom_at(r);
      w.confirm();
    }
    w.E = w.compute_energy();
    return w;
  }
  bool add_atom_at(const Vec3& r) { // wca.rs:104-116
    double dabse = 0.0;
    cell.for_maybe_interacting(r, -1, [&](const Vec3& r1) {
      dabse += potential((r1 - r).norm2());
      return true;
    });
    ch_kind = ADD;
    ch_to = r;
    ch_e = E + dabse;
    ch_dabse = dabse;
    return true;
  }
  bool move_atom(size_t which, const Vec3& r, double* e_out) { // wca.rs:119-140
    double e = E, dabse = 0.0;
    const Vec3 from = cell.positions[which];
    cell.for_maybe_interacting(r, (long)which, [&](const Vec3& r1) {
      const double de = potential((r1 - r).norm2());
      e += de;
      dabse += de;
      return true;
    });
    cell.for_maybe_interacting(from, (long)which, [&](const Vec3& r1) {
      const double de = potential((r1 - from).norm2());
      e -= de;
      dabse += de;
      return true;
    });
    ch_kind = MOVE;
    ch_which = which;
    ch_to = r;
    ch_e = e;
    ch_dabse = dabse;
    *e_out = e;
    return true;
  }
  double expected_accuracy(double newe) const { return std::fabs(newe) * 1e-13 * (double)num_atoms() * (double)num_atoms(); } // wca.rs:178-180
  void set_energy(double new_e, double dabse) { // wca.rs:164-177
    const double n = (double)num_atoms();
    const double single_error = dabse > std::fabs(new_e) ? 1e-14 * dabse * n : 1e-14 * std::fabs(new_e) * n;
    error += single_error * n;
    if (error > expected_accuracy(new_e)) {
      E = compute_energy();
      error = 1e-15 * E * n;
    } else {
      E = new_e;
    }
  }
  double energy() const override { return E; }
  double compute_energy() const override { // wca.rs:222-230
    double e = 0.0;
    for (size_t which = 0; which < num_atoms(); which++) {
      const Vec3 r1 = cell.positions[which];
      cell.for_maybe_interacting(r1, (long)which, [&](const Vec3& r2) {
        e += potential((r1 - r2).norm2());
        return true;
      });
    }
    return e * 0.5;
  }
  bool lowest_possible_energy(double* e) const override { // wca.rs:234-236
    *e = 0.0;
    return true;
  }
  bool verify_energy() const override { // wca.rs:237-251
    const double egood = compute_energy();
    if (std::fabs(egood - E) > expected_accuracy(E)) return egood == E;
    return true;
  }
  bool data_to_collect(uint64_t iter, std::string* key, double* value) const override { // wca.rs:202-218
    const uint64_t n = num_atoms();
    if (iter % (n * n) != 0) return false;
    double p = 0.0;
    for (size_t which = 0; which < n; which++) {
      const Vec3 r1 = cell.positions[which];
      cell.for_maybe_interacting(r1, (long)which, [&](const Vec3& r2) {
        p += potential_pressure((r1 - r2).norm2());
        return true;
      });
    }
    *key = "pressure";
    *value = p / (3.0 * cell.volume());
    return true;
  }
  double randomize(Rng& rng) override { // wca.rs:252-270 (remove all, re-add uniformly through add_atom_at + confirm)
    const size_t n = num_atoms();
    cell.positions.clear(); // == remove_atom(0) n times
    cell.update_caches();
    for (size_t k = 0; k < n; k++) {
      const double a = rng.uniform_f64(0.0, cell.box_diagonal.x);
      const double b = rng.uniform_f64(0.0, cell.box_diagonal.y);
      const double c = rng.uniform_f64(0.0, cell.box_diagonal.z);
      add_atom_at(cell.put_in_cell(Vec3(a, b, c))); // E and error stay stale, as in the reference
      confirm();
    }
    E = compute_energy();
    return E;
  }
  bool plan_move(Rng& rng, double mean_distance, double* e) override { // wca.rs:342-353
    if (num_atoms() == 0) return false;
    const size_t which = (size_t)rng.uniform_usize(0, num_atoms());
    const Vec3 to = cell.put_in_cell(cell.positions[which] + random_vector(rng) * mean_distance);
    return move_atom(which, to, e);
  }
  void confirm() override { // wca.rs:277-306
    if (ch_kind == MOVE) {
      cell.move_atom(ch_which, ch_to);
      set_energy(ch_e, ch_dabse);
    } else if (ch_kind == ADD) {
      cell.add_atom_at(ch_to);
      set_energy(ch_e, ch_dabse);
    }
    ch_kind = NONE;
  }
  std::vector<double> get_state() const override {
    std::vector<double> s;
    for (auto& p : cell.positions) {
      s.push_back(p.x);
      s.push_back(p.y);
      s.push_back(p.z);
    }
    s.push_back(E);
    s.push_back(error);
    return s;
  }
  void set_state(const std::vector<double>& s) override {
    const size_t n = (s.size() - 2) / 3;
    cell.positions.resize(n);
    for (size_t k = 0; k < n; k++) cell.positions[k] = Vec3(s[3 * k], s[3 * k + 1], s[3 * k + 2]);
    cell.update_caches();
    E = s[3 * n];
    error = s[3 * n + 1];
    ch_kind = NONE;
  }
};

// --------------------------------------------------------------------------
// optsquare::SquareWell, src/system/optsquare.rs
// --------------------------------------------------------------------------
struct SquareWell : System {
  double E = 0;
  Cell cell;
  enum { NONE, MOVE, ADD } ch_kind = NONE;
  size_t ch_which = 0;
  Vec3 ch_to;
  double ch_e = 0;

  explicit SquareWell(const Cell& c) : cell(c) { // optsquare.rs:154-170
    if (cell.r_cutoff > cell.box_diagonal.x || cell.r_cutoff > cell.box_diagonal.y || cell.r_cutoff > cell.box_diagonal.z)
      throw std::invalid_argument("The cell is not large enough for the well width, sorry!");
  }
  size_t num_atoms() const { return cell.positions.size(); }
  // From<SquareWellNParams>, optsquare.rs:357-436: random distinct FCC sites
  static SquareWell from_n(size_t n, Vec3 box, double well_width) {
    SquareWell sw(Cell(box, well_width * 1.0));
    const double min_cell_width = 2.0 * std::sqrt(2.0) * 0.5; // units::R = sigma/2
    size_t cells[3] = {(size_t)(sw.cell.box_diagonal.x / min_cell_width), (size_t)(sw.cell.box_diagonal.y / min_cell_width),
                       (size_t)(sw.cell.box_diagonal.z / min_cell_width)};
    const double cw[3] = {sw.cell.box_diagonal.x / (double)cells[0], sw.cell.box_diagonal.y / (double)cells[1],
                          sw.cell.box_diagonal.z / (double)cells[2]};
    for (int i = 0; i < 3; i++)
      if (!(cw[i] >= min_cell_width)) throw std::invalid_argument("sw: cell too small for fcc placement");
    const Vec3 offset[4] = {Vec3(0, 0, 0), Vec3(0.0, cw[1], cw[2]) / 2.0, Vec3(cw[0], 0.0, cw[2]) / 2.0, Vec3(cw[0], cw[1], 0.0) / 2.0};
    const size_t total_spots = 4 * cells[0] * cells[1] * cells[1]; // sic, optsquare.rs:402
    if (total_spots < n) throw std::invalid_argument("sw: not enough fcc spots");
    std::vector<char> reserved(cells[0] * cells[1] * cells[2] * 4, 0);
    Rng rng = Rng::seed_from_u64(0);
    for (size_t a = 0; a < n; a++) {
      for (;;) {
        const size_t i = (size_t)rng.uniform_usize(0, cells[0]);
        const size_t j = (size_t)rng.uniform_usize(0, cells[1]);
        const size_t k = (size_t)rng.uniform_usize(0, cells[2]);
        const size_t l = (size_t)rng.uniform_usize(0, 4);
        char& spot = reserved[((i * cells[1] + j) * cells[2] + k) * 4 + l];
        if (!spot) {
          spot = 1;
          sw.add_atom_at(Vec3((double)i * cw[0], (double)j * cw[1], (double)k * cw[2]) + offset[l]);
          sw.confirm();
          break;
        }
      }
    }
    return sw;
  }
  bool add_atom_at(const Vec3& r) { // optsquare.rs:57-71
    double e = E;
    bool overlap = false;
    const double wsqr = cell.r_cutoff * cell.r_cutoff;
    cell.for_maybe_interacting(r, -1, [&](const Vec3& r1) {
      const double d2 = (r1 - r).norm2();
      if (d2 < 1.0) {
        overlap = true;
        return false;
      } else if (d2 < wsqr) {
        e -= 1.0;
      }
      return true;
    });
    if (overlap) {
      ch_kind = NONE;
      return false;
    }
    ch_kind = ADD;
    ch_to = r;
    ch_e = e;
    return true;
  }
  bool move_atom(size_t which, const Vec3& r, double* e_out) { // optsquare.rs:74-95
    double e = E;
    const double wsqr = cell.r_cutoff * cell.r_cutoff;
    const Vec3 from = cell.positions[which];
    bool overlap = false;
    cell.for_maybe_interacting(r, (long)which, [&](const Vec3& r1) {
      const double d2 = (r1 - r).norm2();
      if (d2 < 1.0) {
        overlap = true;
        return false;
      }
      if (d2 < wsqr) e -= 1.0;
      return true;
    });
    if (overlap) {
      ch_kind = NONE;
      return false;
    }
    cell.for_maybe_interacting(from, (long)which, [&](const Vec3& r1) {
      if ((r1 - from).norm2() < wsqr) e += 1.0;
      return true;
    });
    ch_kind = MOVE;
    ch_which = which;
    ch_to = r;
    ch_e = e;
    *e_out = e;
    return true;
  }
  double energy() const override { return E; }
  double compute_energy() const override { // optsquare.rs:176-186
    double e = 0.0;
    const double wsqr = cell.r_cutoff * cell.r_cutoff;
    for (size_t which = 0; which < num_atoms(); which++) {
      const Vec3 r1 = cell.positions[which];
      cell.for_maybe_interacting(r1, (long)which, [&](const Vec3& r2) {
        if ((r1 - r2).norm2() < wsqr) e -= 1.0;
        return true;
      });
    }
    return e * 0.5;
  }
  // optsquare.rs:108-152: all pairs, all 27 images, no cell list
  double compute_energy_slowly() const {
    double e = 0.0;
    const Vec3 L = cell.box_diagonal;
    const double wsqr = cell.r_cutoff * cell.r_cutoff;
    for (const Vec3& r1 : cell.positions)
      for (const Vec3& r2 : cell.positions) {
        Vec3 d = r1 - r2;
        while (d.x > L.x / 2.0) d.x -= L.x;
        while (d.y > L.y / 2.0) d.y -= L.y;
        while (d.z > L.z / 2.0) d.z -= L.z;
        while (d.x < -L.x / 2.0) d.x += L.x;
        while (d.y < -L.y / 2.0) d.y += L.y;
        while (d.z < -L.z / 2.0) d.z += L.z;
        for (int i = -1; i < 2; i++)
          for (int j = -1; j < 2; j++)
            for (int k = -1; k < 2; k++) {
              const Vec3 r = d + Vec3(L.x * (double)i, L.y * (double)j, L.z * (double)k);
              const double d2 = r.norm2();
              if (d2 < wsqr && d2 > 0.0) e -= 1.0;
            }
      }
    return e * 0.5;
  }
  bool has_delta_energy(double* d) const override { // optsquare.rs:190-192
    *d = 1.0;
    return true;
  }
  static uint64_t max_balls_within(double distance) { // optsquare.rs:290-322
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
  bool lowest_possible_energy(double* e) const override { // optsquare.rs:196-198
    *e = -(double)num_atoms() * (double)max_balls_within(cell.r_cutoff);
    return true;
  }
  bool verify_energy() const override { return E == compute_energy_slowly(); } // optsquare.rs:199-201
  double randomize(Rng&) override { throw std::logic_error("optsquare randomize is todo!() in the reference (optsquare.rs:202-204)"); }
  bool plan_move(Rng& rng, double mean_distance, double* e) override { // optsquare.rs:273-284
    if (num_atoms() == 0) return false;
    const size_t which = (size_t)rng.uniform_usize(0, num_atoms());
    const Vec3 to = cell.put_in_cell(cell.positions[which] + random_vector(rng) * mean_distance);
    return move_atom(which, to, e);
  }
  void confirm() override { // optsquare.rs:213-221
    if (ch_kind == MOVE) {
      cell.move_atom(ch_which, ch_to);
      E = ch_e;
    } else if (ch_kind == ADD) {
      cell.add_atom_at(ch_to);
      E = ch_e;
    }
    ch_kind = NONE;
  }
  std::vector<double> get_state() const override {
    std::vector<double> s;
    for (auto& p : cell.positions) {
      s.push_back(p.x);
      s.push_back(p.y);
      s.push_back(p.z);
    }
    s.push_back(E);
    s.push_back(0.0);
    return s;
  }
  void set_state(const std::vector<double>& s) override {
    const size_t n = (s.size() - 2) / 3;
    cell.positions.resize(n);
    for (size_t k = 0; k < n; k++) cell.positions[k] = Vec3(s[3 * k], s[3 * k + 1], s[3 * k + 2]);
    cell.update_caches();
    E = s[3 * n];
    ch_kind = NONE;
  }
};

// --------------------------------------------------------------------------
// Fake, src/system/fake.rs
// --------------------------------------------------------------------------
struct Fake : System {
  enum Kind { LINEAR = 0, QUADRATIC = 1, PIECES = 2, GAUSSIAN = 3 } kind;
  double a = 0, b = 0, e1 = 0, e2 = 0, sigma = 0;
  std::vector<double> position, possible_change;

  Fake(Kind k, size_t quad_dims, double a_, double b_, double e1_, double e2_, double sigma_)
      : kind(k), a(a_), b(b_), e1(e1_), e2(e2_), sigma(sigma_) { // fake.rs:85-93, dimensions 39-46
    const size_t d = k == LINEAR ? 1 : (k == QUADRATIC ? quad_dims : 3);
    position.assign(d, 0.0);
    possible_change.assign(d, 0.0);
  }
  double f(double r) const { // fake.rs:47-60
    switch (kind) {
      case LINEAR: return r;
      case QUADRATIC: return r * r;
      case GAUSSIAN: return -o_exp(-r * r / (2.0 * sigma * sigma));
      case PIECES:
        if (r < a) return (r * r) / (a * a) * e1 - e1;
        return ((r - b) / (b - a)) * ((r - b) / (b - a)) * e2 - e2;
    }
    return 0;
  }
  static double radius(const std::vector<double>& p) { // iter().map(x*x).sum::<f64>().sqrt(); sum starts at 0.0
    double s = 0.0;
    for (double x : p) s += x * x;
    return std::sqrt(s);
  }
  double energy() const override { return f(radius(position)); } // fake.rs:96-99
  double compute_energy() const override { return energy(); }
  double randomize(Rng& rng) override { // fake.rs:103-112
    double r = 5.0;
    while (r >= 1.0) {
      for (auto& x : position) x = rng.gen_range_f64(0.0, 1.0);
      r = radius(position);
    }
    return energy();
  }
  bool plan_move(Rng& rng, double d, double* e) override { // fake.rs:128-144
    const size_t i = (size_t)rng.gen_range_usize(0, position.size());
    possible_change = position;
    const double v = rng.standard_normal();
    possible_change[i] += v * d;
    const double r = radius(possible_change);
    if (r > 1.0) return false;
    *e = f(r);
    return true;
  }
  void confirm() override { position = possible_change; } // fake.rs:122-124
  std::vector<double> get_state() const override { return position; }
  void set_state(const std::vector<double>& s) override {
    position = s;
    possible_change = s;
  }
};

// --------------------------------------------------------------------------
// TwoWells, src/system/two_wells.rs (the InvCdf sampler, 25-180, is only used by
// `randomize` and is not restated: randomize() is unsupported here)
// --------------------------------------------------------------------------
struct TwoWells : System {
  std::vector<double> position;
  double d_squared;
  size_t N;
  double h2_to_h1, barrier_over_h1, r2, well_position;
  size_t ch_index = 0;
  Vec3 ch_values;

  TwoWells(size_t n, double h2h1, double barrier, double r2_) : N(n), h2_to_h1(h2h1), barrier_over_h1(barrier), r2(r2_) { // two_wells.rs:234-263
    if (N % 3 != 0) throw std::invalid_argument("The number of dimensions is not divisible by three!");
    well_position = std::sqrt(barrier_over_h1) * 1.0 + r2 * std::sqrt(1.0 + barrier_over_h1 - 1.0 / h2_to_h1);
    position.assign(N, 0.0);
    position[0] = -0.99;
    d_squared = 0.0;
    for (double x : position) d_squared += x * x;
  }
  struct Regions {
    double e_1, e_2, e_w, e_i, d_1_squared, d_2_squared;
  };
  Regions regions(double x1, double d_orthog_squared) const { // two_wells.rs:266-293
    const double r1 = 1.0;
    const double rw = well_position;
    const double x2 = x1 - r1 - r2;
    const double xw = x1 - rw;
    const double xi = x2 + rw;
    Regions g;
    g.d_1_squared = d_orthog_squared + x1 * x1;
    g.d_2_squared = d_orthog_squared + x2 * x2;
    const double d_w_squared = d_orthog_squared + xw * xw;
    const double d_i_squared = d_orthog_squared + xi * xi;
    g.e_1 = 1.0 * (g.d_1_squared / (r1 * r1) - 1.0);
    g.e_2 = h2_to_h1 * (g.d_2_squared / (r2 * r2) - 1.0);
    g.e_w = h2_to_h1 * (d_w_squared / (r2 * r2) - 1.0);
    g.e_i = 1.0 * (d_i_squared / (r1 * r1) - 1.0);
    return g;
  }
  bool find_energy(double x1, double d_orthog_squared, double* e) const { // two_wells.rs:266-315
    const Regions g = regions(x1, d_orthog_squared);
    const double r1 = 1.0;
    if (g.d_1_squared <= r1 * r1) {
      *e = g.e_1 < g.e_w ? g.e_1 : g.e_w;
      return true;
    } else if (g.d_2_squared <= r2 * r2) {
      *e = (g.e_i > g.e_2 && g.e_i < 0.0) ? g.e_i : g.e_2;
      return true;
    } else if (d_orthog_squared <= r2 * r2 && x1 > 0.0 && x1 <= r1 + r2) {
      *e = 0.0;
      return true;
    }
    return false;
  }
  double find_which(double x1, double d_orthog_squared) const { // two_wells.rs:317-371
    const Regions g = regions(x1, d_orthog_squared);
    const double r1 = 1.0;
    if (g.d_1_squared <= r1 * r1) return g.e_1 < g.e_w ? 0.0 : 1.0;
    if (g.d_2_squared <= r2 * r2) return (g.e_i > g.e_2 && g.e_i < 0.0) ? 0.0 : 1.0;
    return 0.0;
  }
  double energy() const override { return compute_energy(); } // two_wells.rs:375-377
  double compute_energy() const override {                    // two_wells.rs:378-384
    double e;
    if (!find_energy(position[0], d_squared - position[0] * position[0], &e)) throw std::runtime_error("position is out of bounds");
    return e;
  }
  double randomize(Rng&) override { throw std::logic_error("two-wells randomize (InvCdf sampler) is not restated"); }
  bool data_to_collect(uint64_t, std::string* key, double* value) const override { // two_wells.rs:408-418
    *key = "which";
    *value = find_which(position[0], d_squared - position[0] * position[0]);
    return true;
  }
  bool plan_move(Rng& rng, double d, double* e) override { // two_wells.rs:451-464
    const size_t index = 3 * (size_t)rng.gen_range_usize(0, position.size() / 3);
    const Vec3 old_r(position[index], position[index + 1], position[index + 2]);
    const Vec3 r = random_vector(rng) * d + old_r;
    const double dsq = d_squared - old_r.norm2() + r.norm2();
    const double x1 = index == 0 ? r.x : position[0];
    ch_index = index;
    ch_values = r;
    return find_energy(x1, dsq - x1 * x1, e);
  }
  void confirm() override { // two_wells.rs:437-447
    double s = 0.0;
    for (size_t k = ch_index; k < ch_index + 3; k++) s += position[k] * position[k];
    d_squared -= s;
    position[ch_index] = ch_values.x;
    position[ch_index + 1] = ch_values.y;
    position[ch_index + 2] = ch_values.z;
    d_squared += ch_values.norm2();
  }
  std::vector<double> get_state() const override {
    std::vector<double> s = position;
    s.push_back(d_squared);
    return s;
  }
  void set_state(const std::vector<double>& s) override {
    for (size_t k = 0; k < N; k++) position[k] = s[k];
    d_squared = s[N];
  }
};

// --------------------------------------------------------------------------
// ErfInv, src/system/erfinv.rs.  statrs 0.7 `erf_inv` is an un-vendored
// dependency; restated in oracle_erfinv.hpp.
// --------------------------------------------------------------------------
double erf_inv(double x); // oracle_capi.cpp

struct ErfInv : System {
  std::vector<double> position, possible_change;
  double mean_energy;
  ErfInv(size_t n, double mean) : position(n, 0.5), mean_energy(mean) {} // erfinv.rs:50-58
  double find_energy(const std::vector<double>& p) const {              // erfinv.rs:60-71
    double s = 0.0;
    for (double x : p) s += mean_energy + erf_inv(x);
    return s;
  }
  double energy() const override { return find_energy(position); }
  double compute_energy() const override { return energy(); }
  double randomize(Rng& rng) override { // erfinv.rs:78-83
    for (auto& x : position) x = rng.gen_range_f64(-1.0, 1.0);
    return energy();
  }
  bool plan_move(Rng& rng, double d, double* e) override { // erfinv.rs:101-110
    const size_t i = (size_t)rng.gen_range_usize(0, position.size());
    possible_change = position;
    const double v = rng.standard_normal();
    possible_change[i] += v * d;
    if (possible_change[i] >= 1.0 || possible_change[i] <= -1.0) return false;
    *e = find_energy(possible_change);
    return true;
  }
  void confirm() override { position = possible_change; } // erfinv.rs:95-97
  std::vector<double> get_state() const override { return position; }
  void set_state(const std::vector<double>& s) override { position = s; }
};

} // namespace oracle
