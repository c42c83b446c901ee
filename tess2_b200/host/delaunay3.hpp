// delaunay3.hpp -- serial 3-D Delaunay engine of the host tess() driver (host_tess.cpp).
//
// Plays the part of the reference's local_cells()/gen_delaunay_output() adapters over Qhull or CGAL
// (src/tess-qhull.c:31-165, src/tess-cgal.cpp): points of one block in, tet_t records out
// (include/tess/tet.h:4-7: verts[4], tets[4] with tets[i] opposite verts[i], -1 on the hull).
// Neither library exists in this image, so the engine is the repo's own: incremental insertion
// (Bowyer-Watson) along a Morton curve inside randomised rounds, point location by walking from the
// last tet, conflict region by breadth-first search over the insphere predicate, the hull kept
// closed with ghost tets around one vertex at infinity.  Predicates are exact in sign
// (predicates.hpp), so the result is the Delaunay triangulation whenever the points are in general
// position; points that fall exactly on a circumsphere are kept out of it (any choice is a valid
// triangulation, as with Qhull's 'Qt').  Exact duplicates are skipped (Qhull drops them too; their
// vert_to_tet stays -1 and dense() skips them, src/dense.cpp:251).
#ifndef TESSB200_DELAUNAY3_HPP
#define TESSB200_DELAUNAY3_HPP

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <random>
#include <stdexcept>
#include <vector>

#include "predicates.hpp"

namespace tb_host
{

struct Tet
{
  int v[4];   // vertex ids, INF_V for the vertex at infinity
  int n[4];   // n[i] = tet across the face opposite v[i]
};

class Delaunay3
{
 public:
  static constexpr int INF_V = -1;

  // pts: n x 3 float32.  Returns false when fewer than 4 points in general position exist.
  bool build(const float *pts, int n, uint32_t seed = 12345u)
  {
    n_ = n;
    built_ = false;
    tets_.clear(); free_.clear(); mark_.clear();
    if (n < 4) return false;
    p_ = src_ = pts;
    seed_ = seed;
    insertion_order(order_, seed);
    // work on a copy of the points in insertion order (internal id = rank in the order): the vertices of
    // the tets around the point being inserted then sit close together in memory
    sorted_.resize(3 * (size_t)n);
    for (int i = 0; i < n; i++) memcpy(&sorted_[3 * (size_t)i], pts + 3 * (size_t)order_[i], 12);
    p_ = sorted_.data();
    tets_.reserve(7 * (size_t)n + 64);
    mark_.reserve(7 * (size_t)n + 64);
    // first tet: four points in general position, taken from the front of the order
    int i0 = 0, i1 = -1, i2 = -1, i3 = -1;
    for (int k = 1; k < n; k++) if (!same_point(i0, k)) { i1 = k; break; }
    if (i1 < 0) return false;
    for (int k = 1; k < n; k++) if (k != i1 && !collinear(i0, i1, k)) { i2 = k; break; }
    if (i2 < 0) return false;
    for (int k = 1; k < n; k++)
      if (k != i1 && k != i2 && orient3d(P(i0), P(i1), P(i2), P(k)) != 0) { i3 = k; break; }
    if (i3 < 0) return false;
    const int seed4[4] = {i0, i1, i2, i3};
    if (orient3d(P(i0), P(i1), P(i2), P(i3)) < 0) std::swap(i0, i1);
    first_tet(i0, i1, i2, i3);
    for (int id = 0; id < n; id++) {
      if (id == seed4[0] || id == seed4[1] || id == seed4[2] || id == seed4[3]) continue;
      insert(id);
    }
    built_ = true;
    return true;
  }

  // More points into the finished triangulation: pts is the caller's array grown to n points, its first points
  // unchanged (those already inserted).  Same result as build(pts, n) for points in general position.
  bool add(const float *pts, int n)
  {
    if (!built_) return build(pts, n, seed_);
    const int first = n_;
    if (n <= first) return true;
    src_ = pts;
    n_ = n;
    insertion_order(order_, seed_, first);
    sorted_.resize(3 * (size_t)n);
    for (int i = first; i < n; i++) memcpy(&sorted_[3 * (size_t)i], pts + 3 * (size_t)order_[i], 12);
    p_ = sorted_.data();
    for (int id = first; id < n; id++) insert(id);
    return true;
  }

  // finite tets in tet_t layout (8 ints each: verts, neighbours), hull neighbours -1
  void export_tets(std::vector<int> &out) const
  {
    std::vector<int> newid(tets_.size(), -1);
    int m = 0;
    for (size_t t = 0; t < tets_.size(); t++)
      if (alive(t) && finite(tets_[t])) newid[t] = m++;
    out.assign((size_t)m * 8, 0);
    for (size_t t = 0; t < tets_.size(); t++) {
      if (newid[t] < 0) continue;
      int *o = &out[(size_t)newid[t] * 8];
      for (int i = 0; i < 4; i++) {
        o[i] = order_[tets_[t].v[i]];          // back to the caller's numbering
        o[4 + i] = newid[tets_[t].n[i]];
      }
    }
  }
  size_t num_skipped() const { return skipped_; }
  // diagnostics: walk steps, conflict tests and cavity tets summed over all insertions
  size_t stat_walk = 0, stat_conflict = 0, stat_cavity = 0;

 private:
  const float *p_ = nullptr;         // the points in insertion order (sorted_) once build() has copied them
  const float *src_ = nullptr;       // the caller's array
  float all_lo_[3] = {0, 0, 0}, all_hi_[3] = {0, 0, 0};
  uint32_t seed_ = 12345u;
  bool built_ = false;
  int n_ = 0;
  std::vector<int> order_;           // internal id -> index in the caller's array
  std::vector<float> sorted_;        // the points in insertion order
  std::vector<Tet> tets_;
  std::vector<int> free_;
  std::vector<uint32_t> mark_;       // per tet: epoch * 4 + (in cavity list) * 2 + conflict bit of the current insertion
  uint32_t epoch_ = 0;
  int last_ = 0;
  size_t skipped_ = 0;
  std::mt19937 rng_;
  StaticFilter filter_;
  // scratch of one insertion
  std::vector<int> cavity_, stack_;
  struct BFace { int t, i, nt; };
  std::vector<BFace> boundary_;
  struct EdgeSlot { uint64_t key; int tetslot; uint32_t epoch; };   // tetslot = tet * 4 + slot, -1 once matched
  std::vector<EdgeSlot> edges_;

  const float *P(int i) const { return p_ + 3 * (size_t)i; }
  bool same_point(int a, int b) const { return P(a)[0] == P(b)[0] && P(a)[1] == P(b)[1] && P(a)[2] == P(b)[2]; }
  bool collinear(int a, int b, int c) const
  {
    // cross product of exact differences in double is enough for a seed choice: only used to pick a non-degenerate start
    const double ux = (double)P(b)[0] - P(a)[0], uy = (double)P(b)[1] - P(a)[1], uz = (double)P(b)[2] - P(a)[2];
    const double vx = (double)P(c)[0] - P(a)[0], vy = (double)P(c)[1] - P(a)[1], vz = (double)P(c)[2] - P(a)[2];
    const double cx = uy * vz - uz * vy, cy = uz * vx - ux * vz, cz = ux * vy - uy * vx;
    const double scale = (std::fabs(ux) + std::fabs(uy) + std::fabs(uz)) * (std::fabs(vx) + std::fabs(vy) + std::fabs(vz));
    return std::fabs(cx) + std::fabs(cy) + std::fabs(cz) <= 1e-9 * scale;
  }
  static bool finite(const Tet &t) { return t.v[0] >= 0 && t.v[1] >= 0 && t.v[2] >= 0 && t.v[3] >= 0; }
  bool alive(size_t t) const { return tets_[t].v[0] != -2; }

  // Morton order inside rounds of doubling size (biased randomised insertion order): the walk from
  // the previous tet stays short and no input order can make the insertion quadratic
  // The order covers the caller's points [first, n_): build() passes 0, add() the first new point.
  void insertion_order(std::vector<int> &order, uint32_t seed, int first = 0)
  {
    const float *q = src_;
    float lo[3] = {q[3 * (size_t)first], q[3 * (size_t)first + 1], q[3 * (size_t)first + 2]}, hi[3] = {lo[0], lo[1], lo[2]};
    for (int i = first + 1; i < n_; i++)
      for (int d = 0; d < 3; d++) { lo[d] = std::min(lo[d], q[3 * (size_t)i + d]); hi[d] = std::max(hi[d], q[3 * (size_t)i + d]); }
    double inv[3];
    for (int d = 0; d < 3; d++) inv[d] = hi[d] > lo[d] ? 2097151.0 / ((double)hi[d] - lo[d]) : 0.0;
    // the static filter must hold for every pair of points seen so far
    for (int d = 0; d < 3; d++) { all_lo_[d] = first ? std::min(all_lo_[d], lo[d]) : lo[d]; all_hi_[d] = first ? std::max(all_hi_[d], hi[d]) : hi[d]; }
    double extent = 0.0;
    for (int d = 0; d < 3; d++) extent = std::max(extent, (double)all_hi_[d] - all_lo_[d]);
    filter_.set_extent(extent);
    auto spread = [](uint64_t x) {
      x &= 0x1fffff;
      x = (x | x << 32) & 0x1f00000000ffffull;
      x = (x | x << 16) & 0x1f0000ff0000ffull;
      x = (x | x << 8) & 0x100f00f00f00f00full;
      x = (x | x << 4) & 0x10c30c30c30c30c3ull;
      x = (x | x << 2) & 0x1249249249249249ull;
      return x;
    };
    rng_.seed(seed + (uint32_t)first);
    std::vector<std::pair<uint64_t, int> > key((size_t)(n_ - first));
    for (int i = first; i < n_; i++) {
      uint64_t m = 0;
      for (int d = 0; d < 3; d++) m |= spread((uint64_t)(((double)q[3 * (size_t)i + d] - lo[d]) * inv[d])) << d;
      // round = number of trailing coin flips: the last round holds half the points, the one before a quarter, ...
      uint32_t r = rng_();
      int round = 0;
      while ((r & 1u) && round < 20) { r >>= 1; round++; }
      key[i - first] = std::make_pair(((uint64_t)(20 - round) << 58) | (m >> 5), i);   // rarer rounds first, Morton order inside a round
    }
    std::sort(key.begin(), key.end());
    order.resize(n_);
    for (int i = first; i < n_; i++) order[i] = key[i - first].second;
  }

  int new_tet()
  {
    int t;
    if (!free_.empty()) { t = free_.back(); free_.pop_back(); }
    else { t = (int)tets_.size(); tets_.push_back(Tet()); mark_.push_back(0); }
    return t;
  }

  void first_tet(int a, int b, int c, int d)
  {
    // finite tet 0 (positively oriented) and the four ghosts over its faces.  Convention for a ghost
    // (one vertex is INF_V): the tet with INF_V replaced by a point p is positively oriented exactly
    // when p lies strictly beyond the hull face.  Replacing v[i] of a positive tet by a point on the
    // far side of face i makes it negative, so the ghost swaps two of the remaining vertices.
    tets_.resize(5);
    mark_.assign(5, 0);
    const int v[4] = {a, b, c, d};
    for (int i = 0; i < 4; i++) { tets_[0].v[i] = v[i]; tets_[0].n[i] = 1 + i; }
    for (int i = 0; i < 4; i++) {
      Tet &g = tets_[1 + i];
      for (int j = 0; j < 4; j++) g.v[j] = v[j];
      g.v[i] = INF_V;
      const int s0 = (i + 1) & 3, s1 = (i + 2) & 3;
      std::swap(g.v[s0], g.v[s1]);
      g.n[i] = 0;
      // across the face opposite the finite vertex x = g.v[s] lies the ghost over tet 0's face opposite x
      for (int s = 0; s < 4; s++) {
        if (s == i) continue;
        for (int j = 0; j < 4; j++) if (v[j] == g.v[s]) g.n[s] = 1 + j;
      }
    }
    last_ = 0;
  }

  // > 0: p conflicts with tet t (inside the circumsphere / beyond the hull face)
  int conflict(int t, const float *p) const
  {
    const Tet &T = tets_[t];
    if (finite(T)) return insphere(P(T.v[0]), P(T.v[1]), P(T.v[2]), P(T.v[3]), p, filter_.sphere);
    int k = T.v[0] < 0 ? 0 : (T.v[1] < 0 ? 1 : (T.v[2] < 0 ? 2 : 3));
    const float *q[4];
    for (int i = 0; i < 4; i++) q[i] = i == k ? p : P(T.v[i]);
    const int o = orient3d(q[0], q[1], q[2], q[3], filter_.orient);
    if (o != 0) return o;
    // p in the plane of the hull face: conflict iff inside the face's circumcircle, i.e. inside the
    // circumsphere of the finite tet behind the face
    const Tet &F = tets_[T.n[k]];
    return insphere(P(F.v[0]), P(F.v[1]), P(F.v[2]), P(F.v[3]), p, filter_.sphere);
  }

  // walk towards p from the last tet; returns a tet in conflict with p, or -1 for a duplicate point
  int locate(int id)
  {
    const float *p = P(id);
    int t = last_;
    if (!alive(t)) t = 0;
    for (size_t steps = 0; steps < tets_.size() * 4 + 64; steps++) {
      const Tet &T = tets_[t];
      stat_walk++;
      if (!finite(T)) return t;     // p is outside the hull (or on it): the ghost or a neighbour conflicts
      int start = (int)(rng_() & 3u), moved = 0;
      for (int s = 0; s < 4; s++) {
        const int i = (start + s) & 3;
        const float *q[4] = {P(T.v[0]), P(T.v[1]), P(T.v[2]), P(T.v[3])};
        q[i] = p;
        if (orient3d(q[0], q[1], q[2], q[3], filter_.orient) < 0) { t = T.n[i]; moved = 1; break; }
      }
      if (!moved) return t;
    }
    throw std::runtime_error("delaunay3: point location did not terminate");
  }

  void insert(int id)
  {
    const float *p = P(id);
    int t0 = locate(id);
    // duplicates: p equals a vertex of the located tet
    {
      const Tet &T = tets_[t0];
      for (int i = 0; i < 4; i++)
        if (T.v[i] >= 0 && same_point(T.v[i], id)) { skipped_++; return; }
    }
    epoch_++;
    if (epoch_ >= (1u << 30)) { std::fill(mark_.begin(), mark_.end(), 0u); epoch_ = 1; }
    auto state = [&](int t) -> int {          // 1 conflict, 0 no conflict (cached per insertion)
      if ((mark_[t] >> 2) == epoch_) return (int)(mark_[t] & 1u);
      const int c = conflict(t, p) > 0 ? 1 : 0;
      stat_conflict++;
      mark_[t] = epoch_ << 2 | (uint32_t)c;
      return c;
    };
    if (!state(t0)) {
      // p on the boundary of the located tet, or a hull ghost that does not see p: look around it
      int found = -1;
      for (int i = 0; i < 4 && found < 0; i++) if (state(tets_[t0].n[i])) found = tets_[t0].n[i];
      if (found < 0) {
        // second ring (p on an edge or a vertex of the located tet)
        for (int i = 0; i < 4 && found < 0; i++) {
          const int u = tets_[t0].n[i];
          for (int j = 0; j < 4 && found < 0; j++) if (state(tets_[u].n[j])) found = tets_[u].n[j];
        }
      }
      if (found < 0) { skipped_++; return; }    // cospherical with everything around: leave the point out
      t0 = found;
    }
    // conflict region
    cavity_.clear(); boundary_.clear(); stack_.clear();
    stack_.push_back(t0);
    cavity_.push_back(t0);
    mark_[t0] |= 2u;          // listed (the bit dies with the epoch)
    while (!stack_.empty()) {
      const int t = stack_.back();
      stack_.pop_back();
      for (int i = 0; i < 4; i++) {            // the four neighbours are about to be tested: start their loads together
        __builtin_prefetch(&tets_[tets_[t].n[i]]);
        __builtin_prefetch(&mark_[tets_[t].n[i]]);
      }
      for (int i = 0; i < 4; i++) {
        const int u = tets_[t].n[i];
        if (state(u)) {
          if (!(mark_[u] & 2u)) { mark_[u] |= 2u; cavity_.push_back(u); stack_.push_back(u); }
        } else {
          BFace bf;
          bf.t = t; bf.i = i; bf.nt = -1;
          boundary_.push_back(bf);
        }
      }
    }
    // new tets: boundary face + p.  Pass 1 creates them and leaves, in the cavity tet's neighbour slot of that
    // face, the new tet's id (as ~id < 0); pass 2 links the new tets to each other by turning around each edge
    // of the cavity boundary through the cavity tets until the next boundary face comes up.
    int first_new = -1;
    for (BFace &bf : boundary_) {
      const int nt = new_tet();
      Tet N = tets_[bf.t];
      const int outside = N.n[bf.i];
      N.v[bf.i] = id;
      N.n[0] = N.n[1] = N.n[2] = N.n[3] = -1;
      N.n[bf.i] = outside;
      tets_[nt] = N;
      mark_[nt] = 0;
      // the outside tet now faces the new one
      Tet &O = tets_[outside];
      const int hits = (O.n[0] == bf.t) + (O.n[1] == bf.t) + (O.n[2] == bf.t) + (O.n[3] == bf.t);
      if (hits == 1) {
        O.n[O.n[0] == bf.t ? 0 : (O.n[1] == bf.t ? 1 : (O.n[2] == bf.t ? 2 : 3))] = nt;
      } else {
        for (int j = 0; j < 4; j++) if (O.n[j] == bf.t && shares_face(O, j, N, bf.i)) O.n[j] = nt;
      }
      tets_[bf.t].n[bf.i] = ~nt;
      bf.nt = nt;
      if (first_new < 0) first_new = nt;
    }
    for (const BFace &bf : boundary_) {
      Tet &N = tets_[bf.nt];
      for (int j = 0; j < 4; j++) {
        if (j == bf.i || N.n[j] >= 0) continue;
        // edge (a, b) = the boundary face minus v[j]; leave the cavity tet through the face opposite v[j]
        int a = -2, b = -2;
        for (int m = 0; m < 4; m++)
          if (m != bf.i && m != j) { if (a == -2) a = N.v[m]; else b = N.v[m]; }
        int cur = bf.t, exit = j;
        for (int steps = 0;; steps++) {
          const int nxt = tets_[cur].n[exit];
          if (nxt < 0) break;                                     // (cur, exit) is a boundary face
          if (steps > 100000) throw std::runtime_error("delaunay3: cavity boundary is not a manifold");
          const Tet &X = tets_[nxt];
          int s1 = -1, s2 = -1;
          for (int m = 0; m < 4; m++)
            if (X.v[m] != a && X.v[m] != b) { if (s1 < 0) s1 = m; else s2 = m; }
          exit = X.n[s1] == cur ? s2 : s1;
          cur = nxt;
        }
        const int other = ~tets_[cur].n[exit];
        // in the new tet over (cur, exit) the face that holds p, a, b is opposite the vertex that is neither
        // a nor b nor at slot `exit` (the new tet keeps the slots of the cavity tet it came from)
        const Tet &C = tets_[cur];
        int k = -1;
        for (int m = 0; m < 4; m++) if (m != exit && C.v[m] != a && C.v[m] != b) k = m;
        N.n[j] = other;
        tets_[other].n[k] = bf.nt;
      }
    }
    stat_cavity += cavity_.size();
    for (int t : cavity_) { tets_[t].v[0] = -2; free_.push_back(t); }
    last_ = first_new;
    // walks start from a finite tet when possible
    if (!finite(tets_[last_])) {
      for (int j = 0; j < 4; j++) { const int u = tets_[last_].n[j]; if (u >= 0 && finite(tets_[u])) { last_ = u; break; } }
    }
  }

  // ---- small helpers -----------------------------------------------------------------------------
  uint32_t edge_epoch_ = 0;
  static bool shares_face(const Tet &A, int ia, const Tet &B, int ib)
  {
    // the face of A opposite slot ia equals the face of B opposite slot ib (as vertex sets)
    for (int m = 0; m < 4; m++) {
      if (m == ia) continue;
      bool f = false;
      for (int q = 0; q < 4; q++) if (q != ib && B.v[q] == A.v[m]) f = true;
      if (!f) return false;
    }
    return true;
  }
  void link_edge(int a, int b, int tet, int slot)
  {
    if (a > b) std::swap(a, b);
    const uint64_t key = ((uint64_t)(uint32_t)(a + 1) << 32) | (uint32_t)(b + 1);
    const size_t mask = edges_.size() - 1;
    size_t h = (size_t)((key * 0x9E3779B97F4A7C15ull) >> 20) & mask;
    for (;;) {
      EdgeSlot &s = edges_[h];
      if (s.epoch != edge_epoch_) { s.key = key; s.tetslot = tet * 4 + slot; s.epoch = edge_epoch_; return; }
      if (s.key == key && s.tetslot >= 0) {
        tets_[tet].n[slot] = s.tetslot >> 2;
        tets_[s.tetslot >> 2].n[s.tetslot & 3] = tet;
        s.tetslot = -1;  // matched (an edge of the cavity boundary is shared by exactly two boundary faces)
        return;
      }
      h = (h + 1) & mask;
    }
  }
};

}  // namespace tb_host

#endif
