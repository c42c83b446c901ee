// host_tess.cpp -- C ABI of the host-side tessellation driver (libtess_b200_host.so, no CUDA).
// See include/tess_b200_host.h.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/tess_b200_host.h"
#include "delaunay3.hpp"

static thread_local std::string g_err;

extern "C" const char *tessb200_host_last_error(void) { return g_err.c_str(); }
void tessb200_host_set_error(const std::string &s) { g_err = s; }   // for block_file.cpp

extern "C" int tessb200_host_delaunay(int num_particles, const float *particles, int *num_tets, int **tets)
{
  if (!particles || !num_tets || !tets || num_particles < 0) { g_err = "NULL argument"; return -1; }
  *num_tets = 0;
  *tets = nullptr;
  try {
    tb_host::Delaunay3 d;
    if (!d.build(particles, num_particles)) return 0;    // fewer than 4 points in general position: no tets
    std::vector<int> out;
    d.export_tets(out);
    *num_tets = (int)(out.size() / 8);
    *tets = (int *)malloc(out.size() * sizeof(int) + 8);  // malloc: tess2 frees dblock_t::tets with free() (src/tess.cpp:170)
    if (!*tets) { g_err = "out of memory"; return -2; }
    memcpy(*tets, out.data(), out.size() * sizeof(int));
  } catch (const std::exception &e) {
    g_err = e.what();
    return -3;
  }
  return 0;
}

extern "C" void tessb200_host_free(void *p) { free(p); }

// ---- tess() for one process: blocks = originals + ghosts within a margin, widened until settled --------
#include <atomic>
#include <chrono>
#include <cmath>
#include <thread>

namespace
{

struct Sphere { double c[3], r; bool ok; };

// circumsphere in double (only used to decide whether the ghost region is wide enough)
Sphere circumsphere(const float *a, const float *b, const float *c, const float *d)
{
  const double ux = (double)b[0] - a[0], uy = (double)b[1] - a[1], uz = (double)b[2] - a[2];
  const double vx = (double)c[0] - a[0], vy = (double)c[1] - a[1], vz = (double)c[2] - a[2];
  const double wx = (double)d[0] - a[0], wy = (double)d[1] - a[1], wz = (double)d[2] - a[2];
  const double u2 = ux * ux + uy * uy + uz * uz, v2 = vx * vx + vy * vy + vz * vz, w2 = wx * wx + wy * wy + wz * wz;
  const double cvw[3] = {vy * wz - vz * wy, vz * wx - vx * wz, vx * wy - vy * wx};
  const double cwu[3] = {wy * uz - wz * uy, wz * ux - wx * uz, wx * uy - wy * ux};
  const double cuv[3] = {uy * vz - uz * vy, uz * vx - ux * vz, ux * vy - uy * vx};
  const double det = 2.0 * (ux * cvw[0] + uy * cvw[1] + uz * cvw[2]);
  Sphere s;
  s.ok = det != 0.0;
  if (!s.ok) { s.c[0] = s.c[1] = s.c[2] = 0; s.r = INFINITY; return s; }
  double o[3];
  for (int k = 0; k < 3; k++) o[k] = (u2 * cvw[k] + v2 * cwu[k] + w2 * cuv[k]) / det;
  s.r = std::sqrt(o[0] * o[0] + o[1] * o[1] + o[2] * o[2]);
  for (int k = 0; k < 3; k++) s.c[k] = a[k] + o[k];
  return s;
}

struct BlockJob
{
  int gid;
  double bmin[3], bmax[3];
  std::vector<int> mine;     // global ids of the originals, input order
};

// Tets in Morton order of their centroids.  The dense kernels chase tet records across neighbouring tets (star walks,
// edge links, the neighbour gathers of the circumcenter pass); the incremental engine numbers tets in creation order,
// with freed slots reused, which scatters the star of a site over the whole array.  Sorting them along a space-filling
// curve makes neighbours in space neighbours in memory (measured on a B200: profiles/r02/summary.md, "tet order").
// A renumbering only: same tets, same neighbour relation.
static void morton_order_tets(std::vector<int> &tets, const std::vector<float> &P)
{
  const size_t nt = tets.size() / 8;
  if (nt < 2 || nt >= (size_t)1 << 31) return;
  double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  std::vector<float> cen(3 * nt);
  for (size_t t = 0; t < nt; t++)
    for (int d = 0; d < 3; d++) {
      const int *v = &tets[8 * t];
      const float c = 0.25f * (P[3 * (size_t)v[0] + d] + P[3 * (size_t)v[1] + d] + P[3 * (size_t)v[2] + d] + P[3 * (size_t)v[3] + d]);
      cen[3 * t + d] = c;
      lo[d] = std::min(lo[d], (double)c);
      hi[d] = std::max(hi[d], (double)c);
    }
  auto spread = [](uint32_t x) {
    x &= 0x3ffu;
    x = (x | (x << 16)) & 0x030000ffu;
    x = (x | (x << 8)) & 0x0300f00fu;
    x = (x | (x << 4)) & 0x030c30c3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
  };
  std::vector<uint64_t> key(nt);
  for (size_t t = 0; t < nt; t++) {
    uint32_t q[3];
    for (int d = 0; d < 3; d++) {
      const double ext = std::max(hi[d] - lo[d], 1e-300);
      q[d] = (uint32_t)std::min(1023.0, std::max(0.0, ((double)cen[3 * t + d] - lo[d]) / ext * 1024.0));
    }
    key[t] = ((uint64_t)(spread(q[0]) | (spread(q[1]) << 1) | (spread(q[2]) << 2)) << 32) | (uint64_t)t;
  }
  std::sort(key.begin(), key.end());
  std::vector<int> inv(nt), out(8 * nt);
  for (size_t i = 0; i < nt; i++) inv[(size_t)(key[i] & 0xffffffffull)] = (int)i;
  for (size_t i = 0; i < nt; i++) {
    const int *src = &tets[8 * (size_t)(key[i] & 0xffffffffull)];
    for (int j = 0; j < 4; j++) {
      out[8 * i + j] = src[j];
      out[8 * i + 4 + j] = src[4 + j] < 0 ? src[4 + j] : inv[src[4 + j]];
    }
  }
  tets.swap(out);
}

// `wrap`: the domain is periodic in x, y and z (diy's wrap links, examples/tess/main.cpp:83-88): a block's ghosts are also the
// images of particles -- its own included -- shifted by whole domain extents, with the coordinates of wrap_pt (src/tess.cpp:698-710:
// float `x -= dir * (domain.max - domain.min)`).  An image keeps the id of its particle.  Code 13 = no shift.
constexpr int NO_SHIFT = 13;
inline void image_of(const float *q, int code, const float *ext, float *out)
{
  const int s[3] = {code % 3 - 1, (code / 3) % 3 - 1, code / 9 - 1};
  for (int d = 0; d < 3; d++) out[d] = s[d] == 0 ? q[d] : q[d] - (float)s[d] * ext[d];
}

int tess_block(const float *pts, int n, const int *owner, const BlockJob &job, const double *dmin, const double *dmax, double margin0,
               int max_rounds, double max_growth, bool wrap, const float *ext, tessb200_host_block *out)
{
  const int n_orig = (int)job.mine.size();
  if (margin0 <= 0.0) {
    const double vol = (job.bmax[0] - job.bmin[0]) * (job.bmax[1] - job.bmin[1]) * (job.bmax[2] - job.bmin[2]);
    margin0 = 3.0 * std::cbrt(vol / std::max(n_orig, 1));
  }
  double margin = margin0, prev_margin = -1.0;
  std::vector<float> P;
  std::vector<int> gids(job.mine.begin(), job.mine.end()), tets;
  std::vector<signed char> shift(job.mine.size(), (signed char)NO_SHIFT);     // image of every entry of gids
  int rounds = 0;
  bool settled = false;
  double secs = 0.0;
  tb_host::Delaunay3 dt;       // lives across the rounds: a wider margin only inserts the new ghosts
  for (;;) {
    rounds++;
    for (int i = 0; i < n; i++) {
      for (int code = wrap ? 0 : NO_SHIFT; code <= (wrap ? 26 : NO_SHIFT); code++) {
        if (code == NO_SHIFT && owner[i] == job.gid) continue;       // the original itself
        float qi[3];
        image_of(pts + 3 * (size_t)i, code, ext, qi);
        bool in = true, before = prev_margin >= 0.0;
        for (int d = 0; d < 3; d++) {
          in = in && (double)qi[d] >= job.bmin[d] - margin && (double)qi[d] <= job.bmax[d] + margin;
          before = before && (double)qi[d] >= job.bmin[d] - prev_margin && (double)qi[d] <= job.bmax[d] + prev_margin;
        }
        if (in && !before) { gids.push_back(i); shift.push_back((signed char)code); }
      }
    }
    const int np = (int)gids.size();
    const size_t had = P.size() / 3;
    P.resize(3 * (size_t)np);
    for (size_t i = had; i < (size_t)np; i++) image_of(pts + 3 * (size_t)gids[i], shift[i], ext, &P[3 * i]);
    const auto t0 = std::chrono::steady_clock::now();
    tets.clear();
    if (dt.add(P.data(), np)) dt.export_tets(tets);
    secs += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const size_t nt = tets.size() / 8;
    bool covered = true;
    for (int d = 0; d < 3; d++) covered = covered && job.bmin[d] - margin <= dmin[d] && job.bmax[d] + margin >= dmax[d];
    if (covered && !wrap) { settled = true; break; }     // (a periodic domain has no outside: its images always lie beyond)
    // originals on the local hull have unbounded cells; fine next to the domain boundary, a sign of too
    // few ghosts anywhere else
    std::vector<char> on_hull(np, 0);
    for (size_t t = 0; t < nt; t++)
      for (int k = 0; k < 4; k++)
        if (tets[8 * t + 4 + k] < 0)
          for (int j = 0; j < 4; j++) if (j != k) on_hull[tets[8 * t + j]] = 1;
    bool grow = false;
    for (int i = 0; i < n_orig && !grow; i++) {
      if (!on_hull[i]) continue;
      double dist = INFINITY;
      for (int d = 0; d < 3; d++) dist = std::min(dist, std::min((double)P[3 * (size_t)i + d] - dmin[d], dmax[d] - (double)P[3 * (size_t)i + d]));
      if (dist > margin || wrap) grow = true;            // periodic: no original may stay on the hull
    }
    // circumspheres of the tets at original, finite cells must stay inside the searched region
    // (clipped at the domain: nothing lives beyond it)
    double req = 0.0, req_in = 0.0;
    for (size_t t = 0; t < nt; t++) {
      const int *v = &tets[8 * t];
      bool need = false;
      for (int j = 0; j < 4; j++) need = need || (v[j] < n_orig && !on_hull[v[j]]);
      if (!need) continue;
      const Sphere s = circumsphere(&P[3 * (size_t)v[0]], &P[3 * (size_t)v[1]], &P[3 * (size_t)v[2]], &P[3 * (size_t)v[3]]);
      // a tet whose circumcenter lies outside the domain is a Voronoi vertex outside the data bounds: every cell it
      // belongs to is dropped by dense() (src/dense.cpp:1385-1392), so it cannot unsettle the block (`settled`); it
      // still drives the widening as before (flat tets at the domain boundary have spheres that reach far sideways)
      bool inside = true;
      for (int d = 0; d < 3; d++) inside = inside && (wrap || (s.c[d] >= dmin[d] && s.c[d] <= dmax[d]));
      for (int d = 0; d < 3; d++) {
        const double lo_need = job.bmin[d] - (s.c[d] - s.r), hi_need = (s.c[d] + s.r) - job.bmax[d];
        const double r_lo = wrap ? lo_need : std::min(lo_need, job.bmin[d] - dmin[d]), r_hi = wrap ? hi_need : std::min(hi_need, dmax[d] - job.bmax[d]);
        req = std::max(req, std::max(r_lo, r_hi));
        if (inside) req_in = std::max(req_in, std::max(r_lo, r_hi));
      }
    }
    if (grow) req = std::max(req, 2.0 * margin);
    settled = req_in <= margin && !grow;
    if (req <= margin) break;
    // the limits end the widening; `settled` says whether the test held for every tet that can matter
    if (rounds >= max_rounds || margin >= max_growth * margin0) break;
    prev_margin = margin;
    margin = std::min(req * 1.05, max_growth * margin0);
  }
  const int np = (int)gids.size();
  const size_t nt = tets.size() / 8;
  if (rounds > 1) {
    // ghosts in input order whatever round brought them in (the layout of a single-round run)
    std::vector<int> ord(np), pos(np);
    for (int i = 0; i < np; i++) ord[i] = i;
    std::sort(ord.begin() + n_orig, ord.end(), [&](int a, int b) { return gids[a] != gids[b] ? gids[a] < gids[b] : shift[a] < shift[b]; });
    std::vector<float> P2(P.size());
    std::vector<int> g2(np);
    for (int i = 0; i < np; i++) {
      pos[ord[i]] = i;
      g2[i] = gids[ord[i]];
      memcpy(&P2[3 * (size_t)i], &P[3 * (size_t)ord[i]], 12);
    }
    P.swap(P2);
    gids.swap(g2);
    for (size_t t = 0; t < nt; t++)
      for (int j = 0; j < 4; j++) tets[8 * t + j] = pos[tets[8 * t + j]];
  }
  if (!getenv("TESSB200_HOST_KEEP_TET_ORDER")) morton_order_tets(tets, P);
  out->gid = job.gid;
  for (int d = 0; d < 3; d++) { out->bounds_min[d] = (float)job.bmin[d]; out->bounds_max[d] = (float)job.bmax[d]; }
  out->num_orig_particles = n_orig;
  out->num_particles = np;
  out->num_tets = (int)nt;
  out->particles = (float *)malloc(sizeof(float) * 3 * (size_t)std::max(np, 1));
  out->tets = (int *)malloc(sizeof(int) * 8 * std::max<size_t>(nt, 1));
  out->vert_to_tet = (int *)malloc(sizeof(int) * (size_t)std::max(np, 1));
  out->global_ids = (int *)malloc(sizeof(int) * (size_t)std::max(np, 1));
  if (!out->particles || !out->tets || !out->vert_to_tet || !out->global_ids) return -2;
  if (np > 0) { memcpy(out->particles, P.data(), sizeof(float) * 3 * (size_t)np); memcpy(out->global_ids, gids.data(), sizeof(int) * (size_t)np); }
  if (nt) memcpy(out->tets, tets.data(), sizeof(int) * 8 * nt);
  // fill_vert_to_tet (src/tess.cpp:767-787): the last tet holding the vertex wins
  for (int i = 0; i < np; i++) out->vert_to_tet[i] = -1;
  for (size_t t = 0; t < nt; t++)
    for (int j = 0; j < 4; j++) out->vert_to_tet[tets[8 * t + j]] = (int)t;
  out->ghost_margin = (float)margin;
  out->rounds = rounds;
  out->settled = settled ? 1 : 0;
  out->reserved = 0;
  out->seconds = secs;
  return 0;
}

}  // namespace

extern "C" int tessb200_host_tess(int num_particles, const float *particles, const int *owner, const float *domain_min, const float *domain_max,
                                  int nblocks, const float *block_bounds, int num_gids, const int *gids, float margin0, int max_rounds,
                                  float max_growth, int num_threads, tessb200_host_block *blocks_out)
{
  return tessb200_host_tess_periodic(num_particles, particles, owner, domain_min, domain_max, nblocks, block_bounds, num_gids, gids, margin0, max_rounds,
                                     max_growth, num_threads, 0, blocks_out);
}

extern "C" int tessb200_host_tess_periodic(int num_particles, const float *particles, const int *owner, const float *domain_min, const float *domain_max,
                                           int nblocks, const float *block_bounds, int num_gids, const int *gids, float margin0, int max_rounds,
                                           float max_growth, int num_threads, int wrap, tessb200_host_block *blocks_out)
{
  if (!particles || !domain_min || !domain_max || !block_bounds || !blocks_out || nblocks < 1 || num_particles < 0) { g_err = "bad argument"; return -1; }
  const double dmin[3] = {domain_min[0], domain_min[1], domain_min[2]}, dmax[3] = {domain_max[0], domain_max[1], domain_max[2]};
  const float ext[3] = {domain_max[0] - domain_min[0], domain_max[1] - domain_min[1], domain_max[2] - domain_min[2]};      // float, as wrap_pt has it
  std::vector<BlockJob> jobs(nblocks);
  for (int b = 0; b < nblocks; b++) {
    jobs[b].gid = b;
    for (int d = 0; d < 3; d++) { jobs[b].bmin[d] = block_bounds[6 * b + d]; jobs[b].bmax[d] = block_bounds[6 * b + 3 + d]; }
  }
  std::vector<int> own;
  if (!owner) {
    own.assign(num_particles, -1);
    for (int i = 0; i < num_particles; i++) {
      const float *q = particles + 3 * (size_t)i;
      for (int b = 0; b < nblocks && own[i] < 0; b++) {
        bool in = true;
        for (int d = 0; d < 3; d++) {
          const bool top = jobs[b].bmax[d] >= dmax[d];     // the upper domain faces belong to the last block
          in = in && (double)q[d] >= jobs[b].bmin[d] && ((double)q[d] < jobs[b].bmax[d] || (top && (double)q[d] <= jobs[b].bmax[d]));
        }
        if (in) own[i] = b;
      }
      if (own[i] < 0) { g_err = "a particle lies in no block"; return -1; }
    }
    owner = own.data();
  }
  for (int i = 0; i < num_particles; i++) {
    if (owner[i] < 0 || owner[i] >= nblocks) { g_err = "owner gid out of range"; return -1; }
    jobs[owner[i]].mine.push_back(i);
  }
  std::vector<int> todo;
  if (gids) {
    for (int i = 0; i < num_gids; i++) {
      if (gids[i] < 0 || gids[i] >= nblocks) { g_err = "gid out of range"; return -1; }
      todo.push_back(gids[i]);
    }
  } else {
    for (int b = 0; b < nblocks; b++) todo.push_back(b);
  }
  const int ntodo = (int)todo.size();
  if (max_rounds <= 0) max_rounds = 3;
  if (!(max_growth > 0.0f)) max_growth = 2.5f;
  for (int b = 0; b < ntodo; b++) memset(&blocks_out[b], 0, sizeof(tessb200_host_block));
  int nthreads = num_threads > 0 ? num_threads : (int)std::thread::hardware_concurrency();
  nthreads = std::max(1, std::min(nthreads, ntodo));
  std::atomic<int> next(0), rc(0);
  std::vector<std::string> errs(nthreads);
  auto work = [&](int tid) {
    for (;;) {
      const int b = next.fetch_add(1);
      if (b >= ntodo) return;
      try {
        const int r = tess_block(particles, num_particles, owner, jobs[todo[b]], dmin, dmax, margin0, max_rounds, max_growth, wrap != 0, ext, &blocks_out[b]);
        if (r) { rc = r; errs[tid] = "out of memory"; }
      } catch (const std::exception &e) {
        rc = -3;
        errs[tid] = e.what();
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; t++) th.emplace_back(work, t);
  work(0);
  for (auto &t : th) t.join();
  if (rc) {
    for (auto &e : errs) if (!e.empty()) g_err = e;
    for (int b = 0; b < ntodo; b++) tessb200_host_free_block(&blocks_out[b]);
    return rc;
  }
  return 0;
}

extern "C" void tessb200_host_free_block(tessb200_host_block *b)
{
  if (!b) return;
  free(b->particles); free(b->tets); free(b->vert_to_tet); free(b->global_ids);
  b->particles = nullptr; b->tets = nullptr; b->vert_to_tet = nullptr; b->global_ids = nullptr;
}

// ---- decompositions -----------------------------------------------------------------------------------
#include <algorithm>
#include <numeric>

extern "C" int tessb200_host_regular_blocks(const float *domain_min, const float *domain_max, int nblocks, float *bounds_out)
{
  if (!domain_min || !domain_max || !bounds_out || nblocks < 1) { g_err = "bad argument"; return -1; }
  // factor nblocks as evenly as possible: largest prime factors first onto the smallest dimension
  int dims[3] = {1, 1, 1};
  std::vector<int> factors;
  for (int n = nblocks, f = 2; n > 1; f++)
    while (n % f == 0) { factors.push_back(f); n /= f; }
  std::sort(factors.rbegin(), factors.rend());
  for (int f : factors) *std::min_element(dims, dims + 3) *= f;
  std::sort(dims, dims + 3, std::greater<int>());
  int gid = 0;
  for (int k = 0; k < dims[2]; k++)
    for (int j = 0; j < dims[1]; j++)
      for (int i = 0; i < dims[0]; i++, gid++) {
        const int c[3] = {i, j, k};
        for (int d = 0; d < 3; d++) {
          const float lo = domain_min[d], ext = domain_max[d] - domain_min[d];
          bounds_out[6 * gid + d] = lo + ext * (float)c[d] / (float)dims[d];
          bounds_out[6 * gid + 3 + d] = c[d] == dims[d] - 1 ? domain_max[d] : lo + ext * (float)(c[d] + 1) / (float)dims[d];
        }
      }
  return 0;
}

extern "C" int tessb200_host_kdtree_blocks(int num_particles, const float *particles, const float *domain_min, const float *domain_max, int nblocks,
                                           float *bounds_out, int *owner_out)
{
  if (!particles || !domain_min || !domain_max || !bounds_out || nblocks < 1 || (nblocks & (nblocks - 1)) || num_particles < 0) {
    g_err = "bad argument (nblocks must be a power of two)";
    return -1;
  }
  struct Box { float mn[3], mx[3]; std::vector<int> idx; };
  std::vector<Box> boxes(1);
  for (int d = 0; d < 3; d++) { boxes[0].mn[d] = domain_min[d]; boxes[0].mx[d] = domain_max[d]; }
  boxes[0].idx.resize(num_particles);
  std::iota(boxes[0].idx.begin(), boxes[0].idx.end(), 0);
  // the boxes of a level are independent: split them on host threads (the first levels have fewer boxes than cores)
  auto split_box = [&](const Box &b, int d, Box &l, Box &r) {
    float split;
    const size_t n = b.idx.size();
    if (n >= 2) {
      std::vector<float> x(n);
      for (size_t i = 0; i < n; i++) x[i] = particles[3 * (size_t)b.idx[i] + d];
      const size_t m = n / 2;
      std::nth_element(x.begin(), x.begin() + m, x.end());
      const float hi = x[m], lo = *std::max_element(x.begin(), x.begin() + m);
      split = (float)(((double)lo + (double)hi) * 0.5);
      if (!(lo < split && split < hi)) split = hi;
    } else {
      split = (float)(((double)b.mn[d] + (double)b.mx[d]) * 0.5);
    }
    memcpy(l.mn, b.mn, 12); memcpy(l.mx, b.mx, 12); memcpy(r.mn, b.mn, 12); memcpy(r.mx, b.mx, 12);
    l.mx[d] = split; r.mn[d] = split;
    l.idx.reserve(n / 2 + 1); r.idx.reserve(n / 2 + 1);
    for (int id : b.idx) (particles[3 * (size_t)id + d] < split ? l.idx : r.idx).push_back(id);
  };
  const int hw = std::max(1, (int)std::thread::hardware_concurrency());
  for (int level = 0; (int)boxes.size() < nblocks; level++) {
    const int d = level % 3, nbox = (int)boxes.size();
    std::vector<Box> next((size_t)nbox * 2);
    std::atomic<int> cursor(0);
    auto work = [&]() {
      for (;;) {
        const int k = cursor.fetch_add(1);
        if (k >= nbox) return;
        split_box(boxes[k], d, next[2 * (size_t)k], next[2 * (size_t)k + 1]);
        std::vector<int>().swap(boxes[k].idx);
      }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < std::min(hw, nbox); t++) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
    boxes.swap(next);
  }
  for (int g = 0; g < nblocks; g++) {
    memcpy(bounds_out + 6 * g, boxes[g].mn, 12);
    memcpy(bounds_out + 6 * g + 3, boxes[g].mx, 12);
    if (owner_out) for (int id : boxes[g].idx) owner_out[id] = g;
  }
  return 0;
}
