// cell_core.cuh -- per-tet / per-cell arithmetic of the dense stage, shared by every kernel.
//
// Everything here is written so that the fp32 results equal the reference's default
// x86-64 Release build bit for bit: fp32 add/mul/div/sqrt in the
// reference's operation order, no FMA contraction (the whole library is compiled with
// -fmad=false; the explicit fadd/fmul helpers below make the intent visible and keep the
// host-compiled test build, which uses -ffp-contract=off, identical).
//
// Functions are __host__ __device__ so that tests/emul (a CPU build of the same logic, test
// only, never part of libtess_b200.so) can single-step them against the oracle.
#pragma once
#include <stdint.h>
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define TB_HD __host__ __device__ __forceinline__
#else
#define TB_HD inline
struct int4 { int x, y, z, w; };
struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
#endif

namespace tb
{

// ---- rounding-exact scalar helpers -------------------------------------------------------
TB_HD float fadd(float a, float b)
{
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
TB_HD float fsub(float a, float b)
{
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  return a - b;
#endif
}
TB_HD float fmul(float a, float b)
{
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
TB_HD float fdiv(float a, float b)
{
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}
TB_HD float fsqrt(float a)
{
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(a);
#else
  return sqrtf(a);
#endif
}

// ---- grid <-> physical -------------------------------------------------------------------
struct GridGeom
{
  float gmin[3];   // grid_phys_mins
  float step[3];   // grid_step_size
  float dmin[3];   // data_mins
  float dmax[3];   // data_maxs
  float dext_eps[3]; // (data_max - data_min) * 2.0f * FLT_EPSILON   (src/dense.cpp:1382-1383)
  int gnum[3];     // glo_num_idx
  float eps;
  float mass;
  float div;       // step_x * step_y [* step_z]   (src/dense.cpp:90-91)
  int project;
  int alg;
};

// idx2phys, src/dense.cpp:1098-1106: pos = idx * step + min (two roundings)
TB_HD float idx2phys1(int idx, float step, float gmin) { return fadd(fmul((float)idx, step), gmin); }
// phys2idx, src/dense.cpp:1117-1125: truncation toward zero of (pos - min) / step
TB_HD int phys2idx1(float pos, float step, float gmin) { return (int)fdiv(fsub(pos, gmin), step); }

// ---- circumcenter, src/tet.cpp:37-66 (norm :122-128, cross :131-136, determinant :139-143) ----
TB_HD void circumcenter(const float *a, const float *b, const float *c, const float *d, float *center, float *det_out = nullptr)
{
  float t[3], u[3], v[3];
  for (int i = 0; i < 3; i++) {
    t[i] = fsub(a[i], d[i]);
    u[i] = fsub(b[i], d[i]);
    v[i] = fsub(c[i], d[i]);
  }
  // norm(): res = 0; res += x[i]*x[i]
  float nt = fadd(fadd(fadd(0.0f, fmul(t[0], t[0])), fmul(t[1], t[1])), fmul(t[2], t[2]));
  float nu = fadd(fadd(fadd(0.0f, fmul(u[0], u[0])), fmul(u[1], u[1])), fmul(u[2], u[2]));
  float nv = fadd(fadd(fadd(0.0f, fmul(v[0], v[0])), fmul(v[1], v[1])), fmul(v[2], v[2]));
  // determinant(): t0*u1*v2 + u0*v1*t2 + v0*t1*u2 - v0*u1*t2 - u0*t1*v2 - t0*v1*u2, left to right
  float det = fmul(fmul(t[0], u[1]), v[2]);
  det = fadd(det, fmul(fmul(u[0], v[1]), t[2]));
  det = fadd(det, fmul(fmul(v[0], t[1]), u[2]));
  det = fsub(det, fmul(fmul(v[0], u[1]), t[2]));
  det = fsub(det, fmul(fmul(u[0], t[1]), v[2]));
  det = fsub(det, fmul(fmul(t[0], v[1]), u[2]));
  if (det_out) *det_out = det;
  float den = fmul(2.0f, det);
  float uv[3], vt[3], tu[3];
  uv[0] = fsub(fmul(u[1], v[2]), fmul(u[2], v[1]));
  uv[1] = fsub(fmul(u[2], v[0]), fmul(u[0], v[2]));
  uv[2] = fsub(fmul(u[0], v[1]), fmul(u[1], v[0]));
  vt[0] = fsub(fmul(v[1], t[2]), fmul(v[2], t[1]));
  vt[1] = fsub(fmul(v[2], t[0]), fmul(v[0], t[2]));
  vt[2] = fsub(fmul(v[0], t[1]), fmul(v[1], t[0]));
  tu[0] = fsub(fmul(t[1], u[2]), fmul(t[2], u[1]));
  tu[1] = fsub(fmul(t[2], u[0]), fmul(t[0], u[2]));
  tu[2] = fsub(fmul(t[0], u[1]), fmul(t[1], u[0]));
  for (int i = 0; i < 3; i++) {
    float num = fadd(fadd(fmul(nt, uv[i]), fmul(nu, vt[i])), fmul(nv, tu[i]));
    center[i] = fadd(d[i], fdiv(num, den));
  }
}

// ---- one plane test of PtInCell, src/dense.cpp:1187-1199 --------------------------------
// returns +1 / -1 for a significant distance, 0 when |dist| <= eps (NaN counts as 0)
TB_HD int plane_side(const float *n, const float *f, const float *pt, float eps)
{
  float dist = fadd(fadd(fmul(n[0], fsub(pt[0], f[0])), fmul(n[1], fsub(pt[1], f[1]))), fmul(n[2], fsub(pt[2], f[2])));
  if (fabsf(dist) > eps) return dist >= 0.0f ? 1 : -1;
  return 0;
}

// ---- Newell normal accumulation, src/dense.cpp:1139-1149, one (cur, next) term ----------
TB_HD void newell_term(float *nrm, const float *cur, const float *nxt)
{
  nrm[0] = fadd(nrm[0], fmul(fsub(cur[1], nxt[1]), fadd(cur[2], nxt[2])));
  nrm[1] = fadd(nrm[1], fmul(fsub(cur[2], nxt[2]), fadd(cur[0], nxt[0])));
  nrm[2] = fadd(nrm[2], fmul(fsub(cur[0], nxt[0]), fadd(cur[1], nxt[1])));
}
// normalise + invert (src/dense.cpp:1151-1161), then orient away from the site (:722-731)
TB_HD void newell_finish(float *nrm, const float *v0, const float *site)
{
  float mag = fsqrt(fadd(fadd(fmul(nrm[0], nrm[0]), fmul(nrm[1], nrm[1])), fmul(nrm[2], nrm[2])));
  for (int i = 0; i < 3; i++) nrm[i] = -fdiv(nrm[i], mag);
  float v[3] = {fsub(v0[0], site[0]), fsub(v0[1], site[1]), fsub(v0[2], site[2])};
  float dotp = fadd(fadd(fmul(v[0], nrm[0]), fmul(v[1], nrm[1])), fmul(v[2], nrm[2]));
  if (dotp < 0.0f) { nrm[0] = -nrm[0]; nrm[1] = -nrm[1]; nrm[2] = -nrm[2]; }
}

// ---- star walk -----------------------------------------------------------------------------
// Status of one cell after the topology pass
enum CellStatus
{
  CELL_OK = 0,
  CELL_NO_TET = 1,      // vert_to_tet == -1                       (src/dense.cpp:251)
  CELL_INCOMPLETE = 2,  // complete() == 0                          (src/dense.cpp:252, src/tet.cpp:337-378)
  CELL_OUTSIDE = 3,     // bbox outside the data bounds             (src/dense.cpp:1385-1392)
  CELL_OVERFLOW = 4,    // star / neighbour list exceeds the workspace: retry in the large workspace
  CELL_BAD_MESH = 5     // circulation did not close within the step limit
};

// Workspace: three int arrays reached through accessors so that the same code runs on a
// shared-memory slice (stride = blockDim), a global slice or plain host arrays.
//   star(i) : tets of the star of the site in BFS order (doubles as the FIFO queue)
//   nu(i)   : Delaunay neighbours in first-seen order;  nt(i): the star tet where nu(i) was first seen
template <class WS>
TB_HD int star_and_neighbors(int site, int t0, const int4 *tets, WS &ws, int star_cap, int nbr_cap, int *n_star, int *n_nbr)
{
  // src/tet.cpp:228-270 (neighbor_edges) + :337-378 (complete).  The reference pushes every
  // neighbour tet and skips repeats when they are popped; with a FIFO queue the order of first
  // pops equals the order of first pushes, so repeats are dropped at push time and the queue
  // is the visited list itself.
  int ns = 1, nn = 0;
  ws.star(0) = t0;
  for (int head = 0; head < ns; head++) {
    int t = ws.star(head);
    int4 v = tets[2 * (size_t)t];
    int4 nb = tets[2 * (size_t)t + 1];
    int vv[4] = {v.x, v.y, v.z, v.w};
    int bb[4] = {nb.x, nb.y, nb.z, nb.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int u = vv[i];
      if (u == site) continue;
      bool seen = false;
      for (int j = 0; j < nn; j++)
        if (ws.nu(j) == u) { seen = true; break; }
      if (!seen) {
        if (nn >= nbr_cap) return CELL_OVERFLOW;
        ws.nu(nn) = u;
        ws.nt(nn) = t;
        nn++;
      }
      int next = bb[i];
      if (next < 0) return CELL_INCOMPLETE;
      seen = false;
      for (int j = 0; j < ns; j++)
        if (ws.star(j) == next) { seen = true; break; }
      if (!seen) {
        if (ns >= star_cap) return CELL_OVERFLOW;
        ws.star(ns++) = next;
      }
    }
  }
  *n_star = ns;
  *n_nbr = nn;
  return CELL_OK;
}

// Same result as star_and_neighbors (same BFS order, same first-seen tets), but the two
// "already seen?" questions are answered by 64-slot open-addressing byte tables instead of linear
// searches (57% of k_cell_topo's instructions in the first profile, profiles/r01_*).  The tables
// hold indices into star() / nu(); 0xFF = empty.  Requires star_cap, nbr_cap < 64.
//   ws.vis_hash(h), ws.nbr_hash(h): uint8 slots;  ws.nt_idx(i): uint8 index into star() of the tet
//   where nu(i) was first seen;  ws.hash_clear() empties both tables.
template <class WS>
TB_HD int hash_find_or_insert(WS &ws, bool vertex_table, int key, int *count, int cap)
{
  unsigned h = ((unsigned)key * 0x9E3779B1u) >> 26;
  for (;;) {
    unsigned idx = vertex_table ? ws.nbr_hash(h) : ws.vis_hash(h);
    if (idx == 0xFFu) {
      if (*count >= cap) return -1;
      if (vertex_table) { ws.nbr_hash(h) = (unsigned char)*count; ws.nu(*count) = key; }
      else { ws.vis_hash(h) = (unsigned char)*count; ws.star(*count) = key; }
      (*count)++;
      return 1;
    }
    int have = vertex_table ? ws.nu((int)idx) : ws.star((int)idx);
    if (have == key) return 0;
    h = (h + 1u) & 63u;
  }
}

// In a manifold star the tet popped at `head` shares with its BFS parent the face opposite the slot
// whose neighbour is the parent; the other vertices of that face were recorded when the parent was
// popped, so only the vertex AT that slot can be new, and the parent itself needs no visited test.
// That halves the table operations (3 instead of 6 per popped tet).  ws.parent_idx(i): uint8 index
// into star() of the tet from which star(i) was first pushed.
template <class WS>
TB_HD int star_and_neighbors_hashed(int site, int t0, const int4 *tets, WS &ws, int star_cap, int nbr_cap, int *n_star, int *n_nbr)
{
  ws.hash_clear();
  int ns = 0, nn = 0;
  hash_find_or_insert(ws, false, t0, &ns, star_cap);
  for (int head = 0; head < ns; head++) {
    int t = ws.star(head);
    int4 v = tets[2 * (size_t)t];
    int4 nb = tets[2 * (size_t)t + 1];
    int vv[4] = {v.x, v.y, v.z, v.w};
    int bb[4] = {nb.x, nb.y, nb.z, nb.w};
    const int par = head ? ws.star((int)ws.parent_idx(head)) : -2;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int u = vv[i];
      if (u == site) continue;
      int next = bb[i];
      const bool to_parent = next == par;
      if (head == 0 || to_parent) {
        int before = nn;
        int r = hash_find_or_insert(ws, true, u, &nn, nbr_cap);
        if (r < 0) return CELL_OVERFLOW;
        if (r > 0) ws.nt_idx(before) = (unsigned char)head;
      }
      if (next < 0) return CELL_INCOMPLETE;
      if (to_parent) continue;
      int before_s = ns;
      int r2 = hash_find_or_insert(ws, false, next, &ns, star_cap);
      if (r2 < 0) return CELL_OVERFLOW;
      if (r2 > 0) ws.parent_idx(before_s) = (unsigned char)head;
    }
  }
  *n_star = ns;
  *n_nbr = nn;
  return CELL_OK;
}

// hint: bring a line the walk will need a few visits from now into L1 / L2 (no register, no dependency)
TB_HD void tb_prefetch(const void *p)
{
#if defined(__CUDA_ARCH__)
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

TB_HD int tb_clz(uint32_t x)
{
#if defined(__CUDA_ARCH__)
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}
TB_HD int tb_ffs(uint32_t x)
{
#if defined(__CUDA_ARCH__)
  return __ffs((int)x);
#else
  return __builtin_ffs((int)x);
#endif
}

TB_HD int tb_sel4(int a, int b, int c, int d, int s) { return s == 0 ? a : (s == 1 ? b : (s == 2 ? c : d)); }

// The star walk split in two (k_cell_bfs + k_cell_nbrs): the BFS keeps only the visited set and hands
// every CANDIDATE new neighbour to `sink(k, u, t)` in BFS order -- the root's three vertices, then per
// popped tet the vertex opposite its parent face (the only one that can be new: the other vertices of
// the popped tet belong to the parent, which was popped earlier).  A star that is not a manifold around
// `site` (no slot leads back to the BFS parent) is reported as CELL_OVERFLOW, which reroutes the cell
// to the general (linear-search) kernel.  A second
// pass (nbrs_from_cands) drops the candidates that were seen before; what remains is neighbor_edges'
// list of (u, first tet holding u).  ws needs star(), parent_idx(), vis_hash(), hash_clear_vis().
// Bucketized byte tables: a bucket is one 32-bit word holding four list indices (0xFF = empty, filled
// from byte 0 up; nothing is ever deleted).  One shared-memory load fetches the bucket, the up to
// four list entries it names are compared without data-dependent branching, and a full bucket
// spills to the next one.  Compared with one-byte open addressing this keeps the lanes of a warp on
// the same path (profiles/r01_i: the probe-continuation code ran at 3.6 of 32 lanes).
//   vis table: VIS_BUCKETS buckets over star();  nbr table: NBR_BUCKETS buckets over nu().
#define TB_VIS_BUCKETS 16
#define TB_NBR_BUCKETS 16

template <class WS>
TB_HD int vis_find_or_insert(WS &ws, int key, int *count, int cap)
{
  unsigned h = ((unsigned)key * 0x9E3779B1u) >> 28;      // 16 buckets
  for (int guard = 0; guard < TB_VIS_BUCKETS; guard++) {
    const uint32_t w = ws.vis_word(h);
    const unsigned i0 = w & 0xFFu, i1 = (w >> 8) & 0xFFu, i2 = (w >> 16) & 0xFFu, i3 = w >> 24;
    // entries are filled from byte 0 up, so an empty byte ends the bucket
    // branch-free: the four list entries are loaded unconditionally (an empty byte reads entry 0 and
    // is masked out), so the lanes of a warp do not split on how full their buckets are
    const int e0 = ws.star(i0 == 0xFFu ? 0 : (int)i0), e1 = ws.star(i1 == 0xFFu ? 0 : (int)i1);
    const int e2 = ws.star(i2 == 0xFFu ? 0 : (int)i2), e3 = ws.star(i3 == 0xFFu ? 0 : (int)i3);
    const bool hit = ((i0 != 0xFFu) & (e0 == key)) | ((i1 != 0xFFu) & (e1 == key)) | ((i2 != 0xFFu) & (e2 == key)) | ((i3 != 0xFFu) & (e3 == key));
    if (hit) return 0;
    if (i3 == 0xFFu) {
      if (*count >= cap) return -1;
      const int b = i0 == 0xFFu ? 0 : (i1 == 0xFFu ? 1 : (i2 == 0xFFu ? 2 : 3));
      ws.vis_word(h) = (w & ~(0xFFu << (8 * b))) | ((uint32_t)*count << (8 * b));
      ws.star(*count) = key;
      (*count)++;
      return 1;
    }
    h = (h + 1u) & (TB_VIS_BUCKETS - 1u);
  }
  return -1;
}

template <class WS, class Sink>
TB_HD int star_bfs_cands(int site, int t0, const int4 *tets, const float4 *cc, WS &ws, int star_cap, int *n_star, float *cmin,
                         float *cmax, Sink &sink, double *vsum = nullptr)
{
  ws.hash_clear_vis();
  int ns = 0, ncand = 0;
  vis_find_or_insert(ws, t0, &ns, star_cap);
  {
    int4 v = tets[2 * (size_t)t0];
    int4 nb = tets[2 * (size_t)t0 + 1];
    float4 c = cc[t0];
    cmin[0] = fminf(cmin[0], c.x); cmin[1] = fminf(cmin[1], c.y); cmin[2] = fminf(cmin[2], c.z);
    cmax[0] = fmaxf(cmax[0], c.x); cmax[1] = fmaxf(cmax[1], c.y); cmax[2] = fmaxf(cmax[2], c.z);
    if (vsum) *vsum += (double)c.w;   // tet volume rides in the circumcenter record's 4th component
    int is = v.x == site ? 0 : (v.y == site ? 1 : (v.z == site ? 2 : (v.w == site ? 3 : -1)));
    if (is < 0) return CELL_OVERFLOW;
    for (int q = 0; q < 3; q++) {
      int s = q + (q >= is ? 1 : 0);
      sink(ncand++, tb_sel4(v.x, v.y, v.z, v.w, s), t0);
      int next = tb_sel4(nb.x, nb.y, nb.z, nb.w, s);
      if (next < 0) return CELL_INCOMPLETE;
      int before_s = ns;
      int r2 = vis_find_or_insert(ws, next, &ns, star_cap);
      if (r2 < 0) return CELL_OVERFLOW;
      if (r2 > 0) ws.parent_idx(before_s) = 0;
    }
  }
  int4 v_n = {0, 0, 0, 0}, nb_n = {0, 0, 0, 0};
  float4 c_n = {0, 0, 0, 0};
  bool have_n = false;
  if (ns > 1) {
    int t1 = ws.star(1);
    v_n = tets[2 * (size_t)t1]; nb_n = tets[2 * (size_t)t1 + 1]; c_n = cc[t1];
    have_n = true;
  }
  for (int head = 1; head < ns; head++) {
    const int t = ws.star(head);
    int4 v, nb;
    float4 c;
    if (have_n) { v = v_n; nb = nb_n; c = c_n; }
    else { v = tets[2 * (size_t)t]; nb = tets[2 * (size_t)t + 1]; c = cc[t]; }
    have_n = head + 1 < ns;
    if (have_n) {
      int t1 = ws.star(head + 1);
      v_n = tets[2 * (size_t)t1]; nb_n = tets[2 * (size_t)t1 + 1]; c_n = cc[t1];
    }
    cmin[0] = fminf(cmin[0], c.x); cmin[1] = fminf(cmin[1], c.y); cmin[2] = fminf(cmin[2], c.z);
    cmax[0] = fmaxf(cmax[0], c.x); cmax[1] = fmaxf(cmax[1], c.y); cmax[2] = fmaxf(cmax[2], c.z);
    if (vsum) *vsum += (double)c.w;   // tet volume rides in the circumcenter record's 4th component
    const int par = ws.star((int)ws.parent_idx(head));
    const int is = v.x == site ? 0 : (v.y == site ? 1 : (v.z == site ? 2 : (v.w == site ? 3 : -1)));
    const int ip = nb.x == par ? 0 : (nb.y == par ? 1 : (nb.z == par ? 2 : (nb.w == par ? 3 : -1)));
    if (is < 0 || ip < 0 || is == ip) return CELL_OVERFLOW;
    sink(ncand++, tb_sel4(v.x, v.y, v.z, v.w, ip), t);
    unsigned m = 0xFu & ~(1u << is) & ~(1u << ip);
    const int s1 = tb_ffs(m) - 1;
    m &= m - 1u;
    const int s2 = tb_ffs(m) - 1;
    const int n1 = tb_sel4(nb.x, nb.y, nb.z, nb.w, s1), n2 = tb_sel4(nb.x, nb.y, nb.z, nb.w, s2);
    if (n1 < 0 || n2 < 0) return CELL_INCOMPLETE;
    int before_s = ns;
    int r = vis_find_or_insert(ws, n1, &ns, star_cap);
    if (r < 0) return CELL_OVERFLOW;
    if (r > 0) ws.parent_idx(before_s) = (unsigned char)head;
    before_s = ns;
    r = vis_find_or_insert(ws, n2, &ns, star_cap);
    if (r < 0) return CELL_OVERFLOW;
    if (r > 0) ws.parent_idx(before_s) = (unsigned char)head;
  }
  *n_star = ns;
  return CELL_OK;
}

// second pass: candidates (u, t) in BFS order -> distinct neighbours with their first tet.
// ws needs nu(), nt() (int&), nbr_word(), hash_clear_nbr().  Returns nn, or -1 on overflow.
// Candidates are fetched four at a time (independent loads in flight) before they are consumed.
template <class WS>
TB_HD int nbr_insert(WS &ws, int u, int t, int *nn, int nbr_cap)
{
  unsigned h = ((unsigned)u * 0x9E3779B1u) >> 28;        // 16 buckets
  for (int guard = 0; guard < TB_NBR_BUCKETS; guard++) {
    const uint32_t w = ws.nbr_word(h);
    const unsigned i0 = w & 0xFFu, i1 = (w >> 8) & 0xFFu, i2 = (w >> 16) & 0xFFu, i3 = w >> 24;
    const int e0 = ws.nu(i0 == 0xFFu ? 0 : (int)i0), e1 = ws.nu(i1 == 0xFFu ? 0 : (int)i1);
    const int e2 = ws.nu(i2 == 0xFFu ? 0 : (int)i2), e3 = ws.nu(i3 == 0xFFu ? 0 : (int)i3);
    const bool hit = ((i0 != 0xFFu) & (e0 == u)) | ((i1 != 0xFFu) & (e1 == u)) | ((i2 != 0xFFu) & (e2 == u)) | ((i3 != 0xFFu) & (e3 == u));
    if (hit) return 0;
    if (i3 == 0xFFu) {
      if (*nn >= nbr_cap) return -1;
      const int b = i0 == 0xFFu ? 0 : (i1 == 0xFFu ? 1 : (i2 == 0xFFu ? 2 : 3));
      ws.nbr_word(h) = (w & ~(0xFFu << (8 * b))) | ((uint32_t)*nn << (8 * b));
      ws.nu(*nn) = u;
      ws.nt(*nn) = t;
      (*nn)++;
      return 1;
    }
    h = (h + 1u) & (TB_NBR_BUCKETS - 1u);
  }
  return -1;
}

template <class WS, class Source>
TB_HD int nbrs_from_cands(WS &ws, int n_cand, int nbr_cap, Source &src)
{
  ws.hash_clear_nbr();
  int nn = 0;
  for (int k = 0; k < n_cand; k += 4) {
    int u[4], t[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      u[q] = -1; t[q] = 0;
      if (k + q < n_cand) src(k + q, u[q], t[q]);
    }
#pragma unroll
    for (int q = 0; q < 4; q++)
      if (k + q < n_cand && nbr_insert(ws, u[q], t[q], &nn, nbr_cap) < 0) return -1;
  }
  return nn;
}

#define TB_MAX_LINK 4096

// Walks the tets around Delaunay edge (site, u) starting at ut, in the reference's order
// (fill_edge_link src/tet.cpp:389-409, circulate_start :164-173, circulate_next :186-215) and
// hands every circumcenter to `visit(k, cc)`.  Returns the link length, or -1 if it did not close.
template <class Visit>
TB_HD int walk_edge_link(int site, int u, int ut, const int4 *tets, const float4 *cc, Visit &visit)
{
  int4 v = tets[2 * (size_t)ut];
  int4 nb = tets[2 * (size_t)ut + 1];
  int wi;
  if (v.x != site && v.x != u) wi = 0;
  else if (v.y != site && v.y != u) wi = 1;
  else if (v.z != site && v.z != u) wi = 2;
  else wi = 3;
  int t = ut;
  for (int k = 0; k < TB_MAX_LINK; k++) {
    float4 c = cc[t];
    visit(k, c);
    int vv[4] = {v.x, v.y, v.z, v.w};
    int nv = -1;
#pragma unroll
    for (int i = 3; i >= 0; i--)
      if (i != wi && vv[i] != site && vv[i] != u) nv = vv[i]; // lowest such slot wins, as the reference's break does
    int next_t = wi == 0 ? nb.x : wi == 1 ? nb.y : wi == 2 ? nb.z : nb.w;
    if (next_t == ut || next_t < 0) return k + 1;
    v = tets[2 * (size_t)next_t];
    nb = tets[2 * (size_t)next_t + 1];
    wi = (v.x == nv) ? 0 : (v.y == nv) ? 1 : (v.z == nv) ? 2 : 3;
    t = next_t;
  }
  return -1;
}

// ---- the same circulation on 32-byte walk records ----------------------------------------------------
// WalkRec packs what one circulation step needs into ONE 32-byte sector: the tet's four neighbours,
// its circumcenter and a permutation word: for face i and slot j != i, bits [(4i+j)*2, +2) hold the
// slot that vertex verts[j] occupies in the neighbour across face i.  The walk then tracks the slots
// of the site and of u instead of comparing vertex ids, and never reads a vertex list after the
// first tet.  Same tets in the same order as walk_edge_link (checked in tests/emul).
struct
#if defined(__CUDACC__)
    __align__(16)
#endif
    WalkRec
{
  int nb[4];
  float cx, cy, cz;
  uint32_t perm;
};

TB_HD uint32_t walk_perm(const int *verts, const int *nbrs, const int4 *tets)
{
  uint32_t p = 0;
  for (int i = 0; i < 4; i++) {
    if (nbrs[i] < 0) continue;
    const int4 nv = tets[2 * (size_t)nbrs[i]];
    for (int j = 0; j < 4; j++) {
      if (j == i) continue;
      const int v = verts[j];
      const uint32_t s = nv.x == v ? 0u : (nv.y == v ? 1u : (nv.z == v ? 2u : 3u));
      p |= s << ((4 * i + j) * 2);
    }
  }
  return p;
}

template <class Visit>
TB_HD int walk_edge_link_rec(int s_c, int s_u, int ut, const WalkRec *walk, Visit &visit)
{
  // circulate_start: the first slot holding neither the site nor u
  int wi = 0;
  while (wi == s_c || wi == s_u) wi++;
  int t = ut;
  for (int k = 0; k < TB_MAX_LINK; k++) {
    const WalkRec r = walk[t];
    float4 c;
    c.x = r.cx; c.y = r.cy; c.z = r.cz; c.w = 0.0f;
    visit(k, c);
    const int nvs = 6 - s_c - s_u - wi;                        // slot of the fourth vertex
    const int next_t = wi == 0 ? r.nb[0] : (wi == 1 ? r.nb[1] : (wi == 2 ? r.nb[2] : r.nb[3]));
    if (next_t == ut || next_t < 0) return k + 1;
    const uint32_t p = r.perm >> (8 * wi);                      // the 4 x 2 bits of face wi
    const int n_c = (int)((p >> (2 * s_c)) & 3u), n_u = (int)((p >> (2 * s_u)) & 3u), n_w = (int)((p >> (2 * nvs)) & 3u);
    s_c = n_c; s_u = n_u; wi = n_w;
    t = next_t;
  }
  return -1;
}

// ---- the same BFS on walk records ------------------------------------------------------------------
// k_cell_bfs' version: one 32-byte WalkRec per popped tet (neighbours, circumcenter, slot
// permutation) instead of the tet record and the circumcenter (48 bytes), and no vertex compares:
// a queue entry carries, in the bits above the tet number, the slot of the site in that tet and
// the slot facing the BFS parent, both derived from the popping tet's permutation word when the
// entry is pushed.  That also retires the parent-index bytes of the workspace (272 instead of
// 324 bytes per thread: one more CTA per SM).  Same tets in the same order, same candidates, same
// box as star_bfs_cands (checked cell by cell in tests/emul).  Tet numbers must fit TB_STAR_TET_BITS.
#define TB_STAR_TET_BITS 27
#define TB_STAR_TET_MASK ((1 << TB_STAR_TET_BITS) - 1)

template <class WS>
TB_HD int vis_find_or_insert_tag(WS &ws, int key, int tag, int *count, int cap)
{
  unsigned h = ((unsigned)key * 0x9E3779B1u) >> 28;      // 16 buckets
  for (int guard = 0; guard < TB_VIS_BUCKETS; guard++) {
    const uint32_t w = ws.vis_word(h);
    const unsigned i0 = w & 0xFFu, i1 = (w >> 8) & 0xFFu, i2 = (w >> 16) & 0xFFu, i3 = w >> 24;
    const int e0 = ws.star(i0 == 0xFFu ? 0 : (int)i0) & TB_STAR_TET_MASK, e1 = ws.star(i1 == 0xFFu ? 0 : (int)i1) & TB_STAR_TET_MASK;
    const int e2 = ws.star(i2 == 0xFFu ? 0 : (int)i2) & TB_STAR_TET_MASK, e3 = ws.star(i3 == 0xFFu ? 0 : (int)i3) & TB_STAR_TET_MASK;
    const bool hit = ((i0 != 0xFFu) & (e0 == key)) | ((i1 != 0xFFu) & (e1 == key)) | ((i2 != 0xFFu) & (e2 == key)) | ((i3 != 0xFFu) & (e3 == key));
    if (hit) return 0;
    if (i3 == 0xFFu) {
      if (*count >= cap) return -1;
      const int b = i0 == 0xFFu ? 0 : (i1 == 0xFFu ? 1 : (i2 == 0xFFu ? 2 : 3));
      ws.vis_word(h) = (w & ~(0xFFu << (8 * b))) | ((uint32_t)*count << (8 * b));
      ws.star(*count) = key | (tag << TB_STAR_TET_BITS);
      (*count)++;
      return 1;
    }
    h = (h + 1u) & (TB_VIS_BUCKETS - 1u);
  }
  return -1;
}

// slots, in the neighbour across face s, of the site (at slot `is` here) and of the vertex that is not shared
TB_HD int walk_child_tag(uint32_t perm, int s, int is)
{
  const uint32_t b = (perm >> (8 * s)) & 0xFFu;         // four 2-bit fields, the one for slot s itself is 0
  const int is_n = (int)((b >> (2 * is)) & 3u);
  const int back = 6 - (int)((b & 3u) + ((b >> 2) & 3u) + ((b >> 4) & 3u) + (b >> 6));
  return is_n | (back << 2);
}

template <class WS, class Sink>
TB_HD int star_bfs_rec(int site, int t0, const int4 *tets, const WalkRec *walk, WS &ws, int star_cap, int *n_star, float *cmin, float *cmax,
                       Sink &sink)
{
  ws.hash_clear_vis();
  int ns = 0, ncand = 0;
  if (t0 > TB_STAR_TET_MASK) return CELL_OVERFLOW;
  const int *verts_pf = reinterpret_cast<const int *>(tets);
  {
    const int4 v = tets[2 * (size_t)t0];
    const WalkRec r = walk[t0];
    cmin[0] = fminf(cmin[0], r.cx); cmin[1] = fminf(cmin[1], r.cy); cmin[2] = fminf(cmin[2], r.cz);
    cmax[0] = fmaxf(cmax[0], r.cx); cmax[1] = fmaxf(cmax[1], r.cy); cmax[2] = fmaxf(cmax[2], r.cz);
    const int is = v.x == site ? 0 : (v.y == site ? 1 : (v.z == site ? 2 : (v.w == site ? 3 : -1)));
    if (is < 0) return CELL_OVERFLOW;
    vis_find_or_insert_tag(ws, t0, is, &ns, star_cap);
    for (int q = 0; q < 3; q++) {
      const int s = q + (q >= is ? 1 : 0);
      sink(ncand++, tb_sel4(v.x, v.y, v.z, v.w, s), t0);
      const int next = tb_sel4(r.nb[0], r.nb[1], r.nb[2], r.nb[3], s);
      if (next < 0) return CELL_INCOMPLETE;
      if (next > TB_STAR_TET_MASK) return CELL_OVERFLOW;
      const int tag = walk_child_tag(r.perm, s, is);
      const int ins = vis_find_or_insert_tag(ws, next, tag, &ns, star_cap);
      if (ins < 0) return CELL_OVERFLOW;
      if (ins > 0) { tb_prefetch(&walk[next]); tb_prefetch(&verts_pf[8 * (size_t)next + ((tag >> 2) & 3)]); }
    }
  }
  // software pipeline: the record of the next tet in the queue and the candidate vertex it will hand
  // over (the vertex opposite its parent face) are loaded one visit ahead
  WalkRec r_n;
  r_n.nb[0] = r_n.nb[1] = r_n.nb[2] = r_n.nb[3] = 0; r_n.cx = r_n.cy = r_n.cz = 0.0f; r_n.perm = 0;
  int u_n = 0;
  bool have_n = false;
  const int *verts = reinterpret_cast<const int *>(tets);
  if (ns > 1) {
    const int e1 = ws.star(1);
    r_n = walk[e1 & TB_STAR_TET_MASK];
    u_n = verts[8 * (size_t)(e1 & TB_STAR_TET_MASK) + ((e1 >> (TB_STAR_TET_BITS + 2)) & 3)];
    have_n = true;
  }
  for (int head = 1; head < ns; head++) {
    const int e = ws.star(head);
    const int t = e & TB_STAR_TET_MASK, is = (e >> TB_STAR_TET_BITS) & 3, ip = (e >> (TB_STAR_TET_BITS + 2)) & 3;
    const WalkRec r = have_n ? r_n : walk[t];
    const int u = have_n ? u_n : verts[8 * (size_t)t + ip];
    have_n = head + 1 < ns;
    if (have_n) {
      const int e1 = ws.star(head + 1);
      r_n = walk[e1 & TB_STAR_TET_MASK];
      u_n = verts[8 * (size_t)(e1 & TB_STAR_TET_MASK) + ((e1 >> (TB_STAR_TET_BITS + 2)) & 3)];
    }
    cmin[0] = fminf(cmin[0], r.cx); cmin[1] = fminf(cmin[1], r.cy); cmin[2] = fminf(cmin[2], r.cz);
    cmax[0] = fmaxf(cmax[0], r.cx); cmax[1] = fmaxf(cmax[1], r.cy); cmax[2] = fmaxf(cmax[2], r.cz);
    if (is == ip) return CELL_OVERFLOW;        // inconsistent adjacency: the general walk decides
    // the vertex opposite the parent face is the only one that can be a new neighbour
    sink(ncand++, u, t);
    unsigned m = 0xFu & ~(1u << is) & ~(1u << ip);
    const int s1 = tb_ffs(m) - 1;
    m &= m - 1u;
    const int s2 = tb_ffs(m) - 1;
    const int n1 = tb_sel4(r.nb[0], r.nb[1], r.nb[2], r.nb[3], s1), n2 = tb_sel4(r.nb[0], r.nb[1], r.nb[2], r.nb[3], s2);
    if (n1 < 0 || n2 < 0) return CELL_INCOMPLETE;
    if ((n1 | n2) > TB_STAR_TET_MASK) return CELL_OVERFLOW;
    const int tag1 = walk_child_tag(r.perm, s1, is), tag2 = walk_child_tag(r.perm, s2, is);
    const int ins1 = vis_find_or_insert_tag(ws, n1, tag1, &ns, star_cap);
    if (ins1 < 0) return CELL_OVERFLOW;
    if (ins1 > 0) { tb_prefetch(&walk[n1]); tb_prefetch(&verts[8 * (size_t)n1 + ((tag1 >> 2) & 3)]); }
    const int ins2 = vis_find_or_insert_tag(ws, n2, tag2, &ns, star_cap);
    if (ins2 < 0) return CELL_OVERFLOW;
    if (ins2 > 0) { tb_prefetch(&walk[n2]); tb_prefetch(&verts[8 * (size_t)n2 + ((tag2 >> 2) & 3)]); }
  }
  *n_star = ns;
  return CELL_OK;
}

// One Voronoi face for the dense stage (src/dense.cpp:682-735): circumcenters around the edge ->
// Newell normal, first vertex, running bbox.
struct FaceAccum
{
  float nrm[3], v0[3], prev[3];
  float *cmin, *cmax;   // running bbox of the cell (+inf / -inf to start), or null when the caller has it already
  TB_HD void operator()(int k, const float4 &c)
  {
    float cur[3] = {c.x, c.y, c.z};
    // `if (vv < cell_min) cell_min = vv` (src/dense.cpp:701-714) == fminf for non-NaN vertices
    if (cmin) {
      for (int d = 0; d < 3; d++) {
        cmin[d] = fminf(cmin[d], cur[d]);
        cmax[d] = fmaxf(cmax[d], cur[d]);
      }
    }
    if (k == 0) {
      for (int d = 0; d < 3; d++) { v0[d] = cur[d]; nrm[d] = 0.0f; }
    } else {
      newell_term(nrm, prev, cur);
    }
    for (int d = 0; d < 3; d++) prev[d] = cur[d];
  }
};

// One Voronoi face for volume() (src/volume.cpp:31-47): fan triangulation from the first
// circumcenter; the area sum goes through double because `sqrt(norm(cp))/2` is the double sqrt
// there.
struct AreaAccum
{
  float a[3], b[3];
  float area;
  TB_HD void operator()(int k, const float4 &c)
  {
    float cur[3] = {c.x, c.y, c.z};
    if (k == 0) {
      for (int d = 0; d < 3; d++) a[d] = cur[d];
      area = 0.0f;
    } else if (k == 1) {
      for (int d = 0; d < 3; d++) b[d] = cur[d];
    } else {
      float ab[3], ac[3], cp[3];
      for (int d = 0; d < 3; d++) { ab[d] = fsub(b[d], a[d]); ac[d] = fsub(cur[d], a[d]); }
      cp[0] = fsub(fmul(ab[1], ac[2]), fmul(ab[2], ac[1]));
      cp[1] = fsub(fmul(ab[2], ac[0]), fmul(ab[0], ac[2]));
      cp[2] = fsub(fmul(ab[0], ac[1]), fmul(ab[1], ac[0]));
      float n2 = fadd(fadd(fadd(0.0f, fmul(cp[0], cp[0])), fmul(cp[1], cp[1])), fmul(cp[2], cp[2]));
      area = (float)((double)area + sqrt((double)n2) / 2.0);
      for (int d = 0; d < 3; d++) b[d] = cur[d];
    }
  }
};

// ---- the scan-line state machine, src/dense.cpp:1475-1676 ---------------------------------
// `inside(i,j,k)` answers PtInCell for local bbox index (i,j,k); `line(j,k,min_xi,max_xi)` is
// called for every non-empty scan line in (z,y) order.  Returns the number of interior points.
template <class Inside, class Line>
TB_HD int scan_cell(int nx, int ny, int nz, Inside &inside, Line &line)
{
  int tot = 0;
  int x_left = nx / 2, x_right = nx / 2, y_start = 0, first_x = 0;
  for (int zi = 0; zi < nz; zi++) {
    bool border_found = false, z_step_done = false;
    for (int yi = y_start; yi < ny; yi++) {
      // the reference tests x_left three times per line (init, left walk, right walk when
      // x_right == x_left); PtInCell is pure, so the first answer is reused
      bool in0 = inside(x_left, yi, zi);
      bool x_in = in0;
      int min_xi = nx - 1, max_xi = 0;
      int xl0 = x_left;
      int xi;
      for (xi = x_left; xi >= 0 && xi < nx;) {
        bool in = (xi == xl0) ? in0 : inside(xi, yi, zi);
        if (in) {
          if (xi < min_xi) min_xi = xi;
          if (xi > max_xi) max_xi = xi;
          if (x_in) xi--;
          else { x_left = xi; break; }
        } else {
          if (!x_in) xi++;
          else { x_left = xi; break; }
        }
      }
      for (xi = x_right; xi >= 0 && xi < nx;) {
        bool in = (xi == xl0) ? in0 : inside(xi, yi, zi);
        if (in) {
          if (xi < min_xi) min_xi = xi;
          if (xi > max_xi) max_xi = xi;
          if (x_in) xi++;
          else { x_right = xi; break; }
        } else {
          if (!x_in) xi--;
          else { x_right = xi; break; }
        }
      }
      bool found = min_xi <= max_xi;
      if (found) {
        tot += max_xi - min_xi + 1;
        if (yi == y_start) first_x = (min_xi + max_xi) / 2;
        line(yi, zi, min_xi, max_xi);
      }
      int first_y = y_start;
      if (found && !border_found) { first_y = yi; border_found = true; }
      if (!found && border_found) z_step_done = true;
      if ((yi == ny - 1 || z_step_done) && zi + 1 < nz) {
        int yj;
        for (yj = first_y; yj > 0; yj--)
          if (!inside(first_x, yj, zi + 1)) break;
        y_start = yj;
      }
      if (z_step_done) break;
    }
  }
  return tot;
}

// The same state machine for index boxes at most 32 points wide, on whole scan lines:
// `row(j,k)` returns the inside-bits of line (j,k) (bit i = PtInCell of point (i,j,k)); the two
// x-walks of the reference become find-first-set / count-leading-zeros on masked copies of the row.
template <class Row, class Line>
TB_HD int scan_cell_bits(int nx, int ny, int nz, Row &row, Line &line)
{
  int tot = 0;
  int x_left = nx / 2, x_right = nx / 2, y_start = 0, first_x = 0;
  const uint32_t full = nx >= 32 ? 0xffffffffu : ((1u << nx) - 1u);
  for (int zi = 0; zi < nz; zi++) {
    bool border_found = false, z_step_done = false;
    for (int yi = y_start; yi < ny; yi++) {
      const uint32_t b = row(yi, zi) & full;
      const bool x_in = (b >> x_left) & 1u;          // the reference's single test at x_left (src/dense.cpp:1530-1542)
      int min_xi = nx - 1, max_xi = 0;
      const uint32_t below_left = (1u << x_left) - 1u;        // bits < x_left
      const uint32_t upto_right = (2u << x_right) - 1u;       // bits <= x_right
      if (x_in) {
        // left walk: inside from x_left downwards until the first outside point (src/dense.cpp:1547-1582)
        const uint32_t zb = ~b & below_left;
        int lo_rec = 0;
        const int old_left = x_left;
        if (zb) { int stop = 31 - tb_clz(zb); lo_rec = stop + 1; x_left = stop; }
        if (lo_rec < min_xi) min_xi = lo_rec;
        if (old_left > max_xi) max_xi = old_left;
        // right walk: inside from x_right upwards until the first outside point (src/dense.cpp:1585-1620)
        const uint32_t zr = ~b & full & ~(upto_right >> 1);   // outside points at or above x_right
        int hi_rec = nx - 1;
        const int old_right = x_right;
        if (zr) { int stop = tb_ffs(zr) - 1; hi_rec = stop - 1; x_right = stop; }
        if (hi_rec >= old_right) {
          if (old_right < min_xi) min_xi = old_right;
          if (hi_rec > max_xi) max_xi = hi_rec;
        }
      } else {
        // left walk: outside at x_left, step up to the first inside point
        const uint32_t ob = b & ~below_left;
        if (ob) {
          int xi = tb_ffs(ob) - 1;
          if (xi < min_xi) min_xi = xi;
          if (xi > max_xi) max_xi = xi;
          x_left = xi;
        }
        // right walk: step down from x_right to the first inside point
        const uint32_t orr = b & upto_right;
        if (orr) {
          int xi = 31 - tb_clz(orr);
          if (xi < min_xi) min_xi = xi;
          if (xi > max_xi) max_xi = xi;
          x_right = xi;
        }
      }
      bool found = min_xi <= max_xi;
      if (found) {
        tot += max_xi - min_xi + 1;
        if (yi == y_start) first_x = (min_xi + max_xi) / 2;
        line(yi, zi, min_xi, max_xi);
      }
      int first_y = y_start;
      if (found && !border_found) { first_y = yi; border_found = true; }
      if (!found && border_found) z_step_done = true;
      if ((yi == ny - 1 || z_step_done) && zi + 1 < nz) {
        int yj;
        for (yj = first_y; yj > 0; yj--)
          if (!((row(yj, zi + 1) >> first_x) & 1u)) break;
        y_start = yj;
      }
      if (z_step_done) break;
    }
  }
  return tot;
}

// ---- cloud-in-cell weights, src/dense.cpp:1787-1872 --------------------------------------
// idx0 = truncated index of the point; vals[8] in z-outer, x-inner order = weight * scalar
TB_HD void cic_weights(const float *pt, float scalar, const GridGeom &g, int *idx0, float *vals)
{
  for (int d = 0; d < 3; d++) idx0[d] = phys2idx1(pt[d], g.step[d], g.gmin[d]);
  float w[8];
  float tot = 0.0f, v0 = 0.0f;
  int n = 0;
  float two_eps = fmul(2.0f, g.eps);
  for (int dz = 0; dz < 2; dz++)
    for (int dy = 0; dy < 2; dy++)
      for (int dx = 0; dx < 2; dx++) {
        int ijk[3] = {idx0[0] + dx, idx0[1] + dy, idx0[2] + dz};
        float gp[3], p[3];
        for (int d = 0; d < 3; d++) {
          gp[d] = idx2phys1(ijk[d], g.step[d], g.gmin[d]);
          p[d] = pt[d];
          if (fabsf(fsub(p[d], gp[d])) < g.eps) p[d] = fadd(p[d], two_eps);
        }
        float vol = fabsf(fmul(fmul(fsub(gp[0], p[0]), fsub(gp[1], p[1])), fsub(gp[2], p[2])));
        if (v0 == 0.0f) v0 = vol;
        float v = fdiv(v0, vol);
        w[n++] = v;
        tot = fadd(tot, v);
      }
  for (int i = 0; i < 8; i++) vals[i] = fmul(fdiv(w[i], tot), scalar);
}

// ---- k_cic_gather's merge: the (at most) eight base cells around a grid point, each a list of particle ids in
// ascending order, added in particle order (IterateCellsCic, src/dense.cpp:523-541) ----------------------------
// The lists of one grid point: the heads live in registers (every step compares all eight); position and end of a list
// are looked up by the index of the list that moves on -- plain arrays here (tests/emul), shared memory in k_cic_gather
// (a register array indexed by a run-time value would turn into sixteen selects a step).
struct CicLists
{
  unsigned int pos_[8], end_[8];
  uint32_t head[8];                                  // the next particle id of every list, 0xffffffff at its end
  TB_HD unsigned int pos(int n) const { return pos_[n]; }
  TB_HD unsigned int end(int n) const { return end_[n]; }
  TB_HD void set(int n, unsigned int p, unsigned int e) { pos_[n] = p; end_[n] = e; }
  TB_HD void set_pos(int n, unsigned int p) { pos_[n] = p; }
};

TB_HD uint32_t tb_minu(uint32_t a, uint32_t b) { return a < b ? a : b; }

// one step: the weight of the next particle (corner n of the particle whose id heads list n); past the end -0.0f, which
// `x + (-0.0f)` leaves as it is for every x.  `vals` is in SORTED order (eight weights per position of sorted_ids): the lists
// of a cell are read front to back, and the eight grid points around a cell read the same 32-byte sectors.
template <class Lists>
TB_HD float cic_merge_step(Lists &l, const uint32_t *sorted_ids, const float *vals)
{
  // a particle has one base cell, so the ids of the eight heads are distinct: the smallest is the next particle
  const uint32_t b01 = tb_minu(l.head[0], l.head[1]), b23 = tb_minu(l.head[2], l.head[3]), b45 = tb_minu(l.head[4], l.head[5]),
                 b67 = tb_minu(l.head[6], l.head[7]);
  const uint32_t best = tb_minu(tb_minu(b01, b23), tb_minu(b45, b67));
  if (best == 0xffffffffu) return -0.0f;
  bool e[8];
#pragma unroll
  for (int n = 0; n < 8; n++) e[n] = l.head[n] == best;
  const int bn = (int)(e[1] | e[3] | e[5] | e[7]) | ((int)(e[2] | e[3] | e[6] | e[7]) << 1) | ((int)(e[4] | e[5] | e[6] | e[7]) << 2);
  // exactly one list moves on
  unsigned int ap = l.pos(bn);
  const unsigned int ae = l.end(bn);
  const float m = vals[8 * (size_t)ap + bn];
  ap++;
  l.set_pos(bn, ap);
  const uint32_t nh = ap < ae ? sorted_ids[ap] : 0xffffffffu;
#pragma unroll
  for (int n = 0; n < 8; n++) l.head[n] = e[n] ? nh : l.head[n];
  return m;
}

// the float adds in particle order (src/dense.cpp:539).  The add of a weight is issued CIC_ADD_DELAY steps after its load, so
// that the chain of adds does not wait for the memory while the merge could go on (the pending weights start as -0.0f, the
// neutral element).
constexpr int CIC_ADD_DELAY = 3;
template <class Lists>
TB_HD float cic_merge_sum(Lists &l, unsigned int total, const uint32_t *sorted_ids, const float *vals)
{
  float cur = 0.0f, pend[CIC_ADD_DELAY];
#pragma unroll
  for (int i = 0; i < CIC_ADD_DELAY; i++) pend[i] = -0.0f;
  for (unsigned int done = 0; done < total; done++) {
    const float m = cic_merge_step(l, sorted_ids, vals);
    cur = fadd(cur, pend[0]);
#pragma unroll
    for (int i = 0; i + 1 < CIC_ADD_DELAY; i++) pend[i] = pend[i + 1];
    pend[CIC_ADD_DELAY - 1] = m;
  }
#pragma unroll
  for (int i = 0; i < CIC_ADD_DELAY; i++) cur = fadd(cur, pend[i]);
  return cur;
}

// ---- DTFE, first order (alg 2; not in the reference) ------------------------------
// determinant in tet.cpp:139-143's term order
TB_HD float det3(const float *t, const float *u, const float *v)
{
  float r = fmul(fmul(t[0], u[1]), v[2]);
  r = fadd(r, fmul(fmul(u[0], v[1]), t[2]));
  r = fadd(r, fmul(fmul(v[0], t[1]), u[2]));
  r = fsub(r, fmul(fmul(v[0], u[1]), t[2]));
  r = fsub(r, fmul(fmul(u[0], t[1]), v[2]));
  r = fsub(r, fmul(fmul(t[0], v[1]), u[2]));
  return r;
}
// One tet prepared for rasterisation.  Face i is opposite vertex i; its three vertices are taken in
// ascending index order so that the two tets sharing a face evaluate bit-identical determinants.
struct DtfeTet
{
  float a[4][3], t[4][3], u[4][3];   // per face: first vertex, (second - first), (third - first)
  float sd[4];                       // determinant of the face with the opposite vertex
  float rho[4];
  TB_HD float face_det(int f, const float *x) const
  {
    float v[3] = {fsub(x[0], a[f][0]), fsub(x[1], a[f][1]), fsub(x[2], a[f][2])};
    return det3(t[f], u[f], v);
  }
  // false when the tet is degenerate
  TB_HD bool setup(const int *tv, const float *p0, const float *p1, const float *p2, const float *p3, const float *r)
  {
    const float *pos[4] = {p0, p1, p2, p3};
    bool ok = true;
    for (int i = 0; i < 4; i++) {
      int id[3], k = 0;
      for (int j = 0; j < 4; j++) if (j != i) id[k++] = j;
      // sort the three slots by vertex index
      if (tv[id[0]] > tv[id[1]]) { int s = id[0]; id[0] = id[1]; id[1] = s; }
      if (tv[id[1]] > tv[id[2]]) { int s = id[1]; id[1] = id[2]; id[2] = s; }
      if (tv[id[0]] > tv[id[1]]) { int s = id[0]; id[0] = id[1]; id[1] = s; }
      for (int d = 0; d < 3; d++) {
        a[i][d] = pos[id[0]][d];
        t[i][d] = fsub(pos[id[1]][d], pos[id[0]][d]);
        u[i][d] = fsub(pos[id[2]][d], pos[id[0]][d]);
      }
      sd[i] = face_det(i, pos[i]);
      if (!(sd[i] != 0.0f)) ok = false;
      rho[i] = r[i];
    }
    return ok;
  }
  // set up face i alone (the walk needs only the faces it actually tests)
  TB_HD void setup_face(int i, const int *tv, const float (*pos)[3])
  {
    int id[3], k = 0;
    for (int j = 0; j < 4; j++) if (j != i) id[k++] = j;
    if (tv[id[0]] > tv[id[1]]) { int s = id[0]; id[0] = id[1]; id[1] = s; }
    if (tv[id[1]] > tv[id[2]]) { int s = id[1]; id[1] = id[2]; id[2] = s; }
    if (tv[id[0]] > tv[id[1]]) { int s = id[0]; id[0] = id[1]; id[1] = s; }
    for (int d = 0; d < 3; d++) {
      a[i][d] = pos[id[0]][d];
      t[i][d] = fsub(pos[id[1]][d], pos[id[0]][d]);
      u[i][d] = fsub(pos[id[2]][d], pos[id[0]][d]);
    }
    sd[i] = face_det(i, pos[i]);
  }
  // walk step: faces are set up and tested one at a time; returns -1 if the tet owns x (then all four
  // faces are set up and sp[] holds the point's determinants), -2 if the tet is degenerate, else the
  // face to cross.  Same decisions as setup() + locate().
  TB_HD int locate_lazy(const int *tv, const float (*pos)[3], const float *x, float *sp)
  {
    for (int f = 0; f < 4; f++) {
      setup_face(f, tv, pos);
      if (!(sd[f] != 0.0f)) return -2;
      sp[f] = face_det(f, x);
      if (sp[f] == 0.0f) { if (!(sd[f] > 0.0f)) return f; }
      else if ((sp[f] > 0.0f) != (sd[f] > 0.0f)) return f;
    }
    return -1;
  }
  // eval()'s sum from determinants already computed by locate_lazy
  TB_HD float value_from(const float *sp) const
  {
    float acc = 0.0f;
    for (int f = 0; f < 4; f++) acc = fadd(acc, fmul(fdiv(sp[f], sd[f]), rho[f]));
    return acc;
  }
  // -1 if this tet owns x, else a face through which x lies on the far side (the walk crosses it)
  TB_HD int locate(const float *x) const
  {
    for (int f = 0; f < 4; f++) {
      float sp = face_det(f, x);
      if (sp == 0.0f) { if (!(sd[f] > 0.0f)) return f; }
      else if ((sp > 0.0f) != (sd[f] > 0.0f)) return f;
    }
    return -1;
  }
  // value at x if this tet owns x: for every face the point is on the opposite vertex's side, a point
  // exactly on a face belonging to the tet on the face's positive side
  TB_HD bool eval(const float *x, float *val) const
  {
    float acc = 0.0f;
    for (int f = 0; f < 4; f++) {
      float sp = face_det(f, x);
      if (sp == 0.0f) { if (!(sd[f] > 0.0f)) return false; }
      else if ((sp > 0.0f) != (sd[f] > 0.0f)) return false;
      acc = fadd(acc, fmul(fdiv(sp, sd[f]), rho[f]));
    }
    *val = acc;
    return true;
  }
};

// ---- span records ----------------------------------------------------------------------------
// One x-run of deposits in one row of one block's density array.
//   key  = (((row << 1 | remote) << cell_bits | cell) << z_bits) | zslot      (sorted ascending)
//   data = x0 | len << 16 | float_path << 31 | (uint64)float_bits(value) << 32
//   zslot = z - z_lo, projections only: z index of the deposit (orders one cell's deposits onto the same
//   (x, y) as the reference's z-outer scan does).  z_lo = lowest z index any block's closed bounds hold;
//   it is negative when a given z range is narrower than the data (such points still deposit: the
//   projected index drops z, src/dense.cpp:1047-1090)
struct KeyLayout
{
  int cell_bits, z_bits, z_lo;
};
TB_HD uint32_t z_slot(const KeyLayout &kl, int project, int z) { return project ? (uint32_t)(z - kl.z_lo) : 0u; }
TB_HD uint64_t make_key(const KeyLayout &kl, uint64_t row, int remote, uint32_t cell, uint32_t zslot)
{
  return ((((row << 1) | (uint64_t)remote) << kl.cell_bits | cell) << kl.z_bits) | zslot;
}
TB_HD uint64_t key_row(const KeyLayout &kl, uint64_t key) { return key >> (kl.z_bits + kl.cell_bits + 1); }

TB_HD uint32_t f2u(float f)
{
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  union { float f; uint32_t u; } c;
  c.f = f;
  return c.u;
#endif
}
TB_HD float u2f(uint32_t u)
{
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c;
  c.u = u;
  return c.f;
#endif
}
TB_HD uint64_t make_data(int x0, int len, int float_path, float value)
{
  return (uint64_t)(uint32_t)x0 | ((uint64_t)(uint32_t)len << 16) | ((uint64_t)(float_path & 1) << 31) | ((uint64_t)f2u(value) << 32);
}

// The reference's accumulate step (src/dense.cpp:290,193 double path; :539 float path)
TB_HD float accumulate(float cur, float m, float div, int float_path)
{
  if (float_path) return fadd(cur, fdiv(m, div));
  return (float)((double)cur + (double)m / (double)div);
}

// Integer description of one block for the deposit logic (all boxes inclusive lo, inclusive hi):
//   P = { idx : idx2phys(idx) inside the block's closed bounds }  (src/dense.cpp:279-284)
//   B = the block's sub-grid [block_min_idx, block_min_idx + block_num_idx)   (BlockGridParams)
struct BlockBox
{
  int p_lo[3], p_hi[3];
  int b_lo[3], b_num[3];
  long long row_base;    // first row id of this block
};

// Emits the records of one scan line [xa, xb] x {y} x {z} (global indices) of a cell owned by
// block `e`.  `emit(key, data)` receives every record.  Mirrors src/dense.cpp:276-305: a point
// whose position lies in the emitting block's closed bounds is local, otherwise it goes to every
// other block whose closed bounds contain it; in both cases only indices inside the target's
// sub-grid are kept (the reference writes out of bounds there).
// The blocks other than e that can receive points of the index box [lo, lo + n3): at most `cap`
// indices into `boxes` are written to `cand`; returns their number, or -1 if there are more than cap
// (the caller then lets emit_line scan all blocks).  With many blocks this keeps the per-line loop of
// boundary cells short (64 blocks: 8 GPUs x 8).
TB_HD int candidate_blocks(const BlockBox *boxes, int nblocks, int e, const int *lo, const int *n3, int project, int *cand, int cap)
{
  int n = 0;
  for (int j = 0; j < nblocks; j++) {
    if (j == e) continue;
    const BlockBox &b = boxes[j];
    bool hit = true;
    for (int d = 0; d < 3; d++) {
      const int hi = lo[d] + n3[d] - 1;
      if (hi < b.p_lo[d] || lo[d] > b.p_hi[d]) hit = false;
      if (!(project && d == 2) && (hi < b.b_lo[d] || lo[d] > b.b_lo[d] + b.b_num[d] - 1)) hit = false;
    }
    if (!hit) continue;
    if (n >= cap) return -1;
    cand[n++] = j;
  }
  return n;
}

template <class Emit>
TB_HD void emit_line(const BlockBox *boxes, int nblocks, int e, const KeyLayout &kl, int project, uint32_t cell,
                     int xa, int xb, int y, int z, int float_path_local, float value, Emit &emit,
                     const int *cand = nullptr, int ncand = -1)
{
  const BlockBox &be = boxes[e];
  bool yz_local = y >= be.p_lo[1] && y <= be.p_hi[1] && z >= be.p_lo[2] && z <= be.p_hi[2];
  // local part: x in P_e (and inside B_e)
  int la = xa, lb = xa - 1;
  if (yz_local) {
    la = xa > be.p_lo[0] ? xa : be.p_lo[0];
    lb = xb < be.p_hi[0] ? xb : be.p_hi[0];
  }
  if (la <= lb) {
    int ly = y - be.b_lo[1], lz = z - be.b_lo[2];
    if (ly >= 0 && ly < be.b_num[1] && (project || (lz >= 0 && lz < be.b_num[2]))) {
      int a = la > be.b_lo[0] ? la : be.b_lo[0];
      int b = lb < be.b_lo[0] + be.b_num[0] - 1 ? lb : be.b_lo[0] + be.b_num[0] - 1;
      if (a <= b) {
        uint64_t row = (uint64_t)(be.row_base + (project ? (long long)ly : (long long)lz * be.b_num[1] + ly));
        emit(make_key(kl, row, 0, cell, z_slot(kl, project, z)), make_data(a - be.b_lo[0], b - a + 1, float_path_local, value));
      }
    }
  }
  if (la <= lb && la == xa && lb == xb) return; // whole line was local
  // remote parts: the pieces of [xa, xb] outside [la, lb] (all of it when the line is not local)
  const int nloop = ncand >= 0 ? ncand : nblocks;
  for (int q = 0; q < nloop; q++) {
    const int j = ncand >= 0 ? cand[q] : q;
    if (j == e) continue;
    const BlockBox &bj = boxes[j];
    if (y < bj.p_lo[1] || y > bj.p_hi[1] || z < bj.p_lo[2] || z > bj.p_hi[2]) continue;
    int ly = y - bj.b_lo[1], lz = z - bj.b_lo[2];
    if (ly < 0 || ly >= bj.b_num[1] || (!project && (lz < 0 || lz >= bj.b_num[2]))) continue;
    int a0 = xa > bj.p_lo[0] ? xa : bj.p_lo[0];
    int b0 = xb < bj.p_hi[0] ? xb : bj.p_hi[0];
    a0 = a0 > bj.b_lo[0] ? a0 : bj.b_lo[0];
    b0 = b0 < bj.b_lo[0] + bj.b_num[0] - 1 ? b0 : bj.b_lo[0] + bj.b_num[0] - 1;
    uint64_t row = (uint64_t)(bj.row_base + (project ? (long long)ly : (long long)lz * bj.b_num[1] + ly));
    // pieces of [a0, b0] not in the local range [la, lb]
    int pa[2] = {a0, a0}, pb[2] = {b0, a0 - 1};
    if (la <= lb) {
      pa[0] = a0; pb[0] = b0 < la - 1 ? b0 : la - 1;
      pa[1] = a0 > lb + 1 ? a0 : lb + 1; pb[1] = b0;
    }
    for (int s = 0; s < 2; s++)
      if (pa[s] <= pb[s])
        emit(make_key(kl, row, 1, cell, z_slot(kl, project, z)), make_data(pa[s] - bj.b_lo[0], pb[s] - pa[s] + 1, 0, value));
  }
}

} // namespace tb
