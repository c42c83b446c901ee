// fused.cuh -- K3a as ONE kernel per cell: star walk, Voronoi faces, planes and the inside
// test of the cell's index box never leave the SM.
//
//   k_cell_fused   half a warp (16 lanes) per Voronoi cell, two cells per warp, everything in shared memory:
//                    1. star BFS in the reference's FIFO order (neighbor_edges / complete, src/tet.cpp:228-270,337-378):
//                       the queue is consumed in chunks of 16 entries, one lane per popped tet -- the lanes load the tet
//                       records (32 B) and circumcenters (16 B) of a whole chunk at once, so a cell pays one global
//                       latency per BFS level instead of one per tet.  Children are proposed into a shared-memory hash
//                       table; a child proposed twice inside a chunk goes to the proposal with the lowest (parent
//                       position, slot) -- atomicMin on an order key -- which is exactly the push that a sequential FIFO
//                       would have kept.  Queue positions follow from ballots.
//                    2. Delaunay neighbours: in a manifold star the first popped tet that holds u offers it as "the vertex
//                       opposite my parent face" (cell_core.cuh, star_bfs_cands), so the first tet of neighbor_edges' list
//                       is the minimum queue position over the offers: one more table, one atomicMin per popped tet.
//                    3. faces: one lane per Delaunay neighbour circulates the tets around the edge (fill_edge_link,
//                       src/tet.cpp:389-409) through LOCAL adjacency bytes, Newell normal, orientation (src/dense.cpp:682-735).
//                    4. cell box, data-bounds filter, index box (src/dense.cpp:700-714,1381-1410).
//                    5. PtInCell (src/dense.cpp:1172-1203) over the index box in the axis-separable form: the reference's
//                       dist = (n.x*(x-f.x) + n.y*(y-f.y)) + n.z*(z-f.z) is fadd(fadd(A[i], B[j]), C[k]) with per-axis
//                       products that depend on one index each (same roundings), so a face costs nx+ny+nz products per
//                       cell and one add + min/max per point.
//                  Output per cell: a 32-byte header and the inside bits (32 bytes in line, larger boxes in a pool).
//   k_cell_emit    one LANE per cell: the reference's scan-line state machine (src/dense.cpp:1475-1676) on the bits,
//                  span records out (the phase-3 code of k_cell_scan).
//
// Cells the fast path cannot hold (star > 64 tets, > 32 faces, non-manifold star, index box > 2048 points or > 16
// points along an axis) go to the overflow list and through the general kernels of kernels.cuh.
#pragma once
#include "kernels.cuh"

namespace tb
{

constexpr int FZ_GL = 16;                 // lanes per cell
constexpr int FZ_CPW = 32 / FZ_GL;        // cells per warp
constexpr int FZ_CAP = 64;                // star tets
constexpr int FZ_NBR = 32;                // faces
constexpr int FZ_TH = 128;                // tet table slots
constexpr int FZ_NH = 64;                 // neighbour table slots
constexpr int FZ_AX = 16;                 // index box extent per axis
constexpr int FZ_PTS = 2048;              // index box points
constexpr int FZ_TAB = 512;               // floats of the per-axis product tables (overlay of v + cc)
constexpr int FZ_WARPS = 4;
constexpr int FZ_THREADS = FZ_WARPS * 32;
constexpr int FZ_INLINE_WORDS = 8;        // inside bits kept next to the header (256 points)
constexpr int FZ_PROP = 0x100;            // order keys of proposals sort behind every final queue position

struct __align__(16) FzCell
{
  int4 v[FZ_CAP];                         // verts of the star tets in BFS order            | phase 5: product tables (with cc)
  float4 cc[FZ_CAP];                      // circumcenter; w = 4 bytes: table slots, then local positions, of the 4 neighbours
  float planes[FZ_NBR * 6];               // unit normal + first vertex of every Voronoi face
  int q_id[FZ_CAP];                       // the FIFO queue == the star in BFS order
  int h_id[FZ_TH];                        // tet table: id -> h_val = final queue position, or the lowest proposal key
  int h_val[FZ_TH];
  int n_id[FZ_NH];                        // neighbour table: vertex id -> n_first = lowest queue position that offered it
  int n_first[FZ_NH];
  int f_u[FZ_NBR];                        // faces: neighbour vertex, first tet (queue position)
  unsigned char f_t[FZ_NBR];
  unsigned char q_par[FZ_CAP];            // queue position of the tet that pushed this one
  unsigned char q_tag[FZ_CAP];            // slot of the site | slot facing the parent << 2 (4 = root)
  uint32_t bits[FZ_PTS / 32 + 4];
};
constexpr size_t FZ_SMEM = sizeof(FzCell) * FZ_CPW * FZ_WARPS;

__device__ __forceinline__ int fz_find4(const int4 &a, int key) { return a.x == key ? 0 : (a.y == key ? 1 : (a.z == key ? 2 : (a.w == key ? 3 : -1))); }
__device__ __forceinline__ int fz_get4(const int4 &a, int s) { return s == 0 ? a.x : (s == 1 ? a.y : (s == 2 ? a.z : a.w)); }

// find-or-insert in an open-addressing table of N slots (-1 = empty); returns the slot, -1 when the table is full
template <int N, int LOG2N>
__device__ __forceinline__ int fz_insert(int *ids, int key)
{
  unsigned h = ((unsigned)key * 0x9E3779B1u) >> (32 - LOG2N);
  for (int probe = 0; probe < N; probe++) {
    const int old = atomicCAS(&ids[h], -1, key);
    if (old == -1 || old == key) return (int)h;
    h = (h + 1u) & (unsigned)(N - 1);
  }
  return -1;
}

// PtInCell over the columns (i, j) of one cell's index box, k swept KC at a time (see the file header, step 5).
// T: per face `tot` = nx + ny + nz products: A[i] | B[j] | C[k].  Writes the inside bits, x fastest.
template <int KC>
__device__ __forceinline__ void fz_columns(uint32_t *bits, const float *T, bool tab, int nx, int ny, int nz, int nn, int cols_any, int nz_any,
                                           int nn_any, float eps, int gl, int grp)
{
  const unsigned FULL = 0xffffffffu;
  const unsigned GLMASK = (1u << FZ_GL) - 1u;
  const int nxy = nx * ny, tot = nx + ny + nz;
  const float neg_eps = -eps;
  for (int col0 = 0; col0 < cols_any; col0 += FZ_GL) {
    const int col = col0 + gl;
    const bool cact = tab && col < nxy;
    const int j = cact ? col / nx : 0, i = cact ? col - j * nx : 0;
    for (int k0 = 0; k0 < nz_any; k0 += KC) {
      float dmax[KC], dmin[KC];
#pragma unroll
      for (int kk = 0; kk < KC; kk++) { dmax[kk] = -INFINITY; dmin[kk] = INFINITY; }
      for (int f = 0; f < nn_any; f++) {
        if (cact && f < nn) {
          const float *tf = T + f * tot;
          const float ab = fadd(tf[i], tf[nx + j]);
          const float *tc = tf + nx + ny + k0;
#pragma unroll
          for (int kk = 0; kk < KC; kk++)
            if (k0 + kk < nz) {
              const float dist = fadd(ab, tc[kk]);
              dmax[kk] = fmaxf(dmax[kk], dist);
              dmin[kk] = fminf(dmin[kk], dist);
            }
        }
      }
#pragma unroll
      for (int kk = 0; kk < KC; kk++) {
        const int k = k0 + kk;
        const bool in = cact && k < nz && !(dmax[kk] > eps && dmin[kk] < neg_eps);
        const unsigned m = (__ballot_sync(FULL, in) >> (grp * FZ_GL)) & GLMASK;
        if (gl == 0 && tab && k < nz && m) {
          const int off = k * nxy + col0;
          bits[off >> 5] |= m << (off & 31);
          if ((off & 31) > 32 - FZ_GL) bits[(off >> 5) + 1] |= m >> (32 - (off & 31));
        }
      }
    }
  }
}

struct FusedOut
{
  CellHdr *hdr;                  // [n_slots]; pad == 0: depositing cell, plane_off = word offset in the pool or 0xFFFFFFFF (in line)
  uint32_t *bits_inline;         // [n_slots][FZ_INLINE_WORDS]
  uint32_t *bits_pool;
  unsigned long long pool_words;
  unsigned long long *pool_cursor;
  uint2 *overflow;               // (block index, block-local cell)
  uint32_t cap_overflow;
  Counters *cnt;
};

__global__ void __launch_bounds__(FZ_THREADS) k_cell_fused(const DevBlock *__restrict__ blocks, int blk_begin, int blk_end, uint32_t n_slots,
                                                           const __grid_constant__ GridGeom g, FusedOut out)
{
  extern __shared__ __align__(16) unsigned char fz_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane & (FZ_GL - 1), grp = lane / FZ_GL;
  const unsigned FULL = 0xffffffffu;
  const unsigned GLMASK = (FZ_GL == 32) ? 0xffffffffu : ((1u << FZ_GL) - 1u);
  const unsigned below = (1u << gl) - 1u;
  FzCell &C = reinterpret_cast<FzCell *>(fz_smem)[warp * FZ_CPW + grp];
  unsigned long long c_no_tet = 0, c_incomplete = 0, c_outside = 0, c_bad = 0, c_faces = 0;

  const uint32_t n_pairs = (n_slots + FZ_CPW - 1) / FZ_CPW;
  const uint32_t warps_total = gridDim.x * FZ_WARPS;
  for (uint32_t pair = blockIdx.x * FZ_WARPS + warp; pair < n_pairs; pair += warps_total) {
    const uint32_t slot = pair * FZ_CPW + grp;
    const bool have = slot < n_slots;
    int blk_id = blk_begin;
    for (int b = blk_begin + 1; b < blk_end; b++)
      if (blocks[b].num_orig > 0 && slot >= blocks[b].slot_start) blk_id = b;
    const DevBlock &blk = blocks[blk_id];
    const int4 *__restrict__ tets = blk.tets;
    const float4 *__restrict__ ccs = blk.cc;
    int status = -1, cell = 0, t0 = -1;
    float site[3] = {0.0f, 0.0f, 0.0f};
    if (have && blk.num_orig > 0 && slot - blk.slot_start < (uint32_t)blk.num_orig) {
      cell = (int)blk.order[slot - blk.slot_start];
      t0 = blk.v2t[cell];
      status = t0 < 0 ? CELL_NO_TET : (blk.hull[cell] ? CELL_INCOMPLETE : CELL_OK);
      site[0] = blk.particles[3 * (size_t)cell]; site[1] = blk.particles[3 * (size_t)cell + 1]; site[2] = blk.particles[3 * (size_t)cell + 2];
    }
    const bool real = status != -1;

    // ---- 1 + 2: star BFS, neighbour offers -------------------------------------------------------------
    for (int i = gl; i < FZ_TH; i += FZ_GL) { C.h_id[i] = -1; C.h_val[i] = 0x7fffffff; }
    for (int i = gl; i < FZ_NH; i += FZ_GL) { C.n_id[i] = -1; C.n_first[i] = 0x7fffffff; }
    __syncwarp();
    int head = 0, tail = 0;
    if (status == CELL_OK) {
      if (gl == 0) {
        C.q_id[0] = t0; C.q_par[0] = 0xFF;
        const int hs = fz_insert<FZ_TH, 7>(C.h_id, t0);
        C.h_val[hs] = 0;
      }
      tail = 1;
    }
    __syncwarp();
    float bmin[3] = {INFINITY, INFINITY, INFINITY}, bmax[3] = {-INFINITY, -INFINITY, -INFINITY};
    while (__any_sync(FULL, status == CELL_OK && head < tail)) {
      const bool run = status == CELL_OK && head < tail;
      const int qpos = head + gl;
      const bool act = run && qpos < tail;
      int4 nb = make_int4(0, 0, 0, 0);
      int hs[4] = {-1, -1, -1, -1};
      bool bad = false, inc = false;
      if (act) {
        const int t = C.q_id[qpos];
        const int4 v = __ldg(&tets[2 * (size_t)t]);
        nb = __ldg(&tets[2 * (size_t)t + 1]);
        float4 c = __ldg(&ccs[t]);
        C.v[qpos] = v;
        bmin[0] = fminf(bmin[0], c.x); bmin[1] = fminf(bmin[1], c.y); bmin[2] = fminf(bmin[2], c.z);
        bmax[0] = fmaxf(bmax[0], c.x); bmax[1] = fmaxf(bmax[1], c.y); bmax[2] = fmaxf(bmax[2], c.z);
        const int is = fz_find4(v, cell);
        int ip = 4, par = 0;
        if (qpos > 0) {
          par = (int)C.q_par[qpos];
          ip = fz_find4(nb, C.q_id[par]);
        }
        bad = is < 0 || ip < 0 || ip == is;
        uint32_t packed = 0;
        if (!bad) {
          // neighbour offers: the root hands over its three vertices, every other tet the vertex opposite its parent face
          if (qpos == 0) {
#pragma unroll
            for (int s = 0; s < 4; s++)
              if (s != is) {
                const int h = fz_insert<FZ_NH, 6>(C.n_id, fz_get4(v, s));
                if (h < 0) bad = true; else atomicMin(&C.n_first[h], 0);
              }
          } else {
            const int h = fz_insert<FZ_NH, 6>(C.n_id, fz_get4(v, ip));
            if (h < 0) bad = true; else atomicMin(&C.n_first[h], qpos);
          }
          // children, in slot order (src/tet.cpp:262-266 pushes tets[t].tets[i] for every slot whose vertex is not the site)
#pragma unroll
          for (int s = 0; s < 4; s++)
            if (s != is && s != ip) {
              const int next = fz_get4(nb, s);
              if (next < 0) inc = true;
              else {
                const int h = fz_insert<FZ_TH, 7>(C.h_id, next);
                if (h < 0) bad = true;
                else {
                  hs[s] = h;
                  atomicMin(&C.h_val[h], FZ_PROP | (gl << 2) | s);
                  packed |= (uint32_t)h << (8 * s);
                }
              }
            }
          if (ip < 4) packed |= (uint32_t)par << (8 * ip);       // the parent's queue position is final
        }
        c.w = __uint_as_float(packed);
        C.cc[qpos] = c;
        C.q_tag[qpos] = (unsigned char)((is & 3) | (ip << 2));
      }
      const unsigned b_bad = (__ballot_sync(FULL, act && bad) >> (grp * FZ_GL)) & GLMASK;
      const unsigned b_inc = (__ballot_sync(FULL, act && inc) >> (grp * FZ_GL)) & GLMASK;
      if (run) {
        if (b_inc) status = CELL_INCOMPLETE;        // complete() walks the whole star before anything else (src/dense.cpp:252)
        else if (b_bad) status = CELL_OVERFLOW;
      }
      __syncwarp();
      // which proposals were the first push of their tet
      bool won[4];
      unsigned W[4];
      int total = 0, base = tail;
#pragma unroll
      for (int s = 0; s < 4; s++) {
        won[s] = act && status == CELL_OK && hs[s] >= 0 && C.h_val[hs[s]] == (FZ_PROP | (gl << 2) | s);
        W[s] = (__ballot_sync(FULL, won[s]) >> (grp * FZ_GL)) & GLMASK;
        total += __popc(W[s]);
        base += __popc(W[s] & below);
      }
      if (run && status == CELL_OK && tail + total > FZ_CAP) status = CELL_OVERFLOW;
      __syncwarp();                                  // every lane has read h_val before the winners overwrite it
      if (status == CELL_OK) {
        int p = base;
#pragma unroll
        for (int s = 0; s < 4; s++)
          if (won[s]) {
            C.q_id[p] = fz_get4(nb, s);
            C.q_par[p] = (unsigned char)qpos;
            C.h_val[hs[s]] = p;
            p++;
          }
      }
      if (run) {
        head = head + FZ_GL < tail ? head + FZ_GL : tail;
        tail += total;
      }
      __syncwarp();
    }
    const int ns = tail;

    // ---- neighbour list, local adjacency bytes -----------------------------------------------------------
    int nn = 0;
    {
#pragma unroll
      for (int r = 0; r < FZ_NH / FZ_GL; r++) {
        const int i = r * FZ_GL + gl;
        const bool has = status == CELL_OK && C.n_id[i] != -1;
        const unsigned m = (__ballot_sync(FULL, has) >> (grp * FZ_GL)) & GLMASK;
        const int pos = nn + __popc(m & below);
        if (has && pos < FZ_NBR) { C.f_u[pos] = C.n_id[i]; C.f_t[pos] = (unsigned char)C.n_first[i]; }
        nn += __popc(m);
      }
      if (status == CELL_OK && nn > FZ_NBR) status = CELL_OVERFLOW;
      if (status == CELL_OK)
        for (int q = gl; q < ns; q += FZ_GL) {
          const int tag = C.q_tag[q], is = tag & 3, ip = tag >> 2;
          const uint32_t w = __float_as_uint(C.cc[q].w);
          uint32_t o = 0;
#pragma unroll
          for (int s = 0; s < 4; s++) {
            uint32_t b = (w >> (8 * s)) & 0xffu;
            if (s != is && s != ip) b = (uint32_t)C.h_val[b] & 0xffu;
            o |= b << (8 * s);
          }
          C.cc[q].w = __uint_as_float(o);
        }
    }
    __syncwarp();

    // ---- 3: faces -> planes --------------------------------------------------------------------------------------
    float *planes = C.planes;
    bool face_bad = false;
#pragma unroll 1
    for (int r = 0; r < FZ_NBR / FZ_GL; r++) {
      const int f = r * FZ_GL + gl;
      if (status == CELL_OK && f < nn) {
        const int u = C.f_u[f], ut = (int)C.f_t[f];
        int4 vv = C.v[ut];
        int wi = 0;
        while (wi < 3 && (fz_get4(vv, wi) == cell || fz_get4(vv, wi) == u)) wi++;     // circulate_start, src/tet.cpp:164-173
        FaceAccum fa;
        fa.cmin = nullptr; fa.cmax = nullptr;
        int t = ut, n = -1;
        for (int k = 0; k <= ns; k++) {
          const float4 c = C.cc[t];
          fa(k, c);
          int nv = -1;
#pragma unroll
          for (int i = 3; i >= 0; i--) {
            const int x = fz_get4(vv, i);
            if (i != wi && x != cell && x != u) nv = x;            // the lowest such slot wins, as the reference's break does
          }
          const int next = (int)((__float_as_uint(c.w) >> (8 * wi)) & 0xffu);
          if (next == ut) { n = k + 1; break; }
          vv = C.v[next];
          wi = fz_find4(vv, nv);
          if (wi < 0) break;
          t = next;
        }
        if (n < 0) {
          const float q = __int_as_float(0x7fc00000);               // a NaN plane is never significant in PtInCell
          fa.nrm[0] = fa.nrm[1] = fa.nrm[2] = q; fa.v0[0] = fa.v0[1] = fa.v0[2] = q;
          face_bad = true;
        } else {
          newell_term(fa.nrm, fa.prev, fa.v0);
          newell_finish(fa.nrm, fa.v0, site);
        }
        float2 *dst = reinterpret_cast<float2 *>(planes + 6 * f);
        dst[0] = make_float2(fa.nrm[0], fa.nrm[1]);
        dst[1] = make_float2(fa.nrm[2], fa.v0[0]);
        dst[2] = make_float2(fa.v0[1], fa.v0[2]);
      }
    }
    if (face_bad) c_bad++;

    // ---- 4: cell box, data-bounds filter, index box -----------------------------------------------------------------
#pragma unroll
    for (int d = 8; d >= 1; d >>= 1) {
#pragma unroll
      for (int a = 0; a < 3; a++) {
        bmin[a] = fminf(bmin[a], __shfl_xor_sync(FULL, bmin[a], d));
        bmax[a] = fmaxf(bmax[a], __shfl_xor_sync(FULL, bmax[a], d));
      }
    }
    int lo[3] = {0, 0, 0}, n3[3] = {1, 1, 1};
    int npts = 0;
    if (status == CELL_OK) {
      for (int d = 0; d < 3; d++)
        if (bmin[d] < fsub(g.dmin[d], g.dext_eps[d]) || bmax[d] > fadd(g.dmax[d], g.dext_eps[d])) status = CELL_OUTSIDE;   // src/dense.cpp:1385-1392
    }
    if (status == CELL_OK) {
      for (int d = 0; d < 3; d++) {
        lo[d] = phys2idx1(bmin[d], g.step[d], g.gmin[d]);
        const int hi = phys2idx1(bmax[d], g.step[d], g.gmin[d]);
        n3[d] = hi - lo[d] + 1;
      }
      if (n3[0] < 1 || n3[1] < 1 || n3[2] < 1 || n3[0] > 32767 || n3[1] > 32767 || n3[2] > 32767 || lo[0] < -(1 << 22) || lo[1] < -(1 << 22) ||
          lo[2] < -(1 << 22))
        status = CELL_BAD_MESH;
      else if (n3[0] > FZ_AX || n3[1] > FZ_AX || n3[2] > FZ_AX || n3[0] * n3[1] * n3[2] > FZ_PTS)
        status = CELL_OVERFLOW;                        // a large index box: the general scan kernels
      else
        npts = n3[0] * n3[1] * n3[2];
    }
    __syncwarp();                                      // planes written, v / cc no longer read: the tables may overwrite them

    // ---- 5: inside bits ------------------------------------------------------------------------------------------------
    const int nx = n3[0], ny = n3[1], nz = n3[2], nxy = nx * ny, tot = nx + ny + nz;
    const int nwords = (npts + 31) >> 5;
    if (status == CELL_OK)
      for (int w = gl; w < nwords + 1; w += FZ_GL) C.bits[w] = 0u;
    const bool ok = status == CELL_OK;
    const bool tab = ok && nn * tot <= FZ_TAB;
    float *T = reinterpret_cast<float *>(C.v);
    if (tab) {
      const float bx = idx2phys1(lo[0], g.step[0], g.gmin[0]), by = idx2phys1(lo[1], g.step[1], g.gmin[1]), bz = idx2phys1(lo[2], g.step[2], g.gmin[2]);
      for (int e = gl; e < nn * tot; e += FZ_GL) {
        const int f = e / tot, r = e - f * tot;
        const float *pl = planes + 6 * f;
        float val;
        if (r < nx) val = fmul(pl[0], fsub(fadd(bx, fmul((float)r, g.step[0])), pl[3]));
        else if (r < nx + ny) val = fmul(pl[1], fsub(fadd(by, fmul((float)(r - nx), g.step[1])), pl[4]));
        else val = fmul(pl[2], fsub(fadd(bz, fmul((float)(r - nx - ny), g.step[2])), pl[5]));
        T[e] = val;
      }
    }
    __syncwarp();
    const float neg_eps = -g.eps;
    {
      // columns (i, j) across the lanes, k swept in chunks with the running max / min in registers
      const int cols_any = __reduce_max_sync(FULL, tab ? nxy : 0);
      const int nz_any = __reduce_max_sync(FULL, tab ? nz : 0);
      const int nn_any = __reduce_max_sync(FULL, tab ? nn : 0);
      if (nz_any <= 4) fz_columns<4>(C.bits, T, tab, nx, ny, nz, nn, cols_any, nz_any, nn_any, g.eps, gl, grp);
      else fz_columns<8>(C.bits, T, tab, nx, ny, nz, nn, cols_any, nz_any, nn_any, g.eps, gl, grp);
    }
    {
      // cells whose tables do not fit: points across the lanes, the reference's expression per point and face
      const bool gen = ok && !tab;
      const int pts_any = __reduce_max_sync(FULL, gen ? npts : 0);
      if (pts_any) {
        const float bx = idx2phys1(lo[0], g.step[0], g.gmin[0]), by = idx2phys1(lo[1], g.step[1], g.gmin[1]), bz = idx2phys1(lo[2], g.step[2], g.gmin[2]);
        for (int l0 = 0; l0 < pts_any; l0 += FZ_GL) {
          const int l = l0 + gl;
          const bool pact = gen && l < npts;
          const int k = pact ? l / nxy : 0, r = pact ? l - k * nxy : 0, j = r / (nx > 0 ? nx : 1), i = r - j * nx;
          const float pt[3] = {fadd(bx, fmul((float)i, g.step[0])), fadd(by, fmul((float)j, g.step[1])), fadd(bz, fmul((float)k, g.step[2]))};
          float dmax = -INFINITY, dmin = INFINITY;
          if (pact)
            for (int f = 0; f < nn; f++) {
              const float *pl = planes + 6 * f;
              const float dist = fadd(fadd(fmul(pl[0], fsub(pt[0], pl[3])), fmul(pl[1], fsub(pt[1], pl[4]))), fmul(pl[2], fsub(pt[2], pl[5])));
              dmax = fmaxf(dmax, dist);
              dmin = fminf(dmin, dist);
            }
          const bool in = pact && !(dmax > g.eps && dmin < neg_eps);
          const unsigned m = (__ballot_sync(FULL, in) >> (grp * FZ_GL)) & GLMASK;
          if (gl == 0 && gen && m) {
            C.bits[l0 >> 5] |= m << (l0 & 31);
            if ((l0 & 31) > 32 - FZ_GL) C.bits[(l0 >> 5) + 1] |= m >> (32 - (l0 & 31));
          }
        }
      }
    }
    __syncwarp();

    // ---- output ----------------------------------------------------------------------------------------------------------
    uint32_t pool_off = 0xFFFFFFFFu;
    {
      const bool need = status == CELL_OK && nwords > FZ_INLINE_WORDS;
      unsigned long long at = 0;
      if (need && gl == 0) at = atomicAdd(out.pool_cursor, (unsigned long long)nwords);
      at = __shfl_sync(FULL, at, grp * FZ_GL);
      if (need) {
        if (at + (unsigned long long)nwords > out.pool_words || at + (unsigned long long)nwords >= 0xFFFFFFFFull) status = CELL_OVERFLOW;
        else pool_off = (uint32_t)at;
      }
    }
    if (real) {
      if (status == CELL_OK) {
        if (pool_off == 0xFFFFFFFFu) { if (gl < FZ_INLINE_WORDS) out.bits_inline[(size_t)slot * FZ_INLINE_WORDS + gl] = gl < nwords ? C.bits[gl] : 0u; }
        else for (int w = gl; w < nwords; w += FZ_GL) out.bits_pool[(size_t)pool_off + w] = C.bits[w];
      }
      if (gl == 0) {
        CellHdr h;
        h.cell = blk.cell_base + (uint32_t)cell;
        h.blk_nf = ((uint32_t)blk_id << 16) | (uint32_t)(status == CELL_OK ? nn : 0);
        h.lo[0] = lo[0]; h.lo[1] = lo[1]; h.lo[2] = lo[2];
        h.n3[0] = (uint16_t)n3[0]; h.n3[1] = (uint16_t)n3[1]; h.n3[2] = (uint16_t)n3[2];
        h.pad = status == CELL_OK ? 0 : 0xFFFF;
        h.plane_off = pool_off;
        out.hdr[slot] = h;
        if (status == CELL_OVERFLOW) {
          const uint32_t o = atomicAdd(&out.cnt->n_overflow, 1u);
          if (o < out.cap_overflow) out.overflow[o] = make_uint2((unsigned)blk_id, (unsigned)cell);
        }
        c_no_tet += status == CELL_NO_TET;
        c_incomplete += status == CELL_INCOMPLETE;
        c_outside += status == CELL_OUTSIDE;
        c_bad += status == CELL_BAD_MESH;
        if (status == CELL_OK) c_faces += (unsigned long long)((nn + 1) & ~1);
      }
    } else if (have && gl == 0) {
      CellHdr h;
      h.cell = 0; h.blk_nf = 0; h.lo[0] = h.lo[1] = h.lo[2] = 0; h.n3[0] = h.n3[1] = h.n3[2] = 1; h.pad = 0xFFFF; h.plane_off = 0xFFFFFFFFu;
      out.hdr[slot] = h;
    }
    __syncwarp();
  }
  if (c_no_tet) atomicAdd(&out.cnt->n_no_tet, c_no_tet);
  if (c_incomplete) atomicAdd(&out.cnt->n_incomplete, c_incomplete);
  if (c_outside) atomicAdd(&out.cnt->n_outside, c_outside);
  if (c_bad) atomicAdd(&out.cnt->n_bad, c_bad);
  if (c_faces) atomicAdd(&out.cnt->n_faces_fused, c_faces);
}

// ---- the scan-line walk + span emission of one cell per lane (phase 3 of k_cell_scan) -----------------------------------
// h: this lane's header (any lane may be idle: in_sub == false); bits_w / lines_w: the warp's shared-memory slices;
// row(j, k) / inside(i, j, k): the inside bits of the lane's cell.
// Where the span records of the lanes' cells go, and the two counts: per warp (the callers whose warps run on their own) ...
struct WarpSpanAlloc
{
  __device__ __forceinline__ unsigned long long operator()(const SpanOut &out, int nrec, bool in_sub, bool fallback) const
  {
    const unsigned long long base = warp_alloc<unsigned long long>(&out.cnt->n_spans, (unsigned long long)nrec);
    warp_count(&out.cnt->n_deposit, in_sub);
    warp_count(&out.cnt->n_cic_fallback, fallback);
    return base;
  }
};
// ... or per CTA (k_cell_direct: every thread of the CTA arrives here together): three atomics per CTA instead of per warp
template <int NW>
struct CtaSpanAlloc
{
  unsigned long long *scratch;        // shared memory, NW + 3 entries
  __device__ __forceinline__ unsigned long long operator()(const SpanOut &out, int nrec, bool in_sub, bool fallback) const
  {
    const unsigned nd = (unsigned)__popc(__ballot_sync(0xffffffffu, in_sub)), nf = (unsigned)__popc(__ballot_sync(0xffffffffu, fallback));
    if (threadIdx.x == 0) { scratch[NW + 1] = 0ull; scratch[NW + 2] = 0ull; }
    __syncthreads();
    if (lane_id() == 0 && nd) atomicAdd(&scratch[NW + 1], (unsigned long long)nd);
    if (lane_id() == 0 && nf) atomicAdd(&scratch[NW + 2], (unsigned long long)nf);
    const unsigned long long base = cta_alloc<unsigned long long, NW>(&out.cnt->n_spans, (unsigned long long)nrec, scratch);   // (its barriers order the adds above)
    if (threadIdx.x == 0 && scratch[NW + 1]) atomicAdd(&out.cnt->n_deposit, scratch[NW + 1]);
    if (threadIdx.x == 0 && scratch[NW + 2]) atomicAdd(&out.cnt->n_cic_fallback, scratch[NW + 2]);
    return base;
  }
};

template <class Row, class Inside, class Alloc = WarpSpanAlloc>
__device__ __forceinline__ void scan_emit_lane(const CellHdr &h, bool in_sub, Row row, Inside inside, uint32_t *lines_w, int lane,
                                               const ScanCtx &sc, const GridGeom &g, const DevBlock *__restrict__ blocks, const SpanOut &out,
                                               Alloc alloc = Alloc())
{
  const int e = (int)(h.blk_nf >> 16);
  int tot = 0, nrec = 0, nlines = 0;
  bool local_box = false, bitpath = false;
  float site[3] = {0, 0, 0};
  const int nx = (int)h.n3[0], ny = (int)h.n3[1], nz = (int)h.n3[2];
  if (in_sub) {
    int lo[3] = {h.lo[0], h.lo[1], h.lo[2]}, n3[3] = {nx, ny, nz};
    local_box = box_is_local(sc.boxes[e], lo, n3, sc.project);
    bitpath = nx <= 32 && ny < 2048 && nz < 2048;
    if (bitpath) {
      LineKeeper lk{lines_w + lane, 0};
      tot = scan_cell_bits(nx, ny, nz, row, lk);
      nlines = lk.n;
      if (nlines <= SCAN_LINE_CAP && local_box) nrec = nlines;
      else if (nlines <= SCAN_LINE_CAP) {
        CountEmit ce{0};
        LineEmitter<CountEmit> le{sc, e, h.cell, h.lo, local_box, 0.0f, ce, 0, {0, 0, 0, 0, 0, 0, 0, 0}, -1};
        le.find_candidates(n3);
        for (int q = 0; q < nlines; q++) {
          uint32_t pk = lines_w[q * 32 + lane];
          le((int)(pk & 2047u), (int)((pk >> 11) & 2047u), (int)((pk >> 22) & 31u), (int)(pk >> 27));
        }
        nrec = ce.n;
      } else {
        CountEmit ce{0};
        LineEmitter<CountEmit> le{sc, e, h.cell, h.lo, local_box, 0.0f, ce, 0, {0, 0, 0, 0, 0, 0, 0, 0}, -1};
        le.find_candidates(n3);
        scan_cell_bits(nx, ny, nz, row, le);
        nrec = ce.n;
      }
    } else {
      CountEmit ce{0};
      LineEmitter<CountEmit> le{sc, e, h.cell, h.lo, local_box, 0.0f, ce, 0, {0, 0, 0, 0, 0, 0, 0, 0}, -1};
      le.find_candidates(n3);
      tot = scan_cell(nx, ny, nz, inside, le);
      nrec = ce.n;
    }
    if (tot == 0) {
      const DevBlock &b = blocks[e];
      uint32_t lc = h.cell - b.cell_base;
      site[0] = b.particles[3 * (size_t)lc]; site[1] = b.particles[3 * (size_t)lc + 1]; site[2] = b.particles[3 * (size_t)lc + 2];
      CountEmit ce2{0};
      emit_cic(sc, e, h.cell, site, g, 0, ce2);
      nrec = ce2.n;
    }
  }
  __syncwarp();
  const unsigned long long base = alloc(out, nrec, in_sub, in_sub && tot == 0);
  if (in_sub) {
    StoreEmit se{out.keys, out.data, base, out.capacity};
    if (tot > 0) {
      float m = fdiv(g.mass, (float)tot); // src/dense.cpp:1692
      LineEmitter<StoreEmit> le{sc, e, h.cell, h.lo, local_box, m, se, 0, {0, 0, 0, 0, 0, 0, 0, 0}, -1};
      const int n3e[3] = {nx, ny, nz};
      le.find_candidates(n3e);
      if (bitpath && nlines <= SCAN_LINE_CAP) {
        for (int q = 0; q < nlines; q++) {
          uint32_t pk = lines_w[q * 32 + lane];
          le((int)(pk & 2047u), (int)((pk >> 11) & 2047u), (int)((pk >> 22) & 31u), (int)(pk >> 27));
        }
      } else if (bitpath) {
        scan_cell_bits(nx, ny, nz, row, le);
      } else {
        scan_cell(nx, ny, nz, inside, le);
      }
    } else {
      emit_cic(sc, e, h.cell, site, g, 0, se);
    }
  }
  __syncwarp();
}


// ---- k_cell_direct: one THREAD per cell for the small index boxes ---------------------------------------------------------
// A cell whose index box holds at most M points along every axis (M = 2, 3, 4: two thirds of the cells of a clustered
// set at two grid points per particle spacing) needs no plane storage at all: the thread walks its Voronoi faces one
// after the other (fill_edge_link + NewellNormal, src/tet.cpp:389-409, src/dense.cpp:682-735) and applies every plane to
// all M^3 points of the box at once -- the two "significant" flags of PtInCell (src/dense.cpp:1172-1203) are two 64-bit
// masks with compile-time bit positions, the distance is the axis-separable form fadd(fadd(A[i], B[j]), C[k]) -- then
// runs the scan-line state machine on the mask and writes its span records.  No shared memory besides the kept scan
// lines, no plane pool, no inside-bit buffer: replaces k_cell_faces + k_cell_scan for these cells.
template <int M>
struct RegRow
{
  unsigned long long bits;
  __device__ __forceinline__ uint32_t operator()(int j, int k) const { return (uint32_t)(bits >> ((k * M + j) * M)); }
};
template <int M>
struct RegInside
{
  unsigned long long bits;
  __device__ __forceinline__ bool operator()(int i, int j, int k) const { return (bits >> ((k * M + j) * M + i)) & 1ull; }
};

constexpr int DIRECT_THREADS = 128;
template <int M>
struct DirectSmem
{
  static constexpr int W = (M * M * M + 31) / 32;   // words of one cell's flags
  CellHdr hdr[DIRECT_THREADS];
  int off[DIRECT_THREADS + 1];                      // exclusive prefix of the cells' face counts
  alignas(16) int warp_tot[DIRECT_THREADS / 32];     // own 16 bytes: the compiler reads the four totals with one 128-bit load
  uint32_t pos[DIRECT_THREADS][W], neg[DIRECT_THREADS][W];
  uint32_t lines[DIRECT_THREADS / 32][SCAN_LINE_CAP * 32];
  unsigned long long alloc[DIRECT_THREADS / 32 + 3];  // CtaSpanAlloc's scratch
};

// One CTA = 128 cells of one class.  Step 2 runs one thread per FACE over all faces of the CTA's cells (the lanes of a
// warp then differ only in the length of their edge links), the flags of a cell are OR-ed into shared memory; step 3 is
// one thread per cell.  (A first version with one thread per cell for everything ran at 8 of 32 active lanes: the cells
// of a warp differ in face count AND link length; profiles/r02/direct_a_ncu.txt.)
template <int M>
__global__ void __launch_bounds__(DIRECT_THREADS) k_cell_direct(const CellHdr *__restrict__ hdrs, const FaceRef *__restrict__ faces,
                                                               const DevBlock *__restrict__ blocks, ScanCtx sc, const __grid_constant__ GridGeom g,
                                                               SpanOut out, const unsigned int *range_lo, const unsigned int *range_hi, uint32_t range_cap,
                                                               uint32_t n_fixed)
{
  static_assert(M * M * M <= 64, "the inside flags of a cell live in one 64-bit mask");
  __shared__ DirectSmem<M> S;
  constexpr int W = DirectSmem<M>::W;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t lo_i = 0, hi_i = n_fixed;
  if (range_hi) {
    lo_i = range_lo ? *range_lo : 0u;
    hi_i = *range_hi < range_cap ? *range_hi : range_cap;
  }
  const uint32_t c0 = lo_i + blockIdx.x * DIRECT_THREADS;
  if (c0 >= hi_i) return;                            // the whole CTA leaves together
  const uint32_t idx = c0 + (uint32_t)tid;
  const bool valid = idx < hi_i;
  CellHdr h;
  h.cell = 0; h.blk_nf = 0; h.lo[0] = h.lo[1] = h.lo[2] = 0; h.n3[0] = h.n3[1] = h.n3[2] = 1; h.pad = 0; h.plane_off = 0;
  if (valid) h = hdrs[idx];
  // ---- 1: headers and the prefix of the face counts ----
  S.hdr[tid] = h;
#pragma unroll
  for (int w = 0; w < W; w++) { S.pos[tid][w] = 0u; S.neg[tid][w] = 0u; }
  const int nf_mine = valid ? (int)(h.blk_nf & 0xffffu) : 0;
  const int incl = warp_incl_scan(nf_mine);
  if (lane == 31) S.warp_tot[warp] = incl;
  __syncthreads();
  int wbase = 0;
#pragma unroll
  for (int w = 0; w < DIRECT_THREADS / 32; w++) wbase += w < warp ? S.warp_tot[w] : 0;
  S.off[tid] = wbase + incl - nf_mine;
  if (tid == DIRECT_THREADS - 1) S.off[DIRECT_THREADS] = wbase + incl;
  __syncthreads();
  // ---- 2: one thread per Voronoi face ----
  const int ftot = S.off[DIRECT_THREADS];
  const float neg_eps = -g.eps;
  for (int q0 = 0; q0 < ftot; q0 += DIRECT_THREADS) {
    const int q = q0 + tid;
    const bool fact = q < ftot;
    int c = 0, n = -1;
    float site[3] = {0.0f, 0.0f, 0.0f};
    FaceAccum fa;
    fa.cmin = nullptr; fa.cmax = nullptr;
    if (fact) {
#pragma unroll
      for (int step = DIRECT_THREADS / 2; step >= 1; step >>= 1)
        if (S.off[c + step] <= q) c += step;           // the last cell whose first face is at or before q
      const int f = q - S.off[c];
      const DevBlock &b = blocks[(int)(S.hdr[c].blk_nf >> 16)];
      const FaceRef r = faces[(size_t)S.hdr[c].plane_off * 2 + f];
      site[0] = b.particles[3 * (size_t)r.site]; site[1] = b.particles[3 * (size_t)r.site + 1]; site[2] = b.particles[3 * (size_t)r.site + 2];
      if (b.walk) {
        const int4 v0 = b.tets[2 * (size_t)r.ut];
        const int s_c = v0.x == r.site ? 0 : (v0.y == r.site ? 1 : (v0.z == r.site ? 2 : 3));
        const int s_u = v0.x == r.u ? 0 : (v0.y == r.u ? 1 : (v0.z == r.u ? 2 : 3));
        n = walk_edge_link_rec(s_c, s_u, r.ut, b.walk, fa);
      } else {
        n = walk_edge_link(r.site, r.u, r.ut, b.tets, b.cc, fa);
      }
      if (n < 0) atomicAdd(&out.cnt->n_bad, 1ull);     // the link did not close: a NaN plane, never significant
    }
    // the links of a warp's faces differ in length: meet again before the (uniform) rest of the face
    __syncwarp();
    if (!fact || n < 0) continue;
    const CellHdr &hc = S.hdr[c];
    newell_term(fa.nrm, fa.prev, fa.v0);
    newell_finish(fa.nrm, fa.v0, site);
    // probe positions: cell_min_grid_pos + i * step (src/dense.cpp:1404,1530-1532); per-axis products of this plane
    const float bx = idx2phys1(hc.lo[0], g.step[0], g.gmin[0]), by = idx2phys1(hc.lo[1], g.step[1], g.gmin[1]), bz = idx2phys1(hc.lo[2], g.step[2], g.gmin[2]);
    float A[M], B[M], C[M];
#pragma unroll
    for (int i = 0; i < M; i++) {
      A[i] = fmul(fa.nrm[0], fsub(fadd(bx, fmul((float)i, g.step[0])), fa.v0[0]));
      B[i] = fmul(fa.nrm[1], fsub(fadd(by, fmul((float)i, g.step[1])), fa.v0[1]));
      C[i] = fmul(fa.nrm[2], fsub(fadd(bz, fmul((float)i, g.step[2])), fa.v0[2]));
    }
    unsigned long long pos = 0ull, neg = 0ull;
#pragma unroll
    for (int j = 0; j < M; j++)
#pragma unroll
      for (int i = 0; i < M; i++) {
        const float ab = fadd(A[i], B[j]);
#pragma unroll
        for (int k = 0; k < M; k++) {
          const float dist = fadd(ab, C[k]);
          const unsigned long long bit = 1ull << ((k * M + j) * M + i);
          if (dist > g.eps) pos |= bit;
          if (dist < neg_eps) neg |= bit;
        }
      }
    if ((uint32_t)pos) atomicOr(&S.pos[c][0], (uint32_t)pos);
    if ((uint32_t)neg) atomicOr(&S.neg[c][0], (uint32_t)neg);
    if (W > 1) {
      if ((uint32_t)(pos >> 32)) atomicOr(&S.pos[c][W - 1], (uint32_t)(pos >> 32));
      if ((uint32_t)(neg >> 32)) atomicOr(&S.neg[c][W - 1], (uint32_t)(neg >> 32));
    }
  }
  __syncthreads();
  // ---- 3: one thread per cell: the points the box really has, the scan-line walk, the span records ----
  unsigned long long pos = S.pos[tid][0], neg = S.neg[tid][0];
  if (W > 1) { pos |= (unsigned long long)S.pos[tid][W - 1] << 32; neg |= (unsigned long long)S.neg[tid][W - 1] << 32; }
  unsigned long long vm = 0ull;
  {
    const unsigned long long rowm = (1ull << h.n3[0]) - 1ull;
#pragma unroll
    for (int k = 0; k < M; k++)
#pragma unroll
      for (int j = 0; j < M; j++)
        if (k < (int)h.n3[2] && j < (int)h.n3[1]) vm |= rowm << ((k * M + j) * M);
  }
  const unsigned long long bits = ~(pos & neg) & vm;
  scan_emit_lane(h, valid, RegRow<M>{bits}, RegInside<M>{bits}, S.lines[warp], lane, sc, g, blocks, out, CtaSpanAlloc<DIRECT_THREADS / 32>{S.alloc});
}

constexpr int EMIT_WARPS = 8;
constexpr int EMIT_THREADS = EMIT_WARPS * 32;
constexpr int EMIT_WARP_BYTES = (SCAN_BIT_WORDS + 4) * 4 + SCAN_LINE_CAP * 32 * 4;
constexpr size_t EMIT_SMEM = (size_t)EMIT_WARPS * EMIT_WARP_BYTES;

__global__ void __launch_bounds__(EMIT_THREADS) k_cell_emit(const CellHdr *__restrict__ hdrs, uint32_t n_hdrs, const uint32_t *__restrict__ bits_inline,
                                                             const uint32_t *__restrict__ bits_pool, const DevBlock *__restrict__ blocks, ScanCtx sc,
                                                             const __grid_constant__ GridGeom g, SpanOut out)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t *bits_w = reinterpret_cast<uint32_t *>(smem_raw + (size_t)warp * EMIT_WARP_BYTES);
  uint32_t *lines_w = bits_w + SCAN_BIT_WORDS + 4;
  const uint32_t first = (blockIdx.x * EMIT_WARPS + warp) * 32u;
  if (first >= n_hdrs) return;
  const int ncell = (int)(n_hdrs - first < 32u ? n_hdrs - first : 32u);
  CellHdr h;
  h.cell = 0; h.blk_nf = 0; h.lo[0] = h.lo[1] = h.lo[2] = 0; h.n3[0] = h.n3[1] = h.n3[2] = 1; h.pad = 0xFFFF; h.plane_off = 0xFFFFFFFFu;
  if (lane < ncell) h = hdrs[first + lane];
  const bool valid = lane < ncell && h.pad == 0;
  const int npts = valid ? (int)h.n3[0] * (int)h.n3[1] * (int)h.n3[2] : 0;
  const int pts32 = (npts + 31) & ~31;
  const uint32_t *src = h.plane_off == 0xFFFFFFFFu ? bits_inline + (size_t)(first + lane) * FZ_INLINE_WORDS : bits_pool + (size_t)h.plane_off;
  int c0 = 0;
  while (c0 < ncell) {
    const int incl_p = warp_incl_scan(lane >= c0 ? pts32 : 0);
    const bool fits = lane >= c0 && lane < ncell && incl_p <= SCAN_BIT_WORDS * 32;
    const unsigned fm = __ballot_sync(0xffffffffu, fits);
    const unsigned run = fm >> c0;
    int c1 = c0 + (run == 0xffffffffu ? 32 : __ffs(~run) - 1);
    if (c1 > ncell) c1 = ncell;
    if (c1 == c0) c1 = c0 + 1;                       // cannot happen (a cell holds at most FZ_PTS points): never loop forever
    const bool in_sub = valid && lane >= c0 && lane < c1;
    const int p_off = incl_p - pts32;
    if (in_sub) {
      const int nw = pts32 >> 5;
      for (int w = 0; w < nw; w++) bits_w[(p_off >> 5) + w] = src[w];
    }
    __syncwarp();
    scan_emit_lane(h, in_sub, RowBits{bits_w, (uint32_t)p_off, (int)h.n3[0], (int)h.n3[1]}, BitsInside{bits_w, (uint32_t)p_off, (int)h.n3[0], (int)h.n3[1]},
                   lines_w, lane, sc, g, blocks, out);
    c0 = c1;
  }
}

} // namespace tb
