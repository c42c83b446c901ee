// kernels.cuh -- sm_100a kernels of the dense stage.  See DESIGN.md for the data layout in HBM,
// the algorithmic bytes of each kernel and the roofline that bounds it.
//
//   k_vert_to_tet      fill_vert_to_tet            (src/tess.cpp:767-787)
//   k_circumcenters    fill_circumcenters          (src/volume.cpp:6-11, src/tet.cpp:37-66)
//   k_cell_bfs         complete + neighbor_edges + the cell bbox, data-bounds filter and index box of
//                      CellBounds / CellGridPts (src/tet.cpp:228-270,337-378; src/dense.cpp:700-714,1381-1410)
//   k_cell_faces       fill_edge_link + circumcenters + NewellNormal of CellBounds, one thread per face
//                      (src/tet.cpp:389-409; src/dense.cpp:682-735,1131-1162)
//   k_cell_scan        PtInCell over the cell's index box + CellInteriorGridPts + CIC fallback
//                      (src/dense.cpp:1172-1203,1475-1701,1437-1455) -> span records
//   k_cell_scan_big    same for cells with a large index box / many faces
//   k_cic              IterateCellsCic             (src/dense.cpp:486-562) -> span records
//   k_row_starts, k_rows  the accumulate step of IterateCells / recvd_pts
//                      (src/dense.cpp:286-290,187-193), one warp per grid row, every grid point
//                      written exactly once
//   k_cell_volumes     complete() + volume()       (src/tet.cpp:337-378, src/volume.cpp:13-54)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "cell_core.cuh"

namespace tb
{

// GridGeom::alg value of the per-site volume pipeline (tessb200_cell_volumes): the star kernels skip the data-bounds filter and
// the index box of the dense stage (every complete cell is kept, whatever its extent)
constexpr int TB_ALG_VOLUMES = -1;

// ---- device-side descriptors -------------------------------------------------------------------
struct DevBlock
{
  const float *particles;  // xyz AoS
  const int4 *tets;        // 2 x int4 per tet: verts, neighbours
  const int *v2t;
  const float4 *cc;        // circumcenter per tet, w = tet volume
  const WalkRec *walk;     // 32-byte circulation record per tet (neighbours, circumcenter, slot permutation)
  const unsigned char *hull;   // 1 = the vertex belongs to a hull face, i.e. complete() is false (src/tet.cpp:337-378)
  int num_orig, num_particles, num_tets;
  uint32_t cell_base;      // global number of this block's cell 0 (blocks in ascending gid order)
  const uint32_t *order;   // cells of the block in Morton order of their sites (processing order only)
  uint32_t cta_start;      // first CTA of this block in the all-blocks BFS launch (k_cell_bfs)
  uint32_t slot_start;     // first cell slot of this block in its group (k_cell_fused: slots are dense over the group's blocks)
};

// One accepted cell handed from k_cell_topo to the scan kernels (32 bytes)
struct __align__(16) CellHdr
{
  uint32_t cell;           // global cell number
  uint32_t blk_nf;         // block index << 16 | number of faces
  int lo[3];               // cell_min_grid_idx (src/dense.cpp:1394-1396)
  uint16_t n3[3];          // cell_grid_pts     (src/dense.cpp:1406-1410)
  uint16_t pad;
  uint32_t plane_off;      // offset into the plane pool in units of 2 faces (48 bytes)
};

// The counters every warp (or CTA) of a kernel adds to sit on 128-byte lines of their own: atomics on one line are served one
// after the other by one L2 slice (measured: half a million adds on one address cost k_cic_prepare 1 ms of its 2), so
// the hot ones neither share a line with each other nor with the marks the kernels read.
struct alignas(128) HotU32 { unsigned int v; };
struct Counters
{
  unsigned long long n_no_tet, n_incomplete, n_outside, n_bad;
  unsigned int n_big, n_overflow;              // list lengths (rare appends)
  unsigned long long big_bits;                 // bits needed by the big-cell list (multiples of 32)
  // progress marks kept on the device (k_advance): the cell kernels of a group work on what was
  // appended since the previous group, so the host never has to read a count between launches
  unsigned int pairs_done, small_done, ovf_done;
  unsigned int dir_done[3];
  unsigned long long n_faces_fused;            // Voronoi faces (padded to pairs) of the cells k_cell_fused accepted
  unsigned long long pool_cursor;              // words of the inside-bit pool handed out by k_cell_fused
  unsigned int n_big_points;                   // grid points with more than POINT_SMALL deposits (k_point_apply -> k_point_apply_big)
  unsigned int dep_flags;                      // 1: the segment buffer was too small for the shared deposits (the run falls back to the sorted path)
  alignas(128) unsigned long long n_deposit;
  alignas(128) unsigned long long n_cic_fallback;
  alignas(128) unsigned int n_small;           // list length
  alignas(128) unsigned int plane_cursor;      // units of 2 faces
  alignas(128) unsigned long long n_spans;     // span records requested (may exceed capacity)
  alignas(128) unsigned long long n_cands;     // candidate neighbours written by k_cell_bfs
  alignas(128) unsigned long long n_shared;    // one-point records k_span_place handed to the sorted path
  HotU32 n_dir[3];                             // cells of the three small-box classes handed to k_cell_direct (list lengths)
};

struct FaceRef;
struct TopoOut
{
  CellHdr *small, *big;
  unsigned long long *big_bit_off;
  uint2 *overflow;         // (block index, block-local cell) whose star did not fit the fast workspace
  float *plane_pool;
  struct FaceRef *faces;   // face list, parallel to the plane pool
  Counters *cnt;
  uint32_t cap_small, cap_big, cap_overflow, cap_pairs;
  CellHdr *dir[3];         // cells whose index box is at most 2 / 3 / 4 points along every axis: k_cell_direct
  uint32_t cap_dir;        // 0: no direct classes
};

// class of an accepted cell for k_cell_direct: 0, 1, 2 for index boxes of at most 2, 3, 4 points per axis, else -1
__device__ __forceinline__ int direct_class(const int *n3)
{
  const int mx = n3[0] > n3[1] ? (n3[0] > n3[2] ? n3[0] : n3[2]) : (n3[1] > n3[2] ? n3[1] : n3[2]);
  return mx <= 2 ? 0 : (mx <= 3 ? 1 : (mx <= 4 ? 2 : -1));
}
constexpr int FACE_DIRECT = 0x40000000;   // FaceRef::blk flag: the cell goes through k_cell_direct, k_cell_faces skips the face

struct SpanOut
{
  uint64_t *keys, *data;
  unsigned long long capacity;
  Counters *cnt;
};

// ---- warp helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// every lane of a fully converged warp calls this; returns the slot of lanes with pred set
template <class T>
__device__ __forceinline__ T warp_append(T *counter, bool pred)
{
  unsigned m = __ballot_sync(0xffffffffu, pred);
  T base = 0;
  if (m) {
    int leader = __ffs(m) - 1;
    if ((int)lane_id() == leader) base = atomicAdd(counter, (T)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
  }
  return base + (T)__popc(m & ((1u << lane_id()) - 1u));
}

// warp-aggregated allocation of `want` units per lane (0 allowed); returns the lane's base
template <class T>
__device__ __forceinline__ T warp_alloc(T *counter, T want)
{
  T incl = want;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T o = __shfl_up_sync(0xffffffffu, incl, d);
    if ((int)lane_id() >= d) incl += o;
  }
  T total = __shfl_sync(0xffffffffu, incl, 31);
  T base = 0;
  if (total) {
    if (lane_id() == 31) base = atomicAdd(counter, total);
    base = __shfl_sync(0xffffffffu, base, 31);
  }
  return base + incl - want;
}

__device__ __forceinline__ unsigned long long warp_incl_scan_ull(unsigned long long v)
{
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    unsigned long long o = __shfl_up_sync(0xffffffffu, v, d);
    if ((int)lane_id() >= d) v += o;
  }
  return v;
}

__device__ __forceinline__ void warp_count(unsigned long long *counter, bool pred)
{
  unsigned m = __ballot_sync(0xffffffffu, pred);
  if (m && (int)lane_id() == __ffs(m) - 1) atomicAdd(counter, (unsigned long long)__popc(m));
}

// ---- CTA helpers: one atomic per CTA and counter instead of one per warp -----------------------------------
// Every thread of the CTA calls (NW warps, all converged, no early exits before); `scratch` holds NW + 1 entries of T in
// shared memory and is free again on return.
template <class T, int NW>
__device__ __forceinline__ T cta_alloc(T *counter, T want, T *scratch)
{
  T incl = want;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T o = __shfl_up_sync(0xffffffffu, incl, d);
    if ((int)lane_id() >= d) incl += o;
  }
  const int w = (int)(threadIdx.x >> 5);
  if (lane_id() == 31) scratch[w] = incl;
  __syncthreads();
  if (threadIdx.x == 0) {
    T tot = 0;
#pragma unroll
    for (int i = 0; i < NW; i++) {
      const T t = scratch[i];
      scratch[i] = tot;
      tot += t;
    }
    scratch[NW] = tot ? atomicAdd(counter, tot) : (T)0;
  }
  __syncthreads();
  const T base = scratch[NW] + scratch[w] + incl - want;
  __syncthreads();
  return base;
}

// ---- K0: vert_to_tet ("the last one wins" == highest tet index) ----------------------------------
__global__ void k_fill_i32(int *p, int n, int v)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void k_vert_to_tet(const int4 *__restrict__ tets, int num_tets, int *__restrict__ v2t)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= num_tets) return;
  int4 v = tets[2 * (size_t)t];
  atomicMax(&v2t[v.x], t);
  atomicMax(&v2t[v.y], t);
  atomicMax(&v2t[v.z], t);
  atomicMax(&v2t[v.w], t);
}

// xyz -> one 16-byte record per particle: K1 gathers four particles per tet, and a 16-byte gather is one sector
// where three 4-byte gathers at a 12-byte stride are up to two
__global__ void k_pack_particles(const float *__restrict__ particles, int n, float4 *__restrict__ p4)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p4[i] = make_float4(particles[3 * (size_t)i], particles[3 * (size_t)i + 1], particles[3 * (size_t)i + 2], 0.0f);
}

// ---- K1: circumcenters, one thread per tet ---------------------------------------------------------
// reads 16 B of the tet record (verts only) + 4 gathered particles, writes one float4 (x, y, z, volume)
__global__ void __launch_bounds__(256) k_circumcenters(const int4 *__restrict__ tets, int num_tets,
                                                        const float *__restrict__ particles, const float4 *__restrict__ p4, float4 *__restrict__ cc,
                                                        WalkRec *__restrict__ walk, unsigned char *__restrict__ hull)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= num_tets) return;
  int4 v = __ldg(&tets[2 * (size_t)t]);
  float a[3], b[3], c[3], d[3], o[3];
  if (p4) {
    const float4 pa = __ldg(&p4[v.x]), pb = __ldg(&p4[v.y]), pc = __ldg(&p4[v.z]), pd = __ldg(&p4[v.w]);
    a[0] = pa.x; a[1] = pa.y; a[2] = pa.z; b[0] = pb.x; b[1] = pb.y; b[2] = pb.z;
    c[0] = pc.x; c[1] = pc.y; c[2] = pc.z; d[0] = pd.x; d[1] = pd.y; d[2] = pd.z;
  } else {
#pragma unroll
    for (int i = 0; i < 3; i++) {
      a[i] = __ldg(&particles[3 * (size_t)v.x + i]);
      b[i] = __ldg(&particles[3 * (size_t)v.y + i]);
      c[i] = __ldg(&particles[3 * (size_t)v.z + i]);
      d[i] = __ldg(&particles[3 * (size_t)v.w + i]);
    }
  }
  float det;
  circumcenter(a, b, c, d, o, &det);
  cc[t] = make_float4(o[0], o[1], o[2], fdiv(fabsf(det), 6.0f));   // w = tet volume (used by the DTFE mode)
  if (!walk && !hull) return;
  const int4 nb = __ldg(&tets[2 * (size_t)t + 1]);
  const int vv[4] = {v.x, v.y, v.z, v.w}, nn[4] = {nb.x, nb.y, nb.z, nb.w};
  if (walk) {
    WalkRec r;
    r.nb[0] = nb.x; r.nb[1] = nb.y; r.nb[2] = nb.z; r.nb[3] = nb.w;
    r.cx = o[0]; r.cy = o[1]; r.cz = o[2];
    r.perm = walk_perm(vv, nn, tets);
    walk[t] = r;
  }
  if (hull && (nb.x | nb.y | nb.z | nb.w) < 0) {
    // a face without a neighbour: its three vertices have infinite Voronoi cells.  complete()
    // finds exactly these by walking the star; the flag answers it before the walk starts.
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (nn[i] < 0)
        for (int j = 0; j < 4; j++)
          if (j != i) hull[vv[j]] = 1;
  }
}

// ---- K3a part 1: topology + faces, one thread per cell --------------------------------------------
// Shared-memory workspaces of the fast topology kernels: word w of a thread lives at
// base[w * STRIDE] (stride = threads per CTA, so a warp touching the same word index hits 32
// distinct banks).
//   StarRecWS (k_cell_bfs):    star[52] (tet | site slot | parent slot) | vis buckets[16 words]    = 68 words = 272 B
//   StarWS (k_vertex_density): star[52] | parent_idx bytes[13 words] | vis buckets[16 words]       = 81 words = 324 B
//   NbrWS  (k_cell_nbrs): nu[36] | nt[36] | nbr buckets[16 words]                      = 88 words = 352 B
template <int STRIDE>
struct StarWS
{
  int *base;
  static constexpr int PAR = 52, VH = 65, WORDS = 81;
  __device__ __forceinline__ int &star(int i) { return base[(size_t)i * STRIDE]; }
  __device__ __forceinline__ unsigned char &parent_idx(int i)
  {
    return reinterpret_cast<unsigned char *>(&base[(size_t)(PAR + (i >> 2)) * STRIDE])[i & 3];
  }
  __device__ __forceinline__ uint32_t &vis_word(unsigned h) { return reinterpret_cast<uint32_t *>(base)[(size_t)(VH + (int)h) * STRIDE]; }
  __device__ __forceinline__ void hash_clear_vis()
  {
#pragma unroll
    for (int w = VH; w < WORDS; w++) base[(size_t)w * STRIDE] = -1;
  }
};
template <int STRIDE>
struct NbrWS
{
  int *base;
  static constexpr int NT = 36, NH = 72, WORDS = 88;
  __device__ __forceinline__ int &nu(int i) { return base[(size_t)i * STRIDE]; }
  __device__ __forceinline__ int &nt(int i) { return base[(size_t)(NT + i) * STRIDE]; }
  __device__ __forceinline__ uint32_t &nbr_word(unsigned h) { return reinterpret_cast<uint32_t *>(base)[(size_t)(NH + (int)h) * STRIDE]; }
  __device__ __forceinline__ void hash_clear_nbr()
  {
#pragma unroll
    for (int w = NH; w < WORDS; w++) base[(size_t)w * STRIDE] = -1;
  }
};
struct DynStridedWS
{
  int *base;
  size_t stride;
  int star_cap, nbr_cap;
  __device__ __forceinline__ int &star(int i) { return base[(size_t)i * stride]; }
  __device__ __forceinline__ int &nu(int i) { return base[(size_t)(star_cap + i) * stride]; }
  __device__ __forceinline__ int &nt(int i) { return base[(size_t)(star_cap + nbr_cap + i) * stride]; }
};

constexpr int TOPO_THREADS = 128;
constexpr int TOPO_STAR_CAP = 52;
constexpr int TOPO_NBR_CAP = 36;
template <int STRIDE>
struct StarRecWS
{
  int *base;
  static constexpr int VH = 52, WORDS = 68;
  __device__ __forceinline__ int &star(int i) { return base[(size_t)i * STRIDE]; }
  __device__ __forceinline__ uint32_t &vis_word(unsigned h) { return reinterpret_cast<uint32_t *>(base)[(size_t)(VH + (int)h) * STRIDE]; }
  __device__ __forceinline__ void hash_clear_vis()
  {
#pragma unroll
    for (int w = VH; w < WORDS; w++) base[(size_t)w * STRIDE] = -1;
  }
};
constexpr size_t TOPO_SMEM = (size_t)StarWS<TOPO_THREADS>::WORDS * TOPO_THREADS * sizeof(int);
constexpr size_t BFS_SMEM = (size_t)StarRecWS<TOPO_THREADS>::WORDS * TOPO_THREADS * sizeof(int);
constexpr size_t NBRS_SMEM = (size_t)NbrWS<TOPO_THREADS>::WORDS * TOPO_THREADS * sizeof(int);
constexpr int TOPO_CAND_CAP = TOPO_STAR_CAP + 2;   // candidates per cell: 3 at the root + 1 per further star tet
constexpr int BIG_STAR_CAP = 4096;
constexpr int BIG_NBR_CAP = 1024;
constexpr int BIG_SMEM_STAR = 192, BIG_SMEM_NBR = 96;   // shared-memory part of the general star walk (per warp: 1.5 KB)

constexpr int SCAN_FACE_CAP = 32;     // faces per cell held in shared memory by k_cell_scan
constexpr int SCAN_PTS_CAP = 2048;    // index-box points per cell handled by k_cell_scan

// One Voronoi face handed from the BFS kernels to k_cell_faces (16 bytes); slot f of the face list
// is also slot f of the plane pool
struct __align__(16) FaceRef
{
  int site;   // block-local id of the cell's site
  int u;      // Delaunay neighbour (the face is dual to edge (site, u)); -1 = padding slot
  int ut;     // first star tet in BFS order that holds u: the circulation starts there
  int blk;    // block index
};

// After the star walk: data-bounds filter, index box, classification, header + face list.  All lanes
// of the warp call it (lanes without an accepted cell pass status != CELL_OK).
template <class WS>
__device__ __forceinline__ void bfs_finish(int status, int cell, int n_nbr, WS &ws, const float *cmin, const float *cmax,
                                           const DevBlock &blk, int blk_id, const GridGeom &g, const TopoOut &out)
{
  int lo[3] = {0, 0, 0}, n3[3] = {0, 0, 0};
  const bool vol_only = g.alg == TB_ALG_VOLUMES;
  if (status == CELL_OK && !vol_only) {
    // src/dense.cpp:1385-1392
    for (int d = 0; d < 3; d++)
      if (cmin[d] < fsub(g.dmin[d], g.dext_eps[d]) || cmax[d] > fadd(g.dmax[d], g.dext_eps[d])) status = CELL_OUTSIDE;
  }
  long long npts = 0;
  if (status == CELL_OK && vol_only) { n3[0] = n3[1] = n3[2] = 1; npts = 1; }
  if (status == CELL_OK && !vol_only) {
    for (int d = 0; d < 3; d++) {
      lo[d] = phys2idx1(cmin[d], g.step[d], g.gmin[d]);
      int hi = phys2idx1(cmax[d], g.step[d], g.gmin[d]);
      n3[d] = hi - lo[d] + 1;
    }
    // index boxes may start below the grid (given bounds narrower than the data): the span emitter keeps the points
    // that have a grid element, and a projection keeps every z (phys_box, host_geom.hpp).  The floor only stops
    // garbage (a NaN circumcenter casts to INT_MIN).
    if (n3[0] < 1 || n3[1] < 1 || n3[2] < 1 || n3[0] > 32767 || n3[1] > 32767 || n3[2] > 32767 ||
        lo[0] < -(1 << 22) || lo[1] < -(1 << 22) || lo[2] < -(1 << 22))
      status = CELL_BAD_MESH;
    npts = (long long)n3[0] * n3[1] * n3[2];
  }
  bool ok = status == CELL_OK;
  // plane / face-list space in pairs of faces (48-byte units: every cell's planes start 16-byte aligned)
  const uint32_t want = ok ? (uint32_t)((n_nbr + 1) >> 1) : 0u;
  const uint32_t poff = warp_alloc<unsigned int>(&out.cnt->plane_cursor, want);
  // a pool that is too small drops the cell; the host sees the cursor past the capacity and reports it
  ok = ok && poff + want <= out.cap_pairs;
  const int cls = ok && out.cap_dir ? direct_class(n3) : -1;
  if (ok) {
    FaceRef *fr = out.faces + (size_t)poff * 2;
    for (int k = 0; k < n_nbr; k++) {
      FaceRef r;
      r.site = cell; r.u = ws.nu(k); r.ut = ws.nt(k); r.blk = blk_id | (cls >= 0 ? FACE_DIRECT : 0);
      fr[k] = r;
    }
    if (n_nbr & 1) {
      FaceRef r;
      r.site = cell; r.u = -1; r.ut = 0; r.blk = blk_id;
      fr[n_nbr] = r;
    }
  }
  const bool small = ok && cls < 0 && n_nbr <= SCAN_FACE_CAP && npts <= SCAN_PTS_CAP;
  const bool big = ok && cls < 0 && !small;
  CellHdr h;
  h.cell = blk.cell_base + (uint32_t)cell;
  h.blk_nf = ((uint32_t)blk_id << 16) | (uint32_t)n_nbr;
  h.lo[0] = lo[0]; h.lo[1] = lo[1]; h.lo[2] = lo[2];
  h.n3[0] = (uint16_t)n3[0]; h.n3[1] = (uint16_t)n3[1]; h.n3[2] = (uint16_t)n3[2];
  h.pad = 0;
  h.plane_off = poff;
  uint32_t s_slot = warp_append<unsigned int>(&out.cnt->n_small, small);
  if (small && s_slot < out.cap_small) out.small[s_slot] = h;
  uint32_t b_slot = warp_append<unsigned int>(&out.cnt->n_big, big);
  unsigned long long bits = big ? (unsigned long long)((npts + 31) & ~31LL) : 0ull;
  unsigned long long boff = warp_alloc<unsigned long long>(&out.cnt->big_bits, bits);
  if (big && b_slot < out.cap_big) {
    out.big[b_slot] = h;
    out.big_bit_off[b_slot] = boff;
  }
  if (out.cap_dir) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const uint32_t d_slot = warp_append<unsigned int>(&out.cnt->n_dir[c].v, cls == c);
      if (cls == c && d_slot < out.cap_dir) out.dir[c][d_slot] = h;
    }
  }
  warp_count(&out.cnt->n_no_tet, status == CELL_NO_TET);
  warp_count(&out.cnt->n_incomplete, status == CELL_INCOMPLETE);
  warp_count(&out.cnt->n_outside, status == CELL_OUTSIDE);
  warp_count(&out.cnt->n_bad, status == CELL_BAD_MESH);
}

// Morton key of a site inside its block's bounds, 10 bits per axis (processing order only: cells
// that are close in space share star tets, so consecutive threads reuse L1 / L2 lines)
__device__ __forceinline__ uint32_t spread10(uint32_t x)
{
  x &= 0x3ffu;
  x = (x | (x << 16)) & 0x030000ffu;
  x = (x | (x << 8)) & 0x0300f00fu;
  x = (x | (x << 4)) & 0x030c30c3u;
  x = (x | (x << 2)) & 0x09249249u;
  return x;
}
__global__ void k_morton_keys(const float *__restrict__ particles, int n, float3 bmin, float3 inv_ext, uint32_t blk_tag, int drop_bits,
                              uint32_t *__restrict__ keys, uint32_t *__restrict__ ids)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x = (particles[3 * (size_t)i] - bmin.x) * inv_ext.x, y = (particles[3 * (size_t)i + 1] - bmin.y) * inv_ext.y,
        z = (particles[3 * (size_t)i + 2] - bmin.z) * inv_ext.z;
  uint32_t xi = (uint32_t)fminf(fmaxf(x * 1024.0f, 0.0f), 1023.0f), yi = (uint32_t)fminf(fmaxf(y * 1024.0f, 0.0f), 1023.0f),
           zi = (uint32_t)fminf(fmaxf(z * 1024.0f, 0.0f), 1023.0f);
  // block tag above the Morton bits: ONE sort orders the cells of every block, block by block
  keys[i] = blk_tag | ((spread10(xi) | (spread10(yi) << 1) | (spread10(zi) << 2)) >> drop_bits);
  ids[i] = (uint32_t)i;
}

// src/dense.cpp:1385-1392: a cell whose box leaves the data bounds (plus eps) is skipped
__device__ __forceinline__ bool box_outside_data(const float *cmin, const float *cmax, const GridGeom &g)
{
  if (g.alg == TB_ALG_VOLUMES) return false;
  bool out = false;
  for (int d = 0; d < 3; d++) out = out || cmin[d] < fsub(g.dmin[d], g.dext_eps[d]) || cmax[d] > fadd(g.dmax[d], g.dext_eps[d]);
  return out;
}

// K3a part 1a: star BFS + cell bbox + data-bounds filter + index box, one thread per cell, every block
// of a group (all blocks of this GPU when the inputs are resident; one block at a time when the run is
// pipelined against the host-to-device copies) in one launch: CTAs [cta_start_b, cta_start_{b+1}) work
// on block b.  Hands a pre-header and the candidate neighbours (slot-interleaved, so the lanes of a
// warp write and later read consecutive 8-byte entries) to k_cell_nbrs.
struct CandSink
{
  int2 *cand;          // entry k of this thread at cand[k * n_slots]
  size_t n_slots;
  __device__ __forceinline__ void operator()(int k, int u, int t) { cand[(size_t)k * n_slots] = make_int2(u, t); }
};
struct CandSource
{
  const int2 *cand;
  size_t n_slots;
  __device__ __forceinline__ void operator()(int k, int &u, int &t) const
  {
    int2 e = cand[(size_t)k * n_slots];
    u = e.x;
    t = e.y;
  }
};

__global__ void __launch_bounds__(TOPO_THREADS) k_cell_bfs(const DevBlock *__restrict__ blocks, int blk_begin, int blk_end, const __grid_constant__ GridGeom g,
                                                           TopoOut out, CellHdr *__restrict__ pre, int2 *__restrict__ cand)
{
  extern __shared__ int ws_s[];
  int blk_id = blk_begin;
  for (int b = blk_begin + 1; b < blk_end; b++)
    if (blocks[b].num_orig > 0 && blocks[b].tets != nullptr && blockIdx.x >= blocks[b].cta_start) blk_id = b;
  const DevBlock blk = blocks[blk_id];
  const int slot_in_blk = (int)(blockIdx.x - blk.cta_start) * TOPO_THREADS + (int)threadIdx.x;
  const size_t slot = (size_t)blockIdx.x * TOPO_THREADS + threadIdx.x;
  const size_t n_slots = (size_t)gridDim.x * TOPO_THREADS;
  StarRecWS<TOPO_THREADS> ws{ws_s + threadIdx.x};
  int status = -1, n_star = 0, cell = 0;
  float cmin[3] = {INFINITY, INFINITY, INFINITY}, cmax[3] = {-INFINITY, -INFINITY, -INFINITY};
  if (slot_in_blk < blk.num_orig && blk.tets != nullptr) {
    cell = (int)blk.order[slot_in_blk];
    const int t0 = blk.v2t[cell];
    CandSink sink{cand + slot, n_slots};
    status = t0 < 0 ? CELL_NO_TET : star_bfs_rec(cell, t0, blk.tets, blk.walk, ws, TOPO_STAR_CAP, &n_star, cmin, cmax, sink);
    // A star too large for this kernel usually belongs to a cell at the edge of the data whose Voronoi
    // vertices reach beyond the data bounds: the filter of src/dense.cpp:1385-1392 drops it whatever the
    // rest of the star holds (the box only grows; the hull flag says the cell is complete), so the part seen so far decides.
    if (status == CELL_OVERFLOW && !blk.hull[cell] && box_outside_data(cmin, cmax, g)) status = CELL_OUTSIDE;
  }
  __syncwarp();
  const bool ovf = status == CELL_OVERFLOW;
  uint32_t o_slot = warp_append<unsigned int>(&out.cnt->n_overflow, ovf);
  if (ovf && o_slot < out.cap_overflow) out.overflow[o_slot] = make_uint2((unsigned)blk_id, (unsigned)cell);
  int lo[3] = {0, 0, 0}, n3[3] = {0, 0, 0};
  const bool vol_only = g.alg == TB_ALG_VOLUMES;
  if (status == CELL_OK && !vol_only) {
    // src/dense.cpp:1385-1392
    for (int d = 0; d < 3; d++)
      if (cmin[d] < fsub(g.dmin[d], g.dext_eps[d]) || cmax[d] > fadd(g.dmax[d], g.dext_eps[d])) status = CELL_OUTSIDE;
  }
  if (status == CELL_OK && vol_only) n3[0] = n3[1] = n3[2] = 1;
  if (status == CELL_OK && !vol_only) {
    for (int d = 0; d < 3; d++) {
      lo[d] = phys2idx1(cmin[d], g.step[d], g.gmin[d]);
      int hi = phys2idx1(cmax[d], g.step[d], g.gmin[d]);
      n3[d] = hi - lo[d] + 1;
    }
    // index boxes may start below the grid (given bounds narrower than the data): the span emitter keeps the points
    // that have a grid element, and a projection keeps every z (phys_box, host_geom.hpp).  The floor only stops
    // garbage (a NaN circumcenter casts to INT_MIN).
    if (n3[0] < 1 || n3[1] < 1 || n3[2] < 1 || n3[0] > 32767 || n3[1] > 32767 || n3[2] > 32767 ||
        lo[0] < -(1 << 22) || lo[1] < -(1 << 22) || lo[2] < -(1 << 22))
      status = CELL_BAD_MESH;
  }
  CellHdr h;
  h.cell = blk.cell_base + (uint32_t)cell;
  h.blk_nf = ((uint32_t)blk_id << 16) | (uint32_t)(n_star + 2);     // candidates, replaced by the face count in k_cell_nbrs
  h.lo[0] = lo[0]; h.lo[1] = lo[1]; h.lo[2] = lo[2];
  h.n3[0] = (uint16_t)n3[0]; h.n3[1] = (uint16_t)n3[1]; h.n3[2] = (uint16_t)n3[2];
  h.pad = status == CELL_OK ? 0 : 0xFFFF;
  h.plane_off = 0;
  pre[slot] = h;
  warp_count(&out.cnt->n_no_tet, status == CELL_NO_TET);
  warp_count(&out.cnt->n_incomplete, status == CELL_INCOMPLETE);
  warp_count(&out.cnt->n_outside, status == CELL_OUTSIDE);
  {
    // candidates of accepted cells (for the roofline's algorithmic bytes): one add per CTA
    __shared__ unsigned int cta_cands;
    if (threadIdx.x == 0) cta_cands = 0u;
    __syncthreads();
    unsigned long long nc = warp_incl_scan_ull(status == CELL_OK ? (unsigned long long)(n_star + 2) : 0ull);
    if (lane_id() == 31 && nc) atomicAdd(&cta_cands, (unsigned int)nc);
    __syncthreads();
    if (threadIdx.x == 0 && cta_cands) atomicAdd(&out.cnt->n_cands, (unsigned long long)cta_cands);
  }
  warp_count(&out.cnt->n_bad, status == CELL_BAD_MESH);
}

// K3a part 1a': candidates -> distinct Delaunay neighbours with their first tet (neighbor_edges' list),
// then the warp-aggregated append of the face list, the plane space and the header.  Same launch shape
// as k_cell_bfs (thread <-> slot).
__global__ void __launch_bounds__(TOPO_THREADS) k_cell_nbrs(const DevBlock *__restrict__ blocks, TopoOut out, const CellHdr *__restrict__ pre,
                                                            const int2 *__restrict__ cand)
{
  extern __shared__ int ws_s[];
  const size_t slot = (size_t)blockIdx.x * TOPO_THREADS + threadIdx.x;
  const size_t n_slots = (size_t)gridDim.x * TOPO_THREADS;
  NbrWS<TOPO_THREADS> ws{ws_s + threadIdx.x};
  CellHdr h = pre[slot];
  const int blk_id = (int)(h.blk_nf >> 16);
  bool ok = h.pad == 0;
  int nn = 0;
  if (ok) {
    CandSource src{cand + slot, n_slots};
    nn = nbrs_from_cands(ws, (int)(h.blk_nf & 0xffffu), TOPO_NBR_CAP, src);
  }
  __syncwarp();
  const bool ovf = ok && nn < 0;
  const uint32_t cell_local = ok ? h.cell - blocks[blk_id].cell_base : 0u;
  uint32_t o_slot = warp_append<unsigned int>(&out.cnt->n_overflow, ovf);
  if (ovf && o_slot < out.cap_overflow) out.overflow[o_slot] = make_uint2((unsigned)blk_id, cell_local);
  ok = ok && nn >= 0;
  // plane / face-list space in pairs of faces (48-byte units: every cell's planes start 16-byte aligned)
  const uint32_t want = ok ? (uint32_t)((nn + 1) >> 1) : 0u;
  const uint32_t poff = warp_alloc<unsigned int>(&out.cnt->plane_cursor, want);
  ok = ok && poff + want <= out.cap_pairs;
  const int n3i[3] = {(int)h.n3[0], (int)h.n3[1], (int)h.n3[2]};
  const int cls = ok && out.cap_dir ? direct_class(n3i) : -1;
  if (ok) {
    FaceRef *fr = out.faces + (size_t)poff * 2;
    for (int k = 0; k < nn; k++) {
      FaceRef r;
      r.site = (int)cell_local; r.u = ws.nu(k); r.ut = ws.nt(k); r.blk = blk_id | (cls >= 0 ? FACE_DIRECT : 0);
      fr[k] = r;
    }
    if (nn & 1) {
      FaceRef r;
      r.site = (int)cell_local; r.u = -1; r.ut = 0; r.blk = blk_id;
      fr[nn] = r;
    }
  }
  const long long npts = (long long)h.n3[0] * h.n3[1] * h.n3[2];
  const bool small = ok && cls < 0 && nn <= SCAN_FACE_CAP && npts <= SCAN_PTS_CAP;
  const bool big = ok && cls < 0 && !small;
  h.blk_nf = ((uint32_t)blk_id << 16) | (uint32_t)(ok ? nn : 0);
  h.plane_off = poff;
  uint32_t s_slot = warp_append<unsigned int>(&out.cnt->n_small, small);
  if (small && s_slot < out.cap_small) out.small[s_slot] = h;
  if (out.cap_dir) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const uint32_t d_slot = warp_append<unsigned int>(&out.cnt->n_dir[c].v, cls == c);
      if (cls == c && d_slot < out.cap_dir) out.dir[c][d_slot] = h;
    }
  }
  uint32_t b_slot = warp_append<unsigned int>(&out.cnt->n_big, big);
  unsigned long long bits = big ? (unsigned long long)((npts + 31) & ~31LL) : 0ull;
  unsigned long long boff = warp_alloc<unsigned long long>(&out.cnt->big_bits, bits);
  if (big && b_slot < out.cap_big) {
    out.big[b_slot] = h;
    out.big_bit_off[b_slot] = boff;
  }
}

// The general star walk for the (rare) cells whose star exceeds the shared-memory workspace or is
// not a manifold: one WARP per cell, lists in global memory, the two "already seen?" searches done
// by all lanes at once.  Same order and results as star_and_neighbors (cell_core.cuh).
struct GlobalListWS
{
  int *star_, *nu_, *nt_;
  __device__ __forceinline__ int nu(int i) const { return nu_[i]; }
  __device__ __forceinline__ int nt(int i) const { return nt_[i]; }
};

__device__ __forceinline__ bool warp_contains(const int *list, int n, int key)
{
  bool f = false;
  for (int j = (int)lane_id(); j < n; j += 32) f |= list[j] == key;
  return __any_sync(0xffffffffu, f);
}

__global__ void __launch_bounds__(128) k_cell_bfs_big(const DevBlock *__restrict__ blocks, const __grid_constant__ GridGeom g, TopoOut out,
                                                       const uint2 *__restrict__ cells, int *ws_g)
{
  // persistent warps: the list length is read on the device (no host round trip before the launch),
  // each warp keeps one workspace and takes cells warp_id, warp_id + n_warps, ...
  const int n_warps = (int)((gridDim.x * blockDim.x) >> 5);
  const int warp_id = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = (int)lane_id();
  const unsigned n_list = out.cnt->n_overflow < out.cap_overflow ? out.cnt->n_overflow : out.cap_overflow;
  const unsigned n_first = out.cnt->ovf_done;     // cells of earlier groups are done (k_advance)
  // lists start in shared memory (most oversized stars are only a little over the fast kernels'
  // capacity) and move to the warp's global workspace when they outgrow it
  __shared__ int lists_s[4][BIG_SMEM_STAR + 2 * BIG_SMEM_NBR];
  int *const sm = lists_s[threadIdx.x >> 5];
  int *const gbase = ws_g + (size_t)warp_id * (BIG_STAR_CAP + 2 * BIG_NBR_CAP);
  for (unsigned wi = n_first + (unsigned)warp_id; wi < n_list; wi += (unsigned)n_warps) {
    const int blk_id = (int)cells[wi].x, cell = (int)cells[wi].y;
    const DevBlock blk = blocks[blk_id];
    GlobalListWS ws{sm, sm + BIG_SMEM_STAR, sm + BIG_SMEM_STAR + BIG_SMEM_NBR};
    int star_room = BIG_SMEM_STAR, nbr_room = BIG_SMEM_NBR;
    float cmin[3] = {INFINITY, INFINITY, INFINITY}, cmax[3] = {-INFINITY, -INFINITY, -INFINITY};
    int status = blk.hull[cell] ? CELL_INCOMPLETE : CELL_OK, ns = 1, nn = 0;
    __syncwarp();
    if (lane == 0) ws.star_[0] = blk.v2t[cell];
    __syncwarp();
    for (int head = 0; head < ns && status == CELL_OK; head++) {
      const int t = ws.star_[head];
      const int4 v = blk.tets[2 * (size_t)t], nb = blk.tets[2 * (size_t)t + 1];
      const float4 c = blk.cc[t];
      cmin[0] = fminf(cmin[0], c.x); cmin[1] = fminf(cmin[1], c.y); cmin[2] = fminf(cmin[2], c.z);
      cmax[0] = fmaxf(cmax[0], c.x); cmax[1] = fmaxf(cmax[1], c.y); cmax[2] = fmaxf(cmax[2], c.z);
      if (box_outside_data(cmin, cmax, g)) { status = CELL_OUTSIDE; break; }   // complete and already outside: see k_cell_bfs
      const int vv[4] = {v.x, v.y, v.z, v.w}, bb[4] = {nb.x, nb.y, nb.z, nb.w};
#pragma unroll
      for (int i = 0; i < 4; i++) {
        if (status != CELL_OK) break;
        const int u = vv[i];
        if (u == cell) continue;
        if (!warp_contains(ws.nu_, nn, u)) {
          if (nn >= BIG_NBR_CAP) { status = CELL_BAD_MESH; break; }   // documented limit: > 1024 faces on one cell
          if (nn == nbr_room) {
            int *gu = gbase + BIG_STAR_CAP, *gt = gbase + BIG_STAR_CAP + BIG_NBR_CAP;
            for (int j = lane; j < nn; j += 32) { gu[j] = ws.nu_[j]; gt[j] = ws.nt_[j]; }
            ws.nu_ = gu; ws.nt_ = gt; nbr_room = BIG_NBR_CAP;
            __syncwarp();
          }
          if (lane == 0) { ws.nu_[nn] = u; ws.nt_[nn] = t; }
          nn++;
          __syncwarp();
        }
        const int next = bb[i];
        if (next < 0) { status = CELL_INCOMPLETE; break; }
        if (!warp_contains(ws.star_, ns, next)) {
          if (ns >= BIG_STAR_CAP) { status = CELL_BAD_MESH; break; }  // documented limit: > 4096 tets around one site
          if (ns == star_room) {
            for (int j = lane; j < ns; j += 32) gbase[j] = ws.star_[j];
            ws.star_ = gbase; star_room = BIG_STAR_CAP;
            __syncwarp();
          }
          if (lane == 0) ws.star_[ns] = next;
          ns++;
          __syncwarp();
        }
      }
    }
    __syncwarp();
    // lane 0 carries the cell through the common tail; the other lanes pass "no cell"
    bfs_finish(lane == 0 ? status : -1, cell, nn, ws, cmin, cmax, blk, blk_id, g, out);
  }
}

// K3a part 1b: one thread per Voronoi face: walk the tets around the Delaunay edge in the
// reference's order, Newell normal, orientation, plane = (normal, first vertex) -> plane pool.
// Tiny per-thread state and no shared memory: full occupancy hides the dependent gathers.
// `act`: this thread has a face (every lane of the warp calls: the links of a warp's faces differ in length, and the
// lanes meet again -- __syncwarp -- before the uniform rest of the face: normalisation with one sqrt and three divisions)
__device__ __forceinline__ void cell_face(const FaceRef *__restrict__ faces, size_t f, bool act, const DevBlock *__restrict__ blocks,
                                          float *__restrict__ plane_pool, Counters *cnt)
{
  FaceRef r;
  r.site = 0; r.u = -1; r.ut = 0; r.blk = 0;
  if (act) r = faces[f];
  act = act && r.u >= 0 && !(r.blk & FACE_DIRECT);      // padding slot, or a cell that computes its own planes (k_cell_direct)
  FaceAccum fa;
  fa.cmin = nullptr; fa.cmax = nullptr;
  int n = -1;
  float site[3] = {0.0f, 0.0f, 0.0f};
  if (act) {
    const DevBlock &b = blocks[r.blk];
    site[0] = b.particles[3 * (size_t)r.site]; site[1] = b.particles[3 * (size_t)r.site + 1]; site[2] = b.particles[3 * (size_t)r.site + 2];
    if (b.walk) {
      // slots of the site and of u in the first tet; from there on the walk follows slot permutations
      const int4 v0 = b.tets[2 * (size_t)r.ut];
      const int s_c = v0.x == r.site ? 0 : (v0.y == r.site ? 1 : (v0.z == r.site ? 2 : 3));
      const int s_u = v0.x == r.u ? 0 : (v0.y == r.u ? 1 : (v0.z == r.u ? 2 : 3));
      n = walk_edge_link_rec(s_c, s_u, r.ut, b.walk, fa);
    } else {
      n = walk_edge_link(r.site, r.u, r.ut, b.tets, b.cc, fa);      // no walk records (fused path): tet records + circumcenters
    }
  }
  __syncwarp();
  if (!act) return;
  float2 *dst = reinterpret_cast<float2 *>(plane_pool + f * 6);
  if (n < 0) {
    // the link did not close (malformed mesh): a NaN plane is never significant in PtInCell
    const float q = __int_as_float(0x7fc00000);
    dst[0] = make_float2(q, q); dst[1] = make_float2(q, q); dst[2] = make_float2(q, q);
    atomicAdd(&cnt->n_bad, 1ull);
    return;
  }
  newell_term(fa.nrm, fa.prev, fa.v0);
  newell_finish(fa.nrm, fa.v0, site);
  dst[0] = make_float2(fa.nrm[0], fa.nrm[1]);
  dst[1] = make_float2(fa.nrm[2], fa.v0[0]);
  dst[2] = make_float2(fa.v0[1], fa.v0[2]);
}

__global__ void __launch_bounds__(256) k_cell_faces(const FaceRef *__restrict__ faces, size_t f_begin, size_t f_end, const DevBlock *__restrict__ blocks,
                                                     float *__restrict__ plane_pool, Counters *cnt)
{
  const size_t f = f_begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  cell_face(faces, f, f < f_end, blocks, plane_pool, cnt);
}

// The same over a range kept on the device, in pairs of faces: [*range_lo, min(*range_hi, cap)).  Persistent
// CTAs stride over it, so the launch needs no count from the host.
__global__ void __launch_bounds__(256, 6) k_cell_faces_dev(const FaceRef *__restrict__ faces, const DevBlock *__restrict__ blocks, float *__restrict__ plane_pool,
                                                         Counters *cnt, const unsigned int *range_lo, const unsigned int *range_hi, uint32_t range_cap)
{
  const uint32_t hi = *range_hi < range_cap ? *range_hi : range_cap;
  const size_t f_end = (size_t)hi * 2;
  // every lane of a warp makes the same number of trips (cell_face synchronises the warp)
  for (size_t f0 = (size_t)*range_lo * 2 + (size_t)blockIdx.x * blockDim.x; f0 < f_end; f0 += (size_t)gridDim.x * blockDim.x)
    cell_face(faces, f0 + threadIdx.x, f0 + threadIdx.x < f_end, blocks, plane_pool, cnt);
}

// ---- span emission shared by the scan kernels and k_cic --------------------------------------------
struct CountEmit
{
  int n;
  __device__ __forceinline__ void operator()(uint64_t, uint64_t) { n++; }
};
struct StoreEmit
{
  uint64_t *keys, *data;
  unsigned long long pos, cap;
  __device__ __forceinline__ void operator()(uint64_t k, uint64_t d)
  {
    if (pos < cap) { keys[pos] = k; data[pos] = d; }
    pos++;
  }
};

struct ScanCtx
{
  const BlockBox *boxes;
  int nblocks;
  KeyLayout kl;
  int project;
};

// true when the whole index box of a cell is local to block e and inside its sub-grid: then
// every scan line is exactly one record
__device__ __forceinline__ bool box_is_local(const BlockBox &b, const int *lo, const int *n3, int project)
{
  for (int d = 0; d < 3; d++) {
    int hi = lo[d] + n3[d] - 1;
    if (lo[d] < b.p_lo[d] || hi > b.p_hi[d]) return false;
    if (project && d == 2) continue;
    if (lo[d] < b.b_lo[d] || hi >= b.b_lo[d] + b.b_num[d]) return false;
  }
  return true;
}

template <class Emit>
struct LineEmitter
{
  const ScanCtx &sc;
  int e;
  uint32_t cell;
  const int *lo;
  bool local_box;
  float value;
  Emit &emit;
  int tot;
  int cand[8];       // blocks other than e that the cell's index box can reach (boundary cells)
  int ncand;         // -1: not computed / more than 8 -> emit_line scans every block
  __device__ __forceinline__ void find_candidates(const int *n3)
  {
    ncand = local_box ? 0 : candidate_blocks(sc.boxes, sc.nblocks, e, lo, n3, sc.project, cand, 8);
  }
  __device__ __forceinline__ void operator()(int yi, int zi, int min_xi, int max_xi)
  {
    int y = lo[1] + yi, z = lo[2] + zi, xa = lo[0] + min_xi, xb = lo[0] + max_xi;
    if (local_box) {
      const BlockBox &b = sc.boxes[e];
      int ly = y - b.b_lo[1], lz = z - b.b_lo[2];
      uint64_t row = (uint64_t)(b.row_base + (sc.project ? (long long)ly : (long long)lz * b.b_num[1] + ly));
      emit(make_key(sc.kl, row, 0, cell, z_slot(sc.kl, sc.project, z)), make_data(xa - b.b_lo[0], xb - xa + 1, 0, value));
    } else {
      emit_line(sc.boxes, sc.nblocks, e, sc.kl, sc.project, cell, xa, xb, y, z, 0, value, emit, cand, ncand);
    }
  }
};

// the 8 CIC points of a site as single-point lines (src/dense.cpp:1437-1455 and :512-560)
template <class Emit>
__device__ __forceinline__ void emit_cic(const ScanCtx &sc, int e, uint32_t cell, const float *site, const GridGeom &g,
                                         int float_path_local, Emit &emit)
{
  int i0[3];
  float vals[8];
  cic_weights(site, g.mass, g, i0, vals);
  int n = 0;
  for (int dz = 0; dz < 2; dz++)
    for (int dy = 0; dy < 2; dy++)
      for (int dx = 0; dx < 2; dx++, n++)
        emit_line(sc.boxes, sc.nblocks, e, sc.kl, sc.project, cell, i0[0] + dx, i0[0] + dx, i0[1] + dy, i0[2] + dz,
                  float_path_local, vals[n], emit);
}

// ---- K3a part 2: inside bits + scan-line walk -----------------------------------------------------
// Every warp is autonomous (no CTA barrier): it takes 32 consecutive cell headers, one per lane.
//   phase 2  warp per cell, cell after cell: the lanes test 32 points of the cell's index box per step
//            against the cell's planes.  The planes of cell c+1 are in flight (TMA 1-D bulk copy
//            into the other slot of a two-slot ring in the warp's shared-memory slice, completion on
//            that slot's mbarrier) while cell c is tested; reads are shared-memory broadcasts and the
//            face loop trip count is warp-uniform.  One ballot = one word of inside-bits.
//   phase 3  lane per cell: the reference's scan-line state machine on the bits; the non-empty lines
//            are kept so the walk runs once; span records go out at a warp-aggregated offset.
// A batch whose inside-bits exceed the slice is processed in sub-batches.
constexpr int SCAN_WARPS = 8;
constexpr int SCAN_THREADS = SCAN_WARPS * 32;
constexpr int SCAN_SLOT_FLOATS = SCAN_FACE_CAP * 6;   // one cell's planes: 32 faces * 24 B = 768 B
constexpr int SCAN_BIT_WORDS = 256;                   // 8192 inside bits per warp (+ pad words for the funnel shift)
constexpr int SCAN_LINE_CAP = 24;                     // non-empty scan lines per cell kept from pass 1 (else the walk is redone)
constexpr int SCAN_WARP_BYTES = 2 * SCAN_SLOT_FLOATS * 4 + (SCAN_BIT_WORDS + 4) * 4 + SCAN_LINE_CAP * 32 * 4 + 16;
constexpr size_t SCAN_SMEM = (size_t)SCAN_WARPS * SCAN_WARP_BYTES;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct BitsInside
{
  const uint32_t *bits;
  uint32_t off;
  int nx, ny;
  __device__ __forceinline__ bool operator()(int i, int j, int k) const
  {
    uint32_t b = off + (uint32_t)((k * ny + j) * nx + i);
    return (bits[b >> 5] >> (b & 31u)) & 1u;
  }
};

// inside-bits of one scan line (nx <= 32) out of the packed bit array
struct RowBits
{
  const uint32_t *bits;
  uint32_t off;
  int nx, ny;
  __device__ __forceinline__ uint32_t operator()(int j, int k) const
  {
    uint32_t b = off + (uint32_t)((k * ny + j) * nx);
    return __funnelshift_r(bits[b >> 5], bits[(b >> 5) + 1], b & 31u);
  }
};
// keeps the non-empty lines of pass 1: yi | zi << 11 | min_xi << 22 | max_xi << 27, lane-strided
struct LineKeeper
{
  uint32_t *buf;
  int n;
  __device__ __forceinline__ void operator()(int yi, int zi, int min_xi, int max_xi)
  {
    if (n < SCAN_LINE_CAP) buf[n * 32] = (uint32_t)yi | ((uint32_t)zi << 11) | ((uint32_t)min_xi << 22) | ((uint32_t)max_xi << 27);
    n++;
  }
};

// l / n and l % n for 0 <= l < 2^22, 1 <= n < 2^15 through one float multiply and a fix-up
__device__ __forceinline__ void divmod_small(int l, int n, float inv_n, int &q, int &r)
{
  q = (int)((float)l * inv_n);
  r = l - q * n;
  if (r < 0) { q--; r += n; }
  if (r >= n) { q++; r -= n; }
}

template <class T>
__device__ __forceinline__ T warp_incl_scan(T v)
{
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T o = __shfl_up_sync(0xffffffffu, v, d);
    if ((int)lane_id() >= d) v += o;
  }
  return v;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_cell_scan(const CellHdr *__restrict__ hdrs, uint32_t n_hdrs,
                                                             const float *__restrict__ plane_pool, const DevBlock *__restrict__ blocks,
                                                             ScanCtx sc, const __grid_constant__ GridGeom g, SpanOut out, const unsigned int *range_lo,
                                                             const unsigned int *range_hi, uint32_t range_cap)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (range_hi) {
    // header range kept on the device (the grid is an upper bound): [*range_lo, min(*range_hi, cap))
    const uint32_t lo = range_lo ? *range_lo : 0u, hi = *range_hi < range_cap ? *range_hi : range_cap;
    hdrs += lo;
    n_hdrs = hi > lo ? hi - lo : 0u;
  }
  unsigned char *mine = smem_raw + (size_t)warp * SCAN_WARP_BYTES;
  float *planes_w = reinterpret_cast<float *>(mine);                       // [2][SCAN_SLOT_FLOATS]
  uint32_t *bits_w = reinterpret_cast<uint32_t *>(planes_w + 2 * SCAN_SLOT_FLOATS);
  uint32_t *lines_w = bits_w + SCAN_BIT_WORDS + 4;                          // [SCAN_LINE_CAP][32 lanes]
  uint64_t *bar = reinterpret_cast<uint64_t *>(lines_w + SCAN_LINE_CAP * 32); // [2]

  const uint32_t first = (blockIdx.x * SCAN_WARPS + warp) * 32u;
  if (first >= n_hdrs) return;                       // whole warp leaves together
  const int ncell = (int)(n_hdrs - first < 32u ? n_hdrs - first : 32u);

  if (lane == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  // lane c owns cell c of the batch
  CellHdr h;
  h.cell = 0; h.blk_nf = 0; h.lo[0] = h.lo[1] = h.lo[2] = 0; h.n3[0] = h.n3[1] = h.n3[2] = 1; h.pad = 0; h.plane_off = 0;
  if (lane < ncell) h = hdrs[first + lane];
  const int nf = (int)(h.blk_nf & 0xffffu);
  const int e = (int)(h.blk_nf >> 16);
  const int npts = lane < ncell ? (int)h.n3[0] * (int)h.n3[1] * (int)h.n3[2] : 0;
  const uint32_t my_bytes = lane < ncell ? (uint32_t)((nf + 1) >> 1) * 48u : 0u;   // planes come in 48-byte units
  const int pts32 = (npts + 31) & ~31;

  // the plane ring: cell c uses slot c & 1; its copy is issued by lane c (which holds the header)
  uint32_t phase_bits = 0;                           // bit s = parity to wait for on slot s
  if (lane == 0 && my_bytes) { mbar_arrive_expect_tx(&bar[0], my_bytes); bulk_g2s(planes_w, plane_pool + (size_t)h.plane_off * 12, my_bytes, &bar[0]); }

  int c0 = 0;
  while (c0 < ncell) {
    // sub-batch [c0, c1): the longest run whose inside-bits fit the warp's slice
    const int incl_p = warp_incl_scan(lane >= c0 ? pts32 : 0);
    const bool fits = lane >= c0 && lane < ncell && incl_p <= SCAN_BIT_WORDS * 32;
    const unsigned fm = __ballot_sync(0xffffffffu, fits);
    const unsigned run = fm >> c0;                   // a run of ones from bit 0 (prefix sums are monotone)
    int c1 = c0 + (run == 0xffffffffu ? 32 : __ffs(~run) - 1);
    if (c1 > ncell) c1 = ncell;
    const bool in_sub = lane >= c0 && lane < c1;
    const int p_off = incl_p - pts32;                // bits before this cell in the slice

    // phase 2: warp per cell
    for (int c = c0; c < c1; c++) {
      const int slot = c & 1;
      // prefetch the planes of cell c+1 into the other slot (its previous user, cell c-1, is done)
      if (c + 1 < ncell) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // earlier generic reads of that slot vs. the bulk write
        if (lane == c + 1 && my_bytes) {
          mbar_arrive_expect_tx(&bar[slot ^ 1], my_bytes);
          bulk_g2s(planes_w + (slot ^ 1) * SCAN_SLOT_FLOATS, plane_pool + (size_t)h.plane_off * 12, my_bytes, &bar[slot ^ 1]);
        }
      }
      const int nx = __shfl_sync(0xffffffffu, (int)h.n3[0], c), ny = __shfl_sync(0xffffffffu, (int)h.n3[1], c);
      const int np = __shfl_sync(0xffffffffu, npts, c), cnf = __shfl_sync(0xffffffffu, nf, c);
      const int lx = __shfl_sync(0xffffffffu, h.lo[0], c), ly = __shfl_sync(0xffffffffu, h.lo[1], c), lz = __shfl_sync(0xffffffffu, h.lo[2], c);
      const float *pl = planes_w + slot * SCAN_SLOT_FLOATS;
      const int wbase = __shfl_sync(0xffffffffu, p_off, c) >> 5;
      if (cnf) {
        mbar_wait(&bar[slot], (phase_bits >> slot) & 1u);
        phase_bits ^= 1u << slot;
      }
      // probe position: cell_min_grid_pos + i * step with cell_min_grid_pos = idx2phys(lo)  (src/dense.cpp:1404,1530-1532)
      const float bx = idx2phys1(lx, g.step[0], g.gmin[0]), by = idx2phys1(ly, g.step[1], g.gmin[1]), bz = idx2phys1(lz, g.step[2], g.gmin[2]);
      const float inv_nx = __fdividef(1.0f, (float)nx), inv_ny = __fdividef(1.0f, (float)ny);   // approximate is enough: divmod_small fixes +-1
      for (int l0 = 0; l0 < np; l0 += 32) {
        const int l = l0 + lane;
        const bool valid = l < np;
        int i, r, j, k;
        divmod_small(valid ? l : 0, nx, inv_nx, r, i);
        divmod_small(r, ny, inv_ny, k, j);
        const float pt[3] = {fadd(bx, fmul((float)i, g.step[0])), fadd(by, fmul((float)j, g.step[1])), fadd(bz, fmul((float)k, g.step[2]))};
        // PtInCell (src/dense.cpp:1172-1203): false iff some face has dist > eps and some face has
        // dist < -eps.  Tracked as running max / min of dist (NaN never counts, as in the reference).
        float dmax = -INFINITY, dmin = INFINITY;
        const float neg_eps = -g.eps;
        auto plane_test = [&](int f) {
          const float2 *p = reinterpret_cast<const float2 *>(pl + 6 * f);
          const float2 a = p[0], b = p[1], cc2 = p[2];
          const float dist = fadd(fadd(fmul(a.x, fsub(pt[0], b.y)), fmul(a.y, fsub(pt[1], cc2.x))), fmul(b.x, fsub(pt[2], cc2.y)));
          dmax = fmaxf(dmax, dist);
          dmin = fminf(dmin, dist);
        };
        // groups of four faces without per-face loop tests; after each group the warp leaves early
        // when every point is already outside (a step with no inside point: most steps past the first)
        int f = 0;
        const int cnf4 = cnf & ~3;
        while (f < cnf4) {
          plane_test(f); plane_test(f + 1); plane_test(f + 2); plane_test(f + 3);
          f += 4;
          if (__all_sync(0xffffffffu, !valid || (dmax > g.eps && dmin < neg_eps))) f = 1 << 20;
        }
        for (; f < cnf; f++) plane_test(f);
        const unsigned w = __ballot_sync(0xffffffffu, valid && !(dmax > g.eps && dmin < neg_eps));
        if (lane == 0) bits_w[wbase + (l0 >> 5)] = w;
      }
      __syncwarp();
    }

    // phase 3: lane per cell.  Index boxes at most 32 wide (all but exotic cells) take the
    // bit-parallel walk and keep their non-empty lines, so the walk runs once.
    {
      int tot = 0, nrec = 0, nlines = 0;
      bool local_box = false, bitpath = false;
      float site[3] = {0, 0, 0};
      const int nx = (int)h.n3[0], ny = (int)h.n3[1], nz = (int)h.n3[2];
      BitsInside inside{bits_w, (uint32_t)p_off, nx, ny};
      RowBits row{bits_w, (uint32_t)p_off, nx, ny};
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
      unsigned long long base = warp_alloc<unsigned long long>(&out.cnt->n_spans, (unsigned long long)nrec);
      warp_count(&out.cnt->n_deposit, in_sub);
      warp_count(&out.cnt->n_cic_fallback, in_sub && tot == 0);
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
    c0 = c1;
  }
}

// cells with a large index box or more than 32 faces: one CTA per cell, planes read from global
// (all lanes read the same address), inside bits in a global scratch, thread 0 walks the bits
struct GlobalPlanesInsideBits
{
  const uint32_t *bits;
  int nx, ny;
  __device__ __forceinline__ bool operator()(int i, int j, int k) const
  {
    size_t b = ((size_t)k * ny + j) * nx + i;
    return (bits[b >> 5] >> (b & 31u)) & 1u;
  }
};

__global__ void __launch_bounds__(128) k_cell_scan_big(const CellHdr *__restrict__ hdrs, const unsigned long long *__restrict__ bit_off,
                                                        int n_cells, const float *__restrict__ plane_pool, uint32_t *bits_g, unsigned long long bit_base,
                                                        const DevBlock *__restrict__ blocks, ScanCtx sc,
                                                        const __grid_constant__ GridGeom g, SpanOut out)
{
  int ci = blockIdx.x;
  if (ci >= n_cells) return;
  CellHdr h = hdrs[ci];
  int nf = (int)(h.blk_nf & 0xffffu);
  const float *pl = plane_pool + (size_t)h.plane_off * 12;
  uint32_t *bits = bits_g + ((bit_off[ci] - bit_base) >> 5);
  int nx = h.n3[0], ny = h.n3[1], nz = h.n3[2];
  long long total = (long long)nx * ny * nz;
  long long total32 = (total + 31) & ~31LL;
  float base[3] = {idx2phys1(h.lo[0], g.step[0], g.gmin[0]), idx2phys1(h.lo[1], g.step[1], g.gmin[1]), idx2phys1(h.lo[2], g.step[2], g.gmin[2])};
  for (long long s = threadIdx.x; s < total32; s += blockDim.x) {
    bool in = false;
    if (s < total) {
      int i = (int)(s % nx);
      long long r = s / nx;
      int j = (int)(r % ny);
      int k = (int)(r / ny);
      float pt[3] = {fadd(base[0], fmul((float)i, g.step[0])), fadd(base[1], fmul((float)j, g.step[1])), fadd(base[2], fmul((float)k, g.step[2]))};
      bool pos = false, neg = false;
      for (int f = 0; f < nf && !(pos && neg); f++) {
        float n[3] = {__ldg(pl + 6 * f), __ldg(pl + 6 * f + 1), __ldg(pl + 6 * f + 2)};
        float v[3] = {__ldg(pl + 6 * f + 3), __ldg(pl + 6 * f + 4), __ldg(pl + 6 * f + 5)};
        int sd = plane_side(n, v, pt, g.eps);
        pos |= sd > 0;
        neg |= sd < 0;
      }
      in = !(pos && neg);
    }
    unsigned w = __ballot_sync(0xffffffffu, in);
    if (lane_id() == 0) bits[s >> 5] = w;
  }
  __threadfence_block();
  __syncthreads();
  if (threadIdx.x == 0) {
    int e = (int)(h.blk_nf >> 16);
    int lo[3] = {h.lo[0], h.lo[1], h.lo[2]}, n3[3] = {nx, ny, nz};
    bool local_box = box_is_local(sc.boxes[e], lo, n3, sc.project);
    GlobalPlanesInsideBits inside{bits, nx, ny};
    CountEmit ce{0};
    LineEmitter<CountEmit> le{sc, e, h.cell, h.lo, local_box, 0.0f, ce, 0, {0, 0, 0, 0, 0, 0, 0, 0}, -1};
    le.find_candidates(n3);
    int tot = scan_cell(nx, ny, nz, inside, le);
    int nrec = ce.n;
    float site[3] = {0, 0, 0};
    if (tot == 0) {
      const DevBlock &b = blocks[e];
      uint32_t lc = h.cell - b.cell_base;
      site[0] = b.particles[3 * (size_t)lc]; site[1] = b.particles[3 * (size_t)lc + 1]; site[2] = b.particles[3 * (size_t)lc + 2];
      CountEmit ce2{0};
      emit_cic(sc, e, h.cell, site, g, 0, ce2);
      nrec = ce2.n;
      atomicAdd(&out.cnt->n_cic_fallback, 1ull);
    }
    atomicAdd(&out.cnt->n_deposit, 1ull);
    unsigned long long pos0 = atomicAdd(&out.cnt->n_spans, (unsigned long long)nrec);
    StoreEmit se{out.keys, out.data, pos0, out.capacity};
    if (tot > 0) {
      float m = fdiv(g.mass, (float)tot);
      LineEmitter<StoreEmit> le2{sc, e, h.cell, h.lo, local_box, m, se, 0, {0, 0, 0, 0, 0, 0, 0, 0}, -1};
      le2.find_candidates(n3);
      scan_cell(nx, ny, nz, inside, le2);
    } else {
      emit_cic(sc, e, h.cell, site, g, 0, se);
    }
  }
}

// ---- K4: cloud-in-cell, one thread per original particle, warp-aggregated record allocation ----
__global__ void __launch_bounds__(256) k_cic(DevBlock blk, int blk_id, ScanCtx sc, const __grid_constant__ GridGeom g, SpanOut out)
{
  int cell = blockIdx.x * blockDim.x + threadIdx.x;
  bool act = cell < blk.num_orig;
  float site[3] = {0, 0, 0};
  int nrec = 0;
  uint32_t gc = blk.cell_base + (uint32_t)cell;
  if (act) {
    site[0] = blk.particles[3 * (size_t)cell]; site[1] = blk.particles[3 * (size_t)cell + 1]; site[2] = blk.particles[3 * (size_t)cell + 2];
    CountEmit ce{0};
    emit_cic(sc, blk_id, gc, site, g, 1, ce);
    nrec = ce.n;
  }
  unsigned long long base = warp_alloc<unsigned long long>(&out.cnt->n_spans, (unsigned long long)nrec);
  warp_count(&out.cnt->n_deposit, act);
  if (act) {
    StoreEmit se{out.keys, out.data, base, out.capacity};
    emit_cic(sc, blk_id, gc, site, g, 1, se);
  }
}


// ---- K4 without records for the local deposits ---------------------------------------------------------------
// IterateCellsCic (src/dense.cpp:486-562) adds, at every grid point, first the block's OWN particles in particle order
// (float path, :539) and then the points received from other blocks by source gid (recvd_pts, double path).  The first
// part needs no deposit records at all: k_cic_prepare computes every particle's eight weights once and its base cell
// (the truncated index of DistributeScalarCIC, :1787-1872); a stable radix sort of the particle ids by base cell keeps
// the particle order inside a cell; k_cic_gather, one thread per grid point of a block's sub-grid, merges the (at most
// eight) base cells that touch its point by particle id and adds their weights in that order -- the reference's float
// sum, every grid point written once, no atomics.  Only the deposits that leave their block (block and rank
// boundaries) still travel as records and are added on top in the reference's order (sort + k_rows in its sparse form).
struct CicBlock
{
  int o[3];                  // origin of the block's base-cell box: b_lo - 1
  int d[3];                  // its extent: b_num + 1
  unsigned long long cell0;  // first base cell of this block in the rank-wide numbering
  unsigned long long part0;  // first particle of this block in the rank-wide arrays (vals, keys)
};

// drops the records of the emitting block's own points (key bit `remote` == 0): the gather adds those
template <class Inner>
struct RemoteOnlyEmit
{
  Inner &in;
  KeyLayout kl;
  __device__ __forceinline__ void operator()(uint64_t k, uint64_t d)
  {
    if ((k >> (kl.z_bits + kl.cell_bits)) & 1ull) in(k, d);
  }
};

__global__ void __launch_bounds__(256) k_cic_prepare(DevBlock blk, int blk_id, CicBlock cb, ScanCtx sc, const __grid_constant__ GridGeom g, SpanOut out,
                                                      float *__restrict__ vals_out, uint32_t *__restrict__ keys, uint32_t *__restrict__ ids,
                                                      unsigned int *__restrict__ cell_count)
{
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = cell < blk.num_orig;
  float site[3] = {0, 0, 0};
  int nrec = 0;
  const uint32_t gc = blk.cell_base + (uint32_t)cell;
  if (act) {
    site[0] = blk.particles[3 * (size_t)cell]; site[1] = blk.particles[3 * (size_t)cell + 1]; site[2] = blk.particles[3 * (size_t)cell + 2];
    int i0[3];
    float vals[8];
    cic_weights(site, g.mass, g, i0, vals);
    // what the accumulate step adds is m / div (src/dense.cpp:539): one division per weight here instead of one per merge step
    float4 *vo = reinterpret_cast<float4 *>(vals_out + 8 * (cb.part0 + (size_t)cell));
    vo[0] = make_float4(fdiv(vals[0], g.div), fdiv(vals[1], g.div), fdiv(vals[2], g.div), fdiv(vals[3], g.div));
    vo[1] = make_float4(fdiv(vals[4], g.div), fdiv(vals[5], g.div), fdiv(vals[6], g.div), fdiv(vals[7], g.div));
    // base cell inside the block's box of base cells, or behind every cell (a particle none of whose corners is a point of
    // the block's own sub-grid)
    const long long bx = (long long)i0[0] - cb.o[0], by = (long long)i0[1] - cb.o[1], bz = (long long)i0[2] - cb.o[2];
    uint32_t key = 0xffffffffu;                      // (the host keeps the number of base cells below 2^32 - 1)
    if (bx >= 0 && bx < cb.d[0] && by >= 0 && by < cb.d[1] && bz >= 0 && bz < cb.d[2]) {
      key = (uint32_t)(cb.cell0 + (unsigned long long)((bz * cb.d[1] + by) * cb.d[0] + bx));
      atomicAdd(&cell_count[key], 1u);
    }
    keys[cb.part0 + (size_t)cell] = key;
    ids[cb.part0 + (size_t)cell] = (uint32_t)(cb.part0 + (unsigned long long)cell);       // rank-wide: the position of the particle's weights
    // the whole window inside the block's own points and sub-grid (every interior particle): no record, no second look
    const BlockBox &bb = sc.boxes[blk_id];
    bool interior = true;
#pragma unroll
    for (int d = 0; d < 3; d++) {
      const int lo = bb.p_lo[d] > bb.b_lo[d] ? bb.p_lo[d] : bb.b_lo[d];
      const int hi = bb.p_hi[d] < bb.b_lo[d] + bb.b_num[d] - 1 ? bb.p_hi[d] : bb.b_lo[d] + bb.b_num[d] - 1;
      interior = interior && i0[d] >= lo && i0[d] + 1 <= hi;
    }
    if (!interior) {
      CountEmit ce{0};
      RemoteOnlyEmit<CountEmit> re{ce, sc.kl};
      emit_cic(sc, blk_id, gc, site, g, 1, re);
      nrec = ce.n;
    }
  }
  unsigned long long base = warp_alloc<unsigned long long>(&out.cnt->n_spans, (unsigned long long)nrec);
  // every particle of the block deposits: one add per launch instead of one per warp on the same address
  if (blockIdx.x == 0 && threadIdx.x == 0 && blk.num_orig > 0) atomicAdd(&out.cnt->n_deposit, (unsigned long long)blk.num_orig);
  if (act && nrec) {
    StoreEmit se{out.keys, out.data, base, out.capacity};
    RemoteOnlyEmit<StoreEmit> re{se, sc.kl};
    emit_cic(sc, blk_id, gc, site, g, 1, re);
  }
}

// the weights in sorted order: position q of the sorted ids gets the eight quotients of particle sorted_ids[q] (one
// 32-byte sector in, one out).  The gather then walks a cell's list front to back, and the eight grid points around a
// cell -- lanes of one warp, mostly -- read the same sectors: by particle id they were 8 P random sector reads
// (config 3's input: 4.3 GB from DRAM per step for 0.54 GB of weights, the whole of the gather's time; ncu r02).
__global__ void __launch_bounds__(256) k_cic_permute(const uint32_t *__restrict__ sorted_ids, unsigned long long n, const float4 *__restrict__ vals,
                                                      float4 *__restrict__ vals_sorted)
{
  const unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const size_t id = sorted_ids[q];
  const float4 a = vals[2 * id], b = vals[2 * id + 1];
  vals_sorted[2 * q] = a;
  vals_sorted[2 * q + 1] = b;
}

// k_cic_gather, one thread per grid point of the block's sub-grid.  `cell_start` is the exclusive scan of the particles per
// base cell over cic_cells + 1 entries (the list of cell c ends where that of c + 1 starts); `vals` holds the quotients
// m / div (k_cic_prepare) in sorted order (k_cic_permute), so the merge loop is select -> load -> add.
// Launch shape: a CTA of 4 warps covers 32 x 2 x 2 grid points, a warp a brick of 8 x 2 x 2 (x fastest): the lanes of a
// warp wait for the longest merge among them, and the particle density of a compact brick varies less than that of 32
// points in a row (clustered config 3: 2.4 instead of 3.1 times the work of a perfectly even split); the lanes of a brick
// also share most of their base cells; a warp still writes whole 32-byte sectors.  3-D launch grid: no index divisions.
// (Tried and dropped, both slower on config 3's input: the merge running eight steps ahead of the adds -- the padding of
// the last group costs the many short lists more than the long ones gain; a first pass that compacts the grid points with
// something to add into a work list -- the pass costs more than the idle lanes it removes.)
constexpr int CIC_TILE_X = 32, CIC_TILE_Y = 2, CIC_TILE_Z = 2, CIC_THREADS = 128;
// position and end of the eight lists of every thread of the CTA: column threadIdx.x (no bank conflicts), row = list
struct CicSmem
{
  unsigned int pos[8][CIC_THREADS], end[8][CIC_THREADS];
};
struct CicListsShared
{
  unsigned int *pos_, *end_;                         // this thread's column
  uint32_t head[8];
  __device__ __forceinline__ unsigned int pos(int n) const { return pos_[n * CIC_THREADS]; }
  __device__ __forceinline__ unsigned int end(int n) const { return end_[n * CIC_THREADS]; }
  __device__ __forceinline__ void set(int n, unsigned int p, unsigned int e) { pos_[n * CIC_THREADS] = p; end_[n * CIC_THREADS] = e; }
  __device__ __forceinline__ void set_pos(int n, unsigned int p) { pos_[n * CIC_THREADS] = p; }
};
__global__ void __launch_bounds__(CIC_THREADS) k_cic_gather(CicBlock cb, BlockBox bx, const unsigned int *__restrict__ cell_start,
                                                             const uint32_t *__restrict__ sorted_ids, const float *__restrict__ vals, float *__restrict__ out)
{
  __shared__ CicSmem S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lx = (int)blockIdx.x * CIC_TILE_X + warp * 8 + (lane & 7);
  const int ly = (int)blockIdx.y * CIC_TILE_Y + ((lane >> 3) & 1);
  const int lz = (int)blockIdx.z * CIC_TILE_Z + (lane >> 4);
  if (lx >= bx.b_num[0] || ly >= bx.b_num[1] || lz >= bx.b_num[2]) return;
  const long long p = ((long long)lz * bx.b_num[1] + ly) * bx.b_num[0] + lx;
  const int x = bx.b_lo[0] + lx, y = bx.b_lo[1] + ly, z = bx.b_lo[2] + lz;
  float cur = 0.0f;
  // a point is the block's own only where its position lies in the block's closed bounds (src/dense.cpp:523-531)
  if (x >= bx.p_lo[0] && x <= bx.p_hi[0] && y >= bx.p_lo[1] && y <= bx.p_hi[1] && z >= bx.p_lo[2] && z <= bx.p_hi[2]) {
    // the eight base cells whose window holds the point: base = point - (dx, dy, dz); the point is corner n = dz*4 + dy*2 + dx.
    // Box coordinates of a base cell: (lx + 1 - dx, ly + 1 - dy, lz + 1 - dz), always inside the box.  The cells of dx = 1
    // and dx = 0 follow each other: three scan entries bound both lists.
    const long long row = cb.d[0], slab = (long long)cb.d[0] * cb.d[1];
    const unsigned int *cs11 = cell_start + (cb.cell0 + (unsigned long long)((long long)lz * slab + (long long)ly * row + lx));   // dy = dz = 1
    CicListsShared l{&S.pos[0][threadIdx.x], &S.end[0][threadIdx.x], {0, 0, 0, 0, 0, 0, 0, 0}};
    unsigned int total = 0;
    unsigned int sv[4][3];
#pragma unroll
    for (int h = 0; h < 4; h++) {
      const int dy = h & 1, dz = h >> 1;
      const unsigned int *cs = cs11 + (dy ? 0 : row) + (dz ? 0 : slab);
      sv[h][0] = cs[0]; sv[h][1] = cs[1]; sv[h][2] = cs[2];
      total += sv[h][2] - sv[h][0];
    }
    if (total) {
#pragma unroll
      for (int h = 0; h < 4; h++) {
        const unsigned int s0 = sv[h][0], s1 = sv[h][1], s2 = sv[h][2];
        const int n1 = h * 2 + 1, n0 = h * 2;          // h = dz * 2 + dy
        l.set(n1, s0, s1);
        l.set(n0, s1, s2);
        // (an empty list reads entry 0, which exists: the array holds every particle of the rank)
        const uint32_t f1 = sorted_ids[s0 < s1 ? s0 : 0u], f0 = sorted_ids[s1 < s2 ? s1 : 0u];
        l.head[n1] = s0 < s1 ? f1 : 0xffffffffu;
        l.head[n0] = s1 < s2 ? f0 : 0xffffffffu;
      }
      cur = cic_merge_sum(l, total, sorted_ids, vals);
    }
  }
  out[p] = cur;
}

// after the cell kernels of a group: what was appended so far is done
__global__ void k_advance(Counters *cnt, uint32_t cap_small)
{
  cnt->pairs_done = cnt->plane_cursor;
  cnt->small_done = cnt->n_small < cap_small ? cnt->n_small : cap_small;
  cnt->ovf_done = cnt->n_overflow;
  for (int c = 0; c < 3; c++) cnt->dir_done[c] = cnt->n_dir[c].v;
}

// ---- K3b: deposit.  Spans sorted by (row, remote, cell, z); one warp owns one row -----------------
// row_start[r] = first sorted record of row r (r in [row0, row0 + nrows]); rows without records get
// an empty range
__global__ void k_row_starts(const uint64_t *__restrict__ keys, unsigned long long n, KeyLayout kl, unsigned long long row0,
                             unsigned long long nrows, unsigned long long *__restrict__ row_start)
{
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  // virtual rows: before the first record row0 - 1 + 0 .. ; after the last record row0 + nrows
  long long prev = (i == 0) ? (long long)row0 - 1 : (long long)key_row(kl, keys[i - 1]);
  long long cur = (i == n) ? (long long)(row0 + nrows) : (long long)key_row(kl, keys[i]);
  if (prev < (long long)row0 - 1) prev = (long long)row0 - 1;
  if (cur > (long long)(row0 + nrows)) cur = (long long)(row0 + nrows);
  for (long long r = prev + 1; r <= cur; r++) row_start[r - (long long)row0] = i;
}

struct RowBlock
{
  long long row_base;   // first row id
  long long nrows;
  long long out_off;    // float offset of the block's density array in the output buffer
  int nx;
  int pad;
};

constexpr int ROWS_WARPS = 4;
constexpr int ROWS_SHORT_SPAN = 16;   // spans up to this length are applied one record per lane

__global__ void __launch_bounds__(ROWS_WARPS * 32) k_rows(const uint64_t *__restrict__ data, const unsigned long long *__restrict__ row_start,
                                                          unsigned long long row0, unsigned long long r_begin, unsigned long long nrows, const RowBlock *__restrict__ rblocks,
                                                          int n_rblocks, float div, int nx_max, float *__restrict__ out, int sparse)
{
  // rows [r_begin, r_begin + nrows) of this GPU's row range (row ids row0 + r)
  extern __shared__ float rowbuf_all[];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long r = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + warp;   // 1..ROWS_WARPS warps per CTA (host: what fits shared memory)
  if (r >= nrows) return;
  r += r_begin;
  float *buf = rowbuf_all + (size_t)warp * nx_max;
  long long row = (long long)(row0 + r);
  // owning block: last rblock with row_base <= row
  int lo_b = 0, hi_b = n_rblocks;
  while (hi_b - lo_b > 1) {
    int mid = (lo_b + hi_b) >> 1;
    if (rblocks[mid].row_base <= row) lo_b = mid; else hi_b = mid;
  }
  RowBlock rb = rblocks[lo_b];
  int nx = rb.nx;
  unsigned long long s0 = row_start[r], s1 = row_start[r + 1];
  float *dst = out + rb.out_off + (row - rb.row_base) * (long long)nx;
  // sparse: the row already holds the deposits that are alone on their grid point (k_span_place) and zeros
  // where several meet; rows without a record stay as they are
  if (sparse && s0 == s1) return;
  for (int x = lane; x < nx; x += 32) buf[x] = sparse ? dst[x] : 0.0f;
  __syncwarp();
  for (unsigned long long sb = s0; sb < s1; sb += 32) {
    unsigned long long mine = sb + lane < s1 ? data[sb + lane] : 0ull;
    // the quotient mass / div of the reference's accumulate step is computed once per record by
    // the lane that loaded it (off the serial chain), then broadcast
    const float my_m = u2f((uint32_t)(mine >> 32));
    const int my_fp = (int)((mine >> 31) & 1u);
    const double my_q = my_fp ? (double)fdiv(my_m, div) : (double)my_m / (double)div;
    int cnt = (int)(s1 - sb < 32 ? s1 - sb : 32);
    const unsigned my_lo = (unsigned)mine;
    const int my_x0 = (int)(my_lo & 0xffffu);
    const int my_len = lane < cnt ? (int)((my_lo >> 16) & 0x7fffu) : 0;
    if (__all_sync(0xffffffffu, my_len <= ROWS_SHORT_SPAN)) {
      // short spans (the common case: a cell covers a few points of a row): one record per lane.
      // Only records that share a grid point need the sorted order, so each lane collects the
      // earlier records of the batch its span overlaps and waits for exactly those.
      const int my_x1 = my_x0 + my_len;              // empty spans overlap nothing
      unsigned dep = 0;
#pragma unroll
      for (int d = 1; d < 32; d++) {
        const int o0 = __shfl_up_sync(0xffffffffu, my_x0, d), o1 = __shfl_up_sync(0xffffffffu, my_x1, d);
        if (lane >= d && o0 < my_x1 && my_x0 < o1) dep |= 1u << (lane - d);
      }
      unsigned done = __ballot_sync(0xffffffffu, my_len == 0);
      const int my_fpath = (int)(my_lo >> 31);
      while (done != 0xffffffffu) {
        const bool ready = !((done >> lane) & 1u) && (dep & ~done) == 0u;
        if (ready) {
          for (int x = 0; x < my_len; x++) {
            const int xx = my_x0 + x;
            if (xx < nx) buf[xx] = my_fpath ? fadd(buf[xx], (float)my_q) : (float)((double)buf[xx] + my_q);
          }
        }
        __syncwarp();
        done |= __ballot_sync(0xffffffffu, ready);
      }
      continue;
    }
    for (int j = 0; j < cnt; j++) {
      const unsigned lo32 = __shfl_sync(0xffffffffu, (unsigned)mine, j);
      const double q = __shfl_sync(0xffffffffu, my_q, j);
      const int x0 = (int)(lo32 & 0xffffu);
      const int len = (int)((lo32 >> 16) & 0x7fffu);
      const int fp = (int)(lo32 >> 31);
      for (int x = lane; x < len; x += 32) {
        int xx = x0 + x;
        // double path: (float)((double)cur + q)  (src/dense.cpp:290,193); float path: cur + (m / div)  (:539)
        if (xx < nx) buf[xx] = fp ? fadd(buf[xx], (float)q) : (float)((double)buf[xx] + q);
      }
      __syncwarp();
    }
  }
  for (int x = lane; x < nx; x += 32) dst[x] = buf[x];
}

// ---- K3b without the big sort ------------------------------------------------------------------------
// Only grid points that receive MORE THAN ONE deposit need the reference's accumulation order; a
// point with a single deposit ends up as (float)(0 + m/div) whoever gets there.  So: count the
// deposits per grid point (k_span_count), then write the single ones straight into the grid and
// hand only the deposits on shared points, as one-point records with their original keys, to the
// sorted path (k_span_place -> radix sort -> k_rows in its `sparse` form, which starts from the
// row as it stands instead of zeros and skips rows without records).  3-D runs only: a projection
// stacks every z onto the same point, there everything is shared and the full sort is the shorter way.
__device__ __forceinline__ bool span_target(uint64_t key, uint64_t data, KeyLayout kl, unsigned long long row0, unsigned long long nrows,
                                            const RowBlock *__restrict__ rblocks, int n_rblocks, long long &base, int &x0, int &x1)
{
  const unsigned long long row = key_row(kl, key);
  if (row < row0 || row >= row0 + nrows) return false;        // left for another rank (sentinel key)
  int lo_b = 0, hi_b = n_rblocks;
  while (hi_b - lo_b > 1) {
    const int mid = (lo_b + hi_b) >> 1;
    if (rblocks[mid].row_base <= (long long)row) lo_b = mid; else hi_b = mid;
  }
  const RowBlock rb = rblocks[lo_b];
  const unsigned lo32 = (unsigned)data;
  x0 = (int)(lo32 & 0xffffu);
  x1 = x0 + (int)((lo32 >> 16) & 0x7fffu);
  if (x1 > rb.nx) x1 = rb.nx;
  base = rb.out_off + ((long long)row - rb.row_base) * (long long)rb.nx;
  return x1 > x0;
}

__global__ void k_span_count(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ data, unsigned long long n, KeyLayout kl,
                             unsigned long long row0, unsigned long long nrows, const RowBlock *__restrict__ rblocks, int n_rblocks,
                             unsigned int *__restrict__ count)
{
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long base;
  int x0, x1;
  if (!span_target(keys[i], data[i], kl, row0, nrows, rblocks, n_rblocks, base, x0, x1)) return;
  for (int x = x0; x < x1; x++) atomicAdd(&count[base + x], 1u);
}

constexpr int SPAN_PLACE_THREADS = 256;
__global__ void __launch_bounds__(SPAN_PLACE_THREADS) k_span_place(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ data, unsigned long long n, KeyLayout kl,
                             unsigned long long row0, unsigned long long nrows, const RowBlock *__restrict__ rblocks, int n_rblocks,
                             const unsigned int *__restrict__ count, float div, float *__restrict__ out, uint64_t *__restrict__ shared_keys,
                             uint64_t *__restrict__ shared_data, unsigned long long shared_cap, unsigned long long *n_shared)
{
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long base = 0;
  int x0 = 0, x1 = 0;
  uint64_t key = 0, d = 0;
  bool live = false;
  if (i < n) {
    key = keys[i];
    d = data[i];
    live = span_target(key, d, kl, row0, nrows, rblocks, n_rblocks, base, x0, x1);
  }
  int mine = 0;
  if (live) {
    // double path: (float)((double)0 + (double)m / (double)div)  (src/dense.cpp:290,193); float path: 0 + m / div  (:539)
    const float m = u2f((uint32_t)(d >> 32));
    const float v = ((d >> 31) & 1u) ? fadd(0.0f, fdiv(m, div)) : (float)((double)0.0f + (double)m / (double)div);
    for (int x = x0; x < x1; x++) {
      if (count[base + x] == 1u) out[base + x] = v;
      else mine++;
    }
  }
  // deposits on shared points: one-point records, same key (row, remote, cell, z) as the span they come from
  __shared__ unsigned long long cta_scratch[SPAN_PLACE_THREADS / 32 + 1];
  const unsigned long long at = cta_alloc<unsigned long long, SPAN_PLACE_THREADS / 32>(n_shared, (unsigned long long)mine, cta_scratch);
  if (mine) {
    unsigned long long pos = at;
    for (int x = x0; x < x1; x++) {
      if (count[base + x] == 1u) continue;
      if (pos < shared_cap) {
        shared_keys[pos] = key;
        shared_data[pos] = (d & 0xffffffff80000000ull) | (1ull << 16) | (uint64_t)(unsigned)x;
      }
      pos++;
    }
  }
}


// ---- K3b, shared grid points without a global sort ------------------------------------------------------
// Grid points that receive more than one deposit need the reference's accumulation order: the owner block's own cells in
// cell order, then received points by source gid (cell numbers ascend with the gid; src/dense.cpp:286-296,187-199).
// Instead of sorting every shared deposit by (row, remote, cell), each shared POINT gets a segment: offsets = exclusive
// scan of the per-point counts (k_span_count), k_span_place2 drops each shared deposit into its point's segment as one
// 8-byte record  (remote << 31 | cell) << 32 | value  -- so ascending records ARE the reference's order -- and
// k_point_apply sorts the few records of a point in registers and accumulates them.  Points with many deposits (the
// centres of clumps: hundreds) go to k_point_apply_big, one warp per point, bitonic sort in shared memory.
struct SharedCount
{
  __host__ __device__ __forceinline__ unsigned int operator()(const unsigned int &c) const { return c > 1u ? c : 0u; }
};

constexpr int POINT_SMALL = 8;        // records sorted in registers by one thread
constexpr int POINT_WARP_CAP = 2048;  // records sorted in shared memory by one warp
constexpr int POINT_BIG_WARPS = 2;
constexpr size_t POINT_BIG_SMEM = (size_t)POINT_BIG_WARPS * POINT_WARP_CAP * 8;

__global__ void k_span_place2(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ data, unsigned long long n, KeyLayout kl,
                              unsigned long long row0, unsigned long long nrows, const RowBlock *__restrict__ rblocks, int n_rblocks,
                              const unsigned int *__restrict__ count, const unsigned int *__restrict__ seg_off, unsigned int *__restrict__ fill, float div,
                              float *__restrict__ out, uint64_t *__restrict__ seg, unsigned long long seg_cap, Counters *cnt)
{
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t key = 0, d = 0;
  long long base = 0;
  int x0 = 0, x1 = 0;
  bool live = false;
  if (i < n) {
    key = keys[i];
    d = data[i];
    live = span_target(key, d, kl, row0, nrows, rblocks, n_rblocks, base, x0, x1);
  }
  unsigned long long mine = 0;
  if (live) {
    // double path: (float)((double)0 + (double)m / (double)div)  (src/dense.cpp:290,193); float path: 0 + m / div  (:539)
    const float m = u2f((uint32_t)(d >> 32));
    const float v = ((d >> 31) & 1u) ? fadd(0.0f, fdiv(m, div)) : (float)((double)0.0f + (double)m / (double)div);
    const uint32_t cell = (uint32_t)(key >> kl.z_bits) & (kl.cell_bits >= 32 ? 0xffffffffu : ((1u << kl.cell_bits) - 1u));
    const uint32_t remote = (uint32_t)(key >> (kl.z_bits + kl.cell_bits)) & 1u;
    const uint64_t rec = ((uint64_t)((remote << 31) | cell) << 32) | (uint64_t)f2u(m);
    for (int x = x0; x < x1; x++) {
      const long long gp = base + x;
      if (count[gp] == 1u) out[gp] = v;
      else {
        const unsigned long long p = (unsigned long long)seg_off[gp] + atomicAdd(&fill[gp], 1u);
        if (p < seg_cap) seg[p] = rec;
        else cnt->dep_flags = 1u;
        mine++;
      }
    }
  }
  // deposits that met another one on their grid point (statistics)
  const unsigned long long tot = warp_incl_scan_ull(mine);
  if (lane_id() == 31 && tot) atomicAdd(&cnt->n_shared, tot);
}

__device__ __forceinline__ void cswap64(uint64_t &a, uint64_t &b)
{
  const uint64_t lo = a < b ? a : b, hi = a < b ? b : a;
  a = lo; b = hi;
}

// one thread per grid point of this rank's output
__global__ void __launch_bounds__(256) k_point_apply(const unsigned int *__restrict__ count, const unsigned int *__restrict__ seg_off,
                                                     const uint64_t *__restrict__ seg, unsigned long long seg_cap, unsigned long long npoints, float div,
                                                     int alg_cic, float *__restrict__ out, unsigned int *__restrict__ big_list, unsigned int big_cap, Counters *cnt)
{
  const unsigned long long gp = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned int n = gp < npoints ? count[gp] : 0u;
  const bool big = n > (unsigned)POINT_SMALL;
  const unsigned int slot = warp_append<unsigned int>(&cnt->n_big_points, big);
  if (big && slot < big_cap) big_list[slot] = (unsigned int)gp;
  if (n < 2u || big) return;
  const unsigned long long off = seg_off[gp];
  if (off + n > seg_cap) return;                       // flagged by k_span_place2: the run is redone through the sorted path
  uint64_t r[POINT_SMALL];
#pragma unroll
  for (int i = 0; i < POINT_SMALL; i++) r[i] = (unsigned)i < n ? seg[off + i] : ~0ull;
  // Batcher's odd-even merge sort for 8 keys (19 compare-exchanges)
  cswap64(r[0], r[1]); cswap64(r[2], r[3]); cswap64(r[4], r[5]); cswap64(r[6], r[7]);
  cswap64(r[0], r[2]); cswap64(r[1], r[3]); cswap64(r[4], r[6]); cswap64(r[5], r[7]);
  cswap64(r[1], r[2]); cswap64(r[5], r[6]);
  cswap64(r[0], r[4]); cswap64(r[1], r[5]); cswap64(r[2], r[6]); cswap64(r[3], r[7]);
  cswap64(r[2], r[4]); cswap64(r[3], r[5]);
  cswap64(r[1], r[2]); cswap64(r[3], r[4]); cswap64(r[5], r[6]);
  float cur = 0.0f;
#pragma unroll
  for (int i = 0; i < POINT_SMALL; i++)
    if ((unsigned)i < n) {
      const int remote = (int)(r[i] >> 63);
      cur = accumulate(cur, u2f((uint32_t)r[i]), div, alg_cic && !remote);
    }
  out[gp] = cur;
}

// one warp per grid point with many deposits (persistent warps over the list k_point_apply wrote)
__global__ void __launch_bounds__(POINT_BIG_WARPS * 32) k_point_apply_big(const unsigned int *__restrict__ count, const unsigned int *__restrict__ seg_off,
                                                                          const uint64_t *__restrict__ seg, unsigned long long seg_cap, float div, int alg_cic,
                                                                          float *__restrict__ out, const unsigned int *__restrict__ big_list, unsigned int big_cap,
                                                                          const Counters *cnt)
{
  extern __shared__ __align__(16) uint64_t pbig_s[];
  const int lane = (int)lane_id(), warp = threadIdx.x >> 5;
  uint64_t *s = pbig_s + (size_t)warp * POINT_WARP_CAP;
  const unsigned int n_list = cnt->n_big_points < big_cap ? cnt->n_big_points : big_cap;
  const unsigned int n_warps = gridDim.x * POINT_BIG_WARPS;
  for (unsigned int li = blockIdx.x * POINT_BIG_WARPS + warp; li < n_list; li += n_warps) {
    const unsigned int gp = big_list[li];
    const int n = (int)count[gp];
    const unsigned long long off = seg_off[gp];
    if (off + (unsigned long long)n > seg_cap) continue;
    float cur = 0.0f;
    if (n <= POINT_WARP_CAP) {
      int P = 16;
      while (P < n) P <<= 1;
      for (int i = lane; i < P; i += 32) s[i] = i < n ? seg[off + i] : ~0ull;
      __syncwarp();
      for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int i = lane; i < P; i += 32) {
            const int ixj = i ^ j;
            if (ixj > i) {
              const uint64_t a = s[i], b = s[ixj];
              const bool asc = (i & k) == 0;
              if ((a > b) == asc) { s[i] = b; s[ixj] = a; }
            }
          }
          __syncwarp();
        }
      // the quotients m / div off the serial chain (every lane its records), then one lane adds in order
      int n_local = 0;
      for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        bool loc = false;
        if (i < n) {
          const uint64_t r = s[i];
          loc = !(r >> 63);
          const float m = u2f((uint32_t)r);
          const double q = (alg_cic && loc) ? (double)fdiv(m, div) : (double)m / (double)div;
          s[i] = (uint64_t)__double_as_longlong(q);
        }
        n_local += __popc(__ballot_sync(0xffffffffu, loc));
      }
      __syncwarp();
      if (lane == 0) {
        for (int i = 0; i < n; i++) {
          const double q = __longlong_as_double((long long)s[i]);
          cur = (alg_cic && i < n_local) ? fadd(cur, (float)q) : (float)((double)cur + q);
        }
        out[gp] = cur;
      }
      __syncwarp();
    } else {
      // more records than the warp's slice holds: take them in ascending order, one minimum search per record
      uint64_t last = 0ull;
      bool first = true;
      for (int done = 0; done < n; done++) {
        uint64_t best = ~0ull;
        for (int i = lane; i < n; i += 32) {
          const uint64_t r = seg[off + i];
          if ((first || r > last) && r < best) best = r;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          const uint64_t o = __shfl_xor_sync(0xffffffffu, best, d);
          best = o < best ? o : best;
        }
        const int remote = (int)(best >> 63);
        cur = accumulate(cur, u2f((uint32_t)best), div, alg_cic && !remote);
        last = best;
        first = false;
      }
      if (lane == 0) out[gp] = cur;
    }
  }
}

// sum(value) in double and max(value) over a float array (dense_stats, src/dense.cpp:1298-1326)
__global__ void k_grid_stats(const float *__restrict__ v, unsigned long long n, double *sum, float *mx)
{
  double s = 0.0;
  float m = 0.0f;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    float x = v[i];
    s += (double)x;
    m = fmaxf(m, x);
  }
  for (int d = 16; d > 0; d >>= 1) {
    s += __shfl_down_sync(0xffffffffu, s, d);
    m = fmaxf(m, __shfl_down_sync(0xffffffffu, m, d));
  }
  __shared__ double ss[32];
  __shared__ float sm[32];
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { ss[w] = s; sm[w] = m; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int nw = (blockDim.x + 31) >> 5;
    double S = 0.0;
    float M = 0.0f;
    for (int i = 0; i < nw; i++) { S += ss[i]; M = fmaxf(M, sm[i]); }
    sum[blockIdx.x] = S;   // per-CTA partials, summed in order on the host: deterministic
    mx[blockIdx.x] = M;
  }
}

// workspace of k_cell_volumes (star + neighbours together: 142 words = 568 B per thread)
template <int STRIDE>
struct VolWS
{
  int *base;
  static constexpr int NU = 52, NTI = 88, PAR = 97, VH = 110, NH = 126, WORDS = 142;
  __device__ __forceinline__ int &star(int i) { return base[(size_t)i * STRIDE]; }
  __device__ __forceinline__ int &nu(int i) { return base[(size_t)(NU + i) * STRIDE]; }
  __device__ __forceinline__ unsigned char &byte(int word0, int i)
  {
    return reinterpret_cast<unsigned char *>(&base[(size_t)(word0 + (i >> 2)) * STRIDE])[i & 3];
  }
  __device__ __forceinline__ unsigned char &nt_idx(int i) { return byte(NTI, i); }
  __device__ __forceinline__ unsigned char &parent_idx(int i) { return byte(PAR, i); }
  __device__ __forceinline__ unsigned char &vis_hash(unsigned h) { return byte(VH, (int)h); }
  __device__ __forceinline__ unsigned char &nbr_hash(unsigned h) { return byte(NH, (int)h); }
  __device__ __forceinline__ int nt(int i) { return star((int)nt_idx(i)); }
  __device__ __forceinline__ void hash_clear()
  {
#pragma unroll
    for (int w = VH; w < WORDS; w++) base[(size_t)w * STRIDE] = -1;
  }
};
constexpr size_t VOL_SMEM = (size_t)VolWS<TOPO_THREADS>::WORDS * TOPO_THREADS * sizeof(int);

// ---- K2: per-site complete flag, Voronoi volume and zero-order density ------------------------------
__global__ void __launch_bounds__(TOPO_THREADS) k_cell_volumes(DevBlock blk, int num_sites, float mass, int *__restrict__ complete_out,
                                                               float *__restrict__ volume_out, float *__restrict__ density_out,
                                                               uint32_t *overflow, unsigned int *n_overflow, uint32_t cap_overflow)
{
  extern __shared__ int ws_s[];
  int site = blockIdx.x * TOPO_THREADS + threadIdx.x;
  VolWS<TOPO_THREADS> ws{ws_s + threadIdx.x};
  int status = -1, n_star = 0, n_nbr = 0;
  if (site < num_sites) {
    int t0 = blk.v2t[site];
    status = t0 < 0 ? CELL_NO_TET : star_and_neighbors_hashed(site, t0, blk.tets, ws, TOPO_STAR_CAP, TOPO_NBR_CAP, &n_star, &n_nbr);
  }
  __syncwarp();
  bool ovf = status == CELL_OVERFLOW;
  uint32_t slot = warp_append<unsigned int>(n_overflow, ovf);
  if (ovf && slot < cap_overflow) overflow[slot] = (uint32_t)site;
  if (site >= num_sites || ovf) return;
  int comp = status == CELL_NO_TET ? -1 : (status == CELL_OK ? 1 : 0);
  float vol = status == CELL_NO_TET ? -2.0f : -1.0f;
  if (status == CELL_OK) {
    const float *pv = blk.particles + 3 * (size_t)site;
    vol = 0.0f;
    for (int k = 0; k < n_nbr; k++) {
      AreaAccum aa;
      aa.area = 0.0f;
      int u = ws.nu(k);
      walk_edge_link(site, u, ws.nt(k), blk.tets, blk.cc, aa);
      const float *pu = blk.particles + 3 * (size_t)u;
      // distance(), src/tet.cpp:146-153: n += (u-v)*(u-v); sqrt in double rounded to float == sqrtf
      float n = 0.0f;
      for (int d = 0; d < 3; d++) {
        float df = fsub(pu[d], pv[d]);
        n = fadd(n, fmul(df, df));
      }
      float dist = fsqrt(n);
      vol = fadd(vol, fdiv(fmul(aa.area, dist), 6.0f)); // src/volume.cpp:50
    }
  }
  if (complete_out) complete_out[site] = comp;
  if (volume_out) volume_out[site] = vol;
  if (density_out) density_out[site] = vol > 0.0f ? fdiv(mass, vol) : 0.0f;
}

__global__ void __launch_bounds__(128) k_cell_volumes_big(DevBlock blk, const uint32_t *__restrict__ sites, int n_sites, int *ws_g, float mass,
                                                          int *__restrict__ complete_out, float *__restrict__ volume_out,
                                                          float *__restrict__ density_out)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sites) return;
  DynStridedWS ws{ws_g + i, (size_t)n_sites, BIG_STAR_CAP, BIG_NBR_CAP};
  int site = (int)sites[i];
  int n_star = 0, n_nbr = 0;
  int status = star_and_neighbors(site, blk.v2t[site], blk.tets, ws, BIG_STAR_CAP, BIG_NBR_CAP, &n_star, &n_nbr);
  int comp = status == CELL_OK ? 1 : 0;
  float vol = -1.0f;
  if (status == CELL_OK) {
    const float *pv = blk.particles + 3 * (size_t)site;
    vol = 0.0f;
    for (int k = 0; k < n_nbr; k++) {
      AreaAccum aa;
      aa.area = 0.0f;
      int u = ws.nu(k);
      walk_edge_link(site, u, ws.nt(k), blk.tets, blk.cc, aa);
      const float *pu = blk.particles + 3 * (size_t)u;
      float n = 0.0f;
      for (int d = 0; d < 3; d++) {
        float df = fsub(pu[d], pv[d]);
        n = fadd(n, fmul(df, df));
      }
      vol = fadd(vol, fdiv(fmul(aa.area, fsqrt(n)), 6.0f));
    }
  }
  if (complete_out) complete_out[site] = comp;
  if (volume_out) volume_out[site] = vol;
  if (density_out) density_out[site] = vol > 0.0f ? fdiv(mass, vol) : 0.0f;
}


// ---- K2 on the dense stage's star kernels ------------------------------------------------------------------
// volume() (src/volume.cpp:13-54) = sum over the Delaunay neighbours u, in neighbor_edges' order, of [fan area of the dual face
// from edge_link[0]] * |u - v| / 6.  k_cell_bfs + k_cell_nbrs (+ the general walk) give every complete site its face list in that
// order; the per-face terms are independent -- one thread per FACE (k_face_vol_terms: edge-link walk, fan area through double as
// the reference's sqrt binds to double, src/volume.cpp:45) -- and a site's terms are then added in order by one thread per site
// (k_cell_vol_sum: the float sum of the reference, a segmented reduction that keeps the reference's order; no atomics).
__global__ void __launch_bounds__(256) k_face_vol_terms(const FaceRef *__restrict__ faces, size_t n_faces, const DevBlock *__restrict__ blocks,
                                                         float *__restrict__ terms)
{
  const size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  FaceRef r;
  r.site = 0; r.u = -1; r.ut = 0; r.blk = 0;
  if (f < n_faces) r = faces[f];
  const bool act = f < n_faces && r.u >= 0;
  AreaAccum aa;
  aa.area = 0.0f;
  if (act) {
    const DevBlock &b = blocks[r.blk & ~FACE_DIRECT];
    if (b.walk) {
      const int4 v0 = b.tets[2 * (size_t)r.ut];
      const int s_c = v0.x == r.site ? 0 : (v0.y == r.site ? 1 : (v0.z == r.site ? 2 : 3));
      const int s_u = v0.x == r.u ? 0 : (v0.y == r.u ? 1 : (v0.z == r.u ? 2 : 3));
      walk_edge_link_rec(s_c, s_u, r.ut, b.walk, aa);
    } else {
      walk_edge_link(r.site, r.u, r.ut, b.tets, b.cc, aa);
    }
  }
  __syncwarp();
  if (!act) return;
  const DevBlock &b = blocks[r.blk & ~FACE_DIRECT];
  const float *pv = b.particles + 3 * (size_t)r.site, *pu = b.particles + 3 * (size_t)r.u;
  // distance(), src/tet.cpp:146-153: n += (u-v)*(u-v); sqrt in double rounded to float == sqrtf
  float d2 = 0.0f;
  for (int d = 0; d < 3; d++) {
    const float df = fsub(pu[d], pv[d]);
    d2 = fadd(d2, fmul(df, df));
  }
  terms[f] = fdiv(fmul(aa.area, fsqrt(d2)), 6.0f);      // src/volume.cpp:50
}

__global__ void __launch_bounds__(256) k_cell_vol_sum(const CellHdr *__restrict__ hdrs, uint32_t n_hdrs, const float *__restrict__ terms,
                                                       const DevBlock *__restrict__ blocks, float mass, int *__restrict__ complete_out,
                                                       float *__restrict__ volume_out, float *__restrict__ density_out)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_hdrs) return;
  const CellHdr h = hdrs[i];
  const int nf = (int)(h.blk_nf & 0xffffu);
  const uint32_t site = h.cell - blocks[h.blk_nf >> 16].cell_base;
  const float *t = terms + (size_t)h.plane_off * 2;
  float vol = 0.0f;
  for (int k = 0; k < nf; k++) vol = fadd(vol, t[k]);
  if (complete_out) complete_out[site] = 1;
  if (volume_out) volume_out[site] = vol;
  if (density_out) density_out[site] = vol > 0.0f ? fdiv(mass, vol) : 0.0f;
}

// sites that are in no tet (-1 / -2) or whose cell is infinite (0 / -1): the default the complete cells overwrite
__global__ void k_vol_defaults(const int *__restrict__ v2t, int num_sites, int *__restrict__ complete_out, float *__restrict__ volume_out,
                               float *__restrict__ density_out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num_sites) return;
  const bool none = v2t[i] < 0;
  if (complete_out) complete_out[i] = none ? -1 : 0;
  if (volume_out) volume_out[i] = none ? -2.0f : -1.0f;
  if (density_out) density_out[i] = 0.0f;
}

// ---- DTFE mode (alg 2) -------------------------------------------------------------------------------
struct NoSink
{
  __device__ __forceinline__ void operator()(int, int, int) {}
};

// rho(v) = 4 m / sum of the volumes of the tets of v's star (double sum in BFS order); -1 where the
// star is infinite or v is in no tet.  One thread per particle (ghosts too: tets near the block
// border use their densities).
__global__ void __launch_bounds__(TOPO_THREADS) k_vertex_density(DevBlock blk, float mass, float *__restrict__ rho, uint32_t *overflow,
                                                                 unsigned int *n_overflow, uint32_t cap_overflow)
{
  extern __shared__ int ws_s[];
  const int v = blockIdx.x * TOPO_THREADS + threadIdx.x;
  StarWS<TOPO_THREADS> ws{ws_s + threadIdx.x};
  int status = -1, n_star = 0;
  double sum = 0.0;
  if (v < blk.num_particles) {
    float cmin[3] = {0, 0, 0}, cmax[3] = {0, 0, 0};
    NoSink sink;
    int t0 = blk.v2t[v];
    status = t0 < 0 ? CELL_NO_TET : star_bfs_cands(v, t0, blk.tets, blk.cc, ws, TOPO_STAR_CAP, &n_star, cmin, cmax, sink, &sum);
  }
  __syncwarp();
  const bool ovf = status == CELL_OVERFLOW;
  uint32_t slot = warp_append<unsigned int>(n_overflow, ovf);
  if (ovf && slot < cap_overflow) overflow[slot] = (uint32_t)v;
  if (v < blk.num_particles && !ovf) rho[v] = (status == CELL_OK && sum > 0.0) ? (float)(4.0 * (double)mass / sum) : -1.0f;
}

// the same for stars that do not fit the shared-memory workspace: one warp per vertex
__global__ void __launch_bounds__(128) k_vertex_density_big(DevBlock blk, float mass, float *__restrict__ rho, const uint32_t *__restrict__ verts,
                                                            int n_verts, int *ws_g)
{
  const int wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = (int)lane_id();
  if (wi >= n_verts) return;
  const int v = (int)verts[wi];
  int *star = ws_g + (size_t)wi * BIG_STAR_CAP;
  int ns = 1;
  bool finite = true, bad = false;
  double sum = 0.0;
  if (lane == 0) star[0] = blk.v2t[v];
  __syncwarp();
  for (int head = 0; head < ns && finite && !bad; head++) {
    const int t = star[head];
    const int4 vv4 = blk.tets[2 * (size_t)t], nb = blk.tets[2 * (size_t)t + 1];
    sum += (double)blk.cc[t].w;
    const int vv[4] = {vv4.x, vv4.y, vv4.z, vv4.w}, bb[4] = {nb.x, nb.y, nb.z, nb.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (!finite || bad || vv[i] == v) continue;
      const int next = bb[i];
      if (next < 0) { finite = false; continue; }
      if (!warp_contains(star, ns, next)) {
        if (ns >= BIG_STAR_CAP) { bad = true; continue; }
        if (lane == 0) star[ns] = next;
        ns++;
        __syncwarp();
      }
    }
  }
  if (lane == 0) rho[v] = (finite && !bad && sum > 0.0) ? (float)(4.0 * (double)mass / sum) : -1.0f;
}

// tet-to-grid rasterisation by point location.  A tet of config 2 owns ~1.2 grid points but its
// bounding box holds ~90, so testing boxes (the first version: 20 ms) wastes 75 of 76 tests.  Instead
// every thread takes a short run of consecutive grid points of one grid row and WALKS the
// triangulation: from the tet that owned the previous point (or, for the first point, from a tet of
// a seed particle near it) it crosses the face that separates it from the point until DtfeTet::locate
// says "owner".  Ownership is the same exact partition the per-tet formulation uses (bit-identical
// face determinants on both sides of a face), so the walk ends at the unique owner; the visibility
// walk terminates in a Delaunay triangulation.  Points outside the hull, owners with an invalid
// vertex and (never observed) walks over DTFE_MAX_STEPS leave the zero the buffer was cleared to.
constexpr int DTFE_CHUNK = 32;
constexpr int DTFE_MAX_STEPS = 4096;

// seed particles: one per cell of a coarse grid over the block's bounds (the largest index wins;
// which one is irrelevant for the result)
__global__ void k_dtfe_seed(DevBlock blk, float3 bmin, float3 inv_cell, int3 cg, int *__restrict__ seed)
{
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= blk.num_particles || blk.v2t[v] < 0) return;
  const float x = (blk.particles[3 * (size_t)v] - bmin.x) * inv_cell.x, y = (blk.particles[3 * (size_t)v + 1] - bmin.y) * inv_cell.y,
              z = (blk.particles[3 * (size_t)v + 2] - bmin.z) * inv_cell.z;
  if (x < 0.0f || y < 0.0f || z < 0.0f || x >= (float)cg.x || y >= (float)cg.y || z >= (float)cg.z) {
    atomicMax(&seed[(size_t)cg.x * cg.y * cg.z], v);   // fallback slot: any particle at all
    return;
  }
  atomicMax(&seed[((size_t)(int)z * cg.y + (int)y) * cg.x + (int)x], v);
  atomicMax(&seed[(size_t)cg.x * cg.y * cg.z], v);
}

__device__ __forceinline__ void dtfe_load(const DevBlock &blk, int t, int *tv, int *nb, float (*p)[3])
{
  const int4 v = blk.tets[2 * (size_t)t], n = blk.tets[2 * (size_t)t + 1];
  tv[0] = v.x; tv[1] = v.y; tv[2] = v.z; tv[3] = v.w;
  nb[0] = n.x; nb[1] = n.y; nb[2] = n.z; nb[3] = n.w;
#pragma unroll
  for (int i = 0; i < 4; i++)
    for (int d = 0; d < 3; d++) p[i][d] = blk.particles[3 * (size_t)tv[i] + d];
}

__global__ void __launch_bounds__(128) k_dtfe_raster(DevBlock blk, const float *__restrict__ rho, const __grid_constant__ GridGeom g, int3 b_lo, int3 b_num,
                                                     float3 bmin, float3 inv_cell, int3 cg, const int *__restrict__ seed, float *__restrict__ out,
                                                     unsigned int *n_fail)
{
  const int chunks = (b_num.x + DTFE_CHUNK - 1) / DTFE_CHUNK;
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)chunks * b_num.y * b_num.z;
  if (id >= total) return;
  const int ch = (int)(id % chunks);
  const long long row = id / chunks;
  const int j = (int)(row % b_num.y), k = (int)(row / b_num.y);
  const int i0 = ch * DTFE_CHUNK, i1 = min(i0 + DTFE_CHUNK, b_num.x);
  const float py = idx2phys1(b_lo.y + j, g.step[1], g.gmin[1]), pz = idx2phys1(b_lo.z + k, g.step[2], g.gmin[2]);
  int t = -1;                                       // tet the walk stands in (-1: needs a seed)
  for (int i = i0; i < i1; i++) {
    const float pos[3] = {idx2phys1(b_lo.x + i, g.step[0], g.gmin[0]), py, pz};
    if (t < 0) {
      // seed: a particle of the coarse cell holding the point, else of a neighbouring cell, else any
      int cx = (int)((pos[0] - bmin.x) * inv_cell.x), cy = (int)((pos[1] - bmin.y) * inv_cell.y), cz = (int)((pos[2] - bmin.z) * inv_cell.z);
      cx = max(0, min(cx, cg.x - 1)); cy = max(0, min(cy, cg.y - 1)); cz = max(0, min(cz, cg.z - 1));
      int sv = seed[((size_t)cz * cg.y + cy) * cg.x + cx];
      for (int dz = -1; dz <= 1 && sv < 0; dz++)
        for (int dy = -1; dy <= 1 && sv < 0; dy++)
          for (int dx = -1; dx <= 1 && sv < 0; dx++) {
            int x = cx + dx, y = cy + dy, z = cz + dz;
            if (x >= 0 && y >= 0 && z >= 0 && x < cg.x && y < cg.y && z < cg.z) sv = seed[((size_t)z * cg.y + y) * cg.x + x];
          }
      if (sv < 0) sv = seed[(size_t)cg.x * cg.y * cg.z];
      if (sv < 0) return;                           // no tet in this block at all
      t = blk.v2t[sv];
    }
    // visibility walk
    DtfeTet T;
    int tv[4], nb[4];
    float p[4][3], sp[4];
    bool found = false;
    int last = t;
    for (int step = 0; step < DTFE_MAX_STEPS; step++) {
      last = t;
      dtfe_load(blk, t, tv, nb, p);
      int f = T.locate_lazy(tv, p, pos, sp);
      if (f == -1) { found = true; break; }
      if (f == -2) f = (nb[0] >= 0) ? 0 : (nb[1] >= 0 ? 1 : (nb[2] >= 0 ? 2 : 3));   // degenerate tet: leave through any face
      t = nb[f];
      if (t < 0) break;                             // left the hull: the point is in no tet
    }
    if (found) {
      bool valid = true;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        T.rho[q] = rho[tv[q]];
        valid = valid && T.rho[q] >= 0.0f;
      }
      if (valid) out[((size_t)k * b_num.y + j) * b_num.x + i] = T.value_from(sp);
      // t stays: the next point of the row starts here
    } else {
      if (t >= 0) atomicAdd(n_fail, 1u);
      t = last;                                      // resume from the last tet inside the hull
    }
  }
}

} // namespace tb
