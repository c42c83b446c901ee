// kernels.cuh -- sm_100a kernels of the dense stage.  See DESIGN.md for the data layout in HBM,
// the algorithmic bytes of each kernel and the roofline that bounds it.
//
//   k_vert_to_tet      fill_vert_to_tet            (src/tess.cpp:767-787)
//   k_circumcenters    fill_circumcenters          (src/volume.cpp:6-11, src/tet.cpp:37-66)
//   k_cell_topo        complete + CellBounds + the data-bounds filter and index box of CellGridPts
//                      (src/tet.cpp:228-270,337-409; src/dense.cpp:657-736,1381-1410)
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

// ---- device-side descriptors -------------------------------------------------------------------
struct DevBlock
{
  const float *particles;  // xyz AoS
  const int4 *tets;        // 2 x int4 per tet: verts, neighbours
  const int *v2t;
  const float4 *cc;        // circumcenter per tet (w unused)
  int num_orig, num_particles, num_tets;
  uint32_t cell_base;      // global number of this block's cell 0 (blocks in ascending gid order)
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

struct Counters
{
  unsigned long long n_no_tet, n_incomplete, n_outside, n_bad, n_deposit, n_cic_fallback;
  unsigned int n_small, n_big, n_overflow;     // list lengths
  unsigned int plane_cursor;                   // units of 2 faces
  unsigned long long big_bits;                 // bits needed by the big-cell list (multiples of 32)
  unsigned long long n_spans;                  // span records requested (may exceed capacity)
};

struct TopoOut
{
  CellHdr *small, *big;
  unsigned long long *big_bit_off;
  uint2 *overflow;         // (block index, block-local cell) whose star did not fit the fast workspace
  float *plane_pool;
  Counters *cnt;
  uint32_t cap_small, cap_big, cap_overflow;
};

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

__device__ __forceinline__ void warp_count(unsigned long long *counter, bool pred)
{
  unsigned m = __ballot_sync(0xffffffffu, pred);
  if (m && (int)lane_id() == __ffs(m) - 1) atomicAdd(counter, (unsigned long long)__popc(m));
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

// ---- K1: circumcenters, one thread per tet ---------------------------------------------------------
// reads 16 B of the tet record (verts only) + 4 gathered particles, writes one float4
__global__ void __launch_bounds__(256) k_circumcenters(const int4 *__restrict__ tets, int num_tets,
                                                        const float *__restrict__ particles, float4 *__restrict__ cc)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= num_tets) return;
  int4 v = __ldg(&tets[2 * (size_t)t]);
  float a[3], b[3], c[3], d[3], o[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    a[i] = __ldg(&particles[3 * (size_t)v.x + i]);
    b[i] = __ldg(&particles[3 * (size_t)v.y + i]);
    c[i] = __ldg(&particles[3 * (size_t)v.z + i]);
    d[i] = __ldg(&particles[3 * (size_t)v.w + i]);
  }
  circumcenter(a, b, c, d, o);
  cc[t] = make_float4(o[0], o[1], o[2], 0.0f);
}

// ---- K3a part 1: topology + faces, one thread per cell --------------------------------------------
template <int STRIDE>
struct StridedWS
{
  int *base;
  int star_cap;
  int nbr_cap;
  __device__ __forceinline__ int &star(int i) { return base[(size_t)i * STRIDE]; }
  __device__ __forceinline__ int &nu(int i) { return base[(size_t)(star_cap + i) * STRIDE]; }
  __device__ __forceinline__ int &nt(int i) { return base[(size_t)(star_cap + nbr_cap + i) * STRIDE]; }
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
constexpr int TOPO_STAR_CAP = 56;
constexpr int TOPO_NBR_CAP = 36;
constexpr size_t TOPO_SMEM = (size_t)(TOPO_STAR_CAP + 2 * TOPO_NBR_CAP) * TOPO_THREADS * sizeof(int);
constexpr int BIG_STAR_CAP = 4096;
constexpr int BIG_NBR_CAP = 1024;

constexpr int SCAN_FACE_CAP = 32;     // faces per cell held in shared memory by k_cell_scan
constexpr int SCAN_PTS_CAP = 2048;    // index-box points per cell handled by k_cell_scan

// The part after the star walk, common to the fast and the large-workspace kernels.  All lanes of
// the warp call it (inactive lanes with status != CELL_OK).
template <class WS>
__device__ __forceinline__ void topo_finish(int status, int cell, int n_nbr, WS &ws, const DevBlock &blk, int blk_id,
                                            const GridGeom &g, const TopoOut &out)
{
  // plane space: pairs of faces, so every cell's planes start 16-byte aligned for the bulk copy
  uint32_t want = status == CELL_OK ? (uint32_t)((n_nbr + 1) >> 1) : 0u;
  uint32_t poff = warp_alloc<unsigned int>(&out.cnt->plane_cursor, want);
  float cmin[3] = {0, 0, 0}, cmax[3] = {0, 0, 0};
  if (status == CELL_OK) {
    float site[3] = {blk.particles[3 * (size_t)cell], blk.particles[3 * (size_t)cell + 1], blk.particles[3 * (size_t)cell + 2]};
    bool first = true;
    float2 *dst = reinterpret_cast<float2 *>(out.plane_pool + (size_t)poff * 12);
    for (int k = 0; k < n_nbr; k++) {
      FaceAccum fa;
      fa.cmin = cmin; fa.cmax = cmax; fa.first_of_cell = &first;
      int n = walk_edge_link(cell, ws.nu(k), ws.nt(k), blk.tets, blk.cc, fa);
      if (n < 0) { status = CELL_BAD_MESH; break; }
      newell_term(fa.nrm, fa.prev, fa.v0);
      newell_finish(fa.nrm, fa.v0, site);
      dst[3 * k + 0] = make_float2(fa.nrm[0], fa.nrm[1]);
      dst[3 * k + 1] = make_float2(fa.nrm[2], fa.v0[0]);
      dst[3 * k + 2] = make_float2(fa.v0[1], fa.v0[2]);
    }
  }
  int lo[3] = {0, 0, 0}, n3[3] = {0, 0, 0};
  if (status == CELL_OK) {
    // src/dense.cpp:1385-1392
    for (int d = 0; d < 3; d++)
      if (cmin[d] < fsub(g.dmin[d], g.dext_eps[d]) || cmax[d] > fadd(g.dmax[d], g.dext_eps[d])) status = CELL_OUTSIDE;
  }
  long long npts = 0;
  if (status == CELL_OK) {
    for (int d = 0; d < 3; d++) {
      lo[d] = phys2idx1(cmin[d], g.step[d], g.gmin[d]);
      int hi = phys2idx1(cmax[d], g.step[d], g.gmin[d]);
      n3[d] = hi - lo[d] + 1;
    }
    if (n3[0] < 1 || n3[1] < 1 || n3[2] < 1 || n3[0] > 32767 || n3[1] > 32767 || n3[2] > 32767 ||
        lo[0] < -1 || lo[1] < -1 || lo[2] < -1)
      status = CELL_BAD_MESH;
    npts = (long long)n3[0] * n3[1] * n3[2];
  }
  bool ok = status == CELL_OK;
  bool small = ok && n_nbr <= SCAN_FACE_CAP && npts <= SCAN_PTS_CAP;
  bool big = ok && !small;
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
  warp_count(&out.cnt->n_no_tet, status == CELL_NO_TET);
  warp_count(&out.cnt->n_incomplete, status == CELL_INCOMPLETE);
  warp_count(&out.cnt->n_outside, status == CELL_OUTSIDE);
  warp_count(&out.cnt->n_bad, status == CELL_BAD_MESH);
}

__global__ void __launch_bounds__(TOPO_THREADS) k_cell_topo(DevBlock blk, int blk_id, const __grid_constant__ GridGeom g, TopoOut out)
{
  extern __shared__ int ws_s[];
  int cell = blockIdx.x * TOPO_THREADS + threadIdx.x;
  StridedWS<TOPO_THREADS> ws{ws_s + threadIdx.x, TOPO_STAR_CAP, TOPO_NBR_CAP};
  int status = -1, n_star = 0, n_nbr = 0;
  if (cell < blk.num_orig) {
    int t0 = blk.v2t[cell];
    status = t0 < 0 ? CELL_NO_TET : star_and_neighbors(cell, t0, blk.tets, ws, TOPO_STAR_CAP, TOPO_NBR_CAP, &n_star, &n_nbr);
  }
  __syncwarp();
  bool ovf = status == CELL_OVERFLOW;
  uint32_t o_slot = warp_append<unsigned int>(&out.cnt->n_overflow, ovf);
  if (ovf && o_slot < out.cap_overflow) out.overflow[o_slot] = make_uint2((unsigned)blk_id, (unsigned)cell);
  topo_finish(ovf ? -1 : status, cell, n_nbr, ws, blk, blk_id, g, out);
}

// large-workspace retry for the (rare) cells whose star exceeds the shared-memory workspace
__global__ void __launch_bounds__(128) k_cell_topo_big(const DevBlock *__restrict__ blocks, const __grid_constant__ GridGeom g, TopoOut out,
                                                        const uint2 *__restrict__ cells, int n_cells, int *ws_g)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  DynStridedWS ws{ws_g + (i < n_cells ? i : 0), (size_t)n_cells, BIG_STAR_CAP, BIG_NBR_CAP};
  int status = -1, n_star = 0, n_nbr = 0, cell = 0, blk_id = 0;
  if (i < n_cells) {
    blk_id = (int)cells[i].x;
    cell = (int)cells[i].y;
    const DevBlock &b = blocks[blk_id];
    status = star_and_neighbors(cell, b.v2t[cell], b.tets, ws, BIG_STAR_CAP, BIG_NBR_CAP, &n_star, &n_nbr);
    if (status == CELL_OVERFLOW) status = CELL_BAD_MESH; // documented limit: > 4096 tets around one site
  }
  __syncwarp();
  DevBlock blk = blocks[blk_id];
  topo_finish(status, cell, n_nbr, ws, blk, blk_id, g, out);
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
  __device__ __forceinline__ void operator()(int yi, int zi, int min_xi, int max_xi)
  {
    int y = lo[1] + yi, z = lo[2] + zi, xa = lo[0] + min_xi, xb = lo[0] + max_xi;
    if (local_box) {
      const BlockBox &b = sc.boxes[e];
      int ly = y - b.b_lo[1], lz = z - b.b_lo[2];
      uint64_t row = (uint64_t)(b.row_base + (sc.project ? (long long)ly : (long long)lz * b.b_num[1] + ly));
      emit(make_key(sc.kl, row, 0, cell, sc.project ? (uint32_t)z : 0u), make_data(xa - b.b_lo[0], xb - xa + 1, 0, value));
    } else {
      emit_line(sc.boxes, sc.nblocks, e, sc.kl, sc.project, cell, xa, xb, y, z, 0, value, emit);
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

// ---- K3a part 2: inside bits + scan-line walk, 64 cells per CTA -----------------------------------
constexpr int SCAN_CELLS = 64;
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_SLOT_FLOATS = 196;                 // 32 faces * 6 floats + 4 pad (784 B, 16-B multiple, bank shift 4)
constexpr int SCAN_BIT_WORDS = 2048;                  // 65536 inside bits per round
constexpr size_t SCAN_SMEM = SCAN_CELLS * SCAN_SLOT_FLOATS * 4 + SCAN_BIT_WORDS * 4 + SCAN_CELLS * sizeof(CellHdr) +
                             (SCAN_CELLS + 2) * 4 + 32;

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
struct NullLine
{
  __device__ __forceinline__ void operator()(int, int, int, int) {}
};

__device__ __forceinline__ bool pt_in_cell_smem(const float *pl, int nf, const float *pt, float eps)
{
  bool pos = false, neg = false;
  for (int k = 0; k < nf; k++) {
    const float2 *p = reinterpret_cast<const float2 *>(pl + 6 * k);
    float2 a = p[0], b = p[1], c = p[2];
    float n[3] = {a.x, a.y, b.x}, f[3] = {b.y, c.x, c.y};
    int s = plane_side(n, f, pt, eps);
    pos |= s > 0;
    neg |= s < 0;
    if (pos && neg) return false;
  }
  return true;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_cell_scan(const CellHdr *__restrict__ hdrs, const Counters *cnt_in, uint32_t cap_small,
                                                             const float *__restrict__ plane_pool, const DevBlock *__restrict__ blocks,
                                                             ScanCtx sc, const __grid_constant__ GridGeom g, SpanOut out)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *planes_s = reinterpret_cast<float *>(smem_raw);
  uint32_t *bits_s = reinterpret_cast<uint32_t *>(planes_s + SCAN_CELLS * SCAN_SLOT_FLOATS);
  CellHdr *hdr_s = reinterpret_cast<CellHdr *>(bits_s + SCAN_BIT_WORDS);
  int *pfx_s = reinterpret_cast<int *>(hdr_s + SCAN_CELLS);
  uint64_t *bar = reinterpret_cast<uint64_t *>(pfx_s + SCAN_CELLS + 2);

  uint32_t n_hdrs = cnt_in->n_small < cap_small ? cnt_in->n_small : cap_small;
  uint32_t first = blockIdx.x * SCAN_CELLS;
  if (first >= n_hdrs) return;
  int ncell = (int)(n_hdrs - first < (uint32_t)SCAN_CELLS ? n_hdrs - first : SCAN_CELLS);
  const int tid = threadIdx.x;

  if (tid == 0) {
    mbar_init(bar, SCAN_CELLS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // stage headers and planes: thread c < 64 owns cell c
  if (tid < SCAN_CELLS) {
    uint32_t bytes = 0;
    int npts = 0;
    if (tid < ncell) {
      CellHdr h = hdrs[first + tid];
      hdr_s[tid] = h;
      int nf = (int)(h.blk_nf & 0xffffu);
      bytes = (uint32_t)((nf + 1) >> 1) * 48u;
      npts = (int)h.n3[0] * (int)h.n3[1] * (int)h.n3[2];
      mbar_arrive_expect_tx(bar, bytes);
      if (bytes) bulk_g2s(planes_s + tid * SCAN_SLOT_FLOATS, plane_pool + (size_t)h.plane_off * 12, bytes, bar);
    } else {
      mbar_arrive_expect_tx(bar, 0);
    }
    // inclusive prefix of npts over the 64 cells (two warps)
    int incl = npts;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int o = __shfl_up_sync(0xffffffffu, incl, d);
      if ((int)lane_id() >= d) incl += o;
    }
    pfx_s[tid + 1] = incl;
  }
  if (tid == 0) pfx_s[0] = 0;
  __syncthreads();
  if (tid >= 32 && tid < SCAN_CELLS) pfx_s[tid + 1] += pfx_s[32];
  __syncthreads();
  mbar_wait(bar, 0);

  // rounds of consecutive cells whose inside bits fit the bit buffer
  int c0 = 0;
  while (c0 < ncell) {
    int c1 = c0;
    int base_pts = pfx_s[c0];
    while (c1 < ncell && pfx_s[c1 + 1] - base_pts <= SCAN_BIT_WORDS * 32) c1++;
    int total = pfx_s[c1] - base_pts;
    int total32 = (total + 31) & ~31;

    // phase 2: PtInCell for every point of every index box, one point per lane
    for (int s = tid; s < total32; s += SCAN_THREADS) {
      bool in = false;
      if (s < total) {
        int lo_c = c0, hi_c = c1; // find cell: pfx[c] - base <= s < pfx[c+1] - base
        while (hi_c - lo_c > 1) {
          int mid = (lo_c + hi_c) >> 1;
          if (pfx_s[mid] - base_pts <= s) lo_c = mid; else hi_c = mid;
        }
        const CellHdr &h = hdr_s[lo_c];
        int l = s - (pfx_s[lo_c] - base_pts);
        int nx = h.n3[0], ny = h.n3[1];
        int i = l % nx;
        int r = l / nx;
        int j = r % ny;
        int k = r / ny;
        float pt[3];
        // probe position: cell_min_grid_pos + i * step, cell_min_grid_pos = idx2phys(lo) (src/dense.cpp:1404,1530-1532)
        pt[0] = fadd(idx2phys1(h.lo[0], g.step[0], g.gmin[0]), fmul((float)i, g.step[0]));
        pt[1] = fadd(idx2phys1(h.lo[1], g.step[1], g.gmin[1]), fmul((float)j, g.step[1]));
        pt[2] = fadd(idx2phys1(h.lo[2], g.step[2], g.gmin[2]), fmul((float)k, g.step[2]));
        in = pt_in_cell_smem(planes_s + lo_c * SCAN_SLOT_FLOATS, (int)(h.blk_nf & 0xffffu), pt, g.eps);
      }
      unsigned w = __ballot_sync(0xffffffffu, in);
      if (lane_id() == 0) bits_s[s >> 5] = w;
    }
    __syncthreads();

    // phase 3: the scan-line walk on the bits, one thread per cell; pass 1 counts, pass 2 emits
    if (tid < 64) {
      int c = c0 + tid;
      bool act = c < c1;
      int tot = 0, nrec = 0, e = 0;
      bool local_box = false;
      float site[3] = {0, 0, 0};
      CellHdr h;
      BitsInside inside{bits_s, 0, 1, 1};
      if (act) {
        h = hdr_s[c];
        e = (int)(h.blk_nf >> 16);
        int lo[3] = {h.lo[0], h.lo[1], h.lo[2]}, n3[3] = {h.n3[0], h.n3[1], h.n3[2]};
        local_box = box_is_local(sc.boxes[e], lo, n3, sc.project);
        inside.off = (uint32_t)(pfx_s[c] - base_pts);
        inside.nx = n3[0]; inside.ny = n3[1];
        CountEmit ce{0};
        LineEmitter<CountEmit> le{sc, e, h.cell, h.lo, local_box, 0.0f, ce, 0};
        tot = scan_cell(n3[0], n3[1], n3[2], inside, le);
        nrec = ce.n;
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
      warp_count(&out.cnt->n_deposit, act);
      warp_count(&out.cnt->n_cic_fallback, act && tot == 0);
      if (act) {
        StoreEmit se{out.keys, out.data, base, out.capacity};
        if (tot > 0) {
          float m = fdiv(g.mass, (float)tot); // src/dense.cpp:1692
          LineEmitter<StoreEmit> le{sc, e, h.cell, h.lo, local_box, m, se, 0};
          scan_cell((int)h.n3[0], (int)h.n3[1], (int)h.n3[2], inside, le);
        } else {
          emit_cic(sc, e, h.cell, site, g, 0, se);
        }
      }
    }
    __syncthreads();
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
                                                        int n_cells, const float *__restrict__ plane_pool, uint32_t *bits_g,
                                                        const DevBlock *__restrict__ blocks, ScanCtx sc,
                                                        const __grid_constant__ GridGeom g, SpanOut out)
{
  int ci = blockIdx.x;
  if (ci >= n_cells) return;
  CellHdr h = hdrs[ci];
  int nf = (int)(h.blk_nf & 0xffffu);
  const float *pl = plane_pool + (size_t)h.plane_off * 12;
  uint32_t *bits = bits_g + (bit_off[ci] >> 5);
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
    LineEmitter<CountEmit> le{sc, e, h.cell, h.lo, local_box, 0.0f, ce, 0};
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
      LineEmitter<StoreEmit> le2{sc, e, h.cell, h.lo, local_box, m, se, 0};
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

__global__ void __launch_bounds__(ROWS_WARPS * 32) k_rows(const uint64_t *__restrict__ data, const unsigned long long *__restrict__ row_start,
                                                          unsigned long long row0, unsigned long long nrows, const RowBlock *__restrict__ rblocks,
                                                          int n_rblocks, float div, int nx_max, float *__restrict__ out)
{
  extern __shared__ float rowbuf_all[];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long r = (unsigned long long)blockIdx.x * ROWS_WARPS + warp;
  if (r >= nrows) return;
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
  for (int x = lane; x < nx; x += 32) buf[x] = 0.0f;
  __syncwarp();
  unsigned long long s0 = row_start[r], s1 = row_start[r + 1];
  for (unsigned long long sb = s0; sb < s1; sb += 32) {
    unsigned long long mine = sb + lane < s1 ? data[sb + lane] : 0ull;
    int cnt = (int)(s1 - sb < 32 ? s1 - sb : 32);
    for (int j = 0; j < cnt; j++) {
      unsigned long long d = __shfl_sync(0xffffffffu, mine, j);
      int x0 = (int)(d & 0xffffu);
      int len = (int)((d >> 16) & 0x7fffu);
      int fp = (int)((d >> 31) & 1u);
      float m = u2f((uint32_t)(d >> 32));
      for (int x = lane; x < len; x += 32) {
        int xx = x0 + x;
        if (xx < nx) buf[xx] = accumulate(buf[xx], m, div, fp);
      }
      __syncwarp();
    }
  }
  float *dst = out + rb.out_off + (row - rb.row_base) * (long long)nx;
  for (int x = lane; x < nx; x += 32) dst[x] = buf[x];
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

// ---- K2: per-site complete flag, Voronoi volume and zero-order density ------------------------------
__global__ void __launch_bounds__(TOPO_THREADS) k_cell_volumes(DevBlock blk, int num_sites, float mass, int *__restrict__ complete_out,
                                                               float *__restrict__ volume_out, float *__restrict__ density_out,
                                                               uint32_t *overflow, unsigned int *n_overflow, uint32_t cap_overflow)
{
  extern __shared__ int ws_s[];
  int site = blockIdx.x * TOPO_THREADS + threadIdx.x;
  StridedWS<TOPO_THREADS> ws{ws_s + threadIdx.x, TOPO_STAR_CAP, TOPO_NBR_CAP};
  int status = -1, n_star = 0, n_nbr = 0;
  if (site < num_sites) {
    int t0 = blk.v2t[site];
    status = t0 < 0 ? CELL_NO_TET : star_and_neighbors(site, t0, blk.tets, ws, TOPO_STAR_CAP, TOPO_NBR_CAP, &n_star, &n_nbr);
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

} // namespace tb
