// api.cu -- host side of the C ABI declared in include/tess_b200.h.
//
// The grid bookkeeping the reference does on the host (DataBounds, GridStepParams,
// BlockGridParams; src/dense.cpp:1221-1275, 1712-1767, 575-648) stays on the host here too --
// it is a few dozen flops per block -- in the reference's fp32 operation order (this file is
// compiled with -fmad=false / -ffp-contract=off).  Everything per tet, per cell and per grid
// point runs in the kernels of kernels.cuh.  There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/transform_iterator.h>
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/tess_b200.h"
#include "kernels.cuh"
#include "fused.cuh"
#include "host_geom.hpp"
#ifdef TESSB200_WITH_NCCL
#include <nccl.h>
#include <dlfcn.h>
// NCCL is bound at run time (dlopen) instead of at link time: a process that also imports torch
// must end up with ONE libnccl.so.2 (torch bundles its own, newer than the system's), and the
// loader keys on the SONAME, so whichever copy is already mapped is the one used here.
namespace ncclw
{
struct Api
{
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static Api g;
static const char *load()
{
  if (g.h) return nullptr;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return dlerror();
#define TB_SYM(field, name)                                \
  *(void **)(&g.field) = dlsym(h, name);                   \
  if (!g.field) return "libnccl.so.2 lacks " name;
  TB_SYM(GetUniqueId, "ncclGetUniqueId")
  TB_SYM(CommInitRank, "ncclCommInitRank")
  TB_SYM(CommDestroy, "ncclCommDestroy")
  TB_SYM(AllGather, "ncclAllGather")
  TB_SYM(Send, "ncclSend")
  TB_SYM(Recv, "ncclRecv")
  TB_SYM(GroupStart, "ncclGroupStart")
  TB_SYM(GroupEnd, "ncclGroupEnd")
  TB_SYM(GetErrorString, "ncclGetErrorString")
#undef TB_SYM
  g.h = h;
  return nullptr;
}
} // namespace ncclw
#endif

using namespace tb;

// ---- error plumbing ----------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const char *fmt, ...)
{
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CU(expr)                                                                                             \
  do {                                                                                                       \
    cudaError_t e_ = (expr);                                                                                 \
    if (e_ != cudaSuccess) return fail(e_ == cudaErrorMemoryAllocation ? TESSB200_ENOMEM : TESSB200_ECUDA, \
                                       "%s:%d: %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
  } while (0)
#define COUNT_LAUNCH(c, n) ((c)->launches += (n))
#define TRY(expr)            \
  do {                       \
    int rc_ = (expr);        \
    if (rc_ != 0) return rc_; \
  } while (0)

// grow-only device buffer
struct Buf
{
  void *p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes)
  {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      e = cudaMalloc(&p, bytes);
      want = bytes;
    }
    if (e != cudaSuccess) {
      p = nullptr;
      cudaGetLastError();
      return fail(TESSB200_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    }
    cap = want;
    return 0;
  }
  void release()
  {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T> T *as() { return reinterpret_cast<T *>(p); }
};

struct BlockRes
{
  int gid = 0;
  int num_orig = 0, num_particles = 0, num_tets = 0;
  float bmin[3], bmax[3];
  bool have_v2t = false;
  Buf particles, tets, v2t, cc, rho, walk, hull, p4;
  // geometry of the last run
  int mn[3], num[3];
  long long npts = 0, nrows = 0, row_base = 0, out_off = 0;
  uint32_t cell_base = 0;
};

struct LayoutBlock
{
  int gid;
  float bmin[3], bmax[3];
  int owner;
};

struct tessb200_ctx
{
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t blk_ev[64] = {};
  cudaEvent_t grp_ev[64] = {};   // end of each group's cell kernels (TESSB200_TRACE)
  cudaEvent_t part_ev = nullptr;
  std::vector<cudaEvent_t> tr_ev;          // TESSB200_TRACE: fine-grained marks inside the groups
  std::vector<const char *> tr_name;
  size_t tr_used = 0;
  bool tracing = false;
  std::vector<cudaEvent_t> h2d_ev;
  std::vector<BlockRes *> blocks;   // uploaded blocks of this rank, ascending gid
  std::vector<LayoutBlock> layout;  // every block of the decomposition (multi-GPU), ascending gid
  int nranks = 1, rank = 0;
#ifdef TESSB200_WITH_NCCL
  ncclComm_t comm = nullptr;
#endif
  Buf d_blocks, d_boxes, d_rblocks, d_cnt, plane_pool, face_list, pre_hdr, cand, hdr_small, hdr_big, big_bitoff, overflow, ws_big, bits_big;
  Buf keys[2], data[2], cub_tmp, row_start, out, stat_sum, stat_max, recv_keys, recv_data, mkeys[2], order[2], x_small, pt_count;
  Buf fz_hdr, fz_bits, fz_pool;     // k_cell_fused -> k_cell_emit: headers, in-line inside bits, pool for the larger index boxes
  Buf hdr_dir[3];                   // cells of the small-box classes (k_cell_direct)
  Buf pt_off, pt_fill, big_points;  // shared grid points: segment offsets, fill cursors, the points with many deposits
  Buf cic_vals, cic_keys[2], cic_ids[2], cic_count, cic_start, cic_vals2;   // DENSE_CIC: weights, base cells and ids of the particles, particles per base cell
  bool cic_gather = true;           // TESSB200_CIC_GATHER=0: every CIC deposit as a record (round 1's k_cic)
  // span exchange: per (source, destination) capacities agreed in an exact round; later runs exchange fixed-size,
  // sentinel-padded segments and need no host read-back before the deposit
  std::vector<unsigned long long> xcap;   // [src * nranks + dst], identical on every rank; empty = no agreement yet
  bool x_fast_used = false;               // this run exchanged fixed-size segments: the count matrix is checked at its end
  unsigned long long *h_xall = nullptr;   // pinned copy of the all-gathered (counts | status) matrix
  bool segments = false;            // TESSB200_SEGMENTS=1: shared deposits through per-point segments instead of the radix sort + k_rows
                                    // (faster on uniform input, slower where clumps put hundreds of deposits on one point: profiles/r02)
  bool direct = true;               // TESSB200_DIRECT=0: every cell through k_cell_faces + k_cell_scan (A/B measurements)
  bool fused = false;               // TESSB200_FUSED=1 selects the one-kernel-per-cell path (fused.cuh; A/B measurements)
  int fz_ctas = 0;
  float last_k2_ms = 0.0f;          // device time of the kernels of the last tessb200_cell_volumes call
  Counters *h_cnt = nullptr;        // pinned
  double *h_sum = nullptr;
  float *h_max = nullptr;
  cudaEvent_t ev[20] = {};
  bool ran = false;
  long long launches = 0;           // kernels launched by the current run
  tessb200_dense_params last_params;
  long long out_floats = 0;
  ~tessb200_ctx() {}
};

extern "C" const char *tessb200_last_error(void) { return g_err.c_str(); }
extern "C" int tessb200_version(void) { return TESSB200_VERSION; }

static int create_impl(tessb200_ctx **out, int device);
extern "C" void tessb200_destroy(tessb200_ctx *c);

extern "C" int tessb200_create(tessb200_ctx **out, int device)
{
  if (!out) return fail(TESSB200_EINVAL, "tessb200_create: ctx is NULL");
  *out = nullptr;
  tessb200_ctx *c = nullptr;
  const int rc = create_impl(&c, device);
  if (rc) {
    // a CUDA call failed half way: give back what was created (the message of the failure stays)
    const std::string keep = g_err;
    if (c) tessb200_destroy(c);
    cudaGetLastError();
    g_err = keep;
    return rc;
  }
  *out = c;
  return 0;
}

// `*out` is set as soon as the context exists, so that the caller can free it when a later step fails
static int create_impl(tessb200_ctx **out, int device)
{
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(TESSB200_ECUDA, "no CUDA device available (%s); tess_b200 has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  }
  if (device < 0 || device >= n) return fail(TESSB200_EINVAL, "device %d out of range [0,%d)", device, n);
  CU(cudaSetDevice(device));
  tessb200_ctx *c = new tessb200_ctx;
  *out = c;
  c->device = device;
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  for (auto &ev : c->blk_ev) CU(cudaEventCreate(&ev));
  for (auto &ev : c->grp_ev) CU(cudaEventCreate(&ev));
  CU(cudaEventCreateWithFlags(&c->part_ev, cudaEventDisableTiming));
  CU(cudaMallocHost(&c->h_cnt, sizeof(Counters)));
  CU(cudaMallocHost(&c->h_sum, sizeof(double) * 1024));
  CU(cudaMallocHost(&c->h_max, sizeof(float) * 1024));
  for (auto &ev : c->ev) CU(cudaEventCreate(&ev));
  CU(cudaFuncSetAttribute(k_cell_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCAN_SMEM));
  CU(cudaFuncSetAttribute(k_cell_bfs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BFS_SMEM));
  CU(cudaFuncSetAttribute(k_cell_volumes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VOL_SMEM));
  CU(cudaFuncSetAttribute(k_cell_nbrs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NBRS_SMEM));
  CU(cudaFuncSetAttribute(k_vertex_density, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TOPO_SMEM));
  CU(cudaFuncSetAttribute(k_cell_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FZ_SMEM));
  CU(cudaFuncSetAttribute(k_cell_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EMIT_SMEM));
  {
    const char *f = getenv("TESSB200_FUSED");
    c->fused = f && f[0] == '1';
    const char *d = getenv("TESSB200_DIRECT");
    c->direct = !(d && d[0] == '0');
    const char *cg = getenv("TESSB200_CIC_GATHER");
    c->cic_gather = !(cg && cg[0] == '0');
    const char *sg = getenv("TESSB200_SEGMENTS");
    c->segments = sg && sg[0] == '1';
    CU(cudaFuncSetAttribute(k_point_apply_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POINT_BIG_SMEM));        // opt-in: measured slower than the four-kernel path (profiles/r02/fused_a_*)
    int per_sm = 0, sms = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cell_fused, FZ_THREADS, FZ_SMEM));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    c->fz_ctas = std::max(1, per_sm) * std::max(1, sms);
  }
  return 0;
}

static void free_blocks(tessb200_ctx *c)
{
  for (BlockRes *b : c->blocks) {
    b->particles.release(); b->tets.release(); b->v2t.release(); b->cc.release(); b->rho.release(); b->walk.release(); b->hull.release(); b->p4.release();
    delete b;
  }
  c->blocks.clear();
}

extern "C" void tessb200_destroy(tessb200_ctx *c)
{
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  free_blocks(c);
  Buf *bufs[] = {&c->d_blocks, &c->d_boxes, &c->d_rblocks, &c->d_cnt, &c->plane_pool, &c->face_list, &c->pre_hdr, &c->cand, &c->hdr_small, &c->hdr_big, &c->big_bitoff,
                 &c->overflow, &c->ws_big, &c->bits_big, &c->keys[0], &c->keys[1], &c->data[0], &c->data[1], &c->cub_tmp,
                 &c->row_start, &c->out, &c->stat_sum, &c->stat_max, &c->recv_keys, &c->recv_data, &c->mkeys[0], &c->mkeys[1], &c->order[0], &c->order[1], &c->x_small, &c->pt_count,
                 &c->fz_hdr, &c->fz_bits, &c->fz_pool, &c->hdr_dir[0], &c->hdr_dir[1], &c->hdr_dir[2], &c->pt_off, &c->pt_fill, &c->big_points,
                 &c->cic_vals, &c->cic_keys[0], &c->cic_keys[1], &c->cic_ids[0], &c->cic_ids[1], &c->cic_count, &c->cic_start, &c->cic_vals2};
  for (Buf *b : bufs) b->release();
#ifdef TESSB200_WITH_NCCL
  if (c->comm && ncclw::g.h) ncclw::g.CommDestroy(c->comm);
#endif
  if (c->h_cnt) cudaFreeHost(c->h_cnt);
  if (c->h_sum) cudaFreeHost(c->h_sum);
  if (c->h_max) cudaFreeHost(c->h_max);
  if (c->h_xall) cudaFreeHost(c->h_xall);
  for (auto &ev : c->ev) if (ev) cudaEventDestroy(ev);
  for (auto &ev : c->blk_ev) if (ev) cudaEventDestroy(ev);
  for (auto &ev : c->grp_ev) if (ev) cudaEventDestroy(ev);
  if (c->part_ev) cudaEventDestroy(c->part_ev);
  for (auto &ev : c->h2d_ev) cudaEventDestroy(ev);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

struct Geometry
{
  GridGeom g;
  KeyLayout kl;
  int key_bits;
  std::vector<BlockBox> boxes;      // every block of the decomposition
  std::vector<int> local_of;        // global block index -> index into ctx->blocks, or -1
  std::vector<RowBlock> rblocks;    // this rank's blocks
  unsigned long long row0 = 0, nrows = 0, total_rows = 0;
  long long out_floats = 0;
  int nx_max = 1;
  unsigned long long total_cells = 0;
};

static int check_params(const tessb200_dense_params *p)
{
  if (!p) return fail(TESSB200_EINVAL, "params is NULL");
  if (p->alg != TESSB200_DENSE_TESS && p->alg != TESSB200_DENSE_CIC && p->alg != TESSB200_DENSE_DTFE) return fail(TESSB200_EINVAL, "unknown alg %d", p->alg);
  if (p->alg == TESSB200_DENSE_DTFE && p->project) return fail(TESSB200_EINVAL, "the DTFE mode has no projected variant");
  for (int d = 0; d < 3; d++)
    if (p->glo_num_idx[d] < 2 || p->glo_num_idx[d] > 32767)
      return fail(TESSB200_ELIMIT, "glo_num_idx[%d] = %d outside [2, 32767]", d, p->glo_num_idx[d]);
  if (p->num_given_bounds < 0 || p->num_given_bounds > 3) return fail(TESSB200_EINVAL, "num_given_bounds = %d", p->num_given_bounds);
  if (p->project && !(p->proj_plane[0] == 0.0f && p->proj_plane[1] == 0.0f && p->proj_plane[2] != 0.0f))
    return fail(TESSB200_EINVAL, "projection is supported along z only (the reference asserts the same, src/dense.cpp:1077)");
  if (!(p->eps >= 0.0f)) return fail(TESSB200_EINVAL, "eps must be >= 0");
  return 0;
}

// fills params outputs and every block's grid geometry; `all` = every block of the decomposition
static int make_geometry(tessb200_ctx *c, tessb200_dense_params *p, Geometry *G, const tessb200_block *hb = nullptr, int nhb = 0)
{
  TRY(check_params(p));
  std::vector<LayoutBlock> all;
  if (!c->layout.empty()) all = c->layout;
  else if (hb) {
    // geometry only, straight from the caller's blocks (nothing needs to be uploaded)
    for (int i = 0; i < nhb; i++) {
      LayoutBlock l;
      l.gid = hb[i].gid;
      memcpy(l.bmin, hb[i].bounds_min, 12); memcpy(l.bmax, hb[i].bounds_max, 12);
      l.owner = -1;
      all.push_back(l);
    }
    std::sort(all.begin(), all.end(), [](const LayoutBlock &x, const LayoutBlock &y) { return x.gid < y.gid; });
  } else
    for (BlockRes *b : c->blocks) {
      LayoutBlock l;
      l.gid = b->gid;
      memcpy(l.bmin, b->bmin, 12); memcpy(l.bmax, b->bmax, 12);
      l.owner = c->rank;
      all.push_back(l);
    }
  if (all.empty()) return fail(TESSB200_ESTATE, "no blocks uploaded");
  // DataBounds, src/dense.cpp:1221-1275 (min/max over every block's bounds)
  for (size_t i = 0; i < all.size(); i++)
    for (int d = 0; d < 3; d++) {
      if (i == 0 || all[i].bmin[d] < p->data_mins[d]) p->data_mins[d] = all[i].bmin[d];
      if (i == 0 || all[i].bmax[d] > p->data_maxs[d]) p->data_maxs[d] = all[i].bmax[d];
    }
  grid_step_params(p);
  GridGeom &g = G->g;
  for (int d = 0; d < 3; d++) {
    g.gmin[d] = p->grid_phys_mins[d]; g.step[d] = p->grid_step_size[d];
    g.dmin[d] = p->data_mins[d]; g.dmax[d] = p->data_maxs[d];
    g.dext_eps[d] = (p->data_maxs[d] - p->data_mins[d]) * 2.0f * FLT_EPSILON;
    g.gnum[d] = p->glo_num_idx[d];
    if (!(g.step[d] > 0.0f)) return fail(TESSB200_EINVAL, "grid step %d is not positive (degenerate bounds)", d);
  }
  g.eps = p->eps; g.mass = p->mass; g.project = p->project ? 1 : 0; g.alg = p->alg;
  g.div = p->project ? g.step[0] * g.step[1] : g.step[0] * g.step[1] * g.step[2]; // src/dense.cpp:90-91

  G->boxes.resize(all.size());
  G->local_of.assign(all.size(), -1);
  long long row_base = 0;
  unsigned long long cell_base = 0;
  long long out_off = 0;
  G->rblocks.clear();
  bool seen_local = false, past_local = false;
  for (size_t i = 0; i < all.size(); i++) {
    BlockBox &bx = G->boxes[i];
    block_grid_params(all[i].bmin, all[i].bmax, p, bx.b_lo, bx.b_num);
    phys_box(all[i].bmin, all[i].bmax, p, bx.p_lo, bx.p_hi);
    for (int d = 0; d < 3; d++)
      if (bx.b_num[d] < 1) return fail(TESSB200_EINVAL, "block gid %d owns no grid points along axis %d", all[i].gid, d);
    long long nrows = p->project ? bx.b_num[1] : (long long)bx.b_num[1] * bx.b_num[2];
    bx.row_base = row_base;
    if (all[i].owner == c->rank) {
      if (past_local) return fail(TESSB200_EINVAL, "blocks owned by one rank must be contiguous in gid order");
      int li = -1;
      for (size_t k = 0; k < c->blocks.size(); k++)
        if (c->blocks[k]->gid == all[i].gid) li = (int)k;
      if (li < 0 && hb) { row_base += nrows; continue; }
      if (li < 0) return fail(TESSB200_ESTATE, "block gid %d is owned by this rank but was not uploaded", all[i].gid);
      if (!seen_local) G->row0 = (unsigned long long)row_base;
      seen_local = true;
      G->local_of[i] = li;
      BlockRes *b = c->blocks[li];
      memcpy(b->mn, bx.b_lo, 12); memcpy(b->num, bx.b_num, 12);
      b->nrows = nrows;
      b->npts = nrows * bx.b_num[0];
      b->row_base = row_base;
      b->out_off = out_off;
      b->cell_base = (uint32_t)cell_base;
      RowBlock rb;
      rb.row_base = row_base; rb.nrows = nrows; rb.out_off = out_off; rb.nx = bx.b_num[0]; rb.pad = 0;
      G->rblocks.push_back(rb);
      out_off += b->npts;
      G->nrows += (unsigned long long)nrows;
      if (bx.b_num[0] > G->nx_max) G->nx_max = bx.b_num[0];
      cell_base += (unsigned long long)b->num_orig;
    } else if (seen_local) {
      past_local = true;
    }
    row_base += nrows;
  }
  if (c->layout.empty()) {
    // single rank: cell numbering over the uploaded blocks
  } else {
    // multi rank: cell numbers must be globally ordered by gid; use a fixed 2^26 window per rank
    // (checked against the per-rank cell count) so that no exchange of counts is needed
    if (cell_base > (1ull << 26)) return fail(TESSB200_ELIMIT, "more than 2^26 cells on one rank in a multi-GPU run");
    for (BlockRes *b : c->blocks) b->cell_base += (uint32_t)c->rank << 26;
    cell_base = (unsigned long long)c->nranks << 26;
  }
  G->total_rows = (unsigned long long)row_base;
  G->out_floats = out_off;
  G->total_cells = cell_base;
  G->kl.cell_bits = ceil_log2(cell_base + 1);
  key_z_range(G->boxes.data(), G->boxes.size(), p->project, &G->kl);
  G->key_bits = ceil_log2(G->total_rows + 1) + 1 + G->kl.cell_bits + G->kl.z_bits;
  if (G->key_bits > 64) return fail(TESSB200_ELIMIT, "sort key needs %d bits (> 64): grid rows x cells too large", G->key_bits);
  return 0;
}

// ---- upload ----------------------------------------------------------------------------------------
static int upload_impl(tessb200_ctx *c, int nblocks, const tessb200_block *blocks, bool async)
{
  if (!c) return fail(TESSB200_EINVAL, "ctx is NULL");
  if (nblocks < 1 || !blocks) return fail(TESSB200_EINVAL, "need at least one block");
  if (nblocks > 65535) return fail(TESSB200_ELIMIT, "more than 65535 blocks");
  CU(cudaSetDevice(c->device));
  // async: copies go to the copy stream with one event per block (the one-call path overlaps them
  // with the cell kernels); otherwise they are complete when the call returns
  cudaStream_t cs = async ? c->copy_stream : c->stream;
  CU(cudaEventRecord(c->ev[0], cs));
  while ((int)c->h2d_ev.size() < nblocks) {
    cudaEvent_t e;
    CU(cudaEventCreate(&e));               // timed: the TESSB200_TRACE timeline reads them
    c->h2d_ev.push_back(e);
  }
  if (const char *chk = getenv("TESSB200_CHECK_INPUT"))
    if (chk[0] == '1' || chk[0] == '2')
      for (int i = 0; i < nblocks; i++) TRY(tessb200_check_block(&blocks[i], chk[0] == '2'));
  // reuse device buffers of a previous upload where possible
  std::vector<int> order(nblocks);
  for (int i = 0; i < nblocks; i++) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return blocks[a].gid < blocks[b].gid; });
  for (int i = 1; i < nblocks; i++)
    if (blocks[order[i]].gid == blocks[order[i - 1]].gid) return fail(TESSB200_EINVAL, "duplicate gid %d", blocks[order[i]].gid);
  while ((int)c->blocks.size() > nblocks) {
    BlockRes *b = c->blocks.back();
    b->particles.release(); b->tets.release(); b->v2t.release(); b->cc.release(); b->rho.release(); b->walk.release(); b->hull.release(); b->p4.release();
    delete b;
    c->blocks.pop_back();
  }
  while ((int)c->blocks.size() < nblocks) c->blocks.push_back(new BlockRes);
  // particles and vert_to_tet of every block first (small): the cells' processing order needs only
  // those and is computed while the tets (93 % of the bytes) are still on their way
  for (int k = 0; k < nblocks; k++) {
    const tessb200_block &hb = blocks[order[k]];
    BlockRes *b = c->blocks[k];
    if (hb.num_particles < 0 || hb.num_orig_particles < 0 || hb.num_orig_particles > hb.num_particles || hb.num_tets < 0)
      return fail(TESSB200_EINVAL, "block gid %d: inconsistent counts", hb.gid);
    if ((hb.num_particles && !hb.particles) || (hb.num_tets && !hb.tets)) return fail(TESSB200_EINVAL, "block gid %d: NULL input array", hb.gid);
    b->gid = hb.gid;
    b->num_orig = hb.num_orig_particles; b->num_particles = hb.num_particles; b->num_tets = hb.num_tets;
    memcpy(b->bmin, hb.bounds_min, 12); memcpy(b->bmax, hb.bounds_max, 12);
    for (int d = 0; d < 3; d++)
      if (!(b->bmin[d] <= b->bmax[d])) return fail(TESSB200_EINVAL, "block gid %d: bounds_min > bounds_max on axis %d", hb.gid, d);
    TRY(b->particles.ensure(sizeof(float) * 3 * (size_t)std::max(1, hb.num_particles)));
    TRY(b->tets.ensure(32 * (size_t)std::max(1, hb.num_tets)));
    TRY(b->v2t.ensure(sizeof(int) * (size_t)std::max(1, hb.num_particles)));
    TRY(b->cc.ensure(16 * (size_t)std::max(1, hb.num_tets)));
    if (hb.num_particles) CU(cudaMemcpyAsync(b->particles.p, hb.particles, sizeof(float) * 3 * (size_t)hb.num_particles, cudaMemcpyHostToDevice, cs));
    b->have_v2t = hb.vert_to_tet != nullptr;
    if (b->have_v2t && hb.num_particles)
      CU(cudaMemcpyAsync(b->v2t.p, hb.vert_to_tet, sizeof(int) * (size_t)hb.num_particles, cudaMemcpyHostToDevice, cs));
  }
  CU(cudaEventRecord(c->part_ev, cs));
  for (int k = 0; k < nblocks; k++) {
    const tessb200_block &hb = blocks[order[k]];
    BlockRes *b = c->blocks[k];
    if (hb.num_tets) CU(cudaMemcpyAsync(b->tets.p, hb.tets, 32 * (size_t)hb.num_tets, cudaMemcpyHostToDevice, cs));
    CU(cudaEventRecord(c->h2d_ev[k], cs));
  }
  CU(cudaEventRecord(c->ev[1], cs));
  if (!async) CU(cudaStreamSynchronize(cs));
  c->ran = false;
  return 0;
}

extern "C" int tessb200_dense_upload(tessb200_ctx *c, int nblocks, const tessb200_block *blocks)
{
  return upload_impl(c, nblocks, blocks, false);
}

static DevBlock dev_block(const BlockRes *b)
{
  DevBlock d;
  d.particles = (const float *)b->particles.p;
  d.tets = (const int4 *)b->tets.p;
  d.v2t = (const int *)b->v2t.p;
  d.cc = (const float4 *)b->cc.p;
  d.walk = (const WalkRec *)b->walk.p;
  d.hull = (const unsigned char *)b->hull.p;
  d.num_orig = b->num_orig; d.num_particles = b->num_particles; d.num_tets = b->num_tets;
  d.cell_base = b->cell_base;
  d.order = nullptr;
  d.cta_start = 0;
  d.slot_start = 0;
  return d;
}

static inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

static int prep_block_geometry(tessb200_ctx *c, BlockRes *b, bool want_walk = false, bool want_hull = false)
{
  if (want_walk && b->num_tets) TRY(b->walk.ensure(sizeof(WalkRec) * (size_t)b->num_tets));
  want_hull = want_hull || want_walk;
  if (want_hull) {
    TRY(b->hull.ensure((size_t)std::max(1, b->num_particles)));
    CU(cudaMemsetAsync(b->hull.p, 0, (size_t)std::max(1, b->num_particles), c->stream));
  }
  // vert_to_tet (if not given) and circumcenters for one resident block
  if (!b->have_v2t && b->num_particles) {
    k_fill_i32<<<cdiv(b->num_particles, 256), 256, 0, c->stream>>>((int *)b->v2t.p, b->num_particles, -1);
    COUNT_LAUNCH(c, 1);
    if (b->num_tets) { k_vert_to_tet<<<cdiv(b->num_tets, 256), 256, 0, c->stream>>>((const int4 *)b->tets.p, b->num_tets, (int *)b->v2t.p); COUNT_LAUNCH(c, 1); }
  }
  if (b->num_tets) {
    // 16-byte particle records for the gathers (the packed xyz array stays what every other kernel reads)
    const float4 *p4 = nullptr;
    if (b->num_particles && b->num_tets >= 4096) {
      TRY(b->p4.ensure(sizeof(float4) * (size_t)b->num_particles));
      k_pack_particles<<<cdiv(b->num_particles, 256), 256, 0, c->stream>>>((const float *)b->particles.p, b->num_particles, b->p4.as<float4>());
      COUNT_LAUNCH(c, 1);
      p4 = b->p4.as<float4>();
    }
    k_circumcenters<<<cdiv(b->num_tets, 256), 256, 0, c->stream>>>((const int4 *)b->tets.p, b->num_tets, (const float *)b->particles.p, p4, (float4 *)b->cc.p,
                                                                want_walk ? (WalkRec *)b->walk.p : nullptr, want_hull ? (unsigned char *)b->hull.p : nullptr);
    COUNT_LAUNCH(c, 1);
  }
  CU(cudaGetLastError());
  return 0;
}

// processing order of the cells of local blocks [k0, k1): Morton order of the sites inside each block
// (results do not depend on it).  One radix sort per group: key = block tag | Morton bits.  The
// sorted ids of block k land in order[1] at the block's offset in the rank-wide cell numbering.
static int prep_cell_order(tessb200_ctx *c, int k0, int k1, long long cell_off)
{
  long long cells = 0;
  for (int k = k0; k < k1; k++) cells += c->blocks[k]->num_orig;
  if (cells == 0) return 0;
  const int blk_bits = ceil_log2((unsigned long long)(k1 - k0));
  int morton_bits = 32 - blk_bits;
  if (morton_bits > 30) morton_bits = 30;
  morton_bits -= morton_bits % 3;
  long long off = cell_off;
  for (int k = k0; k < k1; k++) {
    BlockRes *b = c->blocks[k];
    const int n = b->num_orig;
    if (n) {
      float3 bmin = make_float3(b->bmin[0], b->bmin[1], b->bmin[2]);
      float3 inv = make_float3(1.0f / fmaxf(b->bmax[0] - b->bmin[0], 1e-30f), 1.0f / fmaxf(b->bmax[1] - b->bmin[1], 1e-30f),
                               1.0f / fmaxf(b->bmax[2] - b->bmin[2], 1e-30f));
      k_morton_keys<<<cdiv(n, 256), 256, 0, c->stream>>>((const float *)b->particles.p, n, bmin, inv, (uint32_t)(k - k0) << morton_bits, 30 - morton_bits,
                                                       c->mkeys[0].as<uint32_t>() + off, c->order[0].as<uint32_t>() + off);
      COUNT_LAUNCH(c, 1);
    }
    off += n;
  }
  size_t tmp = 0;
  uint32_t *k_in = c->mkeys[0].as<uint32_t>() + cell_off, *k_out = c->mkeys[1].as<uint32_t>() + cell_off;
  uint32_t *v_in = c->order[0].as<uint32_t>() + cell_off, *v_out = c->order[1].as<uint32_t>() + cell_off;
  CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp, k_in, k_out, v_in, v_out, (int)cells, 0, morton_bits + blk_bits, c->stream));
  TRY(c->cub_tmp.ensure(tmp));
  CU(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp, k_in, k_out, v_in, v_out, (int)cells, 0, morton_bits + blk_bits, c->stream));
  return 0;
}

static void trace_mark(tessb200_ctx *c, const char *name)
{
  if (!c->tracing) return;
  if (c->tr_used == c->tr_ev.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    c->tr_ev.push_back(e);
    c->tr_name.push_back(name);
  }
  c->tr_name[c->tr_used] = name;
  cudaEventRecord(c->tr_ev[c->tr_used++], c->stream);
}

static int read_counters(tessb200_ctx *c)
{
  CU(cudaMemcpyAsync(c->h_cnt, c->d_cnt.p, sizeof(Counters), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

#ifdef TESSB200_WITH_NCCL
static int exchange_spans(tessb200_ctx *c, const Geometry &G, int cur, unsigned long long *n_spans);
static void exchange_poison(tessb200_ctx *c);
static int exchange_check(tessb200_ctx *c, bool *redo);
#endif
// A rank that fails before the span exchange still takes part in it (empty, flagged), so that its peers return
// TESSB200_EPEER instead of waiting in a collective forever.
struct ExchangeGuard
{
  tessb200_ctx *c;
  bool armed;
  ~ExchangeGuard()
  {
#ifdef TESSB200_WITH_NCCL
    if (armed) exchange_poison(c);
#endif
  }
};

// ---- run ---------------------------------------------------------------------------------------------
// The stage runs group by group over the local blocks (cell kernels), then once over all span
// records (exchange, sort, deposit).  Resident inputs: one group = all blocks.  One-call path
// (tessb200_dense): one group per block, each waiting for that block's host-to-device copy, so the
// cell kernels of block k overlap the copy of block k+1; the deposit then runs block by block with
// the device-to-host copy of block k overlapping the deposit of block k+1.
struct PipeIO
{
  bool pipelined = false;
  std::vector<cudaEvent_t> *h2d_done = nullptr;   // per local block (ctx->blocks order)
  int nblocks_out = 0;
  tessb200_block *out_blocks = nullptr;           // caller's blocks: density pointers for the streamed download
  float *global_grid = nullptr;
  cudaEvent_t particles_done = nullptr;           // particles (+ vert_to_tet) of every block are on the device
};

static int copy_block_out(tessb200_ctx *c, const tessb200_dense_params &p, BlockRes *b, tessb200_block *ob, float *global_grid, cudaStream_t s)
{
  if (ob) {
    memcpy(ob->block_min_idx, b->mn, 12); memcpy(ob->block_num_idx, b->num, 12);
    ob->num_grid_pts = b->npts;
    if (ob->density) {
      if (ob->density_capacity < b->npts)
        return fail(TESSB200_ECAPACITY, "block gid %d: density_capacity %lld < %lld grid points", b->gid, (long long)ob->density_capacity, b->npts);
      if (b->npts) CU(cudaMemcpyAsync(ob->density, c->out.as<float>() + b->out_off, sizeof(float) * (size_t)b->npts, cudaMemcpyDeviceToHost, s));
    }
  }
  if (global_grid && b->npts) {
    if (p.project) return fail(TESSB200_EINVAL, "global_grid is only assembled for 3-D runs; use tessb200_write_grid for projections");
    const size_t gx = p.glo_num_idx[0], gy = p.glo_num_idx[1];
    cudaMemcpy3DParms cp;
    memset(&cp, 0, sizeof(cp));
    cp.srcPtr = make_cudaPitchedPtr(c->out.as<float>() + b->out_off, sizeof(float) * (size_t)b->num[0], (size_t)b->num[0], (size_t)b->num[1]);
    cp.dstPtr = make_cudaPitchedPtr(global_grid, sizeof(float) * gx, gx, gy);
    cp.dstPos = make_cudaPos(sizeof(float) * (size_t)b->mn[0], (size_t)b->mn[1], (size_t)b->mn[2]);
    cp.extent = make_cudaExtent(sizeof(float) * (size_t)b->num[0], (size_t)b->num[1], (size_t)b->num[2]);
    cp.kind = cudaMemcpyDeviceToHost;
    CU(cudaMemcpy3DAsync(&cp, s));
  }
  return 0;
}

// alg 2 (DTFE, first order): per block, vertex densities from the star volumes, then one thread per
// tet writes the grid points it owns inside the block's sub-grid.  No span records, no exchange: a
// block rasterises its own sub-grid from its own (ghost-padded) tessellation.
static int run_dtfe(tessb200_ctx *c, tessb200_dense_params *p, tessb200_dense_stats *st, const PipeIO &io, const Geometry &G)
{
  cudaStream_t s = c->stream;
  const int nloc = (int)c->blocks.size();
  long long cells = 0, tets = 0;
  for (BlockRes *b : c->blocks) { cells += b->num_orig; tets += b->num_tets; }
  c->launches = 0;
  TRY(c->out.ensure(sizeof(float) * (size_t)std::max<long long>(4, G.out_floats)));
  TRY(c->d_cnt.ensure(sizeof(Counters)));
  Counters *cnt = c->d_cnt.as<Counters>();
  CU(cudaEventRecord(c->ev[2], s));
  CU(cudaMemsetAsync(c->out.p, 0, sizeof(float) * (size_t)G.out_floats, s));
  long long n_slow = 0;
  for (int k = 0; k < nloc; k++) {
    BlockRes *b = c->blocks[k];
    if (io.pipelined && io.h2d_done) CU(cudaStreamWaitEvent(s, (*io.h2d_done)[k], 0));
    TRY(prep_block_geometry(c, b));
    if (b->num_particles && b->num_tets && b->npts) {
      TRY(b->rho.ensure(4 * (size_t)b->num_particles));
      const uint32_t cap = (uint32_t)std::max(1024, b->num_particles / 64);
      TRY(c->overflow.ensure(sizeof(uint2) * (size_t)cap));
      CU(cudaMemsetAsync(c->d_cnt.p, 0, sizeof(Counters), s));
      DevBlock db = dev_block(b);
      k_vertex_density<<<cdiv(b->num_particles, TOPO_THREADS), TOPO_THREADS, TOPO_SMEM, s>>>(db, p->mass, b->rho.as<float>(), c->overflow.as<uint32_t>(),
                                                                                           &cnt->n_overflow, cap);
      COUNT_LAUNCH(c, 1);
      CU(cudaGetLastError());
      TRY(read_counters(c));
      if (c->h_cnt->n_overflow > cap) return fail(TESSB200_ELIMIT, "%u vertices exceed the fast star workspace", c->h_cnt->n_overflow);
      if (c->h_cnt->n_overflow) {
        const int n = (int)c->h_cnt->n_overflow;
        TRY(c->ws_big.ensure(sizeof(int) * (size_t)BIG_STAR_CAP * (size_t)n));
        k_vertex_density_big<<<cdiv((long long)n * 32, 128), 128, 0, s>>>(db, p->mass, b->rho.as<float>(), c->overflow.as<uint32_t>(), n, c->ws_big.as<int>());
        COUNT_LAUNCH(c, 1);
        n_slow += n;
      }
      // seed grid over the block's bounds (about 8 particles per coarse cell), then point location
      int cgn = (int)cbrt((double)b->num_orig / 8.0);
      cgn = std::max(1, std::min(cgn, 256));
      const int3 cg = make_int3(cgn, cgn, cgn);
      const float3 bmin = make_float3(b->bmin[0], b->bmin[1], b->bmin[2]);
      const float3 inv_cell = make_float3((float)cgn / fmaxf(b->bmax[0] - b->bmin[0], 1e-30f), (float)cgn / fmaxf(b->bmax[1] - b->bmin[1], 1e-30f),
                                          (float)cgn / fmaxf(b->bmax[2] - b->bmin[2], 1e-30f));
      const size_t nseed = (size_t)cgn * cgn * cgn + 1;
      TRY(c->hdr_big.ensure(4 * nseed));
      CU(cudaMemsetAsync(c->hdr_big.p, 0xFF, 4 * nseed, s));
      k_dtfe_seed<<<cdiv(b->num_particles, 256), 256, 0, s>>>(db, bmin, inv_cell, cg, c->hdr_big.as<int>());
      const int3 blo = make_int3(b->mn[0], b->mn[1], b->mn[2]), bnum = make_int3(b->num[0], b->num[1], b->num[2]);
      const long long nthreads = (long long)((b->num[0] + DTFE_CHUNK - 1) / DTFE_CHUNK) * b->num[1] * b->num[2];
      k_dtfe_raster<<<cdiv(nthreads, 128), 128, 0, s>>>(db, b->rho.as<float>(), G.g, blo, bnum, bmin, inv_cell, cg, c->hdr_big.as<int>(),
                                                       c->out.as<float>() + b->out_off, &cnt->n_big);
      COUNT_LAUNCH(c, 2);
      CU(cudaGetLastError());
    }
    if (io.pipelined) {
      CU(cudaEventRecord(c->blk_ev[k % 64], s));
      CU(cudaStreamWaitEvent(c->copy_stream, c->blk_ev[k % 64], 0));
      tessb200_block *ob = nullptr;
      for (int j = 0; j < io.nblocks_out; j++) if (io.out_blocks[j].gid == b->gid) ob = &io.out_blocks[j];
      TRY(copy_block_out(c, *p, b, ob, io.global_grid, c->copy_stream));
    }
  }
  CU(cudaEventRecord(c->ev[9], s));
  if (st) {
    memset(st, 0, sizeof(*st));
    const int nb = 592;
    TRY(c->stat_sum.ensure(sizeof(double) * nb));
    TRY(c->stat_max.ensure(sizeof(float) * nb));
    double tot = 0.0;
    float mx = 0.0f;
    if (G.out_floats) {
      k_grid_stats<<<nb, 256, 0, s>>>(c->out.as<float>(), (unsigned long long)G.out_floats, c->stat_sum.as<double>(), c->stat_max.as<float>());
      CU(cudaMemcpyAsync(c->h_sum, c->stat_sum.p, sizeof(double) * nb, cudaMemcpyDeviceToHost, s));
      CU(cudaMemcpyAsync(c->h_max, c->stat_max.p, sizeof(float) * nb, cudaMemcpyDeviceToHost, s));
      CU(cudaStreamSynchronize(s));
      for (int i = 0; i < nb; i++) { tot += c->h_sum[i]; mx = std::max(mx, c->h_max[i]); }
    }
    st->tot_mass = tot * (double)G.g.div;
    st->max_dense = mx;
  }
  CU(cudaStreamSynchronize(s));
  if (io.pipelined) CU(cudaStreamSynchronize(c->copy_stream));
  c->ran = true;
  c->last_params = *p;
  c->out_floats = G.out_floats;
  if (st) {
    st->num_cells = cells;
    st->num_tets = tets;
    st->num_slow_cells = n_slow;
    st->num_kernel_launches = c->launches;
    for (BlockRes *b : c->blocks) st->num_grid_pts += b->npts;
    float m = 0;
    cudaEventElapsedTime(&m, c->ev[2], c->ev[9]);
    st->ms_total_device = m;
    st->ms_cells = m;
  }
  return 0;
}

static int run_impl(tessb200_ctx *c, tessb200_dense_params *p, tessb200_dense_stats *st, const PipeIO &io)
{
  // from here to the span exchange every early return still takes part in the exchange (flagged), see ExchangeGuard
  ExchangeGuard xguard{c, c->nranks > 1 && p && p->alg != TESSB200_DENSE_DTFE};
  if (c->blocks.empty()) return fail(TESSB200_ESTATE, "tessb200_dense_run before tessb200_dense_upload");
  CU(cudaSetDevice(c->device));
  Geometry G;
  TRY(make_geometry(c, p, &G));
  if (p->alg == TESSB200_DENSE_DTFE) return run_dtfe(c, p, st, io, G);
  cudaStream_t s = c->stream;
  c->tracing = io.pipelined && getenv("TESSB200_TRACE") != nullptr;
  c->tr_used = 0;
  const int nloc = (int)c->blocks.size();
  const int nall = (int)G.boxes.size();
  long long cells = 0, tets = 0;
  for (BlockRes *b : c->blocks) { cells += b->num_orig; tets += b->num_tets; }
  c->launches = 0;
  const bool tess = p->alg == TESSB200_DENSE_TESS;
  const bool fused = tess && c->fused;
  c->x_fast_used = false;

  CU(cudaEventRecord(c->ev[2], s));
  // groups of local blocks (indices into c->blocks)
  std::vector<std::pair<int, int> > groups;
  if (io.pipelined) for (int k = 0; k < nloc; k++) groups.push_back(std::make_pair(k, k + 1));
  else groups.push_back(std::make_pair(0, nloc));
  int first_local_all = -1;
  for (int i = 0; i < nall; i++) if (G.local_of[i] == 0) first_local_all = i;   // local blocks are contiguous in `all`
  if (first_local_all < 0) return fail(TESSB200_ESTATE, "no local block in the layout");

  // device descriptors
  if (tess)
    for (int i = 0; i < 2; i++) { TRY(c->mkeys[i].ensure(4 * (size_t)std::max<long long>(1, cells))); TRY(c->order[i].ensure(4 * (size_t)std::max<long long>(1, cells))); }
  std::vector<DevBlock> hblocks(nall);
  memset(hblocks.data(), 0, sizeof(DevBlock) * nall);
  {
    long long order_off = 0;
    for (size_t gi = 0; gi < groups.size(); gi++) {
      uint32_t ctas = 0, slots = 0;
      for (int k = groups[gi].first; k < groups[gi].second; k++) {
        BlockRes *b = c->blocks[k];
        DevBlock &d = hblocks[first_local_all + k];
        if (tess && !fused && b->num_tets) TRY(b->walk.ensure(sizeof(WalkRec) * (size_t)b->num_tets));
        if (tess) TRY(b->hull.ensure((size_t)std::max(1, b->num_particles)));
        d = dev_block(b);
        if (fused) d.walk = nullptr;              // no walk records: the general kernels circulate on tet records + circumcenters
        d.order = c->order[1].as<uint32_t>() + order_off;
        order_off += b->num_orig;
        d.cta_start = ctas;
        ctas += cdiv(b->num_orig, TOPO_THREADS);
        d.slot_start = slots;
        slots += (uint32_t)b->num_orig;
      }
    }
  }
  TRY(c->d_blocks.ensure(sizeof(DevBlock) * nall));
  TRY(c->d_boxes.ensure(sizeof(BlockBox) * nall));
  TRY(c->d_rblocks.ensure(sizeof(RowBlock) * std::max<size_t>(1, G.rblocks.size())));
  TRY(c->d_cnt.ensure(sizeof(Counters)));
  CU(cudaMemcpyAsync(c->d_blocks.p, hblocks.data(), sizeof(DevBlock) * nall, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(c->d_boxes.p, G.boxes.data(), sizeof(BlockBox) * nall, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(c->d_rblocks.p, G.rblocks.data(), sizeof(RowBlock) * G.rblocks.size(), cudaMemcpyHostToDevice, s));
  CU(cudaMemsetAsync(c->d_cnt.p, 0, sizeof(Counters), s));
  Counters *cnt = c->d_cnt.as<Counters>();
  ScanCtx sc;
  sc.boxes = c->d_boxes.as<BlockBox>(); sc.nblocks = nall; sc.kl = G.kl; sc.project = G.g.project;

  unsigned long long span_cap = 0;
  auto ensure_spans = [&](unsigned long long cap) -> int {
    for (int i = 0; i < 2; i++) {
      TRY(c->keys[i].ensure(8 * (size_t)cap));
      TRY(c->data[i].ensure(8 * (size_t)cap));
    }
    span_cap = std::min(c->keys[0].cap, std::min(c->keys[1].cap, std::min(c->data[0].cap, c->data[1].cap))) / 8;
    return 0;
  };

  // scratch that does not depend on the group
  long long n_slow = 0;
  TopoOut to;
  memset(&to, 0, sizeof(to));
  uint32_t cap_ovf = 0;
  uint32_t fast_pairs = 0;
  const int slow_warps = 148 * 4 * 4;          // persistent warps of the general star walk (four 4-warp CTAs per SM)
  if (tess) {
    cap_ovf = fused ? (uint32_t)std::max<long long>(1024, cells) : (uint32_t)std::max<long long>(1024, cells / 64);
    // plane pool / face list: sum of faces <= 4 T, in pairs of faces
    if ((unsigned long long)2 * tets + (unsigned long long)cells + 64 >= 0xffffffffull) return fail(TESSB200_ELIMIT, "plane pool exceeds 2^32 face pairs");
    fast_pairs = (uint32_t)((unsigned long long)2 * tets + (unsigned long long)cells + 64);
    // fused path: the pool holds only the cells the fast kernel hands to the general ones
    if (fused) fast_pairs = (uint32_t)std::max<unsigned long long>(1ull << 20, (unsigned long long)fast_pairs / 4);
    TRY(c->plane_pool.ensure(48 * (size_t)fast_pairs));
    TRY(c->face_list.ensure(32 * (size_t)fast_pairs));
    TRY(c->hdr_small.ensure(sizeof(CellHdr) * (size_t)std::max<long long>(1, cells)));
    TRY(c->hdr_big.ensure(sizeof(CellHdr) * (size_t)std::max<long long>(1, cells)));
    TRY(c->big_bitoff.ensure(8 * (size_t)std::max<long long>(1, cells)));
    TRY(c->overflow.ensure(sizeof(uint2) * (size_t)cap_ovf));
    TRY(c->ws_big.ensure(sizeof(int) * (size_t)(BIG_STAR_CAP + 2 * BIG_NBR_CAP) * (size_t)slow_warps));
    to.small = c->hdr_small.as<CellHdr>(); to.big = c->hdr_big.as<CellHdr>();
    to.big_bit_off = c->big_bitoff.as<unsigned long long>();
    to.overflow = c->overflow.as<uint2>(); to.plane_pool = c->plane_pool.as<float>(); to.faces = c->face_list.as<FaceRef>(); to.cnt = cnt;
    to.cap_pairs = fast_pairs;
    to.cap_small = (uint32_t)cells; to.cap_big = (uint32_t)cells; to.cap_overflow = cap_ovf;
    if (c->direct) {
      for (int q = 0; q < 3; q++) {
        TRY(c->hdr_dir[q].ensure(sizeof(CellHdr) * (size_t)std::max<long long>(1, cells)));
        to.dir[q] = c->hdr_dir[q].as<CellHdr>();
      }
      to.cap_dir = (uint32_t)cells;
    }
    TRY(ensure_spans(std::max<unsigned long long>(1ull << 20, 6ull * (unsigned long long)cells)));
    if (fused) {
      TRY(c->fz_hdr.ensure(sizeof(CellHdr) * (size_t)std::max<long long>(1, cells)));
      TRY(c->fz_bits.ensure(4 * (size_t)FZ_INLINE_WORDS * (size_t)std::max<long long>(1, cells)));
      TRY(c->fz_pool.ensure(4 * (size_t)std::max<long long>(1 << 20, 8 * cells)));
    }
  } else {
    TRY(ensure_spans(std::max<unsigned long long>(1ull << 16, 8ull * (unsigned long long)cells + 1024)));
  }
  // DENSE_CIC, 3-D: the deposits of a block's own points are gathered per grid point (kernels.cuh, "K4 without records for the
  // local deposits"); only the deposits that leave their block are records
  std::vector<CicBlock> cicb(nloc);
  unsigned long long cic_cells = 0;
  bool gather = false;
  if (!tess && !G.g.project && c->cic_gather && cells > 0) {
    unsigned long long part0 = 0;
    for (int k = 0; k < nloc; k++) {
      const BlockBox &bx = G.boxes[first_local_all + k];
      CicBlock &cb = cicb[k];
      for (int d = 0; d < 3; d++) { cb.o[d] = bx.b_lo[d] - 1; cb.d[d] = bx.b_num[d] + 1; }
      cb.cell0 = cic_cells;
      cb.part0 = part0;
      cic_cells += (unsigned long long)cb.d[0] * cb.d[1] * cb.d[2];
      part0 += (unsigned long long)c->blocks[k]->num_orig;
    }
    gather = cic_cells < 0xfffffff0ull && (unsigned long long)cells < 0xfffffff0ull;
    if (gather) {
      TRY(c->cic_vals.ensure(32 * (size_t)cells));
      for (int i = 0; i < 2; i++) { TRY(c->cic_keys[i].ensure(4 * (size_t)cells)); TRY(c->cic_ids[i].ensure(4 * (size_t)cells)); }
      // one entry past the last cell: the scan's last output is where the last list ends
      TRY(c->cic_count.ensure(4 * ((size_t)cic_cells + 1)));
      TRY(c->cic_start.ensure(4 * ((size_t)cic_cells + 1)));
      CU(cudaMemsetAsync(c->cic_count.p, 0, 4 * ((size_t)cic_cells + 1), s));
      TRY(c->cic_vals2.ensure(32 * (size_t)cells));
    }
  }

  CU(cudaEventRecord(c->ev[3], s));
  long long cell_off = 0;
  // one-call path: the processing order of every block was computed while the tets were still on
  // their way (it needs the particles only)
  bool order_ready = false;
  if (tess && io.pipelined && io.particles_done) {
    CU(cudaStreamWaitEvent(s, io.particles_done, 0));
    TRY(prep_cell_order(c, 0, nloc, 0));
    order_ready = true;
  }
  for (size_t gi = 0; gi < groups.size(); gi++) {
    const int k0 = groups[gi].first, k1 = groups[gi].second;
    long long gcells = 0;
    uint32_t gctas = 0;
    for (int k = k0; k < k1; k++) {
      if (io.pipelined && io.h2d_done) CU(cudaStreamWaitEvent(s, (*io.h2d_done)[k], 0));
      gcells += c->blocks[k]->num_orig;
      gctas += cdiv(c->blocks[k]->num_orig, TOPO_THREADS);
    }
    const bool timed = groups.size() == 1;
    if (!tess) {
      if (timed) { CU(cudaEventRecord(c->ev[4], s)); CU(cudaEventRecord(c->ev[5], s)); }
      SpanOut so{c->keys[0].as<uint64_t>(), c->data[0].as<uint64_t>(), span_cap, cnt};
      for (int k = k0; k < k1; k++) {
        const DevBlock &db = hblocks[first_local_all + k];
        if (db.num_orig == 0) continue;
        if (gather)
          k_cic_prepare<<<cdiv(db.num_orig, 256), 256, 0, s>>>(db, first_local_all + k, cicb[k], sc, G.g, so, c->cic_vals.as<float>(), c->cic_keys[0].as<uint32_t>(),
                                                               c->cic_ids[0].as<uint32_t>(), c->cic_count.as<unsigned int>());
        else
          k_cic<<<cdiv(db.num_orig, 256), 256, 0, s>>>(db, first_local_all + k, sc, G.g, so);
        COUNT_LAUNCH(c, 1);
      }
      CU(cudaGetLastError());
      cell_off += gcells;
      continue;
    }
    // K0 / K1 / processing order
    trace_mark(c, "group");
    for (int k = k0; k < k1; k++) TRY(prep_block_geometry(c, c->blocks[k], !fused, true));
    trace_mark(c, "cc");
    if (!order_ready) TRY(prep_cell_order(c, k0, k1, cell_off));
    trace_mark(c, "order");
    cell_off += gcells;
    if (timed) CU(cudaEventRecord(c->ev[4], s));
    // K3a: star BFS, neighbours, faces, scan of the cells the fast kernels can hold.  No host
    // read-back in here: faces and scan take their ranges from the device-side counters, so the
    // launches of a group (and of the next group) queue up behind one another.
    if (gctas) {
      const size_t n_slots = (size_t)gctas * TOPO_THREADS;
      long long gtets = 0;
      for (int k = k0; k < k1; k++) gtets += c->blocks[k]->num_tets;
      if (fused) {
        // one kernel per cell: star walk, faces, planes and inside bits stay in shared memory (fused.cuh); the scan-line
        // walk + span records follow with one lane per cell.  Cells the fast path cannot hold land in the overflow list.
        const long long goff = cell_off - gcells;
        FusedOut fo;
        fo.hdr = c->fz_hdr.as<CellHdr>() + goff;
        fo.bits_inline = c->fz_bits.as<uint32_t>() + (size_t)goff * FZ_INLINE_WORDS;
        fo.bits_pool = c->fz_pool.as<uint32_t>();
        fo.pool_words = c->fz_pool.cap / 4;
        fo.pool_cursor = &cnt->pool_cursor;
        fo.overflow = c->overflow.as<uint2>();
        fo.cap_overflow = cap_ovf;
        fo.cnt = cnt;
        const unsigned pairs = cdiv(gcells, FZ_CPW);
        k_cell_fused<<<std::min<unsigned>(cdiv(pairs, FZ_WARPS), (unsigned)c->fz_ctas), FZ_THREADS, FZ_SMEM, s>>>(c->d_blocks.as<DevBlock>(), first_local_all + k0,
                                                                                                            first_local_all + k1, (uint32_t)gcells, G.g, fo);
        if (timed) CU(cudaEventRecord(c->ev[11], s));
        trace_mark(c, "fused");
        SpanOut so2{c->keys[0].as<uint64_t>(), c->data[0].as<uint64_t>(), span_cap, cnt};
        k_cell_emit<<<cdiv(gcells, EMIT_WARPS * 32), EMIT_THREADS, EMIT_SMEM, s>>>(fo.hdr, (uint32_t)gcells, fo.bits_inline, fo.bits_pool, c->d_blocks.as<DevBlock>(),
                                                                              sc, G.g, so2);
        if (timed) CU(cudaEventRecord(c->ev[12], s));
        trace_mark(c, "emit");
      } else {
      TRY(c->pre_hdr.ensure(sizeof(CellHdr) * n_slots));
      TRY(c->cand.ensure(sizeof(int2) * n_slots * TOPO_CAND_CAP));
      k_cell_bfs<<<gctas, TOPO_THREADS, BFS_SMEM, s>>>(c->d_blocks.as<DevBlock>(), first_local_all + k0, first_local_all + k1, G.g, to,
                                                        c->pre_hdr.as<CellHdr>(), c->cand.as<int2>());
      if (timed) CU(cudaEventRecord(c->ev[11], s));
      trace_mark(c, "bfs");
      k_cell_nbrs<<<gctas, TOPO_THREADS, NBRS_SMEM, s>>>(c->d_blocks.as<DevBlock>(), to, c->pre_hdr.as<CellHdr>(), c->cand.as<int2>());
      if (timed) CU(cudaEventRecord(c->ev[12], s));
      trace_mark(c, "nbrs");
      }
      // stars that did not fit the fast workspace: general BFS, persistent warps, the list range is read on
      // the device; it appends to the same lists, so the faces and scan launches below cover its cells too
      k_cell_bfs_big<<<slow_warps / 4, 128, 0, s>>>(c->d_blocks.as<DevBlock>(), G.g, to, c->overflow.as<uint2>(), c->ws_big.as<int>());
      if (timed) CU(cudaEventRecord(c->ev[13], s));
      trace_mark(c, "big");
      if (!io.pipelined) {
        // inputs resident, one group: nothing to overlap, so read the face count and launch exactly
        TRY(read_counters(c));
        if (timed) CU(cudaEventRecord(c->ev[13], s));
        const size_t f1 = (size_t)std::min(c->h_cnt->plane_cursor, fast_pairs) * 2;
        if (f1) k_cell_faces<<<cdiv((long long)f1, 256), 256, 0, s>>>(c->face_list.as<FaceRef>(), 0, f1, c->d_blocks.as<DevBlock>(), c->plane_pool.as<float>(), cnt);
      } else {
        // one-call path: no host read-back between the launches of a group, the kernel takes its range from the
        // device-side counters; persistent CTAs (one resident wave) stride over it whatever its length
        long long gparts = 0;
        for (int k = k0; k < k1; k++) gparts += c->blocks[k]->num_particles;
        const long long face_est = 2 * gtets + 2 * gparts + gcells + 256;     // faces ~ 2 T + 2 P (Euler)
        static int faces_ctas = 0;
        if (!faces_ctas) {
          int per_sm = 0, sms = 0;
          CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cell_faces_dev, 256, 0));
          CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
          faces_ctas = std::max(1, per_sm * sms);
        }
        k_cell_faces_dev<<<std::min<unsigned>(cdiv(face_est, 256), (unsigned)faces_ctas), 256, 0, s>>>(c->face_list.as<FaceRef>(), c->d_blocks.as<DevBlock>(),
                                                                                                      c->plane_pool.as<float>(), cnt, &cnt->pairs_done,
                                                                                                      &cnt->plane_cursor, fast_pairs);
      }
      if (timed) CU(cudaEventRecord(c->ev[5], s));
      trace_mark(c, "faces");
      SpanOut so{c->keys[0].as<uint64_t>(), c->data[0].as<uint64_t>(), span_cap, cnt};
      k_cell_scan<<<cdiv(gcells, SCAN_WARPS * 32), SCAN_THREADS, SCAN_SMEM, s>>>(c->hdr_small.as<CellHdr>(), 0, c->plane_pool.as<float>(),
                                                                              c->d_blocks.as<DevBlock>(), sc, G.g, so, &cnt->small_done,
                                                                              &cnt->n_small, to.cap_small);
      if (timed) CU(cudaEventRecord(c->ev[16], s));
      if (to.cap_dir) {
        // small index boxes: one thread per cell, planes applied to the whole box as they are produced (no plane storage)
        const unsigned up = cdiv(gcells, DIRECT_THREADS);
        const bool exact = !io.pipelined;      // resident runs read the counters before the faces launch
        const unsigned g2 = exact ? cdiv(c->h_cnt->n_dir[0].v, DIRECT_THREADS) : up, g3 = exact ? cdiv(c->h_cnt->n_dir[1].v, DIRECT_THREADS) : up,
                       g4 = exact ? cdiv(c->h_cnt->n_dir[2].v, DIRECT_THREADS) : up;
        if (g2) k_cell_direct<2><<<g2, DIRECT_THREADS, 0, s>>>(to.dir[0], c->face_list.as<FaceRef>(), c->d_blocks.as<DevBlock>(), sc, G.g, so,
                                                                       &cnt->dir_done[0], &cnt->n_dir[0].v, to.cap_dir, 0u);
        if (g3) k_cell_direct<3><<<g3, DIRECT_THREADS, 0, s>>>(to.dir[1], c->face_list.as<FaceRef>(), c->d_blocks.as<DevBlock>(), sc, G.g, so,
                                                                       &cnt->dir_done[1], &cnt->n_dir[1].v, to.cap_dir, 0u);
        if (g4) k_cell_direct<4><<<g4, DIRECT_THREADS, 0, s>>>(to.dir[2], c->face_list.as<FaceRef>(), c->d_blocks.as<DevBlock>(), sc, G.g, so,
                                                                       &cnt->dir_done[2], &cnt->n_dir[2].v, to.cap_dir, 0u);
        COUNT_LAUNCH(c, (g2 ? 1 : 0) + (g3 ? 1 : 0) + (g4 ? 1 : 0));
      }
      if (timed) CU(cudaEventRecord(c->ev[17], s));
      k_advance<<<1, 1, 0, s>>>(cnt, to.cap_small);
      COUNT_LAUNCH(c, 6);
    } else {
      if (timed) { CU(cudaEventRecord(c->ev[11], s)); CU(cudaEventRecord(c->ev[12], s)); CU(cudaEventRecord(c->ev[13], s)); CU(cudaEventRecord(c->ev[5], s));
                   CU(cudaEventRecord(c->ev[16], s)); CU(cudaEventRecord(c->ev[17], s)); }
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->grp_ev[gi % 64], s));
    trace_mark(c, "scan");
  }
  CU(cudaEventRecord(c->ev[15], s));

  // After the one host read-back of the stage: the cells whose index box or face count exceeds the
  // warp-autonomous scan (one CTA per cell).
  unsigned done_small = 0, done_big = 0;
  auto scan_big_list = [&](const SpanOut &so) -> int {
    TRY(c->bits_big.ensure((size_t)(c->h_cnt->big_bits / 8) + 64));
    k_cell_scan_big<<<done_big, 128, 0, s>>>(to.big, to.big_bit_off, (int)done_big, c->plane_pool.as<float>(), c->bits_big.as<uint32_t>(), 0ull,
                                             c->d_blocks.as<DevBlock>(), sc, G.g, so);
    COUNT_LAUNCH(c, 1);
    return 0;
  };
  if (tess && cells > 0) {
    SpanOut so{c->keys[0].as<uint64_t>(), c->data[0].as<uint64_t>(), span_cap, cnt};
    // resident runs read the counters before the faces launch, and no header is appended after that
    if (io.pipelined) TRY(read_counters(c));
    const Counters &h = *c->h_cnt;
    if (h.n_overflow > cap_ovf) return fail(TESSB200_ELIMIT, "%u cells exceed the fast star workspace (capacity %u)", h.n_overflow, cap_ovf);
    if (h.plane_cursor > fast_pairs) return fail(TESSB200_ELIMIT, "face pool overflow: %u pairs (capacity %u)", h.plane_cursor, fast_pairs);
    done_small = std::min(h.n_small, to.cap_small);
    done_big = std::min(h.n_big, to.cap_big);
    n_slow += h.n_overflow + done_big;
    if (done_big) TRY(scan_big_list(so));
    CU(cudaGetLastError());
  }
  // span count (and the rare regrow: the span buffer was too small -> redo the scans of every group)
  TRY(read_counters(c));
  unsigned long long n_spans = c->h_cnt->n_spans;
  if (n_spans > span_cap) {
    if (!tess) return fail(TESSB200_ECAPACITY, "span buffer overflow in CIC");
    TRY(ensure_spans(n_spans + n_spans / 16 + 1024));
    Counters z = *c->h_cnt;
    z.n_spans = 0; z.n_deposit = 0; z.n_cic_fallback = 0;
    CU(cudaMemcpyAsync(c->d_cnt.p, &z, sizeof(Counters), cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s));
    SpanOut so{c->keys[0].as<uint64_t>(), c->data[0].as<uint64_t>(), span_cap, cnt};
    if (done_small)
      k_cell_scan<<<cdiv(done_small, SCAN_WARPS * 32), SCAN_THREADS, SCAN_SMEM, s>>>(to.small, done_small, c->plane_pool.as<float>(), c->d_blocks.as<DevBlock>(), sc, G.g, so,
                                                                                  nullptr, nullptr, 0);
    if (done_big) TRY(scan_big_list(so));
    if (to.cap_dir) {
      const unsigned n2 = std::min(z.n_dir[0].v, to.cap_dir), n3 = std::min(z.n_dir[1].v, to.cap_dir), n4 = std::min(z.n_dir[2].v, to.cap_dir);
      if (n2) k_cell_direct<2><<<cdiv(n2, DIRECT_THREADS), DIRECT_THREADS, 0, s>>>(to.dir[0], c->face_list.as<FaceRef>(), c->d_blocks.as<DevBlock>(), sc, G.g, so, nullptr, nullptr, 0u, n2);
      if (n3) k_cell_direct<3><<<cdiv(n3, DIRECT_THREADS), DIRECT_THREADS, 0, s>>>(to.dir[1], c->face_list.as<FaceRef>(), c->d_blocks.as<DevBlock>(), sc, G.g, so, nullptr, nullptr, 0u, n3);
      if (n4) k_cell_direct<4><<<cdiv(n4, DIRECT_THREADS), DIRECT_THREADS, 0, s>>>(to.dir[2], c->face_list.as<FaceRef>(), c->d_blocks.as<DevBlock>(), sc, G.g, so, nullptr, nullptr, 0u, n4);
    }
    COUNT_LAUNCH(c, 2);
    CU(cudaGetLastError());
    TRY(read_counters(c));
    n_spans = c->h_cnt->n_spans;
    if (n_spans > span_cap) return fail(TESSB200_ECAPACITY, "span buffer overflow after regrow");
  }
  CU(cudaEventRecord(c->ev[6], s));

  int cur = 0;
  long long n_shared_stat = -1;     // deposits that met another one on their grid point (-1: the full-sort path ran)
#ifdef TESSB200_WITH_NCCL
  xguard.armed = false;
  if (c->nranks > 1) TRY(exchange_spans(c, G, cur, &n_spans));
#endif
  CU(cudaEventRecord(c->ev[7], s));

  TRY(c->row_start.ensure(8 * (size_t)(G.nrows + 2)));
  TRY(c->out.ensure(sizeof(float) * (size_t)std::max<long long>(4, G.out_floats)));
  if (n_spans > 0x7fffffffull) return fail(TESSB200_ELIMIT, "%llu span records exceed the sorter's 2^31 limit", n_spans);
  // sort by (row, remote, cell, z) the records in (keys, data)[cur]
  auto sort_records = [&](unsigned long long n) -> int {
    if (!n) return 0;
    size_t tmp_bytes = 0;
    cub::DoubleBuffer<uint64_t> dk(c->keys[cur].as<uint64_t>(), c->keys[cur ^ 1].as<uint64_t>());
    cub::DoubleBuffer<uint64_t> dv(c->data[cur].as<uint64_t>(), c->data[cur ^ 1].as<uint64_t>());
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, (int)n, 0, G.key_bits, s));
    TRY(c->cub_tmp.ensure(tmp_bytes));
    CU(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp_bytes, dk, dv, (int)n, 0, G.key_bits, s));
    if (dk.Current() != c->keys[cur].as<uint64_t>()) cur ^= 1;
    if (dv.Current() != c->data[cur].as<uint64_t>()) return fail(TESSB200_ECUDA, "sorter returned mismatched buffers");
    return 0;
  };
  // one row buffer per warp; long rows (up to 32767 points) leave room for fewer warps per CTA
  int rw = ROWS_WARPS;
  while (rw > 1 && sizeof(float) * (size_t)rw * (size_t)G.nx_max > 160 * 1024) rw--;
  const size_t rows_smem = sizeof(float) * (size_t)rw * (size_t)G.nx_max;
  if (rows_smem > 48 * 1024) CU(cudaFuncSetAttribute(k_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rows_smem));
  auto copy_out_all = [&]() -> int {
    for (int k = 0; k < nloc; k++) {
      BlockRes *b = c->blocks[k];
      CU(cudaEventRecord(c->blk_ev[k % 64], s));
      CU(cudaStreamWaitEvent(c->copy_stream, c->blk_ev[k % 64], 0));
      tessb200_block *ob = nullptr;
      for (int j = 0; j < io.nblocks_out; j++) if (io.out_blocks[j].gid == b->gid) ob = &io.out_blocks[j];
      TRY(copy_block_out(c, *p, b, ob, io.global_grid, c->copy_stream));
    }
    return 0;
  };

  // 3-D runs: only the deposits that meet on a grid point need the reference's order (kernels.cuh, "K3b without the big sort")
  bool placed = false;
  if (gather && G.out_floats) {
    // DENSE_CIC: per grid point, the block's own particles in particle order (k_cic_gather writes every point of the
    // sub-grids once), then the records that crossed block boundaries on top, in order (sort + k_rows, sparse form)
    {
      size_t tmp = 0;
      CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp, c->cic_count.as<unsigned int>(), c->cic_start.as<unsigned int>(), cic_cells + 1, s));
      size_t tmp2 = 0;
      const int kbits = ceil_log2(cic_cells + 2);
      CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp2, c->cic_keys[0].as<uint32_t>(), c->cic_keys[1].as<uint32_t>(), c->cic_ids[0].as<uint32_t>(),
                                         c->cic_ids[1].as<uint32_t>(), (int)cells, 0, kbits, s));
      TRY(c->cub_tmp.ensure(std::max(tmp, tmp2)));
      tmp = tmp2 = c->cub_tmp.cap;
      CU(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tmp, c->cic_count.as<unsigned int>(), c->cic_start.as<unsigned int>(), cic_cells + 1, s));
      CU(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp2, c->cic_keys[0].as<uint32_t>(), c->cic_keys[1].as<uint32_t>(), c->cic_ids[0].as<uint32_t>(),
                                         c->cic_ids[1].as<uint32_t>(), (int)cells, 0, kbits, s));
    }
    k_cic_permute<<<cdiv(cells, 256), 256, 0, s>>>(c->cic_ids[1].as<uint32_t>(), (unsigned long long)cells, c->cic_vals.as<float4>(), c->cic_vals2.as<float4>());
    COUNT_LAUNCH(c, 1);
    for (int k = 0; k < nloc; k++) {
      BlockRes *b = c->blocks[k];
      if (!b->npts) continue;
      const BlockBox &gb = G.boxes[first_local_all + k];
      const dim3 cgrid(cdiv(gb.b_num[0], CIC_TILE_X), cdiv(gb.b_num[1], CIC_TILE_Y), cdiv(gb.b_num[2], CIC_TILE_Z));
      if (cgrid.y > 65535u || cgrid.z > 65535u)
        return fail(TESSB200_EINVAL, "DENSE_CIC: a block's sub-grid exceeds 131070 points in y or z");
      k_cic_gather<<<cgrid, CIC_THREADS, 0, s>>>(cicb[k], gb, c->cic_start.as<unsigned int>(), c->cic_ids[1].as<uint32_t>(), c->cic_vals2.as<float>(),
                                                 c->out.as<float>() + b->out_off);
      COUNT_LAUNCH(c, 1);
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->ev[8], s));      // ms_sort (7 -> 8): scan + particle sort + gather; ms_deposit: the records that crossed block boundaries
    TRY(sort_records(n_spans));
    if (n_spans) {
      k_row_starts<<<cdiv((long long)n_spans + 1, 256), 256, 0, s>>>(c->keys[cur].as<uint64_t>(), n_spans, G.kl, G.row0, G.nrows, c->row_start.as<unsigned long long>());
      k_rows<<<cdiv((long long)G.nrows, rw), rw * 32, rows_smem, s>>>(c->data[cur].as<uint64_t>(), c->row_start.as<unsigned long long>(), G.row0, 0ull, G.nrows,
                                                                    c->d_rblocks.as<RowBlock>(), (int)G.rblocks.size(), G.g.div, G.nx_max, c->out.as<float>(), 1);
      COUNT_LAUNCH(c, 2);
    }
    placed = true;
    n_shared_stat = (long long)n_spans;
    if (io.pipelined) TRY(copy_out_all());
  } else if (!G.g.project && n_spans && G.out_floats && c->segments && G.kl.cell_bits <= 31 && G.out_floats < 0xffffffffll) {
    // every shared point gets a segment of 8-byte records (offsets = scan of the counts); a point's few records are
    // sorted where they are applied: no global sort, no host read-back before the deposit
    const unsigned long long seg_cap = span_cap;                    // the second half of the double buffer
    const size_t npts = (size_t)G.out_floats;
    const unsigned int big_cap = (unsigned int)std::min<unsigned long long>(0x7fffffffull, std::max<unsigned long long>(1ull << 16, span_cap / 8));
    TRY(c->pt_count.ensure(4 * npts));
    TRY(c->pt_off.ensure(4 * npts));
    TRY(c->pt_fill.ensure(4 * npts));
    TRY(c->big_points.ensure(4 * (size_t)big_cap));
    CU(cudaMemsetAsync(c->pt_count.p, 0, 4 * npts, s));
    CU(cudaMemsetAsync(c->pt_fill.p, 0, 4 * npts, s));
    CU(cudaMemsetAsync(c->out.p, 0, sizeof(float) * npts, s));
    CU(cudaMemsetAsync(&cnt->n_shared, 0, sizeof(unsigned long long), s));
    CU(cudaMemsetAsync(&cnt->n_big_points, 0, 2 * sizeof(unsigned int), s));      // n_big_points, dep_flags
    const unsigned grid = cdiv((long long)n_spans, 256);
    k_span_count<<<grid, 256, 0, s>>>(c->keys[cur].as<uint64_t>(), c->data[cur].as<uint64_t>(), n_spans, G.kl, G.row0, G.nrows, c->d_rblocks.as<RowBlock>(),
                                      (int)G.rblocks.size(), c->pt_count.as<unsigned int>());
    {
      auto it = thrust::make_transform_iterator((const unsigned int *)c->pt_count.as<unsigned int>(), SharedCount());
      size_t tmp = 0;
      CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp, it, c->pt_off.as<unsigned int>(), (int)npts, s));
      TRY(c->cub_tmp.ensure(tmp));
      CU(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tmp, it, c->pt_off.as<unsigned int>(), (int)npts, s));
    }
    k_span_place2<<<grid, 256, 0, s>>>(c->keys[cur].as<uint64_t>(), c->data[cur].as<uint64_t>(), n_spans, G.kl, G.row0, G.nrows, c->d_rblocks.as<RowBlock>(),
                                       (int)G.rblocks.size(), c->pt_count.as<unsigned int>(), c->pt_off.as<unsigned int>(), c->pt_fill.as<unsigned int>(), G.g.div,
                                       c->out.as<float>(), c->keys[cur ^ 1].as<uint64_t>(), seg_cap, cnt);
    CU(cudaEventRecord(c->ev[8], s));      // ms_sort (7 -> 8): count + scan + place; ms_deposit: the ordered accumulation on the shared points
    k_point_apply<<<cdiv((long long)npts, 256), 256, 0, s>>>(c->pt_count.as<unsigned int>(), c->pt_off.as<unsigned int>(), c->keys[cur ^ 1].as<uint64_t>(), seg_cap,
                                                            (unsigned long long)npts, G.g.div, p->alg == TESSB200_DENSE_CIC ? 1 : 0, c->out.as<float>(),
                                                            c->big_points.as<unsigned int>(), big_cap, cnt);
    k_point_apply_big<<<148 * 8, POINT_BIG_WARPS * 32, POINT_BIG_SMEM, s>>>(c->pt_count.as<unsigned int>(), c->pt_off.as<unsigned int>(), c->keys[cur ^ 1].as<uint64_t>(),
                                                                          seg_cap, G.g.div, p->alg == TESSB200_DENSE_CIC ? 1 : 0, c->out.as<float>(),
                                                                          c->big_points.as<unsigned int>(), big_cap, cnt);
    COUNT_LAUNCH(c, 4);
    CU(cudaGetLastError());
    TRY(read_counters(c));
    if (!c->h_cnt->dep_flags && c->h_cnt->n_big_points <= big_cap) {
      placed = true;
      n_shared_stat = (long long)c->h_cnt->n_shared;
      if (io.pipelined) TRY(copy_out_all());
    }
    // else: more shared deposits than the buffer holds (a grid far coarser than the cells): the full sort below
  } else if (!G.g.project && n_spans && G.out_floats) {
    // round 1's form of the same idea: the shared deposits as one-point records through the radix sort and k_rows
    const unsigned long long shared_cap = span_cap;                 // the second half of the double buffer
    TRY(c->pt_count.ensure(sizeof(unsigned int) * (size_t)G.out_floats));
    CU(cudaMemsetAsync(c->pt_count.p, 0, sizeof(unsigned int) * (size_t)G.out_floats, s));
    CU(cudaMemsetAsync(c->out.p, 0, sizeof(float) * (size_t)G.out_floats, s));
    CU(cudaMemsetAsync(&cnt->n_shared, 0, sizeof(unsigned long long), s));
    const unsigned grid = cdiv((long long)n_spans, 256);
    k_span_count<<<grid, 256, 0, s>>>(c->keys[cur].as<uint64_t>(), c->data[cur].as<uint64_t>(), n_spans, G.kl, G.row0, G.nrows, c->d_rblocks.as<RowBlock>(),
                                      (int)G.rblocks.size(), c->pt_count.as<unsigned int>());
    k_span_place<<<grid, SPAN_PLACE_THREADS, 0, s>>>(c->keys[cur].as<uint64_t>(), c->data[cur].as<uint64_t>(), n_spans, G.kl, G.row0, G.nrows, c->d_rblocks.as<RowBlock>(),
                                      (int)G.rblocks.size(), c->pt_count.as<unsigned int>(), G.g.div, c->out.as<float>(), c->keys[cur ^ 1].as<uint64_t>(),
                                      c->data[cur ^ 1].as<uint64_t>(), shared_cap, &cnt->n_shared);
    COUNT_LAUNCH(c, 2);
    CU(cudaGetLastError());
    TRY(read_counters(c));
    const unsigned long long n_shared = c->h_cnt->n_shared;
    if (n_shared <= shared_cap && n_shared <= 0x7fffffffull) {
      placed = true;
      CU(cudaEventRecord(c->ev[8], s));      // ms_sort (7 -> 8) is the count + place pass here; the sort of the few shared records is in ms_deposit
      cur ^= 1;                               // the shared one-point records are the list now
      TRY(sort_records(n_shared));
      if (n_shared) {
        k_row_starts<<<cdiv((long long)n_shared + 1, 256), 256, 0, s>>>(c->keys[cur].as<uint64_t>(), n_shared, G.kl, G.row0, G.nrows, c->row_start.as<unsigned long long>());
        k_rows<<<cdiv((long long)G.nrows, rw), rw * 32, rows_smem, s>>>(c->data[cur].as<uint64_t>(), c->row_start.as<unsigned long long>(), G.row0, 0ull, G.nrows,
                                                                      c->d_rblocks.as<RowBlock>(), (int)G.rblocks.size(), G.g.div, G.nx_max, c->out.as<float>(), 1);
        COUNT_LAUNCH(c, 2);
      }
      n_shared_stat = (long long)n_shared;
      if (io.pipelined) TRY(copy_out_all());
    }
    // else: more shared deposits than the buffer holds (a grid far coarser than the cells): the full sort below
  }
  if (!placed) {
    TRY(sort_records(n_spans));
    CU(cudaEventRecord(c->ev[8], s));
    // deposit: every grid point written exactly once
    k_row_starts<<<cdiv((long long)n_spans + 1, 256), 256, 0, s>>>(c->keys[cur].as<uint64_t>(), n_spans, G.kl, G.row0, G.nrows,
                                                                   c->row_start.as<unsigned long long>());
    COUNT_LAUNCH(c, 1);
    if (!io.pipelined) {
      k_rows<<<cdiv((long long)G.nrows, rw), rw * 32, rows_smem, s>>>(c->data[cur].as<uint64_t>(), c->row_start.as<unsigned long long>(), G.row0, 0ull,
                                                                    G.nrows, c->d_rblocks.as<RowBlock>(), (int)G.rblocks.size(), G.g.div,
                                                                    G.nx_max, c->out.as<float>(), 0);
      COUNT_LAUNCH(c, 1);
    } else {
      // block by block: the device-to-host copy of block k (copy stream) overlaps the deposit of block k+1
      for (int k = 0; k < nloc; k++) {
        BlockRes *b = c->blocks[k];
        if (b->nrows) {
          k_rows<<<cdiv(b->nrows, rw), rw * 32, rows_smem, s>>>(c->data[cur].as<uint64_t>(), c->row_start.as<unsigned long long>(), G.row0,
                                                              (unsigned long long)(b->row_base - (long long)G.row0), (unsigned long long)b->nrows,
                                                              c->d_rblocks.as<RowBlock>(), (int)G.rblocks.size(), G.g.div, G.nx_max, c->out.as<float>(), 0);
          COUNT_LAUNCH(c, 1);
        }
        CU(cudaEventRecord(c->blk_ev[k % 64], s));
        CU(cudaStreamWaitEvent(c->copy_stream, c->blk_ev[k % 64], 0));
        tessb200_block *ob = nullptr;
        for (int j = 0; j < io.nblocks_out; j++) if (io.out_blocks[j].gid == b->gid) ob = &io.out_blocks[j];
        TRY(copy_block_out(c, *p, b, ob, io.global_grid, c->copy_stream));
      }
    }
  }
  CU(cudaGetLastError());
  CU(cudaEventRecord(c->ev[9], s));
  if (io.pipelined) CU(cudaEventRecord(c->ev[14], c->copy_stream));

  if (st) {
    // dense_stats (src/dense.cpp:1284-1333): max density and total mass, from the final grid
    const int nb = 592;  // 4 CTAs per SM
    TRY(c->stat_sum.ensure(sizeof(double) * nb));
    TRY(c->stat_max.ensure(sizeof(float) * nb));
    double tot = 0.0;
    float mx = 0.0f;
    if (G.out_floats) {
      k_grid_stats<<<nb, 256, 0, s>>>(c->out.as<float>(), (unsigned long long)G.out_floats, c->stat_sum.as<double>(), c->stat_max.as<float>());
      CU(cudaMemcpyAsync(c->h_sum, c->stat_sum.p, sizeof(double) * nb, cudaMemcpyDeviceToHost, s));
      CU(cudaMemcpyAsync(c->h_max, c->stat_max.p, sizeof(float) * nb, cudaMemcpyDeviceToHost, s));
      CU(cudaStreamSynchronize(s));
      for (int i = 0; i < nb; i++) { tot += c->h_sum[i]; mx = std::max(mx, c->h_max[i]); }
    }
    st->tot_mass = tot * (double)G.g.div;
    st->max_dense = mx;
  }
  CU(cudaEventRecord(c->ev[10], s));
  CU(cudaStreamSynchronize(s));
  if (io.pipelined) CU(cudaStreamSynchronize(c->copy_stream));

#ifdef TESSB200_WITH_NCCL
  if (c->nranks > 1) {
    // fixed-size exchange: the counts every rank really had arrive with the same all-gather; a segment that was too
    // small (or a peer that failed) is seen by every rank alike, and the run is redone with exact sizes
    bool redo = false;
    TRY(exchange_check(c, &redo));
    if (redo) return run_impl(c, p, st, io);
  }
#endif
  c->ran = true;
  c->last_params = *p;
  c->out_floats = G.out_floats;
  if (io.pipelined && getenv("TESSB200_TRACE")) {
    // timeline of the one-call path, ms since the first host-to-device copy was enqueued
    auto at = [&](cudaEvent_t e) { float m = 0; cudaEventElapsedTime(&m, c->ev[0], e); return m; };
    fprintf(stderr, "[tessb200 trace] h2d done:");
    for (int k = 0; k < nloc; k++) fprintf(stderr, " %.2f", at(c->h2d_ev[k]));
    fprintf(stderr, " | run start %.2f groups:", at(c->ev[2]));
    for (size_t gi = 0; gi < groups.size() && gi < 64; gi++) fprintf(stderr, " %.2f", at(c->grp_ev[gi]));
    fprintf(stderr, " | cells done %.2f exchange %.2f sort %.2f rows:", at(c->ev[6]), at(c->ev[7]), at(c->ev[8]));
    for (int k = 0; k < nloc && k < 64; k++) fprintf(stderr, " %.2f", at(c->blk_ev[k]));
    fprintf(stderr, " | d2h done %.2f end %.2f\n", at(c->ev[14]), at(c->ev[10]));
    fprintf(stderr, "[tessb200 trace] marks:");
    for (size_t i = 0; i < c->tr_used; i++) fprintf(stderr, " %s %.3f", c->tr_name[i], at(c->tr_ev[i]));
    fprintf(stderr, "\n");
  }
  if (st) {
    st->num_cells = cells;
    st->num_no_tet = (int64_t)c->h_cnt->n_no_tet;
    st->num_incomplete = (int64_t)c->h_cnt->n_incomplete;
    st->num_outside = (int64_t)(c->h_cnt->n_outside + c->h_cnt->n_bad);
    st->num_deposit_cells = (int64_t)c->h_cnt->n_deposit;
    st->num_cic_fallback = (int64_t)c->h_cnt->n_cic_fallback;
    st->num_slow_cells = n_slow;
    st->num_spans = (int64_t)n_spans;
    st->num_tets = tets;
    st->num_kernel_launches = c->launches;
    st->num_grid_pts = 0;
    for (BlockRes *b : c->blocks) st->num_grid_pts += b->npts;
    auto ms = [&](int a, int b) { float m = 0; cudaEventElapsedTime(&m, c->ev[a], c->ev[b]); return m; };
    const bool one = groups.size() == 1;
    st->ms_upload = io.pipelined ? ms(0, 1) : 0;
    st->ms_circumcenters = one ? ms(3, 4) : 0;
    st->ms_cells = one ? ms(4, 5) : 0;
    st->ms_scan = one ? ms(5, 15) : ms(3, 15);    // pipelined: all cell stages together (they overlap the copies)
    st->ms_direct = one && tess && cells > 0 ? ms(16, 17) : 0;   // part of ms_scan
    st->ms_slow_path = ms(15, 6) + (one && tess && cells > 0 ? ms(12, 13) : 0);   // oversized stars (general BFS) + oversized index boxes (per-CTA scan)
    st->ms_exchange = ms(6, 7);
    st->ms_sort = ms(7, 8);
    st->ms_deposit = ms(8, 9);
    st->ms_total_device = ms(2, 9);
    st->ms_download = io.pipelined ? ms(9, 14) : 0;
    const bool sub = one && tess && cells > 0;
    st->ms_bfs = sub && !fused ? ms(4, 11) : 0;
    st->ms_nbrs = sub && !fused ? ms(11, 12) : 0;
    st->ms_faces = sub ? ms(13, 5) : 0;
    st->ms_fused = sub && fused ? ms(4, 11) : 0;
    st->ms_emit = sub && fused ? ms(11, 12) : 0;
    st->num_faces = (int64_t)c->h_cnt->plane_cursor * 2 + (int64_t)c->h_cnt->n_faces_fused;
    st->num_candidates = (int64_t)c->h_cnt->n_cands;
    st->num_shared_deposits = n_shared_stat;
  }
  return 0;
}

extern "C" int tessb200_dense_run(tessb200_ctx *c, tessb200_dense_params *p, tessb200_dense_stats *st)
{
  if (!c) return fail(TESSB200_EINVAL, "ctx is NULL");
  PipeIO io;
  return run_impl(c, p, st, io);
}

extern "C" int tessb200_dense_geometry(tessb200_ctx *c, tessb200_dense_params *p, int nblocks, tessb200_block *blocks)
{
  if (!c || !blocks || nblocks < 1) return fail(TESSB200_EINVAL, "NULL argument");
  Geometry G;
  TRY(make_geometry(c, p, &G, blocks, nblocks));
  // G.boxes follows the layout (ascending gid); match the caller's blocks by gid
  std::vector<std::pair<int, int> > gids;   // (gid, index in G.boxes)
  if (!c->layout.empty()) for (size_t i = 0; i < c->layout.size(); i++) gids.push_back(std::make_pair(c->layout[i].gid, (int)i));
  else {
    std::vector<int> g(nblocks);
    for (int i = 0; i < nblocks; i++) g[i] = blocks[i].gid;
    std::sort(g.begin(), g.end());
    for (int i = 0; i < nblocks; i++) gids.push_back(std::make_pair(g[i], i));
  }
  for (int i = 0; i < nblocks; i++) {
    int bi = -1;
    for (size_t k = 0; k < gids.size(); k++) if (gids[k].first == blocks[i].gid) bi = gids[k].second;
    if (bi < 0 || bi >= (int)G.boxes.size()) return fail(TESSB200_EINVAL, "block gid %d is not part of the layout", blocks[i].gid);
    const BlockBox &bx = G.boxes[bi];
    memcpy(blocks[i].block_min_idx, bx.b_lo, 12); memcpy(blocks[i].block_num_idx, bx.b_num, 12);
    blocks[i].num_grid_pts = (p->project ? (int64_t)bx.b_num[1] : (int64_t)bx.b_num[1] * bx.b_num[2]) * bx.b_num[0];
  }
  return 0;
}

extern "C" int tessb200_dense_device_density(tessb200_ctx *c, int gid, void **dptr, int64_t *num_floats)
{
  if (!c || !c->ran) return fail(TESSB200_ESTATE, "no completed run");
  for (BlockRes *b : c->blocks)
    if (b->gid == gid) {
      if (dptr) *dptr = c->out.as<float>() + b->out_off;
      if (num_floats) *num_floats = b->npts;
      return 0;
    }
  return fail(TESSB200_EINVAL, "block gid %d was not uploaded", gid);
}

// ---- download ---------------------------------------------------------------------------------------
extern "C" int tessb200_dense_download(tessb200_ctx *c, int nblocks, tessb200_block *blocks, float *global_grid)
{
  if (!c || !c->ran) return fail(TESSB200_ESTATE, "tessb200_dense_download before a completed tessb200_dense_run");
  CU(cudaSetDevice(c->device));
  const tessb200_dense_params &p = c->last_params;
  for (int i = 0; i < nblocks; i++) {
    BlockRes *b = nullptr;
    for (BlockRes *x : c->blocks) if (x->gid == blocks[i].gid) b = x;
    if (!b) return fail(TESSB200_EINVAL, "block gid %d was not uploaded", blocks[i].gid);
    TRY(copy_block_out(c, p, b, &blocks[i], nullptr, c->stream));
  }
  if (global_grid)
    for (BlockRes *b : c->blocks) TRY(copy_block_out(c, p, b, nullptr, global_grid, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

// One call, host buffers in and out.  The block copies, the cell kernels and the result copies are
// pipelined (see run_impl); with pageable host memory the copies serialise but the result is the same.
extern "C" int tessb200_dense(tessb200_ctx *c, tessb200_dense_params *p, int nblocks, tessb200_block *blocks, float *global_grid,
                              tessb200_dense_stats *st)
{
  if (!c) return fail(TESSB200_EINVAL, "ctx is NULL");
  int rc = upload_impl(c, nblocks, blocks, true);
  if (rc) { cudaStreamSynchronize(c->copy_stream); return rc; }
  PipeIO io;
  io.pipelined = true;
  io.h2d_done = &c->h2d_ev;
  io.particles_done = c->part_ev;
  io.nblocks_out = nblocks;
  io.out_blocks = blocks;
  io.global_grid = global_grid;
  rc = run_impl(c, p, st, io);
  // the inputs are borrowed only for the duration of the call: never return with copies in flight
  cudaStreamSynchronize(c->copy_stream);
  cudaStreamSynchronize(c->stream);
  return rc;
}

// ---- per-tet / per-site entry points ----------------------------------------------------------------
struct TmpBlock
{
  BlockRes b;
  ~TmpBlock() { b.particles.release(); b.tets.release(); b.v2t.release(); b.cc.release(); b.rho.release(); b.walk.release(); b.hull.release(); b.p4.release(); }
};

static int upload_tmp(tessb200_ctx *c, TmpBlock &t, int num_particles, const float *particles, int num_tets, const int *tets, const int *v2t)
{
  if (num_particles < 0 || num_tets < 0) return fail(TESSB200_EINVAL, "negative count");
  if ((num_tets && !tets) || (num_particles && particles == nullptr && false)) return fail(TESSB200_EINVAL, "NULL input");
  BlockRes &b = t.b;
  b.num_particles = num_particles; b.num_tets = num_tets; b.num_orig = num_particles;
  TRY(b.particles.ensure(sizeof(float) * 3 * (size_t)std::max(1, num_particles)));
  TRY(b.tets.ensure(32 * (size_t)std::max(1, num_tets)));
  TRY(b.v2t.ensure(sizeof(int) * (size_t)std::max(1, num_particles)));
  TRY(b.cc.ensure(16 * (size_t)std::max(1, num_tets)));
  if (particles && num_particles) CU(cudaMemcpyAsync(b.particles.p, particles, sizeof(float) * 3 * (size_t)num_particles, cudaMemcpyHostToDevice, c->stream));
  if (num_tets) CU(cudaMemcpyAsync(b.tets.p, tets, 32 * (size_t)num_tets, cudaMemcpyHostToDevice, c->stream));
  b.have_v2t = v2t != nullptr;
  if (v2t && num_particles) CU(cudaMemcpyAsync(b.v2t.p, v2t, sizeof(int) * (size_t)num_particles, cudaMemcpyHostToDevice, c->stream));
  return 0;
}

extern "C" int tessb200_fill_vert_to_tet(tessb200_ctx *c, int num_particles, int num_tets, const int *tets, int *vert_to_tet)
{
  if (!c || !vert_to_tet) return fail(TESSB200_EINVAL, "NULL argument");
  CU(cudaSetDevice(c->device));
  TmpBlock t;
  TRY(upload_tmp(c, t, num_particles, nullptr, num_tets, tets, nullptr));
  if (num_particles) {
    k_fill_i32<<<cdiv(num_particles, 256), 256, 0, c->stream>>>((int *)t.b.v2t.p, num_particles, -1);
    if (num_tets) k_vert_to_tet<<<cdiv(num_tets, 256), 256, 0, c->stream>>>((const int4 *)t.b.tets.p, num_tets, (int *)t.b.v2t.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(vert_to_tet, t.b.v2t.p, sizeof(int) * (size_t)num_particles, cudaMemcpyDeviceToHost, c->stream));
  }
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int tessb200_circumcenters(tessb200_ctx *c, int num_particles, const float *particles, int num_tets, const int *tets, float *out)
{
  if (!c || !out || !particles) return fail(TESSB200_EINVAL, "NULL argument");
  CU(cudaSetDevice(c->device));
  TmpBlock t;
  TRY(upload_tmp(c, t, num_particles, particles, num_tets, tets, nullptr));
  if (num_tets) {
    k_circumcenters<<<cdiv(num_tets, 256), 256, 0, c->stream>>>((const int4 *)t.b.tets.p, num_tets, (const float *)t.b.particles.p, nullptr, (float4 *)t.b.cc.p, nullptr, nullptr);
    CU(cudaGetLastError());
    // float4 -> packed xyz on the way out (the reference's std::vector<float> layout, volume.cpp:8)
    CU(cudaMemcpy2DAsync(out, 12, t.b.cc.p, 16, 12, (size_t)num_tets, cudaMemcpyDeviceToHost, c->stream));
  }
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

// K2 on the dense stage's star kernels (kernels.cuh, "K2 on the dense stage's star kernels"): k_cell_bfs + k_cell_nbrs
// (+ the general walk) in their volumes-only mode give every complete site its face list in neighbor_edges' order, one
// thread per face computes the face's term, one thread per site adds its terms in order.  TESSB200_K2_SIMPLE=1 keeps round 1's
// one-thread-per-site kernel (k_cell_volumes) for A/B measurements.
static int cell_volumes_simple(tessb200_ctx *c, TmpBlock &t, int num_sites, float mass, int *complete, float *volume, float *density);

extern "C" int tessb200_cell_volumes(tessb200_ctx *c, int num_sites, int num_particles, const float *particles, int num_tets, const int *tets,
                                     const int *vert_to_tet, float mass, int *complete, float *volume, float *density)
{
  if (!c || !particles) return fail(TESSB200_EINVAL, "NULL argument");
  if (num_sites < 0 || num_sites > num_particles) return fail(TESSB200_EINVAL, "num_sites out of range");
  CU(cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  TmpBlock t;
  TRY(upload_tmp(c, t, num_particles, particles, num_tets, tets, vert_to_tet));
  c->last_k2_ms = 0.0f;
  const char *simple = getenv("TESSB200_K2_SIMPLE");
  if ((simple && simple[0] == '1') || num_sites == 0 || num_tets == 0) {
    CU(cudaEventRecord(c->ev[18], s));
    TRY(prep_block_geometry(c, &t.b));
    return num_sites ? cell_volumes_simple(c, t, num_sites, mass, complete, volume, density) : 0;
  }
  BlockRes *b = &t.b;
  b->num_orig = num_sites;
  // processing order only: Morton keys inside the particles' bounding box
  for (int d = 0; d < 3; d++) { b->bmin[d] = INFINITY; b->bmax[d] = -INFINITY; }
  for (size_t i = 0; i < (size_t)num_particles; i++)
    for (int d = 0; d < 3; d++) {
      const float x = particles[3 * i + d];
      if (x < b->bmin[d]) b->bmin[d] = x;
      if (x > b->bmax[d]) b->bmax[d] = x;
    }
  const long long cells = num_sites, ntets = num_tets;
  if ((unsigned long long)2 * ntets + (unsigned long long)cells + 64 >= 0xffffffffull) return fail(TESSB200_ELIMIT, "face list exceeds 2^32 face pairs");
  const uint32_t pairs = (uint32_t)(2 * ntets + cells + 64), cap_ovf = (uint32_t)std::max<long long>(1024, cells / 64);
  const int slow_warps = 148 * 4 * 4;
  const unsigned gctas = cdiv(cells, TOPO_THREADS);
  const size_t n_slots = (size_t)gctas * TOPO_THREADS;
  // every allocation before the timed region (tessb200_cell_volumes_ms reports the kernels, not cudaMalloc)
  Buf d_comp, d_vol, d_den, d_terms;
  auto cleanup = [&]() { d_comp.release(); d_vol.release(); d_den.release(); d_terms.release(); };
  int rc = 0;
  size_t sort_tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)cells, 0, 30, s);
  if ((rc = d_comp.ensure(4 * (size_t)cells)) || (rc = d_vol.ensure(4 * (size_t)cells)) || (rc = d_den.ensure(4 * (size_t)cells)) ||
      (rc = d_terms.ensure(8 * (size_t)pairs)) || (rc = c->face_list.ensure(32 * (size_t)pairs)) || (rc = c->hdr_small.ensure(sizeof(CellHdr) * (size_t)cells)) ||
      (rc = c->hdr_big.ensure(sizeof(CellHdr) * (size_t)cells)) || (rc = c->big_bitoff.ensure(8 * (size_t)cells)) || (rc = c->overflow.ensure(sizeof(uint2) * (size_t)cap_ovf)) ||
      (rc = c->ws_big.ensure(sizeof(int) * (size_t)(BIG_STAR_CAP + 2 * BIG_NBR_CAP) * (size_t)slow_warps)) || (rc = c->pre_hdr.ensure(sizeof(CellHdr) * n_slots)) ||
      (rc = c->cand.ensure(sizeof(int2) * n_slots * TOPO_CAND_CAP)) || (rc = c->d_blocks.ensure(sizeof(DevBlock))) || (rc = c->d_cnt.ensure(sizeof(Counters))) ||
      (rc = c->mkeys[0].ensure(4 * (size_t)cells)) || (rc = c->mkeys[1].ensure(4 * (size_t)cells)) || (rc = c->order[0].ensure(4 * (size_t)cells)) ||
      (rc = c->order[1].ensure(4 * (size_t)cells)) || (rc = c->cub_tmp.ensure(sort_tmp)) || (rc = b->walk.ensure(sizeof(WalkRec) * (size_t)ntets)) ||
      (rc = b->hull.ensure((size_t)std::max(1, num_particles))) || (rc = b->p4.ensure(sizeof(float4) * (size_t)std::max(1, num_particles)))) {
    cleanup();
    return rc;
  }
  CU(cudaEventRecord(c->ev[18], s));
  TRY(prep_block_geometry(c, b, true, true));
  {
    const float3 bmin = make_float3(b->bmin[0], b->bmin[1], b->bmin[2]);
    const float3 inv = make_float3(1.0f / fmaxf(b->bmax[0] - b->bmin[0], 1e-30f), 1.0f / fmaxf(b->bmax[1] - b->bmin[1], 1e-30f), 1.0f / fmaxf(b->bmax[2] - b->bmin[2], 1e-30f));
    k_morton_keys<<<cdiv(cells, 256), 256, 0, s>>>((const float *)b->particles.p, (int)cells, bmin, inv, 0u, 0, c->mkeys[0].as<uint32_t>(), c->order[0].as<uint32_t>());
    size_t tmp = c->cub_tmp.cap;
    CU(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp, c->mkeys[0].as<uint32_t>(), c->mkeys[1].as<uint32_t>(), c->order[0].as<uint32_t>(), c->order[1].as<uint32_t>(), (int)cells, 0, 30, s));
  }
  DevBlock db = dev_block(b);
  db.order = c->order[1].as<uint32_t>();
  db.cell_base = 0;
  cudaMemcpyAsync(c->d_blocks.p, &db, sizeof(DevBlock), cudaMemcpyHostToDevice, s);
  cudaMemsetAsync(c->d_cnt.p, 0, sizeof(Counters), s);
  Counters *cnt = c->d_cnt.as<Counters>();
  GridGeom g;
  memset(&g, 0, sizeof(g));
  g.alg = TB_ALG_VOLUMES;
  TopoOut to;
  memset(&to, 0, sizeof(to));
  to.small = c->hdr_small.as<CellHdr>(); to.big = c->hdr_big.as<CellHdr>(); to.big_bit_off = c->big_bitoff.as<unsigned long long>();
  to.overflow = c->overflow.as<uint2>(); to.faces = c->face_list.as<FaceRef>(); to.cnt = cnt;
  to.cap_pairs = pairs; to.cap_small = (uint32_t)cells; to.cap_big = (uint32_t)cells; to.cap_overflow = cap_ovf;
  k_vol_defaults<<<cdiv(cells, 256), 256, 0, s>>>((const int *)b->v2t.p, (int)cells, d_comp.as<int>(), d_vol.as<float>(), d_den.as<float>());
  k_cell_bfs<<<gctas, TOPO_THREADS, BFS_SMEM, s>>>(c->d_blocks.as<DevBlock>(), 0, 1, g, to, c->pre_hdr.as<CellHdr>(), c->cand.as<int2>());
  k_cell_nbrs<<<gctas, TOPO_THREADS, NBRS_SMEM, s>>>(c->d_blocks.as<DevBlock>(), to, c->pre_hdr.as<CellHdr>(), c->cand.as<int2>());
  k_cell_bfs_big<<<slow_warps / 4, 128, 0, s>>>(c->d_blocks.as<DevBlock>(), g, to, c->overflow.as<uint2>(), c->ws_big.as<int>());
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(c->h_cnt, c->d_cnt.p, sizeof(Counters), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) { cleanup(); return fail(TESSB200_ECUDA, "cell volumes (star kernels): %s", cudaGetErrorString(e)); }
  const Counters &h = *c->h_cnt;
  if (h.n_overflow > cap_ovf || h.plane_cursor > pairs) { cleanup(); return fail(TESSB200_ELIMIT, "cell volumes: %u oversized stars / %u face pairs exceed the workspace", h.n_overflow, h.plane_cursor); }
  const size_t nfaces = (size_t)h.plane_cursor * 2;
  if (nfaces) k_face_vol_terms<<<cdiv((long long)nfaces, 256), 256, 0, s>>>(c->face_list.as<FaceRef>(), nfaces, c->d_blocks.as<DevBlock>(), d_terms.as<float>());
  const uint32_t n_small = std::min(h.n_small, to.cap_small), n_big = std::min(h.n_big, to.cap_big);
  if (n_small) k_cell_vol_sum<<<cdiv(n_small, 256), 256, 0, s>>>(to.small, n_small, d_terms.as<float>(), c->d_blocks.as<DevBlock>(), mass, d_comp.as<int>(), d_vol.as<float>(), d_den.as<float>());
  if (n_big) k_cell_vol_sum<<<cdiv(n_big, 256), 256, 0, s>>>(to.big, n_big, d_terms.as<float>(), c->d_blocks.as<DevBlock>(), mass, d_comp.as<int>(), d_vol.as<float>(), d_den.as<float>());
  cudaEventRecord(c->ev[19], s);
  if (complete) cudaMemcpyAsync(complete, d_comp.p, 4 * (size_t)cells, cudaMemcpyDeviceToHost, s);
  if (volume) cudaMemcpyAsync(volume, d_vol.p, 4 * (size_t)cells, cudaMemcpyDeviceToHost, s);
  if (density) cudaMemcpyAsync(density, d_den.p, 4 * (size_t)cells, cudaMemcpyDeviceToHost, s);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cleanup();
  if (e != cudaSuccess) return fail(TESSB200_ECUDA, "cell volumes: %s", cudaGetErrorString(e));
  cudaEventElapsedTime(&c->last_k2_ms, c->ev[18], c->ev[19]);
  return 0;
}

static int cell_volumes_simple(tessb200_ctx *c, TmpBlock &t, int num_sites, float mass, int *complete, float *volume, float *density)
{
  {
    Buf d_comp, d_vol, d_den, d_ovf, d_n, d_ws;
    auto cleanup = [&]() { d_comp.release(); d_vol.release(); d_den.release(); d_ovf.release(); d_n.release(); d_ws.release(); };
    const uint32_t cap = (uint32_t)std::max(1024, num_sites / 64);
    int rc = 0;
    if ((rc = d_comp.ensure(4 * (size_t)num_sites)) || (rc = d_vol.ensure(4 * (size_t)num_sites)) || (rc = d_den.ensure(4 * (size_t)num_sites)) ||
        (rc = d_ovf.ensure(4 * (size_t)cap)) || (rc = d_n.ensure(4))) { cleanup(); return rc; }
    cudaMemsetAsync(d_n.p, 0, 4, c->stream);
    DevBlock db = dev_block(&t.b);
    k_cell_volumes<<<cdiv(num_sites, TOPO_THREADS), TOPO_THREADS, VOL_SMEM, c->stream>>>(db, num_sites, mass, d_comp.as<int>(), d_vol.as<float>(), d_den.as<float>(),
                                                                                  d_ovf.as<uint32_t>(), d_n.as<unsigned int>(), cap);
    unsigned int n_ovf = 0;
    cudaMemcpyAsync(&n_ovf, d_n.p, 4, cudaMemcpyDeviceToHost, c->stream);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { cleanup(); return fail(TESSB200_ECUDA, "k_cell_volumes: %s", cudaGetErrorString(e)); }
    if (n_ovf > cap) { cleanup(); return fail(TESSB200_ELIMIT, "%u sites exceed the fast star workspace", n_ovf); }
    if (n_ovf) {
      if ((rc = d_ws.ensure(sizeof(int) * (size_t)(BIG_STAR_CAP + 2 * BIG_NBR_CAP) * (size_t)n_ovf))) { cleanup(); return rc; }
      k_cell_volumes_big<<<cdiv(n_ovf, 128), 128, 0, c->stream>>>(db, d_ovf.as<uint32_t>(), (int)n_ovf, d_ws.as<int>(), mass, d_comp.as<int>(),
                                                                  d_vol.as<float>(), d_den.as<float>());
    }
    cudaEventRecord(c->ev[19], c->stream);
    if (complete) cudaMemcpyAsync(complete, d_comp.p, 4 * (size_t)num_sites, cudaMemcpyDeviceToHost, c->stream);
    if (volume) cudaMemcpyAsync(volume, d_vol.p, 4 * (size_t)num_sites, cudaMemcpyDeviceToHost, c->stream);
    if (density) cudaMemcpyAsync(density, d_den.p, 4 * (size_t)num_sites, cudaMemcpyDeviceToHost, c->stream);
    e = cudaStreamSynchronize(c->stream);
    cleanup();
    if (e != cudaSuccess) return fail(TESSB200_ECUDA, "cell volumes: %s", cudaGetErrorString(e));
    cudaEventElapsedTime(&c->last_k2_ms, c->ev[18], c->ev[19]);
  }
  return 0;
}

extern "C" int tessb200_cell_volumes_ms(tessb200_ctx *c, float *ms)
{
  if (!c || !ms) return fail(TESSB200_EINVAL, "NULL argument");
  *ms = c->last_k2_ms;
  return 0;
}

// Per-site density of the first-order DTFE mode (SURVEY 8(f) N4 "per-site DTFE density export"): what run_dtfe
// computes for its blocks before rasterising, for one block's vertices.
extern "C" int tessb200_dtfe_vertex_density(tessb200_ctx *c, int num_particles, const float *particles, int num_tets, const int *tets,
                                            const int *vert_to_tet, float mass, float *density)
{
  if (!c || !particles || !density) return fail(TESSB200_EINVAL, "NULL argument");
  CU(cudaSetDevice(c->device));
  TmpBlock t;
  TRY(upload_tmp(c, t, num_particles, particles, num_tets, tets, vert_to_tet));
  if (num_particles == 0) return 0;
  if (num_tets == 0) {
    for (int i = 0; i < num_particles; i++) density[i] = -1.0f;
    return 0;
  }
  TRY(prep_block_geometry(c, &t.b));
  Buf d_rho, d_ovf, d_n, d_ws;
  auto cleanup = [&]() { d_rho.release(); d_ovf.release(); d_n.release(); d_ws.release(); };
  const uint32_t cap = (uint32_t)std::max(1024, num_particles / 64);
  int rc = 0;
  if ((rc = d_rho.ensure(4 * (size_t)num_particles)) || (rc = d_ovf.ensure(4 * (size_t)cap)) || (rc = d_n.ensure(4))) { cleanup(); return rc; }
  cudaMemsetAsync(d_n.p, 0, 4, c->stream);
  DevBlock db = dev_block(&t.b);
  k_vertex_density<<<cdiv(num_particles, TOPO_THREADS), TOPO_THREADS, TOPO_SMEM, c->stream>>>(db, mass, d_rho.as<float>(), d_ovf.as<uint32_t>(),
                                                                                              d_n.as<unsigned int>(), cap);
  unsigned int n_ovf = 0;
  cudaMemcpyAsync(&n_ovf, d_n.p, 4, cudaMemcpyDeviceToHost, c->stream);
  cudaError_t e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) { cleanup(); return fail(TESSB200_ECUDA, "k_vertex_density: %s", cudaGetErrorString(e)); }
  if (n_ovf > cap) { cleanup(); return fail(TESSB200_ELIMIT, "%u vertices exceed the fast star workspace", n_ovf); }
  if (n_ovf) {
    if ((rc = d_ws.ensure(sizeof(int) * (size_t)BIG_STAR_CAP * (size_t)n_ovf))) { cleanup(); return rc; }
    k_vertex_density_big<<<cdiv((long long)n_ovf * 32, 128), 128, 0, c->stream>>>(db, mass, d_rho.as<float>(), d_ovf.as<uint32_t>(), (int)n_ovf, d_ws.as<int>());
  }
  cudaMemcpyAsync(density, d_rho.p, 4 * (size_t)num_particles, cudaMemcpyDeviceToHost, c->stream);
  e = cudaStreamSynchronize(c->stream);
  cleanup();
  if (e != cudaSuccess) return fail(TESSB200_ECUDA, "vertex densities: %s", cudaGetErrorString(e));
  return 0;
}

// ---- input check (host code) --------------------------------------------------------------------------
extern "C" int tessb200_check_block(const tessb200_block *b, int deep)
{
  if (!b) return fail(TESSB200_EINVAL, "block is NULL");
  if (b->num_particles < 0 || b->num_orig_particles < 0 || b->num_orig_particles > b->num_particles || b->num_tets < 0)
    return fail(TESSB200_EINVAL, "block gid %d: inconsistent counts", b->gid);
  if ((b->num_particles && !b->particles) || (b->num_tets && !b->tets)) return fail(TESSB200_EINVAL, "block gid %d: NULL input array", b->gid);
  for (int d = 0; d < 3; d++)
    if (!(b->bounds_min[d] <= b->bounds_max[d])) return fail(TESSB200_EINVAL, "block gid %d: bounds_min > bounds_max on axis %d", b->gid, d);
  const int np = b->num_particles, nt = b->num_tets;
  for (int t = 0; t < nt; t++) {
    const int *v = b->tets + 8 * (size_t)t, *n = v + 4;
    for (int i = 0; i < 4; i++) {
      if (v[i] < 0 || v[i] >= np) return fail(TESSB200_EINVAL, "block gid %d: tet %d has vertex %d outside [0, %d)", b->gid, t, v[i], np);
      if (n[i] < -1 || n[i] >= nt) return fail(TESSB200_EINVAL, "block gid %d: tet %d has neighbour %d outside [-1, %d)", b->gid, t, n[i], nt);
    }
    if (v[0] == v[1] || v[0] == v[2] || v[0] == v[3] || v[1] == v[2] || v[1] == v[3] || v[2] == v[3])
      return fail(TESSB200_EINVAL, "block gid %d: tet %d repeats a vertex", b->gid, t);
    if (!deep) continue;
    for (int i = 0; i < 4; i++) {
      if (n[i] < 0) continue;
      const int *w = b->tets + 8 * (size_t)n[i], *m = w + 4;
      int shared = 0, back = 0;
      for (int j = 0; j < 4; j++) {
        if (j != i && (v[j] == w[0] || v[j] == w[1] || v[j] == w[2] || v[j] == w[3])) shared++;
        if (m[j] == t) back++;
      }
      if (shared != 3 || back != 1 || n[i] == t)
        return fail(TESSB200_EINVAL, "block gid %d: tet %d slot %d: neighbour %d is not the tet across that face", b->gid, t, i, n[i]);
    }
  }
  if (b->vert_to_tet)
    for (int p = 0; p < np; p++) {
      const int t = b->vert_to_tet[p];
      if (t == -1) continue;
      if (t < -1 || t >= nt) return fail(TESSB200_EINVAL, "block gid %d: vert_to_tet[%d] = %d outside [-1, %d)", b->gid, p, t, nt);
      const int *v = b->tets + 8 * (size_t)t;
      if (v[0] != p && v[1] != p && v[2] != p && v[3] != p) return fail(TESSB200_EINVAL, "block gid %d: vert_to_tet[%d] = %d, a tet that does not hold the vertex", b->gid, p, t);
    }
  for (size_t i = 0; i < 3 * (size_t)np; i++)
    if (!std::isfinite(b->particles[i])) return fail(TESSB200_EINVAL, "block gid %d: particle %zu has a non-finite coordinate", b->gid, i / 3);
  return 0;
}

// ---- WriteGrid / ProjectGrid for one process (src/dense.cpp:751-1023) --------------------------------
extern "C" int tessb200_write_grid(const char *outfile, const tessb200_dense_params *p, int nblocks, const tessb200_block *blocks)
{
  if (!outfile || !p || !blocks || nblocks < 1) return fail(TESSB200_EINVAL, "NULL argument");
  const size_t gx = p->glo_num_idx[0], gy = p->glo_num_idx[1], gz = p->glo_num_idx[2];
  std::vector<float> grid(p->project ? gx * gy : gx * gy * gz, 0.0f);
  if (p->project) {
    // ProjectGrid: every block with min_idx z != 0 is added, in block order, into the z = 0 block
    // that shares its (x, y) origin; then only the z = 0 blocks are written
    std::vector<std::vector<float> > acc(nblocks);
    for (int i = 0; i < nblocks; i++)
      if (blocks[i].block_min_idx[2] == 0) {
        if (!blocks[i].density) return fail(TESSB200_EINVAL, "block gid %d has no density", blocks[i].gid);
        acc[i].assign(blocks[i].density, blocks[i].density + (size_t)blocks[i].block_num_idx[0] * blocks[i].block_num_idx[1]);
      }
    for (int i = 0; i < nblocks; i++) {
      if (blocks[i].block_min_idx[2] == 0) continue;
      if (!blocks[i].density) return fail(TESSB200_EINVAL, "block gid %d has no density", blocks[i].gid);
      int root = -1;
      for (int j = 0; j < nblocks; j++)
        if (blocks[j].block_min_idx[2] == 0 && blocks[j].block_min_idx[0] == blocks[i].block_min_idx[0] &&
            blocks[j].block_min_idx[1] == blocks[i].block_min_idx[1]) root = j; // the last match wins, as in dense.cpp:946-954
      if (root < 0) continue;
      size_t n = (size_t)blocks[i].block_num_idx[0] * blocks[i].block_num_idx[1];
      if (acc[root].size() < n) n = acc[root].size();
      for (size_t k = 0; k < n; k++) acc[root][k] += blocks[i].density[k];
    }
    for (int i = 0; i < nblocks; i++) {
      if (blocks[i].block_min_idx[2] != 0) continue;
      const int *mn = blocks[i].block_min_idx, *num = blocks[i].block_num_idx;
      for (int y = 0; y < num[1]; y++)
        memcpy(&grid[(size_t)(mn[1] + y) * gx + mn[0]], &acc[i][(size_t)y * num[0]], sizeof(float) * (size_t)num[0]);
    }
  } else {
    for (int i = 0; i < nblocks; i++) {
      if (!blocks[i].density) return fail(TESSB200_EINVAL, "block gid %d has no density", blocks[i].gid);
      const int *mn = blocks[i].block_min_idx, *num = blocks[i].block_num_idx;
      for (int z = 0; z < num[2]; z++)
        for (int y = 0; y < num[1]; y++)
          memcpy(&grid[((size_t)(mn[2] + z) * gy + (mn[1] + y)) * gx + mn[0]], &blocks[i].density[((size_t)z * num[1] + y) * num[0]],
                 sizeof(float) * (size_t)num[0]);
    }
  }
  FILE *f = fopen(outfile, "wb");
  if (!f) return fail(TESSB200_EIO, "cannot open %s for writing", outfile);
  size_t w = fwrite(grid.data(), sizeof(float), grid.size(), f);
  fclose(f);
  if (w != grid.size()) return fail(TESSB200_EIO, "short write to %s", outfile);
  return 0;
}

// ---- multi-GPU ----------------------------------------------------------------------------------------
extern "C" int tessb200_comm_size(tessb200_ctx *c) { return c ? c->nranks : 0; }

extern "C" int tessb200_dense_set_layout(tessb200_ctx *c, int n, const int *gids, const float *bounds6, const int *owner_rank)
{
  if (!c) return fail(TESSB200_EINVAL, "ctx is NULL");
  c->layout.clear();
  c->xcap.clear();
  if (n == 0) return 0;
  if (!gids || !bounds6 || !owner_rank) return fail(TESSB200_EINVAL, "NULL argument");
  for (int i = 0; i < n; i++) {
    LayoutBlock l;
    l.gid = gids[i];
    memcpy(l.bmin, bounds6 + 6 * i, 12); memcpy(l.bmax, bounds6 + 6 * i + 3, 12);
    l.owner = owner_rank[i];
    if (i && gids[i] <= gids[i - 1]) { c->layout.clear(); return fail(TESSB200_EINVAL, "layout gids must be ascending"); }
    if (owner_rank[i] < 0 || (i && owner_rank[i] < owner_rank[i - 1])) {
      c->layout.clear();
      return fail(TESSB200_EINVAL, "layout owner ranks must be >= 0 and non-decreasing over ascending gid (block gid %d has owner %d)", gids[i], owner_rank[i]);
    }
    c->layout.push_back(l);
  }
  c->xcap.clear();
  return 0;
}

#ifndef TESSB200_WITH_NCCL
extern "C" int tessb200_comm_unique_id(void *) { return fail(TESSB200_ENCCL, "libtess_b200 was built without NCCL"); }
extern "C" int tessb200_comm_init(tessb200_ctx *, int, int, const void *) { return fail(TESSB200_ENCCL, "libtess_b200 was built without NCCL"); }
#else
#define NC(expr)                                                                                        \
  do {                                                                                                  \
    ncclResult_t r_ = (expr);                                                                           \
    if (r_ != ncclSuccess) return fail(TESSB200_ENCCL, "%s:%d: %s: %s", __FILE__, __LINE__, #expr, ncclw::g.GetErrorString(r_)); \
  } while (0)

extern "C" int tessb200_comm_unique_id(void *id128)
{
  if (!id128) return fail(TESSB200_EINVAL, "NULL argument");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  if (const char *e = ncclw::load()) return fail(TESSB200_ENCCL, "cannot load NCCL: %s", e);
  ncclUniqueId id;
  NC(ncclw::g.GetUniqueId(&id));
  memcpy(id128, &id, 128);
  return 0;
}

extern "C" int tessb200_comm_init(tessb200_ctx *c, int nranks, int rank, const void *id128)
{
  if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(TESSB200_EINVAL, "bad argument");
  if (nranks > 64) return fail(TESSB200_ELIMIT, "more than 64 ranks");
  if (const char *e = ncclw::load()) return fail(TESSB200_ENCCL, "cannot load NCCL: %s", e);
  CU(cudaSetDevice(c->device));
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  if (c->comm) { ncclw::g.CommDestroy(c->comm); c->comm = nullptr; }
  NC(ncclw::g.CommInitRank(&c->comm, nranks, id, rank));
  c->nranks = nranks;
  c->rank = rank;
  c->xcap.clear();
  if (!c->h_xall) CU(cudaMallocHost(&c->h_xall, 8 * 65 * 64));
  return 0;
}

// Replaces master.exchange() (src/dense.cpp:98).  Only span records whose row belongs to another
// rank's blocks move (boundary cells: O(surface)): they are counted per destination, the
// nranks x nranks count matrix is all-gathered (one host sync), the records are compacted into
// per-destination segments of a send buffer -- their slots in the local list are overwritten with a
// sentinel key that sorts behind every row -- and the payload crosses NVLink with grouped
// ncclSend / ncclRecv straight into the tail of the local list.
__device__ __forceinline__ int dest_rank_of(uint64_t key, const KeyLayout &kl, const long long *__restrict__ rank_row_end, int nranks)
{
  long long row = (long long)key_row(kl, key);
  int r = 0;
  while (r < nranks - 1 && row >= rank_row_end[r]) r++;
  return r;
}
__global__ void k_count_remote(const uint64_t *__restrict__ keys, unsigned long long n, KeyLayout kl, const long long *__restrict__ rank_row_end,
                               int nranks, int me, unsigned long long *__restrict__ counts)
{
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int r = dest_rank_of(keys[i], kl, rank_row_end, nranks);
  if (r != me) atomicAdd(&counts[r], 1ull);
}
__global__ void k_scatter_remote(uint64_t *__restrict__ keys, const uint64_t *__restrict__ data, unsigned long long n, KeyLayout kl,
                                 const long long *__restrict__ rank_row_end, int nranks, int me, unsigned long long *__restrict__ cursor,
                                 uint64_t *__restrict__ send_k, uint64_t *__restrict__ send_d, uint64_t sentinel)
{
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t k = keys[i];
  int r = dest_rank_of(k, kl, rank_row_end, nranks);
  if (r == me) return;
  unsigned long long pos = atomicAdd(&cursor[r], 1ull);
  send_k[pos] = k;
  send_d[pos] = data[i];
  keys[i] = sentinel;
}
// the same into fixed-capacity segments: seg_start[r] .. seg_start[r] + seg_cap[r]; counts[r] ends up as the number of
// records bound for rank r (those beyond the capacity are not written: the run is redone with exact sizes)
__global__ void k_scatter_fixed(uint64_t *__restrict__ keys, const uint64_t *__restrict__ data, unsigned long long n, KeyLayout kl,
                                const long long *__restrict__ rank_row_end, int nranks, int me, unsigned long long *__restrict__ counts,
                                const unsigned long long *__restrict__ seg_start, const unsigned long long *__restrict__ seg_cap,
                                uint64_t *__restrict__ send_k, uint64_t *__restrict__ send_d, uint64_t sentinel)
{
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t k = keys[i];
  int r = dest_rank_of(k, kl, rank_row_end, nranks);
  if (r == me) return;
  unsigned long long pos = atomicAdd(&counts[r], 1ull);
  if (pos < seg_cap[r]) {
    send_k[seg_start[r] + pos] = k;
    send_d[seg_start[r] + pos] = data[i];
  }
  keys[i] = sentinel;
}

// grow a device buffer, keeping its first keep_bytes
static int grow_keep(Buf &b, size_t bytes, size_t keep_bytes, cudaStream_t s)
{
  if (bytes <= b.cap) return 0;
  Buf nb;
  TRY(nb.ensure(bytes));
  if (keep_bytes && b.p) {
    cudaError_t e = cudaMemcpyAsync(nb.p, b.p, keep_bytes, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { nb.release(); return fail(TESSB200_ECUDA, "grow_keep: %s", cudaGetErrorString(e)); }
  }
  b.release();
  b = nb;
  return 0;
}

// x_small layout (8-byte words): row_end[R] | mine[R + 1] (counts per destination, status) | all[R][R + 1] | seg_start[R] | seg_cap[R]
static int exchange_spans(tessb200_ctx *c, const Geometry &G, int cur, unsigned long long *n_spans)
{
  cudaStream_t s = c->stream;
  const int R = c->nranks, me = c->rank, W = R + 1;
  // row range end of every rank (blocks of a rank are contiguous in gid order, ranks ascend with the gid)
  std::vector<long long> row_end(R, 0);
  {
    const std::vector<LayoutBlock> &L = c->layout;
    if (L.empty()) return fail(TESSB200_ESTATE, "multi-GPU run without tessb200_dense_set_layout");
    for (size_t i = 0; i < L.size(); i++) {
      long long nrows = G.g.project ? G.boxes[i].b_num[1] : (long long)G.boxes[i].b_num[1] * G.boxes[i].b_num[2];
      long long end = G.boxes[i].row_base + nrows;
      if (L[i].owner < 0 || L[i].owner >= R) return fail(TESSB200_EINVAL, "layout owner rank out of range");
      row_end[L[i].owner] = std::max(row_end[L[i].owner], end);
    }
    for (int r = 1; r < R; r++) row_end[r] = std::max(row_end[r], row_end[r - 1]);
  }
  const unsigned long long n = *n_spans;
  TRY(c->x_small.ensure(8 * (size_t)(R + W + (size_t)R * W + 2 * R)));
  long long *d_row_end = c->x_small.as<long long>();
  unsigned long long *d_mine = c->x_small.as<unsigned long long>() + R;
  unsigned long long *d_all = d_mine + W;
  unsigned long long *d_seg_start = d_all + (size_t)R * W, *d_seg_cap = d_seg_start + R;
  CU(cudaMemcpyAsync(d_row_end, row_end.data(), 8 * (size_t)R, cudaMemcpyHostToDevice, s));
  CU(cudaMemsetAsync(d_mine, 0, 8 * (size_t)W, s));
  const uint64_t sentinel = G.key_bits >= 64 ? ~0ull : ((1ull << G.key_bits) - 1ull);
  uint64_t *lk = nullptr, *ld = nullptr;

  if (c->xcap.size() == (size_t)R * R) {
    // ---- steady state: fixed-size, sentinel-padded segments; nothing is read back before the deposit ----
    std::vector<unsigned long long> sstart(R + 1, 0), scap(R, 0), rstart(R + 1, 0);
    for (int r = 0; r < R; r++) {
      scap[r] = r == me ? 0ull : c->xcap[(size_t)me * R + r];
      sstart[r + 1] = sstart[r] + scap[r];
      rstart[r + 1] = rstart[r] + (r == me ? 0ull : c->xcap[(size_t)r * R + me]);
    }
    const unsigned long long send_total = sstart[R], recv_total = rstart[R];
    TRY(c->recv_keys.ensure(8 * (size_t)std::max<unsigned long long>(1, send_total)));
    TRY(c->recv_data.ensure(8 * (size_t)std::max<unsigned long long>(1, send_total)));
    for (int i = 0; i < 2; i++) {
      TRY(grow_keep(c->keys[i], 8 * (size_t)(n + recv_total), i == cur ? 8 * (size_t)n : 0, s));
      TRY(grow_keep(c->data[i], 8 * (size_t)(n + recv_total), i == cur ? 8 * (size_t)n : 0, s));
    }
    uint64_t *sk = c->recv_keys.as<uint64_t>(), *sd = c->recv_data.as<uint64_t>();
    lk = c->keys[cur].as<uint64_t>(); ld = c->data[cur].as<uint64_t>();
    CU(cudaMemcpyAsync(d_seg_start, sstart.data(), 8 * (size_t)R, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(d_seg_cap, scap.data(), 8 * (size_t)R, cudaMemcpyHostToDevice, s));
    if (send_total) CU(cudaMemsetAsync(sk, 0xFF, 8 * (size_t)send_total, s));        // unused slots: a key no row owns
    if (n) {
      k_scatter_fixed<<<cdiv((long long)n, 256), 256, 0, s>>>(lk, ld, n, G.kl, d_row_end, R, me, d_mine, d_seg_start, d_seg_cap, sk, sd, sentinel);
      COUNT_LAUNCH(c, 1);
    }
    NC(ncclw::g.AllGather(d_mine, d_all, (size_t)W, ncclUint64, c->comm, s));
    NC(ncclw::g.GroupStart());
    for (int peer = 0; peer < R; peer++) {
      if (peer == me) continue;
      const unsigned long long ns = scap[peer], nr = c->xcap[(size_t)peer * R + me];
      if (ns) {
        NC(ncclw::g.Send(sk + sstart[peer], (size_t)ns, ncclUint64, peer, c->comm, s));
        NC(ncclw::g.Send(sd + sstart[peer], (size_t)ns, ncclUint64, peer, c->comm, s));
      }
      if (nr) {
        NC(ncclw::g.Recv(lk + n + rstart[peer], (size_t)nr, ncclUint64, peer, c->comm, s));
        NC(ncclw::g.Recv(ld + n + rstart[peer], (size_t)nr, ncclUint64, peer, c->comm, s));
      }
    }
    NC(ncclw::g.GroupEnd());
    CU(cudaMemcpyAsync(c->h_xall, d_all, 8 * (size_t)R * W, cudaMemcpyDeviceToHost, s));   // read at the end of the run (exchange_check)
    CU(cudaGetLastError());
    c->x_fast_used = true;
    *n_spans = n + recv_total;
    return 0;
  }

  // ---- exact round (the first run of a layout, or after a segment proved too small): counts, one host read-back ----
  if (n) {
    k_count_remote<<<cdiv((long long)n, 256), 256, 0, s>>>(c->keys[cur].as<uint64_t>(), n, G.kl, d_row_end, R, me, d_mine);
    COUNT_LAUNCH(c, 1);
  }
  NC(ncclw::g.AllGather(d_mine, d_all, (size_t)W, ncclUint64, c->comm, s));
  std::vector<unsigned long long> all((size_t)R * W);
  CU(cudaMemcpyAsync(all.data(), d_all, 8 * (size_t)R * W, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  for (int r = 0; r < R; r++)
    if (all[(size_t)r * W + R]) return fail(TESSB200_EPEER, "rank %d failed before the span exchange", r);
  // all[src * W + dst]
  std::vector<unsigned long long> soff(R + 1, 0), roff(R + 1, 0);
  for (int r = 0; r < R; r++) {
    soff[r + 1] = soff[r] + all[(size_t)me * W + r];
    roff[r + 1] = roff[r] + all[(size_t)r * W + me];
  }
  const unsigned long long send_total = soff[R], recv_total = roff[R];
  TRY(c->recv_keys.ensure(8 * (size_t)std::max<unsigned long long>(1, send_total)));   // used as the send buffers
  TRY(c->recv_data.ensure(8 * (size_t)std::max<unsigned long long>(1, send_total)));
  for (int i = 0; i < 2; i++) {
    TRY(grow_keep(c->keys[i], 8 * (size_t)(n + recv_total), i == cur ? 8 * (size_t)n : 0, s));
    TRY(grow_keep(c->data[i], 8 * (size_t)(n + recv_total), i == cur ? 8 * (size_t)n : 0, s));
  }
  uint64_t *sk = c->recv_keys.as<uint64_t>(), *sd = c->recv_data.as<uint64_t>();
  lk = c->keys[cur].as<uint64_t>(); ld = c->data[cur].as<uint64_t>();
  if (send_total) {
    CU(cudaMemcpyAsync(d_mine, soff.data(), 8 * (size_t)R, cudaMemcpyHostToDevice, s));   // cursors start at the segment offsets
    k_scatter_remote<<<cdiv((long long)n, 256), 256, 0, s>>>(lk, ld, n, G.kl, d_row_end, R, me, d_mine, sk, sd, sentinel);
    COUNT_LAUNCH(c, 1);
  }
  NC(ncclw::g.GroupStart());
  for (int peer = 0; peer < R; peer++) {
    if (peer == me) continue;
    const unsigned long long ns = all[(size_t)me * W + peer], nr = all[(size_t)peer * W + me];
    if (ns) {
      NC(ncclw::g.Send(sk + soff[peer], (size_t)ns, ncclUint64, peer, c->comm, s));
      NC(ncclw::g.Send(sd + soff[peer], (size_t)ns, ncclUint64, peer, c->comm, s));
    }
    if (nr) {
      NC(ncclw::g.Recv(lk + n + roff[peer], (size_t)nr, ncclUint64, peer, c->comm, s));
      NC(ncclw::g.Recv(ld + n + roff[peer], (size_t)nr, ncclUint64, peer, c->comm, s));
    }
  }
  NC(ncclw::g.GroupEnd());
  CU(cudaGetLastError());
  // capacities for the runs that follow: every rank computes the same table from the same matrix
  c->xcap.assign((size_t)R * R, 0ull);
  for (int a = 0; a < R; a++)
    for (int b = 0; b < R; b++)
      if (a != b) {
        const unsigned long long cnt = all[(size_t)a * W + b];
        c->xcap[(size_t)a * R + b] = cnt ? cnt + cnt / 4 + 4096 : 0ull;
      }
  *n_spans = n + recv_total;
  return 0;
}

// after the run of a fixed-size exchange: did every segment hold its records, did every rank get this far?
static int exchange_check(tessb200_ctx *c, bool *redo)
{
  *redo = false;
  if (!c->x_fast_used) return 0;
  c->x_fast_used = false;
  const int R = c->nranks, W = R + 1;
  bool grow = false;
  for (int a = 0; a < R; a++) {
    if (c->h_xall[(size_t)a * W + R]) { c->xcap.clear(); return fail(TESSB200_EPEER, "rank %d failed before the span exchange", a); }
    for (int b = 0; b < R; b++)
      if (a != b && c->h_xall[(size_t)a * W + b] > c->xcap[(size_t)a * R + b]) grow = true;
  }
  if (grow) {
    c->xcap.clear();        // the next run is an exact round (and sets new capacities)
    *redo = true;
  }
  return 0;
}

// the failing rank's part of the exchange: empty, flagged.  Best effort (an allocation failure here leaves the peers waiting).
static void exchange_poison(tessb200_ctx *c)
{
  if (!c->comm || c->nranks < 2) return;
  cudaStream_t s = c->stream;
  const int R = c->nranks, me = c->rank, W = R + 1;
  std::string keep = g_err;
  if (c->x_small.ensure(8 * (size_t)(R + W + (size_t)R * W + 2 * R))) { g_err = keep; return; }
  unsigned long long *d_mine = c->x_small.as<unsigned long long>() + R, *d_all = d_mine + W;
  std::vector<unsigned long long> mine(W, 0ull);
  mine[R] = 1ull;
  cudaMemcpyAsync(d_mine, mine.data(), 8 * (size_t)W, cudaMemcpyHostToDevice, s);
  if (c->xcap.size() == (size_t)R * R) {
    unsigned long long send_total = 0, recv_total = 0;
    for (int r = 0; r < R; r++)
      if (r != me) { send_total += c->xcap[(size_t)me * R + r]; recv_total += c->xcap[(size_t)r * R + me]; }
    if (c->recv_keys.ensure(8 * (size_t)std::max<unsigned long long>(1, send_total)) || c->recv_data.ensure(8 * (size_t)std::max<unsigned long long>(1, send_total)) ||
        c->keys[0].ensure(8 * (size_t)std::max<unsigned long long>(1, recv_total)) || c->data[0].ensure(8 * (size_t)std::max<unsigned long long>(1, recv_total))) {
      g_err = keep;
      return;
    }
    if (send_total) cudaMemsetAsync(c->recv_keys.p, 0xFF, 8 * (size_t)send_total, s);
    ncclw::g.AllGather(d_mine, d_all, (size_t)W, ncclUint64, c->comm, s);
    ncclw::g.GroupStart();
    unsigned long long so = 0, ro = 0;
    for (int peer = 0; peer < R; peer++) {
      if (peer == me) continue;
      const unsigned long long ns = c->xcap[(size_t)me * R + peer], nr = c->xcap[(size_t)peer * R + me];
      if (ns) {
        ncclw::g.Send(c->recv_keys.as<uint64_t>() + so, (size_t)ns, ncclUint64, peer, c->comm, s);
        ncclw::g.Send(c->recv_data.as<uint64_t>() + so, (size_t)ns, ncclUint64, peer, c->comm, s);
      }
      if (nr) {
        ncclw::g.Recv(c->keys[0].as<uint64_t>() + ro, (size_t)nr, ncclUint64, peer, c->comm, s);
        ncclw::g.Recv(c->data[0].as<uint64_t>() + ro, (size_t)nr, ncclUint64, peer, c->comm, s);
      }
      so += ns; ro += nr;
    }
    ncclw::g.GroupEnd();
    c->xcap.clear();
  } else {
    ncclw::g.AllGather(d_mine, d_all, (size_t)W, ncclUint64, c->comm, s);
  }
  cudaStreamSynchronize(s);
  g_err = keep;
}
#endif
