/* tess_b200.h -- C ABI of the B200-native density-estimation stage of tess2.
 *
 * This is the drop-in boundary for ONE stage of diatomic/tess2: everything that happens
 * inside `dense()` (reference include/tess/dense.hpp:75-89, src/dense.cpp:30-103) plus the
 * per-tet / per-cell geometry of src/tet.cpp and src/volume.cpp.  The serial Delaunay engine
 * (Qhull/CGAL), ghost exchange and DIY stay on the host; their product -- the per-block
 * arrays of `struct dblock_t` (reference include/tess/delaunay.h:38-63) -- is what crosses
 * this boundary.  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions
 *   - every entry point returns 0 on success and a negative TESSB200_E* code on failure;
 *     tessb200_last_error() gives the message of the calling thread's last failure.
 *     Nothing aborts or throws across the boundary (the reference asserts / MPI_Aborts:
 *     src/dense.cpp:1027-1036).
 *   - input arrays are borrowed for the duration of the call and never written.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *     TESSB200_ECUDA.
 *   - one context = one GPU = one host thread at a time.  Multi-GPU runs use one context
 *     per process/GPU joined by tessb200_comm_init() (NCCL over NVLink).
 */
#ifndef TESS_B200_H
#define TESS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TESSB200_VERSION 1

enum
{
  TESSB200_OK = 0,
  TESSB200_EINVAL = -1,    /* bad argument */
  TESSB200_ECUDA = -2,     /* CUDA runtime / driver error, or no device */
  TESSB200_ENOMEM = -3,    /* device or host allocation failed */
  TESSB200_ECAPACITY = -4, /* a caller-provided output buffer is too small */
  TESSB200_ELIMIT = -5,    /* input exceeds a documented limit */
  TESSB200_ENCCL = -6,     /* NCCL error */
  TESSB200_ESTATE = -7,    /* call sequence error (e.g. run before upload) */
  TESSB200_EIO = -8,       /* file error in tessb200_write_grid */
  TESSB200_EPEER = -9      /* multi-GPU run: another rank failed before the span exchange (every rank returns together) */
};

/* estimator algorithm: reference `enum alg`, include/tess/dense.hpp:33-38 */
enum
{
  TESSB200_DENSE_TESS = 0, /* DENSE_TESS: Voronoi-cell deposit (src/dense.cpp:220-313) */
  TESSB200_DENSE_CIC = 1,  /* DENSE_CIC : cloud-in-cell    (src/dense.cpp:486-562) */
  TESSB200_DENSE_DTFE = 2  /* NOT in the reference (its enum stops at DENSE_CIC): first-order DTFE, per-vertex
                              density 4m / (volume of the star) interpolated linearly inside every Delaunay tet.
                              3-D output only.  See DESIGN.md 3.6; checked against oracle/dense_oracle.c only. */
};

/* One DIY block = the fields of `struct dblock_t` (include/tess/delaunay.h:38-63) that the
 * dense stage reads, plus `DBlock::bounds` (include/tess/delaunay.hpp:18-23). */
typedef struct tessb200_block
{
  int gid;                    /* dblock_t::gid; blocks must be unique in gid */
  int num_orig_particles;     /* dblock_t::num_orig_particles: cells [0, num_orig) deposit */
  int num_particles;          /* dblock_t::num_particles: originals first, then ghosts */
  const float *particles;     /* dblock_t::particles, xyz AoS, 3*num_particles floats */
  int num_tets;               /* dblock_t::num_tets */
  const int *tets;            /* dblock_t::tets viewed as int[8] per tet: verts[4], tets[4]
                                 (include/tess/tet.h:4-7; tets[i] opposite verts[i], -1 = hull) */
  const int *vert_to_tet;     /* dblock_t::vert_to_tet, or NULL: recomputed on the device with
                                 fill_vert_to_tet's "last tet wins" rule (src/tess.cpp:767-787) */
  float bounds_min[3];        /* DBlock::bounds.min */
  float bounds_max[3];        /* DBlock::bounds.max */

  /* outputs */
  float *density;             /* dblock_t::density: the block's sub-array of the global grid,
                                 [nz][ny][nx] x-fastest ([ny][nx] when projecting); caller
                                 allocated, may be NULL when only the global grid is wanted */
  int64_t density_capacity;   /* floats available in density */
  int block_min_idx[3];       /* BlockGridParams (src/dense.cpp:575-648): global index of the block's first grid point */
  int block_num_idx[3];       /*                                          grid points per axis */
  int64_t num_grid_pts;       /* dblock_t::num_grid_pts */
} tessb200_block;

/* Scalar arguments of dense() (include/tess/dense.hpp:75-89) == `struct args_t` (:48-61). */
typedef struct tessb200_dense_params
{
  int alg;                    /* TESSB200_DENSE_TESS / TESSB200_DENSE_CIC */
  int num_given_bounds;       /* 0..3: leading axes whose grid extents are given (src/dense.cpp:1725-1735).  Extents narrower than the
                                 data: deposits without a grid element are dropped (the reference writes out of bounds there);
                                 with project != 0 a given z range only sets the z sampling, every z index still deposits */
  float given_mins[3];
  float given_maxs[3];
  int project;                /* != 0: 2-D density, projection along z (the reference asserts xy only) */
  float proj_plane[3];        /* accepted for signature parity; must be (0,0,1) when project != 0 */
  float mass;                 /* mass of one particle */
  float eps;                  /* floating point tolerance (the drivers pass 1e-4) */
  int glo_num_idx[3];         /* global grid size */

  /* outputs (dense()'s float* out-parameters) */
  float data_mins[3], data_maxs[3];          /* DataBounds, src/dense.cpp:1221-1275 */
  float grid_phys_mins[3], grid_phys_maxs[3];/* GridStepParams, src/dense.cpp:1712-1767 */
  float grid_step_size[3];
} tessb200_dense_params;

/* What dense_stats() prints (src/dense.cpp:1284-1333) plus per-stage device times. */
typedef struct tessb200_dense_stats
{
  int64_t num_cells;          /* original particles visited */
  int64_t num_no_tet;         /* vert_to_tet == -1 (src/dense.cpp:251) */
  int64_t num_incomplete;     /* !complete() (src/dense.cpp:252) */
  int64_t num_outside;        /* rejected by the data-bounds filter (src/dense.cpp:1385-1392) */
  int64_t num_deposit_cells;  /* cells that deposited == the reference's check_mass */
  int64_t num_cic_fallback;   /* cells whose scan found no point (src/dense.cpp:1437-1455) */
  int64_t num_slow_cells;     /* cells routed through the large-star / large-bbox kernels */
  int64_t num_spans;          /* x-run records handed to the deposit kernel */
  int64_t num_tets;           /* tets handed over, all blocks */
  int64_t num_grid_pts;       /* grid points owned by this context's blocks */
  int64_t num_kernel_launches;/* this library's own kernels launched by the run (the CUB sort passes are not counted) */
  double tot_mass;            /* sum over the final grid of value * div (== deposited mass) */
  float max_dense;            /* max over the final grid */
  float ms_upload, ms_circumcenters, ms_cells, ms_scan, ms_sort, ms_deposit, ms_exchange, ms_download;
  float ms_total_device;      /* inputs resident -> grid complete in device memory */
  float ms_bfs, ms_nbrs, ms_faces; /* the three kernels inside ms_cells (resident runs, TESSB200_FUSED=0 only) */
  int64_t num_faces;          /* Voronoi faces of the depositing cells (plane records) */
  int64_t num_candidates;     /* candidate neighbours handed from k_cell_bfs to k_cell_nbrs */
  float ms_slow_path;         /* general BFS for oversized stars + per-CTA scan of oversized cells (after the fast kernels) */
  float ms_fused;             /* k_cell_fused: star walk + faces + planes + inside bits of every cell the fast path holds */
  int64_t num_shared_deposits; /* deposits that met another one on their grid point and went through the ordered path;
                                  -1 when every record did (projection, or more shared deposits than the buffer holds) */
  float ms_emit;              /* k_cell_emit: scan-line walk over the inside bits + span records */
  float ms_direct;            /* k_cell_direct (cells with small index boxes: faces + planes + inside test + scan per thread); inside ms_scan */
} tessb200_dense_stats;

typedef struct tessb200_ctx tessb200_ctx;

/* ---- context ---- */
int tessb200_create(tessb200_ctx **ctx, int device);
void tessb200_destroy(tessb200_ctx *ctx);
const char *tessb200_last_error(void);
int tessb200_version(void);

/* ---- the dense stage: replaces dense() (include/tess/dense.hpp:75-89) ----
 * One call = DataBounds + GridStepParams + init_dense + est_dense + exchange + recvd_pts for
 * the given blocks, host buffers in, host buffers out (per-block `density`, and the global
 * C-order grid `global_grid` [gz][gy][gx] when non-NULL and not projecting).  `global_grid` receives the sub-grids of
 * THIS context's blocks only, each copied to its place: where no given block holds a grid point (blocks of other ranks,
 * or a decomposition that does not tile the grid) the caller's array is left as it was -- zero it first if needed. */
int tessb200_dense(tessb200_ctx *ctx, tessb200_dense_params *params, int nblocks, tessb200_block *blocks,
                   float *global_grid, tessb200_dense_stats *stats);

/* The same stage in three steps, so that a caller (or bench.py) can keep inputs resident in
 * HBM:  upload (H2D of particles/tets/vert_to_tet) -> run (device only) -> download (D2H). */
int tessb200_dense_upload(tessb200_ctx *ctx, int nblocks, const tessb200_block *blocks);
int tessb200_dense_run(tessb200_ctx *ctx, tessb200_dense_params *params, tessb200_dense_stats *stats);
int tessb200_dense_download(tessb200_ctx *ctx, int nblocks, tessb200_block *blocks, float *global_grid);
/* grid geometry of the uploaded blocks without running: fills params outputs and the blocks'
 * block_min_idx / block_num_idx / num_grid_pts (BlockGridParams, src/dense.cpp:575-648) */
int tessb200_dense_geometry(tessb200_ctx *ctx, tessb200_dense_params *params, int nblocks, tessb200_block *blocks);
/* device pointer of the density array of the uploaded block with this gid after tessb200_dense_run (for zero-copy
 * consumers); TESSB200_EINVAL when no uploaded block has that gid */
int tessb200_dense_device_density(tessb200_ctx *ctx, int gid, void **dptr, int64_t *num_floats);

/* ---- per-tet / per-site geometry: replaces src/volume.cpp and pieces of src/tet.cpp ---- */
/* fill_vert_to_tet (src/tess.cpp:767-787) */
int tessb200_fill_vert_to_tet(tessb200_ctx *ctx, int num_particles, int num_tets, const int *tets, int *vert_to_tet);
/* fill_circumcenters (src/volume.cpp:6-11): out = 3*num_tets floats */
int tessb200_circumcenters(tessb200_ctx *ctx, int num_particles, const float *particles, int num_tets,
                           const int *tets, float *circumcenters);
/* Per-site quantities for sites [0, num_sites):
 *   complete[i] : complete() (src/tet.cpp:337-378): 1 finite, 0 infinite, -1 site in no tet
 *   volume[i]   : volume()   (src/volume.cpp:13-54): Voronoi cell volume, -1 if infinite, -2 if in no tet
 *   density[i]  : mass / volume where volume > 0, else 0 (zero-order Voronoi density)
 * vert_to_tet may be NULL; any of the outputs may be NULL. */
int tessb200_cell_volumes(tessb200_ctx *ctx, int num_sites, int num_particles, const float *particles, int num_tets,
                          const int *tets, const int *vert_to_tet, float mass, int *complete, float *volume,
                          float *density);

/* device time (ms, CUDA events) of the kernels of the last tessb200_cell_volumes call on this context: circumcenters +
 * the per-site star walk and fan sums, without the host <-> device copies (bench.py's K2 line) */
int tessb200_cell_volumes_ms(tessb200_ctx *ctx, float *ms);

/* Per-site density of the first-order DTFE mode (NOT in the reference, see TESSB200_DENSE_DTFE): for every particle
 * density[v] = 4 * mass / (sum of the volumes of the tets at v), -1 where the star is infinite or v is in no tet.
 * vert_to_tet may be NULL. */
int tessb200_dtfe_vertex_density(tessb200_ctx *ctx, int num_particles, const float *particles, int num_tets, const int *tets,
                                 const int *vert_to_tet, float mass, float *density);

/* ---- input check (host code, no device needed) ----
 * The kernels index `particles` with the tets' vertex ids and `tets` with their neighbour ids without range
 * checks, as the reference does (src/tet.cpp, src/dense.cpp trust tess()).  A caller that does not trust its
 * tessellation can ask first: every vertex id in [0, num_particles), every neighbour id in [-1, num_tets),
 * every vert_to_tet entry -1 or a tet that holds the vertex, counts consistent; with `deep` != 0 also that
 * tets[i] really is the tet across the face opposite verts[i] (shares the other three vertices and points
 * back).  Returns 0, or TESSB200_EINVAL with the first finding in tessb200_last_error().  With the environment
 * variable TESSB200_CHECK_INPUT=1 (2 = deep) tessb200_dense / tessb200_dense_upload run it on every block. */
int tessb200_check_block(const tessb200_block *block, int deep);

/* ---- output: replaces WriteGrid (src/dense.cpp:751-870) for one process ----
 * Writes the raw C-order float32 grid (x fastest, no header) that the reference's MPI-IO
 * subarray writes produce.  With project != 0 the z-stacked blocks are summed into the z=0
 * blocks first (ProjectGrid, src/dense.cpp:881-1023).  Uses the blocks' host `density`. */
int tessb200_write_grid(const char *outfile, const tessb200_dense_params *params, int nblocks,
                        const tessb200_block *blocks);

/* ---- multi-GPU: replaces master.exchange() + recvd_pts (src/dense.cpp:98-102, 166-202) ----
 * One context per rank.  Rank 0 calls tessb200_comm_unique_id, the 128 bytes travel to the
 * other ranks by any host channel, every rank calls tessb200_comm_init.  Afterwards
 * tessb200_dense_run treats `blocks` as this rank's share of `all_blocks` (see
 * tessb200_dense_set_layout) and exchanges boundary spans with NCCL. */
int tessb200_comm_unique_id(void *id128);
int tessb200_comm_init(tessb200_ctx *ctx, int nranks, int rank, const void *id128);
/* number of ranks of the communicator this context joined (1: none) */
int tessb200_comm_size(tessb200_ctx *ctx);
/* Global layout for multi-GPU runs: bounds and owner rank of EVERY block of the decomposition,
 * in ascending gid order (bounds: 6 floats per block: min xyz, max xyz).  owner_rank must be
 * non-decreasing over ascending gid (diy::ContiguousAssigner's rule: rank r owns one contiguous run
 * of gids, and lower ranks own lower gids): the exchange routes a record by the row range of its
 * rank and the accumulation order of received points follows the rank-windowed cell numbers.
 * Anything else is refused with TESSB200_EINVAL.  Every rank must own at least one block. */
int tessb200_dense_set_layout(tessb200_ctx *ctx, int nblocks_global, const int *gids, const float *bounds6,
                              const int *owner_rank);

#ifdef __cplusplus
}
#endif
#endif /* TESS_B200_H */
