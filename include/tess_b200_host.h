/* tess_b200_host.h -- host side of the tessellation stage (CPU only, libtess_b200_host.so).
 *
 * the serial Delaunay engine stays on the host (north_star).  The reference plugs
 * Qhull or CGAL behind local_cells() / gen_delaunay_output() (src/tess-qhull.c:31-165,
 * src/tess-cgal.cpp); neither library is installed here, so the engine is this repo's own
 * (tess2_b200/host/delaunay3.hpp).  Output layout = struct tet_t (include/tess/tet.h:4-7).
 */
#ifndef TESS_B200_HOST_H
#define TESS_B200_HOST_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Delaunay tetrahedralisation of num_particles float32 points (xyz AoS).
 * *tets receives a malloc'ed array of *num_tets records {int verts[4]; int tets[4]} with tets[i]
 * opposite verts[i] and -1 on the convex hull (free with tessb200_host_free or free()).
 * Returns 0 on success (also when the points admit no tet: *num_tets = 0), < 0 on error. */
int tessb200_host_delaunay(int num_particles, const float *particles, int *num_tets, int **tets);

/* One block after tessellation: the dblock_t fields dense() reads (include/tess/delaunay.h:38-63).
 * Arrays are malloc'ed (tess2 frees them with free(), src/tess.cpp:166-178). */
typedef struct tessb200_host_block {
  int gid;
  float bounds_min[3], bounds_max[3];
  int num_orig_particles;    /* particles owned by the block: first in `particles`, in input order */
  int num_particles;         /* originals + ghosts */
  int num_tets;
  float *particles;          /* [3 * num_particles] */
  int *tets;                 /* [8 * num_tets]: verts[4], tets[4] */
  int *vert_to_tet;          /* [num_particles], last tet holding the vertex (src/tess.cpp:767-787), -1 if none */
  int *global_ids;           /* [num_particles] index of each local particle in the input array */
  float ghost_margin;        /* width of the ghost region that was finally used */
  int rounds;                /* tessellation rounds (the region grows until every original cell is settled) */
  double seconds;            /* time spent in the Delaunay engine for this block */
  int settled;               /* 1: the ghost region covers the circumsphere of every tet at an original, finite cell (or the whole
                                domain); 0: the rounds / growth limits stopped the widening first -- the stars near the block border
                                may differ from the global Delaunay triangulation and dense() may deposit wrong or missing mass there */
  int reserved;
} tessb200_host_block;

/* tess() for one process holding all particles (replaces the round loop of src/tess.cpp:52-116 with
 * its ghost exchange, :351-457, for the single-node case): every block takes the particles it owns
 * plus the ghosts inside a margin around its bounds, tessellates them, and widens the margin until
 * the circumsphere of every tet at an original, finite cell stays inside the searched region (the
 * test of incomplete_cells, src/tess.cpp:492-628, against the region instead of neighbour bounds).
 *   owner        [num_particles] gid of the owning block, or NULL: ownership by containment in
 *                block_bounds (min <= x < max; the upper domain faces are inclusive)
 *   block_bounds [6 * nblocks]: min xyz, max xyz of block gid = index
 *   gids         [num_gids] the blocks to tessellate (a rank's share), or NULL: all nblocks
 *   margin0      first ghost margin; <= 0: three mean particle spacings of the block
 *   max_rounds   tessellation rounds per block at most (<= 0: 3)
 *   max_growth   the margin never exceeds max_growth * margin0 (<= 0: 2.5)
 *   num_threads  blocks in flight (<= 0: hardware concurrency)
 * blocks_out[num_gids (or nblocks)] is filled in gids order; release with tessb200_host_free_block. */
int tessb200_host_tess(int num_particles, const float *particles, const int *owner, const float *domain_min, const float *domain_max,
                       int nblocks, const float *block_bounds, int num_gids, const int *gids, float margin0, int max_rounds,
                       float max_growth, int num_threads, tessb200_host_block *blocks_out);
/* the same with a periodic domain (wrap != 0: x, y and z, as the reference drivers set all three, examples/tess/main.cpp:83-88):
 * a block's ghosts are also the images of particles -- its own included -- shifted by whole domain extents, with the
 * coordinates of wrap_pt (src/tess.cpp:698-710); an image keeps the global id of its particle.  The margin then widens until
 * no original is left on the hull and every circumsphere at an original cell fits (no clipping at the domain). */
int tessb200_host_tess_periodic(int num_particles, const float *particles, const int *owner, const float *domain_min, const float *domain_max,
                       int nblocks, const float *block_bounds, int num_gids, const int *gids, float margin0, int max_rounds,
                       float max_growth, int num_threads, int wrap, tessb200_host_block *blocks_out);
void tessb200_host_free_block(tessb200_host_block *b);

/* Block decompositions (the reference delegates both to DIY: RegularDecomposer,
 * examples/tess-dense/main.cpp:190-195, and diy::kdtree, src/tess-kdtree.cpp:83-107; DIY is not
 * vendored, so the split positions are this repo's: SURVEY 8(c) "parity unpinned").
 * bounds_out: [6 * nblocks] min xyz, max xyz of block gid = index.
 *   regular: nblocks factored as evenly as possible (8 -> 2x2x2, 64 -> 4x4x4), gid x-fastest.
 *   kdtree : nblocks a power of two; exact-median splits cycling x, y, z per level;
 *            owner_out [num_particles] receives the gid of every particle (may be NULL). */
int tessb200_host_regular_blocks(const float *domain_min, const float *domain_max, int nblocks, float *bounds_out);
int tessb200_host_kdtree_blocks(int num_particles, const float *particles, const float *domain_min, const float *domain_max, int nblocks,
                                float *bounds_out, int *owner_out);

/* ---- the hand-off file between the stages ("del.out"), SURVEY 8(f) N3 ---------------------------------
 * tess_save (src/tess.cpp:126-137) = diy::io::write_blocks + save_block_light (src/tess.cpp:198-221);
 * examples/dense/main.cpp:158-161 = diy::io::read_blocks + load_block_light (src/tess.cpp:223-259).
 * One record per block with every field load_block_light fills, in dblock_t's terms
 * (include/tess/delaunay.h:38-63) plus DBlock's three boxes (include/tess/delaunay.hpp:18-23).
 * The payload order is the reference's; the container (footer of {gid, offset, count}, trailing
 * footer size) restates DIY's published block-file layout and is parity-unpinned, DIY not being
 * vendored (tess2_b200/host/block_file.cpp has the grammar). */
typedef struct tessb200_host_dblock {
  int gid;
  float bounds_min[3], bounds_max[3];   /* DBlock::bounds: local block extents */
  float box_min[3], box_max[3];         /* DBlock::box: box of the last redistribution round */
  float data_min[3], data_max[3];       /* DBlock::data_bounds: global data extents */
  int num_orig_particles, num_particles;
  float *particles;                     /* [3 * num_particles] */
  int *rem_gids, *rem_lids;             /* [num_particles - num_orig_particles] owner block / index there of each ghost */
  int num_grid_pts;
  float *density;                       /* [num_grid_pts] (0 points straight after tess()) */
  int complete;
  int num_tets;
  int *tets;                            /* [8 * num_tets] */
  int *vert_to_tet;                     /* [num_particles] */
} tessb200_host_dblock;

enum {
  TESSB200_DIY_BOUNDS_DYNAMIC = 0,      /* Bounds = {size_t 3, float[3]} x 2: DIY with run-time point dimension (`Bounds b {3}`) */
  TESSB200_DIY_BOUNDS_STATIC4 = 1       /* Bounds = float[4] x 2: DIY with DIY_MAX_DIM = 4 points */
};

/* Writes the blocks in the order given; the footer lists them by gid.  NULL rem_gids / rem_lids are
 * written as -1.  extra: opaque user bytes of tess_save's `extra` buffer (may be NULL / 0). */
int tessb200_host_write_blocks(const char *path, int nblocks, const tessb200_host_dblock *blocks, int bounds_layout,
                               const void *extra, size_t extra_size);
/* Reads every block of the file, in gid order.  The link record in front of each payload is
 * skipped whatever its class; both Bounds layouts are recognised (*bounds_layout, may be NULL,
 * receives the one found).  Arrays are malloc'ed; release with tessb200_host_free_dblocks. */
int tessb200_host_read_blocks(const char *path, int *nblocks, tessb200_host_dblock **blocks, int *bounds_layout);
void tessb200_host_free_dblocks(int nblocks, tessb200_host_dblock *blocks);

void tessb200_host_free(void *p);
const char *tessb200_host_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
