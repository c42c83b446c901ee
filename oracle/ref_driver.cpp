// C-ABI driver around the UNMODIFIED reference sources (TEST INFRASTRUCTURE ONLY).
//
// Linked together with /root/reference/src/{dense,tet,volume}.cpp (compiled where
// they lie, see oracle/Makefile) into oracle/_ref/libtess_ref.so.  The product
// never loads this library; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs do, as the checker / reported baseline.
//
// What is the reference's and what is restated here:
//   * dense(), BlockGridParams(), CellBounds(), CellGridPts(), WriteGrid()
//     (src/dense.cpp), fill_circumcenters(), volume() (src/volume.cpp),
//     complete() (src/tet.cpp) are called as they are.
//   * fill_vert_to_tet (src/tess.cpp:767-787, "the last one wins") is restated
//     because tess.cpp needs libqhull, which is not installed.
//   * the diy::Master / links are built by hand: every block is linked to every
//     other block, so diy::in() (dense.cpp:302) forwards a grid point to whichever
//     block's closed bounds contain it.
#include <vector>
#include <cstring>
#include <cstdio>
#include "tess/dense.hpp"
#include "tess/volume.h"

// -DTESSB200_DROPIN builds the SAME driver with one line changed: the call to the reference's dense()
// becomes tessb200::dense() from include/tess_b200_diy.hpp (same signature + a GPU context).  That
// library (oracle/_ref/libtess_dropin.so, entry point drp_dense) is the drop-in test: identical
// DBlock objects in an identical diy::Master, only the dense stage swapped (tests/test_dropin.py).
#ifdef TESSB200_DROPIN
#include "tess_b200_diy.hpp"
#define ref_dense drp_dense
#define ref_fill_vert_to_tet drp_fill_vert_to_tet
#define ref_circumcenters drp_circumcenters
#define ref_complete drp_complete
#define ref_volumes drp_volumes
#define ref_cell_points drp_cell_points
#endif

extern "C" {

struct ref_block_t
{
  int gid;
  int num_orig_particles;
  int num_particles;
  const float *particles;   // xyz AoS, originals first
  int num_tets;
  const int *tets;          // num_tets x 8 ints = tet_t layout (include/tess/tet.h:4-7)
  const int *vert_to_tet;   // may be NULL -> recomputed (last tet wins)
  float bounds_min[3];
  float bounds_max[3];
  float *density;           // out, caller allocated
  long long density_capacity; // floats available in density
  int block_min_idx[3];     // out
  int block_num_idx[3];     // out
  int num_grid_pts;         // out
};

struct ref_params_t
{
  int alg;                  // 0 = DENSE_TESS, 1 = DENSE_CIC (include/tess/dense.hpp:33-38)
  int num_given_bounds;
  float given_mins[3], given_maxs[3];
  int project;
  float proj_plane[3];
  float mass;
  float eps;
  int glo_num_idx[3];
  // outputs
  float data_mins[3], data_maxs[3];
  float grid_phys_mins[3], grid_phys_maxs[3];
  float grid_step_size[3];
  double seconds;           // wall time of dense() alone (the reference's COMP_TIME interval)
};

// src/tess.cpp:767-787
void ref_fill_vert_to_tet(int num_particles, int num_tets, const int *tets, int *vert_to_tet)
{
  for (int p = 0; p < num_particles; ++p)
    vert_to_tet[p] = -1;
  for (int t = 0; t < num_tets; ++t)
    for (int v = 0; v < 4; ++v)
      vert_to_tet[tets[8 * t + v]] = t; // the last one wins
}

// src/volume.cpp:6-11
void ref_circumcenters(int num_tets, const int *tets, const float *particles, float *out)
{
  std::vector<float> cc;
  fill_circumcenters(cc, (tet_t *)tets, num_tets, (float *)particles);
  memcpy(out, cc.data(), sizeof(float) * 3 * (size_t)num_tets);
}

// src/tet.cpp:337-378 for every vertex; -1 where the vertex is in no tet
void ref_complete(int num_verts, int num_tets, const int *tets, const int *vert_to_tet, int *out)
{
  for (int v = 0; v < num_verts; ++v)
    out[v] = vert_to_tet[v] < 0 ? -1 : complete(v, (tet_t *)tets, num_tets, vert_to_tet[v]);
}

// src/volume.cpp:13-54 for every vertex in [0, num_verts); -2 where the vertex is in no tet
void ref_volumes(int num_verts, int num_tets, const int *tets, const float *particles,
                 const int *vert_to_tet, float *out)
{
  std::vector<float> cc;
  fill_circumcenters(cc, (tet_t *)tets, num_tets, (float *)particles);
  for (int v = 0; v < num_verts; ++v)
    out[v] = vert_to_tet[v] < 0 ? -2.0f
                                : volume(v, (int *)vert_to_tet, (tet_t *)tets, num_tets, (float *)particles, cc);
}

static void fill_dblock(DBlock *b, const ref_block_t *rb, std::vector<int> &v2t_storage)
{
  memset(static_cast<dblock_t *>(b), 0, sizeof(dblock_t));
  b->gid = rb->gid;
  b->num_orig_particles = rb->num_orig_particles;
  b->num_particles = rb->num_particles;
  b->particles = (float *)rb->particles;
  b->num_tets = rb->num_tets;
  b->tets = (tet_t *)rb->tets;
  if (rb->vert_to_tet)
    b->vert_to_tet = (int *)rb->vert_to_tet;
  else {
    v2t_storage.resize(rb->num_particles);
    ref_fill_vert_to_tet(rb->num_particles, rb->num_tets, rb->tets, v2t_storage.data());
    b->vert_to_tet = v2t_storage.data();
  }
  b->complete = 1;
  for (int i = 0; i < 3; i++) {
    b->bounds.min[i] = rb->bounds_min[i];
    b->bounds.max[i] = rb->bounds_max[i];
  }
}

// the reference's dense() (src/dense.cpp:30-103) over nblocks in-process blocks.
// only_gid >= 0: every other block is given zero original particles (its cells are
// skipped) -- used to time one block per OS process for the multi-core CPU baseline.
// outfile != NULL: also run the reference's WriteGrid (dense.cpp:751-870).
// max_cells >= 0: only the first max_cells cells of every (selected) block are visited.
// ref_set_cell_window(first): the next ref_dense call also skips the cells below `first` (they are handed to the
// reference with vert_to_tet = -1, which it skips at src/dense.cpp:251): with max_cells this gives a window of cells,
// so that one block can be timed on several host cores at once.
static int g_first_cell = 0;
extern "C" void ref_set_cell_window(int first) { g_first_cell = first > 0 ? first : 0; }
int ref_dense(ref_params_t *p, int nblocks, ref_block_t *blocks, int only_gid, const char *outfile, int max_cells)
{
  diy::Master master;
  std::vector<DBlock *> dblocks(nblocks);
  std::vector<std::vector<int> > v2t(nblocks);

  float dmin[3], dmax[3];
  for (int i = 0; i < nblocks; i++)
    for (int d = 0; d < 3; d++) {
      if (i == 0 || blocks[i].bounds_min[d] < dmin[d]) dmin[d] = blocks[i].bounds_min[d];
      if (i == 0 || blocks[i].bounds_max[d] > dmax[d]) dmax[d] = blocks[i].bounds_max[d];
    }

  for (int i = 0; i < nblocks; i++) {
    DBlock *b = new DBlock;
    fill_dblock(b, &blocks[i], v2t[i]);
    if (only_gid >= 0 && blocks[i].gid != only_gid)
      b->num_orig_particles = 0;
    if (max_cells >= 0 && b->num_orig_particles > max_cells)
      b->num_orig_particles = max_cells; // bounded sample for the timed CPU baseline: cells [0, max_cells)
    if (g_first_cell > 0 && b->num_orig_particles > 0) {
      if (b->vert_to_tet != v2t[i].data()) v2t[i].assign(b->vert_to_tet, b->vert_to_tet + b->num_particles);   // the caller's array stays as it is
      b->vert_to_tet = v2t[i].data();
      for (int c = 0; c < g_first_cell && c < b->num_orig_particles; c++) v2t[i][c] = -1;
    }
    for (int d = 0; d < 3; d++) {
      b->data_bounds.min[d] = dmin[d];
      b->data_bounds.max[d] = dmax[d];
    }
    RCLink *l = new RCLink(3, b->bounds, b->bounds);
    for (int j = 0; j < nblocks; j++) {
      if (j == i) continue;
      diy::BlockID id; id.gid = blocks[j].gid; id.proc = 0;
      diy::ContinuousBounds nb(3);
      for (int d = 0; d < 3; d++) { nb.min[d] = blocks[j].bounds_min[d]; nb.max[d] = blocks[j].bounds_max[d]; }
      l->add_neighbor(id);
      l->add_bounds(nb);
      l->add_direction(diy::Direction());
      l->add_wrap(diy::Direction());
    }
    master.add(blocks[i].gid, b, l);
    dblocks[i] = b;
  }

  double t0 = MPI_Wtime();
#ifdef TESSB200_DROPIN
  static tessb200_ctx *ctx = 0;
  try {
    if (!ctx) tessb200::check(tessb200_create(&ctx, 0));
    tessb200::dense((alg)p->alg, p->num_given_bounds, p->given_mins, p->given_maxs, p->project != 0, p->proj_plane,
                    p->mass, p->data_mins, p->data_maxs, p->grid_phys_mins, p->grid_phys_maxs, p->grid_step_size,
                    p->eps, p->glo_num_idx, master, ctx);
  } catch (const tessb200::Error &e) {
    fprintf(stderr, "drp_dense: %s\n", e.what());
    return e.code;
  }
#else
  dense((alg)p->alg, p->num_given_bounds, p->given_mins, p->given_maxs, p->project != 0, p->proj_plane,
        p->mass, p->data_mins, p->data_maxs, p->grid_phys_mins, p->grid_phys_maxs, p->grid_step_size,
        p->eps, p->glo_num_idx, master);
#endif
  p->seconds = MPI_Wtime() - t0;

  int rc = 0;
  for (int i = 0; i < nblocks; i++) {
    int mx[3];
    BlockGridParams(dblocks[i], blocks[i].block_min_idx, mx, blocks[i].block_num_idx, p->grid_phys_mins,
                    p->grid_step_size, p->eps, p->data_mins, p->data_maxs, p->glo_num_idx);
    blocks[i].num_grid_pts = dblocks[i]->num_grid_pts;
    if (blocks[i].density) {
      if (blocks[i].density_capacity >= dblocks[i]->num_grid_pts)
        memcpy(blocks[i].density, dblocks[i]->density, sizeof(float) * (size_t)dblocks[i]->num_grid_pts);
      else
        rc = -1;
    }
  }

  if (outfile) {
    diy::ContiguousAssigner assigner(1, nblocks);
    WriteGrid(nblocks, nblocks, (char *)outfile, p->project != 0, p->glo_num_idx, p->eps, p->data_mins,
              p->data_maxs, p->num_given_bounds, p->given_mins, p->given_maxs, master, assigner);
  }

  for (int i = 0; i < nblocks; i++) {
    delete[] dblocks[i]->density;
    delete dblocks[i];
  }
  g_first_cell = 0;
  return rc;
}

// Per-cell probe: the reference's CellBounds (dense.cpp:657-736) + CellGridPts
// (dense.cpp:1363-1458) for one cell of one block.  Returns the number of grid
// points (0 = rejected by the data-bounds filter, -1 = incomplete / no tet),
// writes up to cap (idx[3], mass) entries, the cell bbox and the face count.
int ref_cell_points(const ref_block_t *rb, int cell, const float *data_mins, const float *data_maxs,
                    const float *grid_phys_mins, const float *grid_step_size, float mass, float eps,
                    int cap, int *out_idx, float *out_mass, float *cell_min, float *cell_max, int *num_faces)
{
  DBlock b;
  std::vector<int> v2t;
  fill_dblock(&b, rb, v2t);
  if (b.vert_to_tet[cell] == -1 || !complete(cell, b.tets, b.num_tets, b.vert_to_tet[cell]))
    return -1;
  vector<float> normals;
  vector<vector<float> > face_verts;
  CellBounds(&b, cell, cell_min, cell_max, normals, face_verts);
  *num_faces = (int)face_verts.size();
  grid_pt_t *grid_pts = NULL;
  int *border = NULL;
  int alloc = 0;
  int n = CellGridPts(cell_min, cell_max, grid_pts, border, alloc, normals, face_verts, (float *)data_mins,
                      (float *)data_maxs, (float *)grid_phys_mins, (float *)grid_step_size, mass, eps,
                      &b.particles[3 * cell]);
  for (int i = 0; i < n && i < cap; i++) {
    out_idx[3 * i] = grid_pts[i].idx[0];
    out_idx[3 * i + 1] = grid_pts[i].idx[1];
    out_idx[3 * i + 2] = grid_pts[i].idx[2];
    out_mass[i] = (float)grid_pts[i].mass;
  }
  if (grid_pts) free(grid_pts);
  if (border) free(border);
  return n;
}

} // extern "C"
