// tess_b200_diy.hpp -- source-compatible replacement of tess2's dense() for DIY builds.
//
// Include this from a tess2 driver (examples/dense/main.cpp:170, examples/tess-dense/main.cpp:216)
// instead of calling ::dense(): the signature, the out-parameters and the ownership of
// DBlock::density are those of include/tess/dense.hpp:75-89 / src/dense.cpp:30-127.  All compute
// happens on the GPU behind the C ABI in tess_b200.h; this header only walks the diy::Master.
//
//   reference                                      here
//   DataBounds + GridStepParams (dense.cpp:53-57)  tessb200_dense_geometry (host, same fp32 order)
//   init_dense: b->density = new float[npts]       same (new[]: destroy_block delete[]s it, tess.cpp:177)
//   est_dense / exchange / recvd_pts               tessb200_dense_run (+ NCCL when a communicator was joined)
//   b->num_grid_pts                                same
//
// It needs the reference's headers (<tess/dense.hpp>, which pulls in DIY and mpi.h) and is therefore
// only compiled where those exist; the repo's tests build it against oracle/stub's single-process
// DIY stand-in (tests/test_dropin.py).
#ifndef TESS_B200_DIY_HPP
#define TESS_B200_DIY_HPP

#include <stdexcept>
#include <string>
#include <vector>
#include "tess/dense.hpp"
#include "tess_b200.h"

namespace tessb200
{

struct Error : std::runtime_error
{
  int code;
  Error(int c, const char *msg) : std::runtime_error(std::string("tess_b200: ") + msg), code(c) {}
};

inline void check(int rc)
{
  if (rc != 0) throw Error(rc, tessb200_last_error());
}

// Same parameters, same order, same meaning as ::dense (include/tess/dense.hpp:75-89), plus the
// GPU context (one per rank).  Every local block of `master` is processed.
inline void dense(alg alg_type, int num_given_bounds, float *given_mins, float *given_maxs, bool project, float *proj_plane,
                  float mass, float *data_mins, float *data_maxs, float *grid_phys_mins, float *grid_phys_maxs,
                  float *grid_step_size, float eps, int *glo_num_idx, diy::Master &master, tessb200_ctx *ctx,
                  tessb200_dense_stats *stats = 0)
{
  const int nblocks = (int)master.size();
  std::vector<tessb200_block> blocks(nblocks);
  for (int i = 0; i < nblocks; i++) {
    DBlock *b = master.block<DBlock>(i);
    tessb200_block &t = blocks[i];
    t = tessb200_block();
    t.gid = b->gid;
    t.num_orig_particles = b->num_orig_particles;
    t.num_particles = b->num_particles;
    t.particles = b->particles;
    t.num_tets = b->num_tets;
    t.tets = reinterpret_cast<const int *>(b->tets);   // struct tet_t { int verts[4]; int tets[4]; }
    t.vert_to_tet = b->vert_to_tet;
    for (int d = 0; d < 3; d++) {
      t.bounds_min[d] = b->bounds.min[d];
      t.bounds_max[d] = b->bounds.max[d];
    }
  }
  tessb200_dense_params p = tessb200_dense_params();
  p.alg = alg_type == DENSE_CIC ? TESSB200_DENSE_CIC : TESSB200_DENSE_TESS;
  p.num_given_bounds = num_given_bounds;
  for (int d = 0; d < 3; d++) {
    p.given_mins[d] = given_mins ? given_mins[d] : 0.0f;
    p.given_maxs[d] = given_maxs ? given_maxs[d] : 0.0f;
    p.proj_plane[d] = proj_plane ? proj_plane[d] : (d == 2 ? 1.0f : 0.0f);
    p.glo_num_idx[d] = glo_num_idx[d];
  }
  p.project = project ? 1 : 0;
  p.mass = mass;
  p.eps = eps;

  check(tessb200_dense_upload(ctx, nblocks, blocks.data()));
  check(tessb200_dense_geometry(ctx, &p, nblocks, blocks.data()));
  // init_dense (src/dense.cpp:106-127): the block owns a new[]-allocated density array
  for (int i = 0; i < nblocks; i++) {
    DBlock *b = master.block<DBlock>(i);
    b->density = new float[blocks[i].num_grid_pts > 0 ? blocks[i].num_grid_pts : 1];
    b->num_grid_pts = (int)blocks[i].num_grid_pts;
    blocks[i].density = b->density;
    blocks[i].density_capacity = blocks[i].num_grid_pts;
  }
  check(tessb200_dense_run(ctx, &p, stats));
  check(tessb200_dense_download(ctx, nblocks, blocks.data(), 0));
  for (int d = 0; d < 3; d++) {
    data_mins[d] = p.data_mins[d];
    data_maxs[d] = p.data_maxs[d];
    grid_phys_mins[d] = p.grid_phys_mins[d];
    grid_phys_maxs[d] = p.grid_phys_maxs[d];
    grid_step_size[d] = p.grid_step_size[d];
  }
}

} // namespace tessb200

#endif
