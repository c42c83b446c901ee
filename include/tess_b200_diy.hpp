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
//   est_dense / exchange / recvd_pts               tessb200_dense_run (+ NCCL span exchange between the ranks of master.communicator())
//   DataBounds' MPI_Allreduce, DIY links           join(): layout of every block gathered over master.communicator()
//   b->num_grid_pts                                same
//
// It needs the reference's headers (<tess/dense.hpp>, which pulls in DIY and mpi.h) and is therefore
// only compiled where those exist; the repo's tests build it against oracle/stub's single-process
// DIY stand-in (tests/test_dropin.py).
#ifndef TESS_B200_DIY_HPP
#define TESS_B200_DIY_HPP

#include <cstdlib>
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

// Multi-rank runs (mpiexec -n N, one GPU per rank): what the reference gets from DIY's links and MPI -- every block's
// bounds for DataBounds (src/dense.cpp:1221-1275: min / max all-reduced over the ranks) and the neighbours a deposit can
// be sent to (src/dense.cpp:299-305) -- the library gets as a layout: gid, bounds and owner rank of EVERY block, gathered
// here over the master's communicator.  The NCCL communicator is joined on the first call (unique id from rank 0 by
// MPI_Bcast).  The owner ranks must ascend with the gid (diy::ContiguousAssigner, the assigner the drivers use for
// `tot_blocks >= nprocs`, examples/tess-dense/main.cpp:184-195); a round-robin assignment is refused by the library.
inline void join(diy::Master &master, tessb200_ctx *ctx, const std::vector<tessb200_block> &blocks)
{
  MPI_Comm comm = master.communicator();
  int size = 1, rank = 0;
  MPI_Comm_size(comm, &size);
  MPI_Comm_rank(comm, &rank);
  // one rank: the uploaded blocks are the whole decomposition (TESSB200_DIY_FORCE_JOIN=1 runs the gather and the
  // communicator set-up all the same: the repo's single-process test of this function)
  const char *force = getenv("TESSB200_DIY_FORCE_JOIN");
  if (size == 1 && !(force && force[0] == '1')) return;
  int nloc = (int)blocks.size();
  std::vector<int> counts(size), displs(size + 1, 0);
  MPI_Allgather(&nloc, 1, MPI_INT, counts.data(), 1, MPI_INT, comm);
  for (int r = 0; r < size; r++) displs[r + 1] = displs[r] + counts[r];
  const int total = displs[size];
  std::vector<int> my_gids(nloc > 0 ? nloc : 1), gids(total > 0 ? total : 1), c6(size), d6(size);
  std::vector<float> my_b6(6 * (nloc > 0 ? nloc : 1)), b6(6 * (total > 0 ? total : 1));
  for (int i = 0; i < nloc; i++) {
    my_gids[i] = blocks[i].gid;
    for (int d = 0; d < 3; d++) { my_b6[6 * i + d] = blocks[i].bounds_min[d]; my_b6[6 * i + 3 + d] = blocks[i].bounds_max[d]; }
  }
  for (int r = 0; r < size; r++) { c6[r] = 6 * counts[r]; d6[r] = 6 * displs[r]; }
  MPI_Allgatherv(my_gids.data(), nloc, MPI_INT, gids.data(), counts.data(), displs.data(), MPI_INT, comm);
  MPI_Allgatherv(my_b6.data(), 6 * nloc, MPI_FLOAT, b6.data(), c6.data(), d6.data(), MPI_FLOAT, comm);
  // ascending gid order, owner = the rank the block came from
  std::vector<int> order(total), owner(total);
  for (int r = 0; r < size; r++)
    for (int i = displs[r]; i < displs[r + 1]; i++) { order[i] = i; owner[i] = r; }
  for (int i = 1; i < total; i++)             // insertion sort: the lists are short and nearly sorted
    for (int j = i; j > 0 && gids[order[j]] < gids[order[j - 1]]; j--) { int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t; }
  std::vector<int> s_gids(total), s_owner(total);
  std::vector<float> s_b6(6 * (total > 0 ? total : 1));
  for (int i = 0; i < total; i++) {
    s_gids[i] = gids[order[i]];
    s_owner[i] = owner[order[i]];
    for (int k = 0; k < 6; k++) s_b6[6 * i + k] = b6[6 * order[i] + k];
  }
  if (tessb200_comm_size(ctx) != size || size == 1) {
    unsigned char id[128] = {0};
    if (rank == 0) check(tessb200_comm_unique_id(id));
    MPI_Bcast(id, 128, MPI_BYTE, 0, comm);
    check(tessb200_comm_init(ctx, size, rank, id));
  }
  check(tessb200_dense_set_layout(ctx, total, s_gids.data(), s_b6.data(), s_owner.data()));
}

// Same parameters, same order, same meaning as ::dense (include/tess/dense.hpp:75-89), plus the
// GPU context (one per rank).  Every local block of `master` is processed.  Unlike the reference (void, asserts,
// MPI_Abort) it throws tessb200::Error on failure: catch it at the call site.
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

  join(master, ctx, blocks);
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
