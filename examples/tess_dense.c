/* tess_dense.c -- the flow of tess2's examples/tess-dense/main.cpp (particles -> blocks -> tess() ->
 * dense() -> dense.raw) for one process, through the two C ABIs of this repository only:
 * include/tess_b200_host.h (CPU: decomposition, ghosts, Delaunay) and include/tess_b200.h (GPU: dense).
 *
 *   tess_dense <points per block side> <blocks (power of two for kd)> <grid size> <outfile> [alg=0] [project=0] [kd=0]
 *
 * Particles as the reference's test driver makes them (src/tess.cpp:264-293): block gid draws
 * (side)^3 points in its bounds with srand(gid) / rand().  Build: see tests/test_gpu_parity.py.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tess_b200.h"
#include "tess_b200_host.h"

#define CHECK(x) do { int rc_ = (x); if (rc_) { fprintf(stderr, "%s failed (%d): %s / %s\n", #x, rc_, tessb200_last_error(), tessb200_host_last_error()); return 1; } } while (0)

int main(int argc, char **argv)
{
  if (argc < 5) { fprintf(stderr, "usage: %s side nblocks gsize outfile [alg] [project] [kd]\n", argv[0]); return 2; }
  const int side = atoi(argv[1]), nblocks = atoi(argv[2]), gsize = atoi(argv[3]);
  const char *outfile = argv[4];
  const int alg = argc > 5 ? atoi(argv[5]) : 0, project = argc > 6 ? atoi(argv[6]) : 0, kd = argc > 7 ? atoi(argv[7]) : 0;

  /* domain and a regular decomposition; each block generates its own particles (gen_particles) */
  int dims[3] = {1, 1, 1};
  for (int n = nblocks, d = 0; n > 1; n /= 2, d = (d + 1) % 3) dims[d] *= 2;
  const float dmin[3] = {0, 0, 0}, dmax[3] = {(float)(dims[0] * side - 1), (float)(dims[1] * side - 1), (float)(dims[2] * side - 1)};
  float *bounds = (float *)malloc(sizeof(float) * 6 * nblocks);
  CHECK(tessb200_host_regular_blocks(dmin, dmax, nblocks, bounds));
  size_t cap = 0;
  for (int g = 0; g < nblocks; g++) {
    size_t n = 1;
    for (int d = 0; d < 3; d++) n *= (size_t)(int)((bounds[6 * g + 3 + d] - bounds[6 * g + d]) + 1.0f);
    cap += n;
  }
  float *xyz = (float *)malloc(sizeof(float) * 3 * cap);
  int *owner = (int *)malloc(sizeof(int) * cap);
  size_t np = 0;
  for (int g = 0; g < nblocks; g++) {
    int sz[3];
    for (int d = 0; d < 3; d++) sz[d] = (int)((bounds[6 * g + 3 + d] - bounds[6 * g + d]) + 1.0f);
    srand((unsigned)g);
    const size_t n = (size_t)sz[0] * sz[1] * sz[2];
    for (size_t i = 0; i < n; i++, np++) {
      for (int d = 0; d < 3; d++) {
        const float t = (float)rand() / (float)RAND_MAX;
        xyz[3 * np + d] = t * (bounds[6 * g + 3 + d] - bounds[6 * g + d]) + bounds[6 * g + d];
      }
      owner[np] = g;
    }
  }
  if (kd) CHECK(tessb200_host_kdtree_blocks((int)np, xyz, dmin, dmax, nblocks, bounds, owner));   /* tess2's -kd option */

  /* tess(): blocks with ghosts, tets, vert_to_tet */
  tessb200_host_block *hb = (tessb200_host_block *)calloc(nblocks, sizeof(*hb));
  CHECK(tessb200_host_tess((int)np, xyz, owner, dmin, dmax, nblocks, bounds, 0, NULL, 0.0f, 0, 0.0f, 0, hb));

  /* dense() */
  tessb200_ctx *ctx;
  CHECK(tessb200_create(&ctx, 0));
  tessb200_block *blk = (tessb200_block *)calloc(nblocks, sizeof(*blk));
  long long ntets = 0;
  for (int g = 0; g < nblocks; g++) {
    blk[g].gid = hb[g].gid;
    blk[g].num_orig_particles = hb[g].num_orig_particles; blk[g].num_particles = hb[g].num_particles; blk[g].particles = hb[g].particles;
    blk[g].num_tets = hb[g].num_tets; blk[g].tets = hb[g].tets; blk[g].vert_to_tet = hb[g].vert_to_tet;
    memcpy(blk[g].bounds_min, hb[g].bounds_min, 12); memcpy(blk[g].bounds_max, hb[g].bounds_max, 12);
    ntets += hb[g].num_tets;
  }
  tessb200_dense_params p;
  memset(&p, 0, sizeof(p));
  p.alg = alg; p.project = project; p.proj_plane[2] = 1.0f; p.mass = 1.0f; p.eps = 1e-4f;
  p.glo_num_idx[0] = p.glo_num_idx[1] = p.glo_num_idx[2] = gsize;
  CHECK(tessb200_dense_geometry(ctx, &p, nblocks, blk));       /* BlockGridParams: sizes of the per-block arrays */
  for (int g = 0; g < nblocks; g++) {
    blk[g].density = (float *)malloc(sizeof(float) * (size_t)(blk[g].num_grid_pts > 0 ? blk[g].num_grid_pts : 1));
    blk[g].density_capacity = blk[g].num_grid_pts;
  }
  tessb200_dense_stats st;
  CHECK(tessb200_dense(ctx, &p, nblocks, blk, NULL, &st));
  CHECK(tessb200_write_grid(outfile, &p, nblocks, blk));
  printf("particles %zu tets %lld cells deposited %lld total mass %.6f max density %.6g device ms %.3f\n", np, ntets,
         (long long)st.num_deposit_cells, st.tot_mass, st.max_dense, st.ms_total_device);
  for (int g = 0; g < nblocks; g++) { free(blk[g].density); tessb200_host_free_block(&hb[g]); }
  tessb200_destroy(ctx);
  free(blk); free(hb); free(xyz); free(owner); free(bounds);
  return 0;
}
