/* common.h -- shared by the three drivers in this directory, which take the command lines of the
 * reference's example drivers so that its run scripts work with the executable swapped:
 *   tess        examples/tess/main.cpp        (TESS_TEST)        particles -> tess() -> del.out
 *   dense       examples/dense/main.cpp       (DENSE_TEST)       del.out -> dense() -> dense.raw
 *   tess-dense  examples/tess-dense/main.cpp  (TESS_DENSE_TEST)  particles -> tess() -> dense() -> dense.raw
 * One process, plain C over include/tess_b200_host.h (CPU) and include/tess_b200.h (GPU); no MPI, no DIY.
 */
#ifndef TESSB200_EXAMPLE_COMMON_H
#define TESSB200_EXAMPLE_COMMON_H
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "tess_b200_host.h"

#define HCHECK(x) do { int rc_ = (x); if (rc_) { fprintf(stderr, "%s failed (%d): %s\n", #x, rc_, tessb200_host_last_error()); exit(1); } } while (0)

static double now_s(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* the grid arguments shared by `dense` and `tess-dense` from argv[a] on: gx gy gz, then "!" (3-D) or the
 * normal of the projection plane, mass, number of given bounds and the bounds (mins first, then maxs);
 * examples/dense/main.cpp:24-96, examples/tess-dense/main.cpp:60-133 */
typedef struct grid_args { int gsize[3]; int project; float proj_plane[3]; float mass; int ng; float gmin[3], gmax[3]; } grid_args;

__attribute__((unused)) static int parse_grid_args(int argc, char **argv, int a, grid_args *g)
{
  memset(g, 0, sizeof(*g));
  g->proj_plane[2] = 1.0f;
  if (argc < a + 6) return -1;
  for (int d = 0; d < 3; d++) g->gsize[d] = atoi(argv[a + d]);
  if (g->gsize[0] < 2 || g->gsize[1] < 2 || g->gsize[2] < 2) return -1;
  a += 3;
  if (strcmp(argv[a], "!")) {
    if (argc < a + 5) return -1;
    g->project = 1;
    for (int d = 0; d < 3; d++) g->proj_plane[d] = (float)atof(argv[a + d]);
    /* unit length, as both drivers do before the call (examples/dense/main.cpp:124-130) */
    const float len = sqrtf(g->proj_plane[0] * g->proj_plane[0] + g->proj_plane[1] * g->proj_plane[1] + g->proj_plane[2] * g->proj_plane[2]);
    for (int d = 0; d < 3; d++) g->proj_plane[d] /= len;
    a += 3;
  } else {
    a += 1;
  }
  g->mass = (float)atof(argv[a]);
  g->ng = atoi(argv[a + 1]);
  a += 2;
  if (g->ng < 0 || g->ng > 3 || argc < a + 2 * g->ng) return -1;
  for (int d = 0; d < g->ng; d++) { g->gmin[d] = (float)atof(argv[a + d]); g->gmax[d] = (float)atof(argv[a + g->ng + d]); }
  return 0;
}

/* AddAndGenerate + tess() of the two generating drivers: domain [0, dsize-1]^3, tb regular blocks, every block
 * draws (extent + 1)^3 points in its bounds with srand(gid) / rand() (gen_particles, src/tess.cpp:264-293; the
 * jitter argument is unused there).  Returns tb blocks in dblock form, ready for tess_save or dense(). */
__attribute__((unused)) static tessb200_host_dblock *generate_and_tess(int tb, const int *dsize, int wrap, int walls, float minvol, float maxvol, double *seconds)
{
  if (tb < 1 || dsize[0] < 2 || dsize[1] < 2 || dsize[2] < 2) { fprintf(stderr, "need tot_blocks >= 1 and a domain of at least 2 x 2 x 2\n"); exit(2); }
  /* walls: parsed by the reference drivers and read by nothing (examples/tess/main.cpp:46, examples/tess-dense/main.cpp:121; no
   * wall_particles in src/): accepted and without effect here too.  wrap: periodic neighbours -- the images of particles shifted
   * by whole domain extents become ghosts (tessb200_host_tess_periodic). */
  if (walls) fprintf(stderr, "note: walls has no effect (the reference parses it and never reads it)\n");
  if (minvol > 0.0f || maxvol > 0.0f) fprintf(stderr, "note: minvol / maxvol do not act on the tets handed to dense(); ignored\n");
  const float dmin[3] = {0, 0, 0}, dmax[3] = {dsize[0] - 1.0f, dsize[1] - 1.0f, dsize[2] - 1.0f};
  float *bounds = (float *)malloc(sizeof(float) * 6 * (size_t)tb);
  HCHECK(tessb200_host_regular_blocks(dmin, dmax, tb, bounds));
  size_t cap = 0;
  for (int g = 0; g < tb; g++) {
    size_t n = 1;
    for (int d = 0; d < 3; d++) n *= (size_t)(int)(bounds[6 * g + 3 + d] - bounds[6 * g + d] + 1);
    cap += n;
  }
  float *xyz = (float *)malloc(sizeof(float) * 3 * (cap ? cap : 1));
  int *owner = (int *)malloc(sizeof(int) * (cap ? cap : 1));
  size_t np = 0;
  for (int g = 0; g < tb; g++) {
    size_t n = 1;
    for (int d = 0; d < 3; d++) n *= (size_t)(int)(bounds[6 * g + 3 + d] - bounds[6 * g + d] + 1);
    srand((unsigned)g);
    for (size_t i = 0; i < n; i++, np++) {
      for (int d = 0; d < 3; d++) {
        const float t = (float)rand() / RAND_MAX;
        xyz[3 * np + d] = t * (bounds[6 * g + 3 + d] - bounds[6 * g + d]) + bounds[6 * g + d];
      }
      owner[np] = g;
    }
  }
  const double t0 = now_s();
  tessb200_host_block *hb = (tessb200_host_block *)calloc((size_t)tb, sizeof(*hb));
  HCHECK(tessb200_host_tess_periodic((int)np, xyz, owner, dmin, dmax, tb, bounds, 0, NULL, 0.0f, wrap ? 6 : 0, wrap ? 8.0f : 0.0f, 0, wrap ? 1 : 0, hb));
  if (seconds) *seconds = now_s() - t0;
  tessb200_host_dblock *db = (tessb200_host_dblock *)calloc((size_t)tb, sizeof(*db));
  long long ntets = 0, nghost = 0;
  /* a block keeps its originals in input order, so a particle's local id is its rank among its owner's */
  int *lid = (int *)malloc(sizeof(int) * (np ? np : 1)), *count = (int *)calloc((size_t)tb, sizeof(int));
  for (size_t i = 0; i < np; i++) lid[i] = count[owner[i]]++;
  for (int g = 0; g < tb; g++) {
    db[g].gid = hb[g].gid;
    for (int d = 0; d < 3; d++) {
      db[g].bounds_min[d] = db[g].box_min[d] = hb[g].bounds_min[d];
      db[g].bounds_max[d] = db[g].box_max[d] = hb[g].bounds_max[d];
      db[g].data_min[d] = dmin[d];
      db[g].data_max[d] = dmax[d];
    }
    db[g].num_orig_particles = hb[g].num_orig_particles;
    db[g].num_particles = hb[g].num_particles;
    db[g].num_tets = hb[g].num_tets;
    db[g].particles = hb[g].particles;          /* ownership moves to the dblock (all malloc'ed) */
    db[g].tets = hb[g].tets;
    db[g].vert_to_tet = hb[g].vert_to_tet;
    /* rem_gids / rem_lids of the ghosts (src/tess.cpp:654-673): owner block and index there */
    const int ng = hb[g].num_particles - hb[g].num_orig_particles;
    db[g].rem_gids = (int *)malloc(sizeof(int) * (size_t)(ng ? ng : 1));
    db[g].rem_lids = (int *)malloc(sizeof(int) * (size_t)(ng ? ng : 1));
    for (int i = 0; i < ng; i++) {
      const int gi = hb[g].global_ids[hb[g].num_orig_particles + i];
      db[g].rem_gids[i] = owner[gi];
      db[g].rem_lids[i] = lid[gi];
    }
    free(hb[g].global_ids);
    hb[g].global_ids = NULL;
    db[g].complete = 1;
    if (!hb[g].settled)
      fprintf(stderr, "tess: WARNING: block %d: the ghost region stopped growing (margin %g, %d rounds) before every original cell was settled; "
                      "cells near its border may be wrong\n", hb[g].gid, (double)hb[g].ghost_margin, hb[g].rounds);
    ntets += hb[g].num_tets;
    nghost += ng;
  }
  free(count);
  free(lid);
  fprintf(stderr, "tess: %zu particles in %d blocks, %lld ghosts, %lld tets\n", np, tb, nghost, ntets);
  free(hb); free(xyz); free(owner); free(bounds);
  return db;
}

#ifdef TESS_B200_H
#define GCHECK(x) do { int rc_ = (x); if (rc_) { fprintf(stderr, "%s failed (%d): %s\n", #x, rc_, tessb200_last_error()); exit(1); } } while (0)

/* dense() + WriteGrid + dense_stats of the two density drivers (examples/dense/main.cpp:170-193) */
__attribute__((unused)) static void dense_and_write(int alg, const grid_args *g, int nblocks, const tessb200_host_dblock *db, const char *outfile)
{
  tessb200_ctx *ctx;
  GCHECK(tessb200_create(&ctx, 0));
  tessb200_block *blk = (tessb200_block *)calloc((size_t)(nblocks ? nblocks : 1), sizeof(*blk));
  for (int i = 0; i < nblocks; i++) {
    blk[i].gid = db[i].gid;
    blk[i].num_orig_particles = db[i].num_orig_particles;
    blk[i].num_particles = db[i].num_particles;
    blk[i].particles = db[i].particles;
    blk[i].num_tets = db[i].num_tets;
    blk[i].tets = db[i].tets;
    blk[i].vert_to_tet = db[i].vert_to_tet;
    memcpy(blk[i].bounds_min, db[i].bounds_min, 12);
    memcpy(blk[i].bounds_max, db[i].bounds_max, 12);
  }
  tessb200_dense_params p;
  memset(&p, 0, sizeof(p));
  p.alg = alg ? TESSB200_DENSE_CIC : TESSB200_DENSE_TESS;
  p.num_given_bounds = g->ng;
  memcpy(p.given_mins, g->gmin, 12);
  memcpy(p.given_maxs, g->gmax, 12);
  p.project = g->project;
  memcpy(p.proj_plane, g->proj_plane, 12);
  p.mass = g->mass;
  p.eps = 0.0001f;                               /* both drivers' eps */
  memcpy(p.glo_num_idx, g->gsize, 12);
  const double t0 = now_s();
  GCHECK(tessb200_dense_geometry(ctx, &p, nblocks, blk));
  for (int i = 0; i < nblocks; i++) {
    blk[i].density = (float *)malloc(sizeof(float) * (size_t)(blk[i].num_grid_pts > 0 ? blk[i].num_grid_pts : 1));
    blk[i].density_capacity = blk[i].num_grid_pts;
  }
  tessb200_dense_stats st;
  GCHECK(tessb200_dense(ctx, &p, nblocks, blk, NULL, &st));
  const double t1 = now_s();
  if (outfile && outfile[0]) GCHECK(tessb200_write_grid(outfile, &p, nblocks, blk));
  const double t2 = now_s();
  /* dense_stats (src/dense.cpp:1284-1333) */
  fprintf(stderr, "----------------- global stats ------------------\n");
  fprintf(stderr, "comp time = %.3lf s (device %.3lf ms), output time = %.3lf s\n", t1 - t0, (double)st.ms_total_device, t2 - t1);
  fprintf(stderr, "grid size = %d x %d x %d, step = %.4e %.4e %.4e, min corner = %.4e %.4e %.4e\n", p.glo_num_idx[0], p.glo_num_idx[1],
          p.glo_num_idx[2], p.grid_step_size[0], p.grid_step_size[1], p.grid_step_size[2], p.grid_phys_mins[0], p.grid_phys_mins[1],
          p.grid_phys_mins[2]);
  fprintf(stderr, "max density = %.4e, total mass = %.6lf, depositing cells = %lld (x mass %g = %.6lf)\n", (double)st.max_dense, st.tot_mass,
          (long long)st.num_deposit_cells, (double)g->mass, (double)st.num_deposit_cells * (double)g->mass);
  fprintf(stderr, "-------------------------------------------------\n");
  for (int i = 0; i < nblocks; i++) free(blk[i].density);
  free(blk);
  tessb200_destroy(ctx);
}
#endif
#endif
