/* bench_dense.c -- stage times of the resident dense path without Python: BASELINE config 2's particles
 * (gen_particles per block, srand(gid)), the repo's host tess(), upload once, then timed tessb200_dense_run calls.
 * For A/B experiments on the GPU box in seconds (bench.py is the number that counts; its tets come from SciPy's
 * Qhull, these from the repo's engine: the same set, numbered differently).
 *   bench_dense [side=128] [blocks=8] [gsize=2*side] [steps=5] [warmup=3] [alg=0] [project=0]
 * Build: make -C profiles/quick */
#include "tess_b200.h"
#include "../../examples/drivers/common.h"

int main(int argc, char **argv)
{
  const int side = argc > 1 ? atoi(argv[1]) : 128, tb = argc > 2 ? atoi(argv[2]) : 8;
  const int gsize = argc > 3 && atoi(argv[3]) > 0 ? atoi(argv[3]) : 2 * side, steps = argc > 4 ? atoi(argv[4]) : 5, warmup = argc > 5 ? atoi(argv[5]) : 3;
  const int alg = argc > 6 ? atoi(argv[6]) : 0, project = argc > 7 ? atoi(argv[7]) : 0;
  const int dsize[3] = {side, side, side};
  double tess_s = 0.0;
  tessb200_host_dblock *db = generate_and_tess(tb, dsize, 0, 0, -1.0f, -1.0f, &tess_s);
  tessb200_ctx *ctx;
  GCHECK(tessb200_create(&ctx, 0));
  tessb200_block *blk = (tessb200_block *)calloc((size_t)tb, sizeof(*blk));
  for (int i = 0; i < tb; i++) {
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
  p.alg = alg; p.project = project; p.proj_plane[2] = 1.0f; p.mass = 1.0f; p.eps = 0.0001f;
  p.glo_num_idx[0] = p.glo_num_idx[1] = p.glo_num_idx[2] = gsize;
  GCHECK(tessb200_dense_upload(ctx, tb, blk));
  tessb200_dense_stats st, acc;
  memset(&acc, 0, sizeof(acc));
  for (int i = 0; i < warmup; i++) GCHECK(tessb200_dense_run(ctx, &p, &st));
  for (int i = 0; i < steps; i++) {
    GCHECK(tessb200_dense_run(ctx, &p, &st));
    acc.ms_circumcenters += st.ms_circumcenters; acc.ms_bfs += st.ms_bfs; acc.ms_nbrs += st.ms_nbrs; acc.ms_faces += st.ms_faces;
    acc.ms_scan += st.ms_scan; acc.ms_sort += st.ms_sort; acc.ms_deposit += st.ms_deposit; acc.ms_slow_path += st.ms_slow_path;
    acc.ms_total_device += st.ms_total_device; acc.ms_fused += st.ms_fused; acc.ms_emit += st.ms_emit;
  }
  const double k = 1.0 / (steps > 0 ? steps : 1), ms = acc.ms_total_device * k;
  const double g = project ? (double)gsize * gsize : (double)gsize * gsize * gsize;
  printf("{\"side\": %d, \"blocks\": %d, \"gsize\": %d, \"alg\": %d, \"project\": %d, \"steps\": %d, \"tess_s\": %.3f, \"tets\": %lld, \"cells\": %lld, "
         "\"deposit_cells\": %lld, \"spans\": %lld, \"shared_deposits\": %lld, \"slow_cells\": %lld, \"launches\": %lld, \"tot_mass\": %.6f, "
         "\"ms\": {\"fused\": %.3f, \"emit\": %.3f, \"cc\": %.3f, \"bfs\": %.3f, \"nbrs\": %.3f, \"faces\": %.3f, \"scan\": %.3f, \"sort\": %.3f, \"deposit\": %.3f, \"slow\": %.3f, \"total\": %.3f}, "
         "\"grid_points_per_sec\": %.4e}\n",
         side, tb, gsize, alg, project, steps, tess_s, (long long)st.num_tets, (long long)st.num_cells, (long long)st.num_deposit_cells, (long long)st.num_spans,
         (long long)st.num_shared_deposits, (long long)st.num_slow_cells, (long long)st.num_kernel_launches, st.tot_mass, acc.ms_fused * k, acc.ms_emit * k, acc.ms_circumcenters * k, acc.ms_bfs * k,
         acc.ms_nbrs * k, acc.ms_faces * k, acc.ms_scan * k, acc.ms_sort * k, acc.ms_deposit * k, acc.ms_slow_path * k, ms, ms > 0 ? g / (ms * 1e-3) : 0.0);
  tessb200_destroy(ctx);
  free(blk);
  tessb200_host_free_dblocks(tb, db);
  return 0;
}
