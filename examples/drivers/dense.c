/* dense.c -- command line of the reference's examples/dense/main.cpp (run script DENSE_TEST):
 *   dense <infile> <outfile> <alg> <gx gy gz> <! | px py pz> <mass> <ng> [given mins] [given maxs]
 * Reads the tessellation from a DIY block file (tess_load, src/tess.cpp:139-152), runs the dense stage
 * on GPU 0 and writes the raw C-order float32 grid (WriteGrid, src/dense.cpp:751-870). */
#include "tess_b200.h"
#include "common.h"

int main(int argc, char **argv)
{
  grid_args g;
  if (argc < 10 || parse_grid_args(argc, argv, 4, &g)) {
    fprintf(stderr, "usage: %s infile outfile alg gx gy gz !|px py pz mass ng [mins] [maxs]\n", argv[0]);
    return 2;
  }
  const double t0 = now_s();
  int nblocks = 0;
  tessb200_host_dblock *db = NULL;
  HCHECK(tessb200_host_read_blocks(argv[1], &nblocks, &db, NULL));
  fprintf(stderr, "input time = %.3lf s, %d blocks\n", now_s() - t0, nblocks);
  dense_and_write(atoi(argv[3]), &g, nblocks, db, argv[2]);
  tessb200_host_free_dblocks(nblocks, db);
  return 0;
}
