/* tess-dense.c -- command line of the reference's examples/tess-dense/main.cpp (run script TESS_DENSE_TEST):
 *   tess-dense <alg> <tot_blocks> <dsize x y z> <jitter> <minvol> <maxvol> <wrap> <walls> <outfile | !>
 *              <gx gy gz> <! | px py pz> <mass> <ng> [given mins] [given maxs]
 * Particles -> host tess() -> GPU dense() -> raw grid, without the intermediate file. */
#include "tess_b200.h"
#include "common.h"

int main(int argc, char **argv)
{
  grid_args g;
  if (argc < 18 || parse_grid_args(argc, argv, 12, &g)) {
    fprintf(stderr, "usage: %s alg tot_blocks dx dy dz jitter minvol maxvol wrap walls outfile|! gx gy gz !|px py pz mass ng [mins] [maxs]\n", argv[0]);
    return 2;
  }
  const int alg = atoi(argv[1]), tb = atoi(argv[2]);
  const int dsize[3] = {atoi(argv[3]), atoi(argv[4]), atoi(argv[5])};
  const char *outfile = argv[11][0] == '!' ? "" : argv[11];
  const double t0 = now_s();
  double tess_s = 0.0;
  tessb200_host_dblock *db = generate_and_tess(tb, dsize, atoi(argv[9]), atoi(argv[10]), (float)atof(argv[7]), (float)atof(argv[8]), &tess_s);
  const double t1 = now_s();
  dense_and_write(alg, &g, tb, db, outfile);
  const double t2 = now_s();
  fprintf(stderr, "Overall time = %.3lf s = %.3lf s tess + %.3lf s dense\n", t2 - t0, t1 - t0, t2 - t1);
  tessb200_host_free_dblocks(tb, db);
  return 0;
}
