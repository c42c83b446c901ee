/* tess.c -- command line of the reference's examples/tess/main.cpp (run script TESS_TEST):
 *   tess <tot_blocks> <mem_blocks> <dsize x y z> <jitter> <minvol> <maxvol> <wrap> <walls> <outfile | !>
 * Generates the test particles, tessellates them block by block on the host and writes the DIY block
 * file the reference's `dense` driver reads (tess_save, src/tess.cpp:126-137).  CPU only. */
#include "common.h"

int main(int argc, char **argv)
{
  if (argc < 12) {
    fprintf(stderr, "usage: %s tot_blocks mem_blocks dx dy dz jitter minvol maxvol wrap walls outfile|!\n", argv[0]);
    return 2;
  }
  const int tb = atoi(argv[1]);
  const int dsize[3] = {atoi(argv[3]), atoi(argv[4]), atoi(argv[5])};
  const float minvol = (float)atof(argv[7]), maxvol = (float)atof(argv[8]);
  const int wrap = atoi(argv[9]), walls = atoi(argv[10]);
  const char *outfile = argv[11][0] == '!' ? "" : argv[11];
  double tess_s = 0.0;
  tessb200_host_dblock *db = generate_and_tess(tb, dsize, wrap, walls, minvol, maxvol, &tess_s);
  const double t0 = now_s();
  if (outfile[0]) HCHECK(tessb200_host_write_blocks(outfile, tb, db, TESSB200_DIY_BOUNDS_DYNAMIC, NULL, 0));
  fprintf(stderr, "tess time = %.3lf s, output time = %.3lf s\n", tess_s, now_s() - t0);
  tessb200_host_free_dblocks(tb, db);
  return 0;
}
