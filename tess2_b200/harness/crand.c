/* glibc srand()/rand() stream into a buffer: the reference's gen_particles (src/tess.cpp:281-293)
 * draws 3 * n values after srand(gid).  Host-side input staging only. */
#include <stdlib.h>
void crand_fill(unsigned seed, long long n, int *out)
{
  srand(seed);
  for (long long i = 0; i < n; i++) out[i] = rand();
}
