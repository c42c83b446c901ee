// host_geom.hpp -- the grid bookkeeping the reference does on the host, restated in its fp32
// operation order.  Shared by api.cu and by the test-only CPU emulation in tests/emul.
// Must be compiled without FMA contraction (-fmad=false / -ffp-contract=off).
#pragma once
#include <math.h>
#include <float.h>
#include "../../include/tess_b200.h"
#include "cell_core.cuh"

namespace tb
{
// ---- host-side grid bookkeeping -----------------------------------------------------------------
// GridStepParams, src/dense.cpp:1712-1767
inline void grid_step_params(tessb200_dense_params *p)
{
  float mds = p->data_maxs[0] - p->data_mins[0];
  if (p->data_maxs[1] - p->data_mins[1] > mds) mds = p->data_maxs[1] - p->data_mins[1];
  if (p->data_maxs[2] - p->data_mins[2] > mds) mds = p->data_maxs[2] - p->data_mins[2];
  for (int i = 0; i < 3; i++) {
    float pad = mds - (p->data_maxs[i] - p->data_mins[i]);
    p->grid_phys_mins[i] = (float)((double)p->data_mins[i] - (double)pad / 2.0);
    p->grid_phys_maxs[i] = (float)((double)p->data_maxs[i] + (double)pad / 2.0);
  }
  for (int i = 0; i < 3 && i < p->num_given_bounds; i++) {
    p->grid_phys_mins[i] = p->given_mins[i];
    p->grid_phys_maxs[i] = p->given_maxs[i];
  }
  for (int i = 0; i < 3; i++)
    p->grid_step_size[i] = (p->grid_phys_maxs[i] - p->grid_phys_mins[i]) / (float)(p->glo_num_idx[i] - 1);
}

// BlockGridParams, src/dense.cpp:575-648
inline void block_grid_params(const float *bmin, const float *bmax, const tessb200_dense_params *p, int *mn, int *num)
{
  const float *step = p->grid_step_size, *gmin = p->grid_phys_mins;
  int mx[3];
  for (int i = 0; i < 3; i++) {
    mn[i] = phys2idx1(bmin[i], step[i], gmin[i]);
    if (idx2phys1(mn[i], step[i], gmin[i]) < bmin[i]) mn[i]++;
    mx[i] = phys2idx1(bmax[i], step[i], gmin[i]);
    if (idx2phys1(mx[i], step[i], gmin[i]) + step[i] <= bmax[i]) mx[i]++;
    if (fabsf(p->data_mins[i] + (float)mx[i] * step[i] - bmax[i]) < p->eps && fabsf(bmax[i] - p->data_maxs[i]) > step[i]) mx[i]--;
    if (fabsf(bmin[i] - p->data_mins[i]) < step[i]) mn[i] = 0;
    if (fabsf(bmax[i] - p->data_maxs[i]) < step[i]) mx[i] = p->glo_num_idx[i] - 1;
    num[i] = mx[i] - mn[i] + 1;
  }
}

// { idx : bmin <= idx2phys(idx) <= bmax } as an inclusive index interval: the integer form of the
// closed-bounds test of src/dense.cpp:279-284 (idx2phys is monotone in idx).  Indices are kept within
// two points of the grid -- a point further out has no array element to land in -- except along z when
// projecting: the projected index drops z (src/dense.cpp:1047-1090), so a point of any z index deposits
// as long as its position lies in the block (a given z range narrower than the data).
inline void phys_box(const float *bmin, const float *bmax, const tessb200_dense_params *p, int *lo, int *hi)
{
  for (int d = 0; d < 3; d++) {
    float step = p->grid_step_size[d], gmin = p->grid_phys_mins[d];
    const bool free_z = p->project && d == 2;
    const int lim_lo = free_z ? -(1 << 22) : -2, lim_hi = free_z ? (1 << 22) : p->glo_num_idx[d] + 2;
    float ea = (bmin[d] - gmin) / step, eb = (bmax[d] - gmin) / step;      // estimates, clamped before the cast
    ea = fminf(fmaxf(ea, (float)lim_lo), (float)lim_hi);
    eb = fminf(fmaxf(eb, (float)lim_lo), (float)lim_hi);
    int a = (int)ea - 2;
    if (a < lim_lo) a = lim_lo;
    if (a > lim_hi) a = lim_hi;
    while (a > lim_lo && idx2phys1(a - 1, step, gmin) >= bmin[d]) a--;
    while (a <= lim_hi && idx2phys1(a, step, gmin) < bmin[d]) a++;
    int b = (int)eb + 2;
    if (b > lim_hi) b = lim_hi;
    if (b < lim_lo) b = lim_lo;
    while (b < lim_hi && idx2phys1(b + 1, step, gmin) <= bmax[d]) b++;
    while (b >= lim_lo && idx2phys1(b, step, gmin) > bmax[d]) b--;
    lo[d] = a;
    hi[d] = b;
  }
}

inline int ceil_log2(unsigned long long v)
{
  int b = 0;
  while (b < 64 && (1ull << b) < v) b++;
  return b;
}

// z slots of the sort key (projections): the span of z indices that any block's closed bounds hold
inline void key_z_range(const BlockBox *boxes, size_t nblocks, int project, KeyLayout *kl)
{
  kl->z_bits = 0;
  kl->z_lo = 0;
  if (!project || nblocks == 0) return;
  int zlo = boxes[0].p_lo[2], zhi = boxes[0].p_hi[2];
  for (size_t i = 1; i < nblocks; i++) {
    if (boxes[i].p_lo[2] < zlo) zlo = boxes[i].p_lo[2];
    if (boxes[i].p_hi[2] > zhi) zhi = boxes[i].p_hi[2];
  }
  if (zhi < zlo) zhi = zlo;
  kl->z_lo = zlo;
  kl->z_bits = ceil_log2((unsigned long long)((long long)zhi - zlo) + 1);
  if (kl->z_bits < 1) kl->z_bits = 1;
}



} // namespace tb
