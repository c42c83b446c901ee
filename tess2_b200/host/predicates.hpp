// predicates.hpp -- exact-sign orient3d / insphere for float32 points (host side of the tess() driver).
//
// The reference hands its points to Qhull or CGAL (src/tess-qhull.c, src/tess-cgal.cpp); neither is
// installed here, so the serial engine behind tessb200_tess() is this repo's own (delaunay3.hpp) and
// needs its own predicates.  Evaluation in double with a forward error bound; when the bound cannot
// decide the sign, the determinant is re-evaluated exactly as a sum of doubles (two_sum / two_product
// building blocks as in Shewchuk's "Adaptive Precision Floating-Point Arithmetic"; the sign by
// repeated distillation).  Every coordinate difference is captured exactly by two doubles.
#ifndef TESSB200_PREDICATES_HPP
#define TESSB200_PREDICATES_HPP

#include <cmath>
#include <vector>

namespace tb_host
{

// ---- exact arithmetic on sums of doubles -----------------------------------------------------------
// An Expansion is a bag of doubles whose exact sum is the value.  Every operation below is exact
// (two_sum and two_prod lose nothing); one two_sum chain per operation drops the zero terms.  The
// sign comes from repeated distillation: a chain leaves q = fl(sum) and the round-off terms; the sum
// of the |round-off| bounds the distance of q from the true value, and each further pass shrinks the
// round-off terms by about 2^-53.
struct Expansion
{
  std::vector<double> c;
  Expansion() {}
};

inline void two_sum(double a, double b, double &x, double &y)
{
  x = a + b;
  const double bv = x - a, av = x - bv;
  y = (a - av) + (b - bv);
}
inline void two_prod(double a, double b, double &x, double &y)
{
  x = a * b;
  y = std::fma(a, b, -x);
}

// one distillation pass in place: afterwards c = {round-off terms..., fl(sum)}, zeros dropped
inline void distill(std::vector<double> &c)
{
  if (c.size() < 2) return;
  double q = c[0];
  size_t k = 0;
  for (size_t m = 1; m < c.size(); m++) {
    double x, y;
    two_sum(q, c[m], x, y);
    if (y != 0.0) c[k++] = y;
    q = x;
  }
  if (q != 0.0) c[k++] = q;
  c.resize(k);
}

inline int sign(Expansion e)
{
  for (int pass = 0; pass < 64; pass++) {
    distill(e.c);
    if (e.c.empty()) return 0;
    const double q = e.c.back();
    double rest = 0.0;
    for (size_t m = 0; m + 1 < e.c.size(); m++) rest += std::fabs(e.c[m]);
    if (std::fabs(q) > rest * (1.0 + 1e-12)) return q > 0.0 ? 1 : -1;
  }
  return 0;   // not reached for finite inputs
}

inline Expansion operator+(const Expansion &e, const Expansion &f)
{
  Expansion h;
  h.c.reserve(e.c.size() + f.c.size());
  h.c.insert(h.c.end(), e.c.begin(), e.c.end());
  h.c.insert(h.c.end(), f.c.begin(), f.c.end());
  distill(h.c);
  return h;
}
inline Expansion operator-(const Expansion &e)
{
  Expansion h = e;
  for (double &v : h.c) v = -v;
  return h;
}
inline Expansion operator-(const Expansion &e, const Expansion &f) { return e + (-f); }

inline Expansion operator*(const Expansion &e, const Expansion &f)
{
  Expansion h;
  h.c.reserve(2 * e.c.size() * f.c.size());
  for (double a : e.c)
    for (double b : f.c) {
      double x, y;
      two_prod(a, b, x, y);
      if (y != 0.0) h.c.push_back(y);
      if (x != 0.0) h.c.push_back(x);
    }
  distill(h.c);
  return h;
}

inline Expansion diff(double a, double b)   // a - b exactly
{
  double x, y;
  two_sum(a, -b, x, y);
  Expansion h;
  if (y != 0.0) h.c.push_back(y);
  if (x != 0.0) h.c.push_back(x);
  return h;
}

// ---- orient3d: sign of det[a-d, b-d, c-d] --------------------------------------------------------
inline int orient3d_exact(const float *a, const float *b, const float *c, const float *d)
{
  Expansion ax = diff(a[0], d[0]), ay = diff(a[1], d[1]), az = diff(a[2], d[2]);
  Expansion bx = diff(b[0], d[0]), by = diff(b[1], d[1]), bz = diff(b[2], d[2]);
  Expansion cx = diff(c[0], d[0]), cy = diff(c[1], d[1]), cz = diff(c[2], d[2]);
  Expansion det = ax * (by * cz - bz * cy) + ay * (bz * cx - bx * cz) + az * (bx * cy - by * cx);
  return sign(det);
}

// static_bound: 0, or a bound on the rounding error valid for every call of a run (StaticFilter below):
// most calls are decided by one comparison, before the per-call bound is even computed
inline int orient3d(const float *a, const float *b, const float *c, const float *d, double static_bound = 0.0)
{
  const double adx = (double)a[0] - d[0], ady = (double)a[1] - d[1], adz = (double)a[2] - d[2];
  const double bdx = (double)b[0] - d[0], bdy = (double)b[1] - d[1], bdz = (double)b[2] - d[2];
  const double cdx = (double)c[0] - d[0], cdy = (double)c[1] - d[1], cdz = (double)c[2] - d[2];
  const double bdxcdy = bdx * cdy, cdxbdy = cdx * bdy, cdxady = cdx * ady, adxcdy = adx * cdy, adxbdy = adx * bdy, bdxady = bdx * ady;
  const double det = adz * (bdxcdy - cdxbdy) + bdz * (cdxady - adxcdy) + cdz * (adxbdy - bdxady);
  if (det > static_bound && static_bound > 0.0) return 1;
  if (det < -static_bound && static_bound > 0.0) return -1;
  const double permanent = (std::fabs(bdxcdy) + std::fabs(cdxbdy)) * std::fabs(adz) + (std::fabs(cdxady) + std::fabs(adxcdy)) * std::fabs(bdz) +
                           (std::fabs(adxbdy) + std::fabs(bdxady)) * std::fabs(cdz);
  // float32 inputs: the differences above are exact unless the exponents are far apart; the bound
  // below covers the rounding of the products and sums (Shewchuk's o3derrboundA, doubled for slack)
  const double errbound = 1.6e-15 * permanent;
  if (det > errbound) return 1;
  if (det < -errbound) return -1;
  return orient3d_exact(a, b, c, d);
}

// ---- insphere: > 0 iff e lies inside the sphere through a, b, c, d (orient3d(a,b,c,d) > 0) ---------
inline int insphere_exact(const float *a, const float *b, const float *c, const float *d, const float *e)
{
  Expansion p[4][3];
  const float *v[4] = {a, b, c, d};
  for (int i = 0; i < 4; i++)
    for (int k = 0; k < 3; k++) p[i][k] = diff(v[i][k], e[k]);
  Expansion lift[4];
  for (int i = 0; i < 4; i++) lift[i] = p[i][0] * p[i][0] + p[i][1] * p[i][1] + p[i][2] * p[i][2];
  auto m2 = [&](int i, int j) { return p[i][0] * p[j][1] - p[j][0] * p[i][1]; };   // xy minors
  const Expansion ab = m2(0, 1), bc = m2(1, 2), cd = m2(2, 3), da = m2(3, 0), ac = m2(0, 2), bd = m2(1, 3);
  const Expansion abc = p[0][2] * bc - p[1][2] * ac + p[2][2] * ab;
  const Expansion bcd = p[1][2] * cd - p[2][2] * bd + p[3][2] * bc;
  const Expansion cda = p[2][2] * da + p[3][2] * ac + p[0][2] * cd;
  const Expansion dab = p[3][2] * ab + p[0][2] * bd + p[1][2] * da;
  const Expansion det = (lift[3] * abc - lift[2] * dab) + (lift[1] * cda - lift[0] * bcd);
  return sign(det);
}

inline int insphere(const float *a, const float *b, const float *c, const float *d, const float *e, double static_bound = 0.0)
{
  const double aex = (double)a[0] - e[0], aey = (double)a[1] - e[1], aez = (double)a[2] - e[2];
  const double bex = (double)b[0] - e[0], bey = (double)b[1] - e[1], bez = (double)b[2] - e[2];
  const double cex = (double)c[0] - e[0], cey = (double)c[1] - e[1], cez = (double)c[2] - e[2];
  const double dex = (double)d[0] - e[0], dey = (double)d[1] - e[1], dez = (double)d[2] - e[2];
  const double aexbey = aex * bey, bexaey = bex * aey, ab = aexbey - bexaey;
  const double bexcey = bex * cey, cexbey = cex * bey, bc = bexcey - cexbey;
  const double cexdey = cex * dey, dexcey = dex * cey, cd = cexdey - dexcey;
  const double dexaey = dex * aey, aexdey = aex * dey, da = dexaey - aexdey;
  const double aexcey = aex * cey, cexaey = cex * aey, ac = aexcey - cexaey;
  const double bexdey = bex * dey, dexbey = dex * bey, bd = bexdey - dexbey;
  const double abc = aez * bc - bez * ac + cez * ab;
  const double bcd = bez * cd - cez * bd + dez * bc;
  const double cda = cez * da + dez * ac + aez * cd;
  const double dab = dez * ab + aez * bd + bez * da;
  const double alift = aex * aex + aey * aey + aez * aez, blift = bex * bex + bey * bey + bez * bez;
  const double clift = cex * cex + cey * cey + cez * cez, dlift = dex * dex + dey * dey + dez * dez;
  const double det = (dlift * abc - clift * dab) + (blift * cda - alift * bcd);
  if (det > static_bound && static_bound > 0.0) return 1;
  if (det < -static_bound && static_bound > 0.0) return -1;
  const double aezp = std::fabs(aez), bezp = std::fabs(bez), cezp = std::fabs(cez), dezp = std::fabs(dez);
  const double aexbeyp = std::fabs(aexbey), bexaeyp = std::fabs(bexaey), bexceyp = std::fabs(bexcey), cexbeyp = std::fabs(cexbey);
  const double cexdeyp = std::fabs(cexdey), dexceyp = std::fabs(dexcey), dexaeyp = std::fabs(dexaey), aexdeyp = std::fabs(aexdey);
  const double aexceyp = std::fabs(aexcey), cexaeyp = std::fabs(cexaey), bexdeyp = std::fabs(bexdey), dexbeyp = std::fabs(dexbey);
  const double permanent = ((cexdeyp + dexceyp) * bezp + (dexbeyp + bexdeyp) * cezp + (bexceyp + cexbeyp) * dezp) * alift +
                           ((dexaeyp + aexdeyp) * cezp + (aexceyp + cexaeyp) * dezp + (cexdeyp + dexceyp) * aezp) * blift +
                           ((aexbeyp + bexaeyp) * dezp + (bexdeyp + dexbeyp) * aezp + (dexaeyp + aexdeyp) * bezp) * clift +
                           ((bexceyp + cexbeyp) * aezp + (cexaeyp + aexceyp) * bezp + (aexbeyp + bexaeyp) * cezp) * dlift;
  const double errbound = 3.6e-15 * permanent;     // isperrboundA (1.78e-15) doubled for slack
  if (det > errbound) return 1;
  if (det < -errbound) return -1;
  return insphere_exact(a, b, c, d, e);
}

// Error bounds that hold for every predicate call on points whose coordinate differences are at
// most D in magnitude: the per-call bounds above are (coefficient) x permanent, and the permanents
// are at most 6 D^3 (orient3d) and 72 D^5 (insphere).
struct StaticFilter
{
  double orient = 0.0, sphere = 0.0;
  void set_extent(double D)
  {
    orient = 1.6e-15 * 6.0 * D * D * D;
    sphere = 3.6e-15 * 72.0 * D * D * D * D * D;
  }
};

}  // namespace tb_host

#endif
