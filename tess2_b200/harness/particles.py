"""Synthetic particle sets for the dense stage (host side, input staging only).

* ``gen_particles`` restates the reference's test-particle generator
  (src/tess.cpp:264-293): ``srand(gid)`` then, per particle and per axis,
  ``t = (float)rand() / RAND_MAX`` and ``p = t * (max - min) + min`` in fp32, with
  ``n = (int)(dx + 1) * (int)(dy + 1) * (int)(dz + 1)`` particles per block.  glibc's
  ``rand`` is called through ctypes so the sequence is the reference's own.
* ``clustered_particles`` is the Gaussian-clump distribution SURVEY.md 8(d) defines
  for the clustered configs (the reference ships no such generator).
"""
import ctypes
import numpy as np

_libc = None


def _glibc():
    global _libc
    if _libc is None:
        _libc = ctypes.CDLL("libc.so.6")
        _libc.rand.restype = ctypes.c_int
        _libc.srand.argtypes = [ctypes.c_uint]
    return _libc


RAND_MAX = 2147483647
_crand = None


def _rand_stream(seed, n):
    """n values of glibc rand() after srand(seed): through the small C helper crand.c when it has
    been built (build() does), else one ctypes call per value."""
    global _crand
    import os
    so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_crand.so")
    if _crand is None and os.path.exists(so):
        _crand = ctypes.CDLL(so)
        _crand.crand_fill.argtypes = [ctypes.c_uint, ctypes.c_longlong, ctypes.c_void_p]
    if _crand is not None:
        out = np.empty(n, dtype=np.int32)
        _crand.crand_fill(seed, n, out.ctypes.data)
        return out.astype(np.int64)
    libc = _glibc()
    libc.srand(seed)
    return np.fromiter((libc.rand() for _ in range(n)), dtype=np.int64, count=n)


def gen_particles(gid, bounds_min, bounds_max):
    """Uniform random particles of one block, bit-identical to src/tess.cpp:264-293."""
    bmin = np.asarray(bounds_min, dtype=np.float32)
    bmax = np.asarray(bounds_max, dtype=np.float32)
    sizes = [int(np.float32(bmax[i] - bmin[i]) + np.float32(1)) for i in range(3)]
    n = sizes[0] * sizes[1] * sizes[2]
    r = _rand_stream(gid, 3 * n)
    # (float)rand() / RAND_MAX : int -> float (rounded), RAND_MAX -> float (2^31), float division
    t = r.astype(np.float32) / np.float32(RAND_MAX)
    t = t.reshape(n, 3)
    ext = (bmax - bmin).astype(np.float32)
    p = (t * ext[None, :]).astype(np.float32) + bmin[None, :]
    return np.ascontiguousarray(p.astype(np.float32))


def uniform_particles(n, domain_min, domain_max, seed):
    """n uniform particles in the open domain (numpy generator; for property tests)."""
    rng = np.random.default_rng(seed)
    lo = np.asarray(domain_min, dtype=np.float64)
    hi = np.asarray(domain_max, dtype=np.float64)
    p = (rng.random((n, 3)) * (hi - lo) + lo).astype(np.float32)
    return _dedup_open(p, lo, hi)


def clustered_particles(n, domain_min, domain_max, seed, n_clumps=64, clump_frac=0.8,
                        sigma_lo=0.01, sigma_hi=0.05):
    """Gaussian-clump particles: clump_frac of the points in n_clumps
    isotropic Gaussians (centres uniform in the domain, sigma ~ U[sigma_lo, sigma_hi] * extent,
    equal weights), the rest uniform; rejection to the open domain; fp32; exact duplicates dropped."""
    rng = np.random.default_rng(seed)
    lo = np.asarray(domain_min, dtype=np.float64)
    hi = np.asarray(domain_max, dtype=np.float64)
    ext = float(np.max(hi - lo))
    centres = rng.random((n_clumps, 3)) * (hi - lo) + lo
    sigmas = rng.uniform(sigma_lo, sigma_hi, n_clumps) * ext
    n_cl = int(round(n * clump_frac))
    out = []
    need = n_cl
    while need > 0:
        m = int(need * 1.3) + 16
        which = rng.integers(0, n_clumps, m)
        q = centres[which] + rng.standard_normal((m, 3)) * sigmas[which][:, None]
        ok = np.all((q > lo) & (q < hi), axis=1)
        q = q[ok][:need]
        out.append(q)
        need -= len(q)
    out.append(rng.random((n - n_cl, 3)) * (hi - lo) + lo)
    p = np.concatenate(out).astype(np.float32)
    p = p[rng.permutation(len(p))]
    return _dedup_open(p, lo, hi)


def _dedup_open(p, lo, hi):
    lo32 = lo.astype(np.float32)
    hi32 = hi.astype(np.float32)
    ok = np.all((p > lo32) & (p < hi32), axis=1)
    p = p[ok]
    # Qhull leaves exact duplicates out of the triangulation; the reference's pread
    # drivers deduplicate first (examples/pread-voronoi/common.h:19-54)
    return _drop_duplicates(p)


def _drop_duplicates(p, large=1 << 22):
    """First occurrence of every distinct row, in input order.  Large sets (configs 3-5: up to 512^3 rows) go through a
    64-bit hash of the three float32 bit patterns: rows with a unique hash are unique, and only the rows that share a
    hash (normally none) are compared exactly -- the same result as the row-wise np.unique, without sorting 12-byte rows."""
    p = np.ascontiguousarray(p, dtype=np.float32)
    if len(p) <= large:
        _, first = np.unique(p, axis=0, return_index=True)
        return np.ascontiguousarray(p[np.sort(first)])
    u = (p + np.float32(0.0)).view(np.uint32)            # -0.0 and 0.0 are the same point
    h = u[:, 0].astype(np.uint64) | (u[:, 1].astype(np.uint64) << np.uint64(32))
    h ^= u[:, 2].astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    hs = np.sort(h)                                      # the usual case ends here: no two rows share a hash
    if not (hs[1:] == hs[:-1]).any():
        return p
    del hs
    order = np.argsort(h, kind="stable")
    hs = h[order]
    eq = hs[1:] == hs[:-1]
    member = np.zeros(len(p), dtype=bool)
    member[1:] |= eq
    member[:-1] |= eq
    idx = order[member]                                  # rows that share a hash, grouped by hash, input order inside a group
    _, first = np.unique(p[idx], axis=0, return_index=True)
    keep = np.ones(len(p), dtype=bool)
    keep[idx] = False
    keep[idx[first]] = True
    return np.ascontiguousarray(p[keep])
