"""Block decompositions (host side).  The reference delegates both to DIY
(examples/tess-dense/main.cpp:190-195 RegularDecomposer; src/tess-kdtree.cpp:83-107 diy::kdtree),
which is not vendored, so split positions are "parity unpinned" (SURVEY.md 8(c)); the dense
stage's global grid depends on them only through fp32 summation order.
"""
import numpy as np


def factor_blocks(nblocks):
    """nblocks -> (bx, by, bz), as even as possible (2x2x2 for 8, 4x4x4 for 64)."""
    dims = [1, 1, 1]
    n = nblocks
    f = 2
    factors = []
    while n > 1:
        while n % f == 0:
            factors.append(f)
            n //= f
        f += 1
    for fac in sorted(factors, reverse=True):
        dims[int(np.argmin(dims))] *= fac
    return tuple(sorted(dims, reverse=True))


def regular_blocks(domain_min, domain_max, nblocks):
    """Regular decomposition of a continuous domain: block (i,j,k) spans
    min + ext * i / b  ..  min + ext * (i+1) / b  per axis, gid = x-fastest order."""
    lo = np.asarray(domain_min, dtype=np.float32)
    hi = np.asarray(domain_max, dtype=np.float32)
    b = factor_blocks(nblocks)
    bounds = []
    for k in range(b[2]):
        for j in range(b[1]):
            for i in range(b[0]):
                c = (i, j, k)
                mn = [np.float32(lo[d] + (hi[d] - lo[d]) * np.float32(c[d]) / np.float32(b[d])) for d in range(3)]
                mx = [np.float32(lo[d] + (hi[d] - lo[d]) * np.float32(c[d] + 1) / np.float32(b[d])) for d in range(3)]
                for d in range(3):
                    if c[d] == b[d] - 1:
                        mx[d] = hi[d]
                bounds.append((np.array(mn, dtype=np.float32), np.array(mx, dtype=np.float32)))
    return bounds


def kdtree_blocks(points, domain_min, domain_max, nblocks):
    """kd-tree decomposition into nblocks (a power of two) by exact-median splits cycling
    x, y, z per level (stand-in for diy::kdtree's 1024-bin histogram medians).
    Returns (bounds list, owner gid per point)."""
    assert nblocks & (nblocks - 1) == 0
    lo = np.asarray(domain_min, dtype=np.float32)
    hi = np.asarray(domain_max, dtype=np.float32)
    boxes = [(lo.copy(), hi.copy(), np.arange(len(points)))]
    level = 0
    while len(boxes) < nblocks:
        d = level % 3
        nxt = []
        for mn, mx, idx in boxes:
            x = points[idx, d]
            if len(idx) >= 2:
                s = np.sort(x)
                m = len(s) // 2
                split = np.float32((np.float64(s[m - 1]) + np.float64(s[m])) * 0.5)
                if not (s[m - 1] < split < s[m]):
                    split = s[m]
            else:
                split = np.float32((np.float64(mn[d]) + np.float64(mx[d])) * 0.5)
            left = idx[x < split]
            right = idx[x >= split]
            lmx = mx.copy(); lmx[d] = split
            rmn = mn.copy(); rmn[d] = split
            nxt.append((mn, lmx, left))
            nxt.append((rmn, mx, right))
        boxes = nxt
        level += 1
    owner = np.empty(len(points), dtype=np.int32)
    bounds = []
    for gid, (mn, mx, idx) in enumerate(boxes):
        owner[idx] = gid
        bounds.append((mn.astype(np.float32), mx.astype(np.float32)))
    return bounds, owner


def assign_regular(points, bounds):
    """owner gid of each point for a list of (min, max) boxes (half-open, last box closed)."""
    owner = np.full(len(points), -1, dtype=np.int32)
    for gid, (mn, mx) in enumerate(bounds):
        inside = np.all((points >= mn) & (points <= mx), axis=1) & (owner < 0)
        owner[inside] = gid
    return owner
