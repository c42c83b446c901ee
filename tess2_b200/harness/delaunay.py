"""Host-side serial Delaunay per block (the stage north_star keeps on the host).

The reference runs Qhull with options "d Qt" per DIY block (src/tess-qhull.c:46-50) and
exchanges ghost particles until every cell is final (src/tess.cpp:79-98, 492-628).
libqhull is not installed here; SciPy bundles the same Qhull, and
``scipy.spatial.Delaunay(qhull_options="Qt")`` returns simplices / neighbors with the
tet_t convention (neighbor i opposite vertex i, -1 on the hull; include/tess/tet.h:4-7).

Ghost particles: every non-owned particle within a margin of the block bounds; the margin
grows until the circumsphere of every tet incident to an original particle with a finite
cell lies inside the region that was searched for ghosts (the reference's
incomplete_cells test, src/tess.cpp:492-628, in its simplest form).  Any scheme is
parity-safe for the dense stage: the oracle and the GPU path consume the same tets.
"""
import multiprocessing as mp
import os
import numpy as np


def fill_vert_to_tet(num_particles, tets):
    """'the last one wins' (src/tess.cpp:767-787) == highest tet index containing the vertex."""
    v2t = np.full(num_particles, -1, dtype=np.int32)
    nt = len(tets)
    if nt:
        np.maximum.at(v2t, tets[:, :4].ravel(), np.repeat(np.arange(nt, dtype=np.int32), 4))
    return v2t


def _circumspheres64(p, simplices):
    a = p[simplices[:, 0]]; b = p[simplices[:, 1]]; c = p[simplices[:, 2]]; d = p[simplices[:, 3]]
    t = a - d; u = b - d; v = c - d
    nt = np.einsum("ij,ij->i", t, t); nu = np.einsum("ij,ij->i", u, u); nv = np.einsum("ij,ij->i", v, v)
    den = 2.0 * np.einsum("ij,ij->i", t, np.cross(u, v))
    den = np.where(den == 0.0, 1e-300, den)
    num = nt[:, None] * np.cross(u, v) + nu[:, None] * np.cross(v, t) + nv[:, None] * np.cross(t, u)
    rel = num / den[:, None]
    return d + rel, np.sqrt(np.einsum("ij,ij->i", rel, rel))


def tessellate_block(points, owner, gid, bmin, bmax, dmin, dmax, margin0=None, max_rounds=3, max_growth=2.5):
    """Delaunay of block gid's originals + ghosts.  Returns dict(particles f32 [n,3] originals
    first, num_orig, tets int32 [T,8], margin, rounds)."""
    from scipy.spatial import Delaunay
    bmin = np.asarray(bmin, np.float64); bmax = np.asarray(bmax, np.float64)
    dmin = np.asarray(dmin, np.float64); dmax = np.asarray(dmax, np.float64)
    mine = np.flatnonzero(owner == gid)
    others = np.flatnonzero(owner != gid)
    orig = points[mine]
    n_orig = len(orig)
    if margin0 is None:
        vol = float(np.prod(bmax - bmin))
        margin0 = 3.0 * (vol / max(n_orig, 1)) ** (1.0 / 3.0)
    margin = float(margin0)
    po = points[others].astype(np.float64)
    rounds = 0
    while True:
        rounds += 1
        sel = np.all((po >= bmin - margin) & (po <= bmax + margin), axis=1)
        ghosts = points[others[sel]]
        allp = np.concatenate([orig, ghosts]).astype(np.float32)
        tri = Delaunay(allp.astype(np.float64), qhull_options="Qt")
        simp = tri.simplices.astype(np.int32)
        nbr = tri.neighbors.astype(np.int32)
        covered = bool(np.all(bmin - margin <= dmin) and np.all(bmax + margin >= dmax))
        if covered or rounds >= max_rounds or len(others) == 0:
            break
        # finite-cell originals: not a vertex of any hull facet
        on_hull = np.zeros(len(allp), dtype=bool)
        for k in range(4):
            hull_t = nbr[:, k] < 0
            for j in range(4):
                if j != k:
                    on_hull[simp[hull_t, j]] = True
        need_v = np.zeros(len(allp), dtype=bool)
        need_v[:n_orig] = ~on_hull[:n_orig]
        # an original on the local hull that is not near the true domain boundary was not
        # surrounded by ghosts: the searched region must grow
        grow = False
        hull_orig = np.flatnonzero(on_hull[:n_orig])
        if len(hull_orig):
            q = orig[hull_orig].astype(np.float64)
            dist_dom = np.minimum(q - dmin, dmax - q).min(axis=1)
            grow = bool(np.any(dist_dom > margin))
        tsel = need_v[simp].any(axis=1)
        cen, rad = _circumspheres64(allp.astype(np.float64), simp[tsel])
        lo_need = (bmin - (cen - rad[:, None]))          # how far the sphere pokes below the block
        hi_need = ((cen + rad[:, None]) - bmax)
        # no ghosts exist beyond the domain, so clip the requirement there
        lo_cap = bmin - dmin
        hi_cap = dmax - bmax
        req = max(float(np.minimum(lo_need, lo_cap).max(initial=0.0)),
                  float(np.minimum(hi_need, hi_cap).max(initial=0.0)), 2.0 * margin if grow else 0.0)
        if req <= margin or margin >= max_growth * margin0:
            break
        margin = min(req * 1.05, max_growth * margin0)
    tets = np.ascontiguousarray(np.concatenate([simp, nbr], axis=1).astype(np.int32))
    return dict(gid=gid, particles=np.ascontiguousarray(allp), num_orig=n_orig, tets=tets,
                margin=margin, rounds=rounds,
                bounds_min=np.asarray(bmin, np.float32), bounds_max=np.asarray(bmax, np.float32))


_G = {}


def _worker(gid):
    g = _G
    mn, mx = g["bounds"][gid]
    return tessellate_block(g["points"], g["owner"], gid, mn, mx, g["dmin"], g["dmax"], g["margin0"],
                            max_growth=g.get("max_growth", 2.5))


def tessellate(points, owner, bounds, domain_min, domain_max, workers=None, margin0=None, max_growth=2.5):
    """Tessellate every block (one process per block, up to `workers`)."""
    nblocks = len(bounds)
    _G.update(points=points, owner=owner, bounds=bounds, dmin=domain_min, dmax=domain_max, margin0=margin0, max_growth=max_growth)
    if workers is None:
        workers = min(nblocks, os.cpu_count() or 1)
    if workers <= 1 or nblocks == 1:
        return [_worker(g) for g in range(nblocks)]
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        return pool.map(_worker, range(nblocks), chunksize=1)


def tessellate_gids(gids, workers=None):
    """Tessellate the given gids using the shared state already placed in _G (bounds keyed by gid)."""
    gids = list(gids)
    if workers is None:
        workers = min(len(gids), os.cpu_count() or 1)
    if workers <= 1 or len(gids) == 1:
        return [_worker(g) for g in gids]
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        return pool.map(_worker, gids, chunksize=1)
