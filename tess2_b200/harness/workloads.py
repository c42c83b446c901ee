"""Named synthetic workloads (BASELINE.json configs) with an on-disk cache of the host-side
tessellation, so that bench.py's two arms and repeated runs on one box tessellate once."""
import hashlib
import os
import time

import numpy as np

from . import particles, decomp, delaunay

CACHE_DIR = os.environ.get("TESSB200_CACHE", "/tmp/tess2_b200_cache")
# host tessellation time of the last workload built or loaded (reported beside the dense numbers)
LAST_TESS = {"seconds": None, "workers": None, "cached": False}


def _cache_path(key):
    return os.path.join(CACHE_DIR, hashlib.sha1(key.encode()).hexdigest()[:16] + ".npz")


def _save(path, blocks, tess_seconds=0.0, workers=0):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    out = {"n": np.array(len(blocks)), "tess_seconds": np.array(float(tess_seconds)), "tess_workers": np.array(int(workers))}
    for i, b in enumerate(blocks):
        out[f"{i}_gid"] = np.array(b["gid"])
        out[f"{i}_particles"] = b["particles"]
        out[f"{i}_tets"] = b["tets"]
        out[f"{i}_num_orig"] = np.array(b["num_orig"])
        out[f"{i}_bounds"] = np.stack([b["bounds_min"], b["bounds_max"]]).astype(np.float32)
        out[f"{i}_v2t"] = b["vert_to_tet"]
    tmp = path + f".tmp{os.getpid()}.npz"
    np.savez(tmp, **out)
    os.replace(tmp, path)


def _load(path):
    z = np.load(path)
    LAST_TESS.update(seconds=float(z["tess_seconds"]) if "tess_seconds" in z else None,
                     workers=int(z["tess_workers"]) if "tess_workers" in z else None, cached=True)
    blocks = []
    for i in range(int(z["n"])):
        blocks.append(dict(gid=int(z[f"{i}_gid"]), particles=z[f"{i}_particles"], tets=z[f"{i}_tets"],
                           num_orig=int(z[f"{i}_num_orig"]), bounds_min=z[f"{i}_bounds"][0], bounds_max=z[f"{i}_bounds"][1],
                           vert_to_tet=z[f"{i}_v2t"]))
    return blocks


def uniform_regular(n_side, blocks_xyz, gids=None, workers=None, cache=True, log=None, engine="qhull"):
    """`gen_particles` workload (TESS_DENSE_TEST semantics): domain
    [0, n_x-1] x [0, n_y-1] x [0, n_z-1] with n = n_side * blocks per axis / ... each block holds
    (block extent + 1)^3-ish particles drawn with srand(gid).  blocks_xyz = (bx, by, bz); every
    block spans n_side/..: the domain is (bx*h, by*h, bz*h) - 1 wide with h = n_side.
    gids: the blocks to tessellate (default all); ghosts come from the neighbouring blocks.
    engine: "qhull" (SciPy's Qhull with 'Qt', the reference's engine and options, one process per
    block) or "native" (the package's C++ driver and Delaunay engine, tess2_b200/host): same
    particles, same set of tets, different tet numbering.
    Returns (blocks, layout) where layout lists (gid, bounds_min, bounds_max) of EVERY block."""
    bx, by, bz = blocks_xyz
    h = n_side
    dom_min = np.zeros(3, np.float32)
    dom_max = np.array([bx * h - 1, by * h - 1, bz * h - 1], np.float32)
    layout = []
    for k in range(bz):
        for j in range(by):
            for i in range(bx):
                c = (i, j, k)
                b = (bx, by, bz)
                mn = np.array([dom_min[d] + (dom_max[d] - dom_min[d]) * np.float32(c[d]) / np.float32(b[d]) for d in range(3)], np.float32)
                mx = np.array([dom_max[d] if c[d] == b[d] - 1 else
                               dom_min[d] + (dom_max[d] - dom_min[d]) * np.float32(c[d] + 1) / np.float32(b[d]) for d in range(3)], np.float32)
                layout.append((len(layout), mn, mx))
    if gids is None:
        gids = list(range(len(layout)))
    key = f"uniform_regular:{n_side}:{blocks_xyz}:{sorted(gids)}:v5" + ("" if engine == "qhull" else ":" + engine)
    path = _cache_path(key)
    if cache and os.path.exists(path):
        if log:
            log(f"tessellation cache hit {path}")
        return _load(path), layout, dom_min, dom_max
    t0 = time.time()
    # particles of the wanted blocks and of every block that touches one of them
    need = set()
    for g in gids:
        _, mn, mx = layout[g]
        for g2, mn2, mx2 in layout:
            if np.all(mn2 <= mx + 1e-3) and np.all(mx2 >= mn - 1e-3):
                need.add(g2)
    need = sorted(need)
    ps = {g: particles.gen_particles(g, layout[g][1], layout[g][2]) for g in need}
    allp = np.concatenate([ps[g] for g in need])
    owner = np.concatenate([np.full(len(ps[g]), g, np.int32) for g in need])
    bounds = {g: (layout[g][1], layout[g][2]) for g in gids}
    # tessellate only the wanted gids
    if engine == "native":
        from .. import host_tess
        blocks = host_tess.tess(allp, owner, [(mn, mx) for _, mn, mx in layout], dom_min, dom_max, gids=list(gids), max_growth=1.0,
                                threads=workers or 0)
    else:
        delaunay._G.update(points=allp, owner=owner, bounds=bounds, dmin=dom_min, dmax=dom_max, margin0=None, max_growth=1.0)
        blocks = delaunay.tessellate_gids(gids, workers)
        for b in blocks:   # dblock_t::vert_to_tet is part of what tess hands to dense (src/tess.cpp:767-787)
            b["vert_to_tet"] = delaunay.fill_vert_to_tet(len(b["particles"]), b["tets"])
    if log:
        log(f"tessellated {len(gids)} blocks ({sum(b['num_orig'] for b in blocks)} particles, "
            f"{sum(len(b['tets']) for b in blocks)} tets) in {time.time() - t0:.1f} s")
    import os as _os
    nworkers = workers if workers is not None else min(len(gids), _os.cpu_count() or 1)
    LAST_TESS.update(seconds=time.time() - t0, workers=nworkers, cached=False)
    if cache:
        _save(path, blocks, LAST_TESS["seconds"], nworkers)
    return blocks, layout, dom_min, dom_max
