"""Does the order of the tets in memory matter?  Renumbers the tets of every block along a Morton
curve of their centroids (neighbour ids and vert_to_tet remapped; the slots inside a tet, which is
what the reference's walk order depends on, stay as they are) and times the dense stage on both
layouts.  The densities must be identical bit for bit.  GPU box: python profiles/probe_tetorder.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import tess2_b200  # noqa: E402


def morton_perm(b):
    t = b["tets"]
    p = b["particles"]
    c = p[t[:, :4]].mean(axis=1)
    lo, hi = c.min(0), c.max(0)
    q = np.clip(((c - lo) / np.maximum(hi - lo, 1e-30) * 1023).astype(np.uint64), 0, 1023)

    def spread(x):
        x = (x | (x << 16)) & 0x030000FF
        x = (x | (x << 8)) & 0x0300F00F
        x = (x | (x << 4)) & 0x030C30C3
        x = (x | (x << 2)) & 0x09249249
        return x
    key = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
    return np.argsort(key, kind="stable")


def renumber(b):
    order = morton_perm(b)                 # new position -> old id
    newid = np.empty(len(order), np.int32)
    newid[order] = np.arange(len(order), dtype=np.int32)
    t = b["tets"][order].copy()
    nb = t[:, 4:]
    t[:, 4:] = np.where(nb >= 0, newid[np.maximum(nb, 0)], -1)
    v2t = b["vert_to_tet"]
    out = dict(b)
    out["tets"] = np.ascontiguousarray(t)
    out["vert_to_tet"] = np.where(v2t >= 0, newid[np.maximum(v2t, 0)], -1).astype(np.int32)
    return out


blocks, layout, owner, dmin, dmax, gsize = bench.build_workload(1, 0)
ctx = tess2_b200.Context(0)
params = ctx.make_params(tess2_b200.DENSE_TESS, 0, dmin, dmax, False, (0.0, 0.0, 1.0), 1.0, 1e-4, gsize)
res = {}
dens = {}
for name, bl in (("as given (Qhull order)", blocks), ("Morton order", [renumber(b) for b in blocks])):
    ctx.upload(bl)
    for _ in range(3):
        st = ctx.run(params)
    acc = {}
    for _ in range(5):
        st = ctx.run(params)
        for k in ("ms_circumcenters", "ms_bfs", "ms_nbrs", "ms_faces", "ms_scan", "ms_sort", "ms_deposit", "ms_slow_path", "ms_total_device"):
            acc[k] = acc.get(k, 0.0) + getattr(st, k) / 5
    res[name] = {k: round(v, 3) for k, v in acc.items()}
    dens[name] = [d.copy() for d in ctx.download(params, want_grid=False).block_density]
a, b = list(dens.values())
res["identical densities"] = all(np.array_equal(x.view(np.uint32), y.view(np.uint32)) for x, y in zip(a, b))
print(json.dumps(res, indent=1))
