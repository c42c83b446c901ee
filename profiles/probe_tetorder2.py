#!/usr/bin/env python
"""Does the ORDER of the tets (and of the particles) in memory matter?  The same workload three ways: as the host engine
numbers them, tets renumbered along a Morton curve of their circumcenters, and tets + particles renumbered.  Results are
identical up to the accumulation order of equal cells (cell numbers are kept), stage times are printed.

    python profiles/probe_tetorder2.py --config 3 --scale 2"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def morton(q):
    def spread(x):
        x = x.astype(np.uint64) & np.uint64(0x1fffff)
        x = (x | (x << np.uint64(32))) & np.uint64(0x1f00000000ffff)
        x = (x | (x << np.uint64(16))) & np.uint64(0x1f0000ff0000ff)
        x = (x | (x << np.uint64(8))) & np.uint64(0x100f00f00f00f00f)
        x = (x | (x << np.uint64(4))) & np.uint64(0x10c30c30c30c30c3)
        x = (x | (x << np.uint64(2))) & np.uint64(0x1249249249249249)
        return x
    return spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)) | (spread(q[:, 2]) << np.uint64(2))


def reorder_tets(b):
    p, t = b["particles"], b["tets"]
    c = p[t[:, :4]].mean(axis=1)                    # centroid stands in for the circumcenter
    lo, hi = c.min(axis=0), c.max(axis=0)
    q = ((c - lo) / np.maximum(hi - lo, 1e-30) * 2097151.0).astype(np.uint64)
    order = np.argsort(morton(q), kind="stable")
    inv = np.empty(len(t), np.int64)
    inv[order] = np.arange(len(t))
    nt = t[order].copy()
    nb = nt[:, 4:]
    nt[:, 4:] = np.where(nb >= 0, inv[np.maximum(nb, 0)], -1)
    v2t = b["vert_to_tet"]
    # vert_to_tet must stay "a tet that holds the vertex" AND the same tet as before (it fixes the BFS order)
    nv = np.where(v2t >= 0, inv[np.maximum(v2t, 0)], -1).astype(np.int32)
    return dict(b, tets=np.ascontiguousarray(nt.astype(np.int32)), vert_to_tet=nv)


def run(ctx, blocks, w, tag):
    import tess2_b200
    params = ctx.make_params(0, w["ng"], w["dmin"], w["dmax"], False, (0.0, 0.0, 1.0), 1.0, 1e-4, w["gsize"])
    ctx.upload(blocks)
    for _ in range(3):
        ctx.run(params)
    acc = {}
    for _ in range(3):
        st = ctx.run(params)
        for k, _t in tess2_b200.lib.DenseStats._fields_:
            if k.startswith("ms_"):
                acc[k] = acc.get(k, 0.0) + getattr(st, k) / 3
    res = ctx.download(params, want_grid=True)
    print(tag, json.dumps({k: round(v, 3) for k, v in acc.items() if v > 0}), "mass", st.tot_mass, flush=True)
    return res.grid


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--scale", type=int, default=2)
    args = ap.parse_args()
    import tess2_b200
    w = bench.build_workload(args.config, 1, 0, args.scale)
    ctx = tess2_b200.Context(0)
    g0 = run(ctx, w["blocks"], w, "as numbered by the host engine ")
    b2 = [reorder_tets(b) for b in w["blocks"]]
    g1 = run(ctx, b2, w, "tets along a Morton curve      ")
    print("same grid bits:", bool(np.array_equal(g0.view(np.uint32), g1.view(np.uint32))))
    ctx.close()


if __name__ == "__main__":
    main()
