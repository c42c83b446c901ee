"""Generates tests/golden/*.npz by running the UNMODIFIED reference sources
(oracle/_ref/libtess_ref.so, built from /root/reference by oracle/Makefile).

Run from the repo root in the development container (the reference does not exist on the
GPU box):   python tests/golden/make_golden.py

The reference ships no golden vectors of its own, so these files are
the pinned known answers: inputs (particles, tets, bounds) and, from the reference itself,
fill_circumcenters, complete, volume, and dense() grids for DENSE_TESS / DENSE_CIC, 3-D and
xy-projected, plus the bytes WriteGrid writes.
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from tess2_b200.harness import particles, decomp, delaunay  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    R = ref.Checker("reference")
    # ---- small: everything stored -------------------------------------------------------------
    dom = ([0, 0, 0], [9, 9, 9])
    p = particles.clustered_particles(700, *dom, seed=424242, n_clumps=5, sigma_lo=0.05, sigma_hi=0.15)
    bounds, owner = decomp.kdtree_blocks(p, *dom, 2)
    blocks = delaunay.tessellate(p, owner, bounds, *dom, workers=1)
    out = {}
    gs = (24, 24, 24)
    out["gsize"] = np.array(gs, np.int32)
    out["nblocks"] = np.array(len(blocks), np.int32)
    for i, b in enumerate(blocks):
        v2t = R.fill_vert_to_tet(len(b["particles"]), b["tets"])
        b["vert_to_tet"] = v2t
        out[f"b{i}_particles"] = b["particles"]
        out[f"b{i}_tets"] = b["tets"]
        out[f"b{i}_num_orig"] = np.array(b["num_orig"], np.int32)
        out[f"b{i}_bounds"] = np.stack([b["bounds_min"], b["bounds_max"]]).astype(np.float32)
        out[f"b{i}_v2t"] = v2t
        out[f"b{i}_cc"] = R.circumcenters(b["tets"], b["particles"])
        out[f"b{i}_complete"] = R.complete(len(b["particles"]), b["tets"], v2t)
        out[f"b{i}_volume"] = R.volumes(len(b["particles"]), b["tets"], b["particles"], v2t)
    for alg in (0, 1):
        for proj in (0, 1):
            with tempfile.TemporaryDirectory() as td:
                f = os.path.join(td, "dense.raw")
                o = R.dense(blocks, gs, alg=alg, project=bool(proj), outfile=f)
                out[f"alg{alg}_proj{proj}_raw"] = np.fromfile(f, dtype=np.float32)
            for i in range(len(blocks)):
                out[f"alg{alg}_proj{proj}_b{i}_density"] = o["block_density"][i]
                out[f"alg{alg}_proj{proj}_b{i}_min_idx"] = np.array(o["block_min_idx"][i], np.int32)
            out[f"alg{alg}_proj{proj}_step"] = o["step"]
            out[f"alg{alg}_proj{proj}_gmin"] = o["grid_phys_mins"]
    np.savez_compressed(os.path.join(HERE, "dense_small.npz"), **out)

    # ---- config 1 (32^3 uniform, 1 block, 64^3 grid): hashes + a sample of values --------------------
    dom = ([0, 0, 0], [31, 31, 31])
    p = particles.gen_particles(0, *dom)
    b = decomp.regular_blocks(*dom, 1)
    blocks = delaunay.tessellate(p, decomp.assign_regular(p, b), b, *dom, workers=1)
    blk = blocks[0]
    v2t = R.fill_vert_to_tet(len(blk["particles"]), blk["tets"])
    blk["vert_to_tet"] = v2t
    c1 = dict(particles_sha=sha(blk["particles"]), tets_sha=sha(blk["tets"]))
    c1["particles_head"] = blk["particles"][:8]
    c1["cc_sha"] = sha(R.circumcenters(blk["tets"], blk["particles"]))
    c1["volume_sha"] = sha(R.volumes(blk["num_orig"], blk["tets"], blk["particles"], v2t))
    c1["complete_sum"] = np.array(int((R.complete(blk["num_orig"], blk["tets"], v2t) == 1).sum()), np.int64)
    rng = np.random.default_rng(1)
    sample = rng.integers(0, 64 ** 3, 4096)
    c1["sample_idx"] = sample
    for alg in (0, 1):
        o = R.dense(blocks, (64, 64, 64), alg=alg)
        g = o["grid"].reshape(-1)
        c1[f"alg{alg}_grid_sha"] = sha(g)
        c1[f"alg{alg}_sample"] = g[sample]
        c1[f"alg{alg}_sum"] = np.array(g.astype(np.float64).sum())
        c1[f"alg{alg}_nonzero"] = np.array(int((g != 0).sum()), np.int64)
    np.savez_compressed(os.path.join(HERE, "config1.npz"), **{k: np.asarray(v) for k, v in c1.items()})
    for f in ("dense_small.npz", "config1.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
