#!/usr/bin/env python
"""Full-size parity of a BASELINE config: the GPU grid of `bench.py --config C` (one GPU) against the CPU oracle on ALL cells.

The oracle (oracle/dense_oracle.c, pinned to the unmodified reference) needs ~30 us per cell, so the cells are cut into one window
per host core and every window runs in its own process (cells are independent in the reference, src/dense.cpp:245-312).  A grid
point's value depends on the ORDER of the deposits it receives, so windows cannot be added up bit for bit; what can be compared:

  * a point that received deposits from exactly one window (the great majority): the GPU value must equal that window's value
    BIT FOR BIT (both started from zero and added the same deposits in the same order);
  * a point no window touched: the GPU value must be +0.0;
  * a point several windows touched: the GPU value against the float64 sum of the windows, relative tolerance 1e-5 (north_star's).

    python profiles/full_parity.py --config 3 [--scale 1]        (about two minutes of 16 host cores for config 3)
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

_W = {}


def _job(args):
    bi, lo, hi, alg, tag = args
    from oracle import ref
    port = ref.Checker("port")
    w = _W
    o = port.dense(w["blocks"], w["gsize"], alg=alg, given_bounds=None, only_gid=w["blocks"][bi]["gid"], first_cell=lo, max_cells=hi, assemble=False)
    paths = []
    for k, d in enumerate(o["block_density"]):
        if not d.any():
            paths.append(None)
            continue
        p = f"/dev/shm/tessb200_fullparity_{tag}_{bi}_{lo}_{k}.npy"
        np.save(p, d)
        paths.append(p)
    return bi, lo, hi, paths, int(o["out_of_range"]), [list(x) for x in o["block_num_idx"]]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--scale", type=int, default=1)
    ap.add_argument("--alg", type=int, default=0)
    args = ap.parse_args()
    import tess2_b200
    w = bench.build_workload(args.config, 1, 0, args.scale)
    blocks = bench.plain_blocks(w["blocks"])
    gs = w["gsize"]
    t0 = time.time()
    ctx = tess2_b200.Context(0)
    res = ctx.dense(args.alg, 0, None, None, False, (0.0, 0.0, 1.0), 1.0, 1e-4, gs, blocks, want_grid=False)
    gpu = {g: np.ascontiguousarray(d) for g, d in zip(res.gids, res.block_density)}
    ctx.close()
    t_gpu = time.time() - t0
    cores = os.cpu_count() or 1
    parts = max(1, cores // len(blocks))
    jobs = []
    for bi, b in enumerate(blocks):
        n = b["num_orig"]
        for k in range(parts):
            lo, hi = (k * n) // parts, ((k + 1) * n) // parts
            if hi > lo:
                jobs.append((bi, lo, hi, args.alg, os.getpid()))
    _W.update(blocks=blocks, gsize=gs)
    nb = len(blocks)
    first = [None] * nb
    count = [None] * nb
    total = [None] * nb
    oor = 0
    t0 = time.time()
    with mp.get_context("fork").Pool(min(cores, len(jobs))) as pool:
        for bi, lo, hi, paths, o_r, nums in pool.imap_unordered(_job, jobs, chunksize=1):
            oor += o_r
            for k, p in enumerate(paths):
                if p is None:
                    continue
                d = np.load(p)
                os.remove(p)
                if first[k] is None:
                    first[k] = np.zeros(d.shape, np.float32)
                    count[k] = np.zeros(d.shape, np.uint8)
                    total[k] = np.zeros(d.shape, np.float64)
                nz = d != 0
                new = nz & (count[k] == 0)
                first[k][new] = d[new]
                count[k] += nz.astype(np.uint8)
                total[k] += d
    t_cpu = time.time() - t0
    out = dict(config=bench.workload_name(args.config, 1, args.scale), alg=args.alg, cells=int(sum(b["num_orig"] for b in blocks)), windows=len(jobs),
               host_cores=cores, cpu_seconds=round(t_cpu, 1), gpu_seconds_incl_copies=round(t_gpu, 2), oracle_out_of_range_deposits=oor,
               grid_points=0, untouched_points=0, untouched_nonzero_on_gpu=0, single_window_points=0, single_window_bit_mismatches=0,
               multi_window_points=0, multi_window_max_rel_err=0.0)
    out["multi_window_beyond_1e-5"] = 0
    for k, b in enumerate(blocks):
        g = gpu[b["gid"]]
        if first[k] is None:
            first[k] = np.zeros(g.shape, np.float32); count[k] = np.zeros(g.shape, np.uint8); total[k] = np.zeros(g.shape, np.float64)
        assert g.shape == first[k].shape, (g.shape, first[k].shape)
        c = count[k]
        out["grid_points"] += int(g.size)
        z = c == 0
        out["untouched_points"] += int(z.sum())
        out["untouched_nonzero_on_gpu"] += int((g[z].view(np.uint32) != 0).sum())
        s = c == 1
        out["single_window_points"] += int(s.sum())
        out["single_window_bit_mismatches"] += int((g[s].view(np.uint32) != first[k][s].view(np.uint32)).sum())
        m = c > 1
        out["multi_window_points"] += int(m.sum())
        if m.any():
            rel = np.abs(g[m].astype(np.float64) - total[k][m]) / np.maximum(np.abs(total[k][m]), 1e-30)
            out["multi_window_beyond_1e-5"] += int((rel > 1e-5).sum())
            out["multi_window_max_rel_err"] = max(out["multi_window_max_rel_err"], float(rel.max()))
    out["ok"] = out["untouched_nonzero_on_gpu"] == 0 and out["single_window_bit_mismatches"] == 0 and out["multi_window_beyond_1e-5"] == 0
    print(json.dumps(out))


if __name__ == "__main__":
    main()
