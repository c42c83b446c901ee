"""Where the one-call (host buffers in, host buffers out) time goes: raw pinned copy rates of this
box beside the stats of tessb200_dense() on the bench workload.  GPU box: python profiles/probe_e2e.py"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import tess2_b200  # noqa: E402

blocks, layout, owner, dmin, dmax, gsize = bench.build_workload(1, 0)
keep = []
for b in blocks:
    for k in ("particles", "tets", "vert_to_tet"):
        t, b[k] = bench.pinned_copy(b[k])
        keep.append(t)
h2d = sum(b["particles"].nbytes + b["tets"].nbytes + b["vert_to_tet"].nbytes for b in blocks)
out = {"h2d_bytes": h2d}

# raw copy rates, pinned memory
src = torch.empty(h2d, dtype=torch.uint8).pin_memory()
dst = torch.empty(h2d, dtype=torch.uint8, device="cuda")
back = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
for name, fn, nbytes in (("h2d", lambda: dst.copy_(src, non_blocking=True), h2d),
                         ("d2h", lambda: back.copy_(dst[:64 << 20], non_blocking=True), 64 << 20)):
    best = 1e9
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    out[name + "_GBps"] = nbytes / best / 1e9
    out[name + "_ms"] = best * 1e3

ctx = tess2_b200.Context(0)
params = ctx.make_params(tess2_b200.DENSE_TESS, 0, dmin, dmax, False, (0.0, 0.0, 1.0), 1.0, 1e-4, gsize)
ctx.upload(blocks)
geo = ctx.geometry(params)
outs = [torch.empty(npts, dtype=torch.float32).pin_memory() for (_, _, npts) in geo]
out_blocks = [t.numpy() for t in outs]
steps = []
for i in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = ctx.dense_params(params, blocks, want_grid=False, out_blocks=out_blocks, want_stats=True)
    wall = 1e3 * (time.perf_counter() - t0)
    st = res.stats
    steps.append({"wall_ms": wall, "ms_upload": st.ms_upload, "ms_total_device": st.ms_total_device, "ms_download": st.ms_download,
                  "ms_cells": st.ms_cells, "ms_scan": st.ms_scan, "ms_sort": st.ms_sort, "ms_deposit": st.ms_deposit})
out["steps"] = steps[2:]
print(json.dumps(out, indent=1))
