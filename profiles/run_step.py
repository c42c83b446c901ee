#!/usr/bin/env python
"""One resident dense() step of a bench.py workload, nothing else: the command the ncu launch lists are taken from.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \\
        --log-file gpurun_out/launches.csv python profiles/run_step.py --config 3 --steps 1 --warmup 1

(a number printed by a run under ncu is never a bench value; profiles/launches_r02.py turns the csv into the per-kernel
table and profiles/traffic.json)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--scale", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--alg", type=int, default=0)
    args = ap.parse_args()
    import tess2_b200
    w = bench.build_workload(args.config, 1, 0, args.scale)
    ctx = tess2_b200.Context(0)
    params = ctx.make_params(args.alg, w["ng"], w["dmin"], w["dmax"], False, (0.0, 0.0, 1.0), 1.0, 1e-4, w["gsize"])
    ctx.upload(w["blocks"])
    st = None
    for _ in range(args.warmup + args.steps):
        st = ctx.run(params)
    print(json.dumps({"config": args.config, "ms_total_device": st.ms_total_device, "launches": int(st.num_kernel_launches), "tets": int(st.num_tets),
                      "cells": int(st.num_cells), "grid_points": int(st.num_grid_pts), "spans": int(st.num_spans),
                      "particles": int(sum(len(b["particles"]) for b in w["blocks"]))}))
    ctx.close()


if __name__ == "__main__":
    main()
