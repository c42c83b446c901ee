#!/usr/bin/env python
"""bench.py -- the dense stage of tess2 on B200: grid points/s (and tets/s), device-resident and
end to end, with the kernel roofline and the CPU reference timed beside it.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1] at N = 1): 128^3 uniform particles drawn by the reference's
gen_particles (srand(gid)) in 8 regular blocks, Delaunay tets from the host engine (SciPy's
Qhull, options "Qt"), DENSE_TESS onto a 256^3 grid, mass 1, eps 1e-4, 3-D output.  At N > 1 every
rank owns one such 8-block slab of a (2 x 2 x 2N)-block domain (weak scaling: per-GPU work fixed),
grid 256 x 256 x 256N with the grid bounds given as the domain; boundary spans cross ranks with
NCCL inside the library.  A "step" is one pass of dense() over the resident blocks.

`value` is timed on the device (CUDA events on the library's stream, recorded inside
tessb200_dense_run: inputs resident in HBM -> grids complete in HBM), max over ranks.
`e2e` is the same metric through tessb200_dense() with pinned HOST buffers: H2D of particles and
tets, the run, D2H of every block's density, all inside the timed region.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H = int(os.environ.get("TESSB200_BENCH_H", "64"))          # particles per block per axis (64 -> 128^3 total at N = 1)
METRIC = "dense_grid_points_per_sec"
UNIT = "grid points/s"


def log(msg):
    print(f"[bench] {msg}", file=sys.stderr, flush=True)


def build_workload(n_ranks, rank):
    from tess2_b200.harness import workloads
    from tess2_b200 import multi
    blocks_xyz = (2, 2, 2 * n_ranks)
    nblocks = 8 * n_ranks
    owner = multi.assign_blocks(nblocks, n_ranks)
    gids = [g for g in range(nblocks) if owner[g] == rank]
    blocks, layout, dmin, dmax = workloads.uniform_regular(H, blocks_xyz, gids=gids, log=log if rank == 0 else None)
    gsize = (4 * H, 4 * H, 4 * H * n_ranks)
    return blocks, layout, owner, dmin, dmax, gsize


def workload_name(n):
    per = f"{2 * H}^3" if n == 1 else f"{2 * H}x{2 * H}x{2 * H * n}"
    return (f"tess-dense DENSE_TESS: {per} uniform gen_particles, {8 * n} regular blocks (8 per GPU), "
            f"gsize {4 * H}x{4 * H}x{4 * H * n}, SciPy-Qhull 'Qt' tets")


# ---- clocks ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for nme, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ---- CPU reference arm ---------------------------------------------------------------------------------
_CPU = {}


def _cpu_worker(args):
    gid, first, last, kind = args
    from oracle import ref
    chk = ref.Checker(kind)
    w = _CPU
    t0 = time.perf_counter()
    o = chk.dense(w["blocks"], w["gsize"], alg=0, given_bounds=w["given"], only_gid=gid, first_cell=first, max_cells=last)
    return gid, time.perf_counter() - t0, o["seconds"]


def cpu_jobs(blocks, max_cells, kind, cores):
    """The sample = the first max_cells cells of every block.  One OS process per block is the stand-in for one MPI
    rank per block; when the box has more cores than blocks every block's sample is cut into windows of cells
    (cells are independent in the reference: src/dense.cpp:245-312), so that all host cores work."""
    parts = max(1, cores // max(1, len(blocks)))
    smallest = min((min(b["num_orig"], max_cells) if max_cells >= 0 else b["num_orig"]) for b in blocks) if blocks else 0
    parts = max(1, min(parts, smallest // 4096))      # a window below ~4096 cells measures process start-up, not dense()
    jobs = []
    for b in blocks:
        n = min(b["num_orig"], max_cells) if max_cells >= 0 else b["num_orig"]
        for k in range(parts):
            lo, hi = (k * n) // parts, ((k + 1) * n) // parts
            if hi > lo:
                jobs.append((b["gid"], lo, hi, kind))
    return jobs, parts


def cpu_kind():
    return "reference" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libtess_ref.so")) else "port"


def cpu_dense_sample(blocks, gsize, given, max_cells, procs):
    """One bounded sample of the reference's dense() on the host cores: one OS process per block
    (the stand-in for one MPI rank per block), each visiting the first max_cells cells of its
    block.  Returns (wall seconds of the slowest process, cells visited)."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libtess_oracle.so")):
        subprocess.run(["make", "--no-print-directory", "port"], cwd=os.path.join(ROOT, "oracle"), check=True, stdout=subprocess.DEVNULL)
    kind = cpu_kind()
    _CPU.update(blocks=blocks, gsize=gsize, given=given)
    jobs, _ = cpu_jobs(blocks, max_cells, kind, procs)
    t0 = time.perf_counter()
    if procs <= 1:
        res = [_cpu_worker(j) for j in jobs]
    else:
        with mp.get_context("fork").Pool(min(procs, len(jobs))) as pool:
            res = pool.map(_cpu_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    cells = sum(min(b["num_orig"], max_cells) if max_cells >= 0 else b["num_orig"] for b in blocks)
    slowest = max(r[2] for r in res)
    return wall, slowest, cells, kind


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU dense() (unmodified sources when oracle/_ref was
    built, else the C restatement) on this box's host cores, bounded sample per step."""
    if rank != 0:
        return
    n = args.gpus
    blocks, layout, owner, dmin, dmax, gsize = build_workload(n, 0)
    given = (dmin, dmax) if n > 1 else None
    cores = os.cpu_count() or 1
    procs = cores
    cells_total = sum(b["num_orig"] for b in blocks) * n
    G_total = gsize[0] * gsize[1] * gsize[2]
    max_cells = int(os.environ.get("TESSB200_CPU_SAMPLE_CELLS", str(max(1024, (H ** 3) // 8))))
    # layout-wide bounds are needed for DataBounds: hand every block of the decomposition to the checker,
    # non-local ones empty
    allb = list(blocks)
    have = {b["gid"] for b in blocks}
    for gid, mn, mx in layout:
        if gid not in have:
            allb.append(dict(gid=gid, particles=np.zeros((0, 3), np.float32), tets=np.zeros((0, 8), np.int32), num_orig=0,
                             bounds_min=mn, bounds_max=mx, vert_to_tet=np.zeros(0, np.int32)))
    allb.sort(key=lambda b: b["gid"])
    _CPU.update(blocks=allb, gsize=gsize, given=given)
    kind = cpu_kind()
    times = []
    for it in range(args.warmup + args.steps):
        jobs, parts = cpu_jobs(blocks, max_cells, kind, cores)
        procs = min(cores, len(jobs))
        t0 = time.perf_counter()
        with mp.get_context("fork").Pool(procs) as pool:
            pool.map(_cpu_worker, jobs, chunksize=1)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    cells_step = sum(min(b["num_orig"], max_cells) for b in blocks)
    t = float(np.mean(times))
    value = G_total * (cells_step / cells_total) / t
    tets_total = sum(len(b["tets"]) for b in blocks) * n
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(n), "alg": "DENSE_TESS", "sample": f"first {max_cells} cells of each of {len(blocks)} blocks per step"},
        "tets_per_sec": tets_total * (cells_step / cells_total) / t,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind,
                         "sample": f"{cells_step} of {cells_total} cells per step ({len(blocks)} blocks x first {max_cells} cells, {parts} window(s) of cells "
                                   f"per block), {procs} processes on {cores} host cores, throughput scaled by cells"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- our arm ----------------------------------------------------------------------------------------------
def pinned_copy(a):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t, t.numpy()


def run_ours(args, rank, world, local_rank):
    # stdout carries exactly one JSON line: anything libraries print there (NCCL's version banner)
    # is diverted to stderr for the duration of the run
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import tess2_b200
    from tess2_b200 import multi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; tess2_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = world
    blocks, layout, owner, dmin, dmax, gsize = build_workload(n, rank)
    from tess2_b200.harness import workloads as _wl
    host_tess = dict(_wl.LAST_TESS)
    # the package's own host driver + Delaunay engine on the same particles (same set of tets), timed beside SciPy's Qhull
    native_tess = None
    if n == 1 and not args.no_cpu_baseline:
        t0 = time.perf_counter()
        nb_blocks, _, _, _ = _wl.uniform_regular(H, (2, 2, 2), gids=list(range(8)), cache=False, engine="native")
        native_tess = {"engine": "tess2_b200/host (C++ incremental Delaunay, exact predicates), one thread per block", "seconds": time.perf_counter() - t0,
                       "threads": min(8, os.cpu_count() or 1), "tets": int(sum(len(b["tets"]) for b in nb_blocks)),
                       "same_tet_count_as_qhull": int(sum(len(b["tets"]) for b in nb_blocks)) == int(sum(len(b["tets"]) for b in blocks))}
        del nb_blocks
    keep = []
    for b in blocks:                      # pinned host buffers: the e2e leg copies from these
        for k in ("particles", "tets", "vert_to_tet"):
            t, b[k] = pinned_copy(b[k])
            keep.append(t)
    ctx = tess2_b200.Context(local_rank)
    if world > 1:
        multi.init_comm(ctx, layout, owner)
    ng = 3 if n > 1 else 0
    params = ctx.make_params(tess2_b200.DENSE_TESS, ng, dmin, dmax, False, (0.0, 0.0, 1.0), 1.0, 1e-4, gsize)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------------
    ctx.upload(blocks)
    sampler = ClockSampler(local_rank)
    sampler.start()             # sampled every 20 ms from the warm-up to the end of the e2e leg (both timed regions)
    for _ in range(args.warmup):
        ctx.run(params)
    barrier()
    t0 = time.perf_counter()
    stage = {k: 0.0 for k in ("ms_circumcenters", "ms_cells", "ms_bfs", "ms_nbrs", "ms_faces", "ms_scan", "ms_exchange", "ms_sort", "ms_deposit",
                              "ms_slow_path", "ms_total_device")}
    launches = 0
    st = None
    for _ in range(args.steps):
        st = ctx.run(params)
        for k in stage:
            stage[k] += getattr(st, k)
        launches += st.num_kernel_launches
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = multi.max_over_ranks(stage["ms_total_device"] / args.steps)
    wall_ms = multi.max_over_ranks(1e3 * wall / args.steps)
    G_local = int(st.num_grid_pts)
    G_total = gsize[0] * gsize[1] * gsize[2]
    T_local = int(st.num_tets)
    T_total = int(multi.sum_over_ranks(T_local))
    P_local = sum(len(b["particles"]) for b in blocks)
    cells_local = int(st.num_cells)
    spans = int(st.num_spans)
    value = G_total / (dev_ms * 1e-3)

    # ---- end to end through tessb200_dense(): pinned host buffers in, host densities out -----------
    geo = ctx.geometry(params)
    out_t = []
    out_blocks = []
    for (_, _, npts) in geo:
        t = torch.empty(npts, dtype=torch.float32).pin_memory()
        out_t.append(t)
        out_blocks.append(t.numpy())
    h2d = sum(b["particles"].nbytes + b["tets"].nbytes + b["vert_to_tet"].nbytes for b in blocks)
    d2h = sum(o.nbytes for o in out_blocks)

    def e2e_step():
        # the call a user makes: tessb200_dense(), host buffers in, host buffers out
        ctx.dense_params(params, blocks, want_grid=False, out_blocks=out_blocks, want_stats=False)

    for _ in range(max(1, min(args.warmup, 3))):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_ms = multi.max_over_ranks(1e3 * (time.perf_counter() - t0) / args.steps)
    e2e_value = G_total / (e2e_ms * 1e-3)
    clocks = sampler.stop()
    checksum = float(sum(float(o.astype(np.float64).sum()) for o in out_blocks))

    # ---- the other two estimators on the same resident inputs (BASELINE config 5 compares CIC with the
    # tessellation estimator; DTFE is the repo's first-order mode): a few steps each, reported beside ----
    other = {}
    ctx.upload(blocks)
    for name, alg in (("DENSE_CIC", tess2_b200.DENSE_CIC), ("DENSE_DTFE (not in the reference)", tess2_b200.DENSE_DTFE)):
        pa = ctx.make_params(alg, ng, dmin, dmax, False, (0.0, 0.0, 1.0), 1.0, 1e-4, gsize)
        for _ in range(2):
            ctx.run(pa)
        barrier()
        ms = 0.0
        reps = 3
        for _ in range(reps):
            ms += ctx.run(pa).ms_total_device
        barrier()
        ms = multi.max_over_ranks(ms / reps)
        other[name] = {"ms_per_step": ms, "grid_points_per_sec": G_total / (ms * 1e-3)}

    # ---- roofline of the dominant kernel (algorithmic bytes: DESIGN.md section 4) -----------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak = 6650.0; peak_src = "fallback (B200_PROFILING.md)"
    P_orig = cells_local
    F = int(st.num_faces)              # plane records (Voronoi faces of accepted cells, padded to pairs)
    Cn = int(st.num_candidates)        # candidate neighbours = sum over accepted cells of (star tets + 2)
    star = max(Cn - 2 * int(st.num_deposit_cells), 0)   # star tets visited by the BFS (~27 per cell)
    alg_bytes = {
        "k_circumcenters": 32 * T_local + 12 * P_local,            # 16 B verts + 16 B float4 out per tet, particles once
        "k_cell_bfs": 48 * T_local + 16 * P_orig + 8 * Cn + 32 * P_orig,   # each tet record + circumcenter once, site + v2t, candidates + pre-header out
        "k_cell_nbrs": 8 * Cn + 32 * P_orig + 16 * F + 32 * P_orig,       # candidates + pre-header in, face list + header out
        "k_cell_faces": 16 * F + 48 * T_local + 24 * F,                   # face refs in, tet records + circumcenters once, planes out
        "k_cell_scan": 24 * F + 32 * P_orig + 16 * spans,                 # planes + headers in, span records out
        "sort (cub radix, 64-bit key + 64-bit payload)": 2 * 16 * spans,
        "k_rows": 16 * spans + 4 * G_local,                               # span records in, every grid point written once
    }
    stage_ms = {"k_circumcenters": stage["ms_circumcenters"], "k_cell_bfs": stage["ms_bfs"], "k_cell_nbrs": stage["ms_nbrs"],
                "k_cell_faces": stage["ms_faces"], "k_cell_scan": stage["ms_scan"],
                "sort (cub radix, 64-bit key + 64-bit payload)": stage["ms_sort"], "k_rows": stage["ms_deposit"]}
    n_shared = int(st.num_shared_deposits)
    if n_shared >= 0:
        # 3-D runs: deposits that are alone on their grid point are written directly, only the shared ones are sorted
        del stage_ms["sort (cub radix, 64-bit key + 64-bit payload)"], stage_ms["k_rows"]
        del alg_bytes["sort (cub radix, 64-bit key + 64-bit payload)"], alg_bytes["k_rows"]
        stage_ms["k_span_count + k_span_place"] = stage["ms_sort"]
        alg_bytes["k_span_count + k_span_place"] = 2 * 16 * spans + 3 * 4 * G_local      # records read twice; count grid cleared + grid cleared and written
        stage_ms["sort + k_rows of the shared grid points"] = stage["ms_deposit"]
        alg_bytes["sort + k_rows of the shared grid points"] = 3 * 16 * n_shared
    stage_ms["nccl span exchange"] = stage["ms_exchange"]
    alg_bytes["nccl span exchange"] = 0
    # oversized stars / index boxes (k_cell_bfs_big, k_cell_scan_big and their faces), run once after the fast kernels
    alg_bytes["slow path (oversized cells)"] = 0
    stage_ms["slow path (oversized cells)"] = stage["ms_slow_path"]
    stage_ms = {k: v / args.steps for k, v in stage_ms.items()}
    dom = max((k for k in stage_ms if k.startswith("k_")), key=lambda k: stage_ms[k])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(dom)
    ach = alg_bytes[dom] / (stage_ms[dom] * 1e-3) / 1e9 if stage_ms[dom] > 0 else 0.0
    whole = 32 * T_local + 16 * P_local + 4 * G_local
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes": alg_bytes[dom], "kernel_ms": stage_ms[dom],
                "note": "the cell kernels are latency / L2-sector / issue bound, not HBM bound: see DESIGN.md 3 and profiles/",
                "stages": {k: {"ms": stage_ms[k], "algorithmic_GBps": (alg_bytes[k] / (stage_ms[k] * 1e-3) / 1e9 if stage_ms[k] > 0 else None)} for k in stage_ms},
                "whole_stage": {"algorithmic_bytes": whole, "achieved": whole / (dev_ms * 1e-3) / 1e9, "frac": whole / (dev_ms * 1e-3) / 1e9 / peak}}

    # ---- CPU baseline beside it (rank 0, N = 1 only) -------------------------------------------------------
    cpu = None
    if rank == 0 and n == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        max_cells = int(os.environ.get("TESSB200_CPU_SAMPLE_CELLS", str(max(1024, (H ** 3) // 4))))
        plain = [dict(b, particles=np.array(b["particles"]), tets=np.array(b["tets"]), vert_to_tet=np.array(b["vert_to_tet"])) for b in blocks]
        wall_s, slowest, cells, kind = cpu_dense_sample(plain, gsize, None, max_cells, cores)
        njobs = len(cpu_jobs(plain, max_cells, kind, cores)[0])
        cpu_value = G_total * (cells / (cells_local)) / wall_s
        cpu = {"value": cpu_value, "unit": UNIT, "cores": min(cores, njobs), "kind": kind,
               "sample": f"{cells} of {cells_local} cells ({len(blocks)} blocks x first {max_cells} cells, cut into {njobs} windows of cells), "
                         f"{min(cores, njobs)} processes on {cores} host cores, {wall_s:.1f} s wall, throughput scaled by cells"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(n), "alg": "DENSE_TESS", "mass": 1.0, "eps": 1e-4,
                       "l2": "inputs larger than L2 (tets %.0f MB per GPU vs 126 MB L2), no flush" % (32 * T_local / 1e6),
                       "timing": "CUDA events on the library stream inside tessb200_dense_run, max over ranks"},
            "tets_per_sec": T_total / (dev_ms * 1e-3), "cells_per_sec": cells_local * n / (dev_ms * 1e-3),
            "ms_per_step_wall": wall_ms,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms,
                    "checksum_sum_density": checksum},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "other_algs": other,
            "host_tess": {"engine": "SciPy Qhull 'Qt', one process per block", "seconds": host_tess.get("seconds"), "workers": host_tess.get("workers"),
                          "from_cache": host_tess.get("cached"),
                          "tess_plus_dense_seconds": (host_tess["seconds"] + e2e_ms * 1e-3) if host_tess.get("seconds") else None,
                          "native": native_tess,
                          "native_tess_plus_dense_seconds": (native_tess["seconds"] + e2e_ms * 1e-3) if native_tess else None},
            "stats": {"shared_deposits": int(st.num_shared_deposits), "cells": cells_local, "tets": T_local, "particles_with_ghosts": P_local, "grid_points": G_local, "spans": spans, "faces": F, "candidates": Cn,
                      "deposit_cells": int(st.num_deposit_cells), "cic_fallback_cells": int(st.num_cic_fallback), "slow_cells": int(st.num_slow_cells),
                      "tot_mass": float(st.tot_mass)},
        }
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus:
        if args.gpus > 1 and world == 1:
            # convenience: re-launch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                   "--master-port", os.environ.get("MASTER_PORT", "29531"), os.path.abspath(__file__), "--gpus", str(args.gpus),
                   "--steps", str(args.steps), "--warmup", str(args.warmup)] + (["--no-cpu-baseline"] if args.no_cpu_baseline else [])
            raise SystemExit(subprocess.call(cmd))
        raise SystemExit(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
