#!/usr/bin/env python
"""bench.py -- the dense stage of tess2 on B200: grid points/s (and tets/s), device-resident and
end to end, with the kernel roofline and the CPU reference timed beside it.

    python bench.py [--config 2|3|4|5] --gpus N --steps K --warmup W [--impl reference]

Workloads = BASELINE.json configs; the default is config 3, the one BASELINE quotes at 1/2/4/8 GPUs:

  3 (default)  256^3 clustered Gaussian-clump particles, kd-tree decomposition into 8 blocks, gsize 512^3, DENSE_TESS.
               STRONG scaling: the same particles, blocks and grid whatever N is; rank r owns blocks
               [8r/N, 8(r+1)/N) and their grid slabs, boundary span records cross ranks with NCCL inside the library.
  4            512^3 clustered, kd-tree 64 blocks, gsize 1024^3 (8 GPUs: 8 blocks per rank).  Strong scaling.
  5            config 3's input through DENSE_TESS and DENSE_CIC (the comparison BASELINE names); `value` is the
               CIC estimator, the tessellation estimator is reported beside it.
  2            128^3 uniform gen_particles (srand(gid)) in 8 regular blocks, gsize 256^3, 1 GPU; at N > 1 every rank
               owns one such 8-block slab (weak scaling, round 1's line).

Tets come from the package's own host tess() (tess2_b200/host: C++ incremental Delaunay, exact predicates), each rank
tessellating its own blocks on its share of the host cores.  A "step" is one pass of dense() over the resident blocks.

`value` is timed on the device (CUDA events on the library's stream inside tessb200_dense_run: inputs resident in
HBM -> grids complete in HBM), max over ranks.  `e2e` is the same metric through tessb200_dense() with pinned HOST
buffers: H2D of particles and tets, the run, D2H of every block's density, all inside the timed region.
`parity` is checked in the same process before anything is timed: a small clone of the workload (same generator, same
decomposition, same multi-GPU path) compared bit for bit with the CPU oracle, the mass balance of the full-size run
summed over ranks, and a sha256 of the full-size grid that must not depend on N.
"""
import argparse
import hashlib
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "dense_grid_points_per_sec"
UNIT = "grid points/s"
CACHE_DIR = os.environ.get("TESSB200_CACHE", "/tmp/tess2_b200_cache")
H = int(os.environ.get("TESSB200_BENCH_H", "64"))          # config 2: particles per block per axis

# side = particles per axis, nb = kd-tree blocks, g = grid points per axis
CLUSTERED = {3: dict(side=256, nb=8, g=512, seed=2027), 4: dict(side=512, nb=64, g=1024, seed=2028), 5: dict(side=256, nb=8, g=512, seed=2027)}


def log(msg):
    print(f"[bench] {msg}", file=sys.stderr, flush=True)


# ---- workloads ---------------------------------------------------------------------------------------
def clustered_spec(cfg, scale):
    s = dict(CLUSTERED[cfg])
    if scale > 1:       # development only: never a bench line (the workload name says so)
        s["side"] = max(8, s["side"] // scale)
        s["g"] = max(16, s["g"] // scale)
    return s


def workload_name(cfg, n, scale=1):
    if cfg == 2:
        per = f"{2 * H}^3" if n == 1 else f"{2 * H}x{2 * H}x{2 * H * n}"
        return (f"config 2: tess-dense DENSE_TESS, {per} uniform gen_particles, {8 * n} regular blocks (8 per GPU), "
                f"gsize {4 * H}x{4 * H}x{4 * H * n}")
    s = clustered_spec(cfg, scale)
    what = "DENSE_CIC vs DENSE_TESS" if cfg == 5 else "DENSE_TESS"
    return (f"config {cfg}: tess-dense {what}, {s['side']}^3 clustered Gaussian-clump particles, kd-tree {s['nb']} blocks, "
            f"gsize {s['g']}^3" + (f" [REDUCED by {scale}: development run, not a bench line]" if scale > 1 else ""))


def _wait_for(path, timeout=3600.0):
    t0 = time.time()
    while not os.path.exists(path):
        if time.time() - t0 > timeout:
            raise RuntimeError(f"timed out waiting for {path}")
        time.sleep(0.2)


def clustered_inputs(spec, rank, tag):
    """Particles + kd-tree decomposition, generated once per box (rank 0) and shared through the cache directory:
    every rank needs all particles to find the ghosts of its blocks."""
    from tess2_b200 import host_tess
    from tess2_b200.harness import particles
    d = os.path.join(CACHE_DIR, f"clustered_{spec['side']}_{spec['nb']}_{spec['seed']}")
    done = os.path.join(d, "inputs.done")
    t0 = time.time()
    if rank == 0 and not os.path.exists(done):
        os.makedirs(d, exist_ok=True)
        n = spec["side"]
        dom = (np.zeros(3, np.float32), np.full(3, n - 1, np.float32))
        p = particles.clustered_particles(n ** 3, *dom, seed=spec["seed"])
        bounds, owner = host_tess.kdtree_blocks(p, *dom, spec["nb"])
        np.save(os.path.join(d, "p.npy"), p)
        np.save(os.path.join(d, "owner.npy"), owner)
        np.save(os.path.join(d, "bounds.npy"), np.array([np.concatenate([mn, mx]) for mn, mx in bounds], np.float32))
        with open(done + f".tmp{tag}", "w") as f:
            f.write(f"{time.time() - t0:.2f}\n")
        os.replace(done + f".tmp{tag}", done)
    _wait_for(done)
    gen_s = float(open(done).read().split()[0])
    p = np.load(os.path.join(d, "p.npy"), mmap_mode="r")
    owner = np.load(os.path.join(d, "owner.npy"), mmap_mode="r")
    b6 = np.load(os.path.join(d, "bounds.npy"))
    bounds = [(b6[g, :3].copy(), b6[g, 3:].copy()) for g in range(len(b6))]
    return d, p, owner, bounds, gen_s, time.time() - t0


def clustered_blocks(spec, d, p, owner, bounds, gids, threads):
    """The host tess() of this rank's blocks, cached per block on the box (both arms and every N reuse them)."""
    from tess2_b200 import host_tess
    n = spec["side"]
    dom = (np.zeros(3, np.float32), np.full(3, n - 1, np.float32))
    keys = ("particles", "tets", "vert_to_tet")
    max_gb = float(os.environ.get("TESSB200_CACHE_MAX_GB", "12"))
    blocks, missing = {}, []
    for g in gids:
        meta = os.path.join(d, f"blk{g}.json")
        if os.path.exists(meta):
            try:
                m = json.load(open(meta))
                b = dict(gid=g, num_orig=m["num_orig"], bounds_min=bounds[g][0], bounds_max=bounds[g][1], rounds=m["rounds"], seconds=m["seconds"],
                         settled=m.get("settled", True), cached=True, made_in=m.get("made_in"), made_with=m.get("made_with"))
                for k in keys:
                    b[k] = np.load(os.path.join(d, f"blk{g}_{k}.npy"))
                blocks[g] = b
                continue
            except Exception as e:       # a half-written cache entry: tessellate again
                log(f"cache entry of block {g} unusable ({e})")
        missing.append(g)
    t0 = time.time()
    if missing:
        made = host_tess.tess(np.asarray(p), np.asarray(owner), bounds, *dom, threads=threads, gids=missing)
        wall = time.time() - t0
        for b in made:
            b["cached"] = False
            b["made_in"], b["made_with"] = wall, f"{len(missing)} blocks on {threads} threads"
            b.pop("global_ids", None)
            blocks[b["gid"]] = b
            nbytes = sum(b[k].nbytes for k in keys)
            if nbytes * len(bounds) < max_gb * 1e9:
                try:
                    for k in keys:
                        np.save(os.path.join(d, f"blk{b['gid']}_{k}.npy"), b[k])
                    tmp = os.path.join(d, f"blk{b['gid']}.json.tmp{os.getpid()}")
                    json.dump(dict(num_orig=int(b["num_orig"]), rounds=int(b["rounds"]), seconds=float(b["seconds"]), settled=bool(b["settled"]),
                                   made_in=b["made_in"], made_with=b["made_with"]), open(tmp, "w"))
                    os.replace(tmp, os.path.join(d, f"blk{b['gid']}.json"))
                except OSError as e:
                    log(f"block cache not written ({e})")
    return [blocks[g] for g in gids], time.time() - t0, len(missing)


def build_workload(cfg, n_ranks, rank, scale=1, gids=None, threads=None):
    """Returns dict(blocks, layout, owner, dmin, dmax, gsize, ng, scaling, host)."""
    from tess2_b200 import multi
    cores = os.cpu_count() or 1
    if cfg == 2:
        from tess2_b200.harness import workloads
        blocks_xyz = (2, 2, 2 * n_ranks)
        nblocks = 8 * n_ranks
        owner = multi.assign_blocks(nblocks, n_ranks)
        mine = [g for g in range(nblocks) if owner[g] == rank] if gids is None else gids
        t0 = time.time()
        blocks, layout, dmin, dmax = workloads.uniform_regular(H, blocks_xyz, gids=mine, log=log if rank == 0 else None, engine="native",
                                                               workers=threads or max(1, cores // n_ranks))
        host = dict(engine="tess2_b200/host (C++ incremental Delaunay, exact predicates)", tess_seconds=workloads.LAST_TESS.get("seconds"),
                    from_cache=bool(workloads.LAST_TESS.get("cached")), wall_seconds=time.time() - t0)
        return dict(blocks=blocks, layout=layout, owner=owner, dmin=dmin, dmax=dmax, gsize=(4 * H, 4 * H, 4 * H * n_ranks),
                    ng=3 if n_ranks > 1 else 0, scaling="weak", host=host, particles_total=(2 * H) ** 3 * n_ranks)
    spec = clustered_spec(cfg, scale)
    d, p, powner, bounds, gen_s, wait_s = clustered_inputs(spec, rank, f"{os.getpid()}")
    nb = spec["nb"]
    owner = multi.assign_blocks(nb, n_ranks)
    mine = [g for g in range(nb) if owner[g] == rank] if gids is None else gids
    thr = threads or max(1, cores // n_ranks)
    blocks, tess_s, n_made = clustered_blocks(spec, d, p, powner, bounds, mine, thr)
    n = spec["side"]
    dmin, dmax = np.zeros(3, np.float32), np.full(3, n - 1, np.float32)
    layout = [(g, bounds[g][0], bounds[g][1]) for g in range(nb)]
    # wall time of the host tess() of this rank's blocks: now, or -- blocks from the box's cache -- when they were made
    if n_made == 0 and blocks:
        tess_s = max((b.get("made_in") or 0.0) for b in blocks)
    host = dict(engine="tess2_b200/host (C++ incremental Delaunay, exact predicates)", generate_seconds=gen_s, tess_seconds=tess_s,
                tess_made=sorted({str(b.get("made_with")) for b in blocks}), engine_seconds_per_block=[round(float(b["seconds"]), 2) for b in blocks],
                blocks_tessellated_now=n_made, blocks_from_cache=len(mine) - n_made, threads=thr,
                max_ghost_rounds=int(max([b["rounds"] for b in blocks], default=0)), all_blocks_settled=bool(all(b.get("settled", True) for b in blocks)))
    # ng = 0: DataBounds comes from the layout (every block's bounds), so the grid is the same at every N
    return dict(blocks=blocks, layout=layout, owner=owner, dmin=dmin, dmax=dmax, gsize=(spec["g"],) * 3, ng=0, scaling="strong", host=host,
                particles_total=int(len(p)))


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


def cpu_kind():
    return "reference" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libtess_ref.so")) else "port"


def _cpu_worker(job):
    """One window of cells of ONE block through the reference's dense().  The worker hands the reference only that
    block: blocks it does not work on would still be allocated and zeroed by init_dense (src/dense.cpp:106-127).  Two
    empty one-percent slabs at the bottom and the top of the domain keep DataBounds (src/dense.cpp:1221-1275: min / max
    over the blocks' bounds) equal to the full run's, so the data-bounds filter of CellGridPts sees the same cells.
    Deposits that leave the block have no neighbour to go to and are dropped after diy::in -- the enqueue the reference
    would do is a vector push_back."""
    bi, first, last, kind, alg = job
    from oracle import ref
    chk = ref.Checker(kind)
    w = _CPU
    o = chk.dense([w["blocks"][bi]] + w["phantoms"], w["gsize"], alg=alg, given_bounds=w["given"], first_cell=first, max_cells=last, assemble=False)
    maps = ""
    try:
        for line in open("/proc/self/maps"):
            if "libtess_ref" in line or "libtess_oracle" in line:
                maps = line.split()[-1]
                break
    except OSError:
        pass
    return o["seconds"], maps


def cpu_jobs(blocks, max_cells, kind, cores, alg=0):
    """The sample = max_cells cells of every block, cut into windows of cells (cells are independent in the reference,
    src/dense.cpp:245-312; the cells before a window reach it with vert_to_tet = -1 and are skipped at :251) so that
    exactly one window runs on every host core at the same time."""
    nb = max(1, len(blocks))
    parts = max(1, cores // nb)
    jobs = []
    for bi, b in enumerate(blocks):
        n = min(b["num_orig"], max_cells)
        for k in range(parts):
            lo, hi = (k * n) // parts, ((k + 1) * n) // parts
            if hi > lo:
                jobs.append((bi, lo, hi, kind, alg))
    return jobs[:cores] if len(jobs) > cores else jobs


def cpu_dense_sample(blocks, gsize, given, max_cells, cores, alg=0):
    """One bounded sample on the host cores.  Returns (seconds, cells visited, kind, processes, maps line): seconds =
    the slowest worker's COMP_TIME (the interval the reference's drivers time around dense(),
    examples/tess-dense/main.cpp:213-220); every worker runs at the same time on its own core."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libtess_oracle.so")):
        subprocess.run(["make", "--no-print-directory", "port"], cwd=os.path.join(ROOT, "oracle"), check=True, stdout=subprocess.DEVNULL)
    kind = cpu_kind()
    from oracle import ref
    ref.Checker(kind)            # mapped in this process too: the driver's record of loaded native code sees it
    dmin, dmax = np.asarray(given[0], np.float32), np.asarray(given[1], np.float32)
    thick = np.float32(0.01) * (dmax[2] - dmin[2])
    empty = dict(num_orig=0, particles=np.zeros((0, 3), np.float32), tets=np.zeros((0, 8), np.int32), vert_to_tet=np.zeros(0, np.int32))
    top = max(b["gid"] for b in blocks)
    phantoms = [dict(empty, gid=top + 1, bounds_min=dmin, bounds_max=np.array([dmax[0], dmax[1], dmin[2] + thick], np.float32)),
                dict(empty, gid=top + 2, bounds_min=np.array([dmin[0], dmin[1], dmax[2] - thick], np.float32), bounds_max=dmax)]
    _CPU.update(blocks=blocks, gsize=gsize, given=given, phantoms=phantoms)
    jobs = cpu_jobs(blocks, max_cells, kind, cores, alg)
    with mp.get_context("fork").Pool(len(jobs)) as pool:
        res = pool.map(_cpu_worker, jobs, chunksize=1)
    cells = sum(j[2] - j[1] for j in jobs)
    return max(r[0] for r in res), cells, kind, len(jobs), res[0][1]


def plain_blocks(blocks):
    return [dict(gid=b["gid"], num_orig=b["num_orig"], bounds_min=b["bounds_min"], bounds_max=b["bounds_max"],
                 particles=np.asarray(b["particles"]), tets=np.asarray(b["tets"]), vert_to_tet=np.asarray(b["vert_to_tet"])) for b in blocks]


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU dense() (unmodified sources when oracle/_ref was built, else the C
    restatement) on this box's host cores, a bounded sample of the same workload per step."""
    if rank != 0:
        return
    n, cfg = args.gpus, args.config
    cores = os.cpu_count() or 1
    # every block of the workload (config 2 at N > 1: the first slab stands for all, they are statistically alike)
    if cfg == 2:
        w = build_workload(cfg, n, 0, args.scale, threads=cores)
        cells_total = sum(b["num_orig"] for b in w["blocks"]) * n
        tets_total = sum(len(b["tets"]) for b in w["blocks"]) * n
    else:
        nb = clustered_spec(cfg, args.scale)["nb"]
        use = list(range(nb)) if nb <= 16 else list(range(0, nb, nb // 8))       # config 4: 8 of the 64 blocks
        w = build_workload(cfg, 1, 0, args.scale, gids=use, threads=cores)
        frac = len(use) / nb
        cells_total = int(round(sum(b["num_orig"] for b in w["blocks"]) / frac))
        tets_total = int(round(sum(len(b["tets"]) for b in w["blocks"]) / frac))
    blocks = plain_blocks(w["blocks"])
    gsize = w["gsize"]
    given = (w["dmin"], w["dmax"])
    G_total = gsize[0] * gsize[1] * gsize[2]
    max_cells = int(os.environ.get("TESSB200_CPU_SAMPLE_CELLS", "65536"))
    alg = 1 if cfg == 5 else 0
    times = []
    for it in range(args.warmup + args.steps):
        t, cells_step, kind, procs, maps = cpu_dense_sample(blocks, gsize, given, max_cells, cores, alg)
        if it >= args.warmup:
            times.append(t)
    t = float(np.mean(times))
    value = G_total * (cells_step / cells_total) / t
    sample = (f"{cells_step} of {cells_total} cells per step ({procs} windows of cells over {len(blocks)} blocks, one process per window, all at the "
              f"same time on {cores} host cores); seconds = the slowest process's dense() interval (COMP_TIME); throughput scaled by cells")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(cfg, n, args.scale), "alg": "DENSE_CIC" if alg else "DENSE_TESS", "mass": 1.0, "eps": 1e-4},
        "tets_per_sec": tets_total * (cells_step / cells_total) / t,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample, "native_so": maps},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- parity, checked before anything is timed -------------------------------------------------------------
def small_clone(cfg):
    """A reduced copy of the workload: same generator and decomposition, sizes the CPU oracle finishes in seconds."""
    from tess2_b200 import host_tess
    from tess2_b200.harness import particles
    if cfg == 2:
        dom = (np.zeros(3, np.float32), np.full(3, 23, np.float32))
        bounds = host_tess.regular_blocks(*dom, 8)
        ps = [particles.gen_particles(g, mn, mx) for g, (mn, mx) in enumerate(bounds)]
        p = np.concatenate(ps)
        owner = np.concatenate([np.full(len(q), g, np.int32) for g, q in enumerate(ps)])
        name, gs = "24^3 gen_particles, 8 regular blocks, gsize 48^3", (48, 48, 48)
    else:
        nb = CLUSTERED[cfg]["nb"]
        side = 48
        dom = (np.zeros(3, np.float32), np.full(3, side - 1, np.float32))
        p = particles.clustered_particles(side ** 3, *dom, seed=CLUSTERED[cfg]["seed"], n_clumps=16)
        bounds, owner = host_tess.kdtree_blocks(p, *dom, nb)
        name, gs = f"{side}^3 clustered, kd-tree {nb} blocks, gsize {2 * side}^3", (2 * side,) * 3
    blocks = host_tess.tess(p, owner, bounds, *dom, threads=2)
    for b in blocks:
        b.pop("global_ids", None)
    layout = [(g, bounds[g][0], bounds[g][1]) for g in range(len(bounds))]
    return name, blocks, layout, gs


def check_small_clone(ctx, cfg, world, rank, algs):
    """Every rank runs its share of the clone's blocks through the multi-GPU path and compares its blocks with the
    CPU oracle (which runs all blocks: deposits cross blocks) bit for bit."""
    import tess2_b200
    from tess2_b200 import multi
    from oracle import ref
    name, blocks, layout, gs = small_clone(cfg)
    owner = multi.assign_blocks(len(blocks), world)
    if world > 1:
        multi.set_layout(ctx, layout, owner)
    mine = [b for b in blocks if owner[b["gid"]] == rank]
    port = ref.Checker("port")
    differing, compared = 0, 0
    for alg in algs:
        o = port.dense(blocks, gs, alg=alg, assemble=False)
        params = ctx.make_params(alg, 0, None, None, False, (0.0, 0.0, 1.0), 1.0, 1e-4, gs)
        ctx.upload(mine)
        ctx.run(params)
        res = ctx.download(params, want_grid=False)
        for gid, d in zip(res.gids, res.block_density):
            want = o["block_density"][gid]
            a, b = np.ascontiguousarray(d).view(np.uint32), np.ascontiguousarray(want).view(np.uint32)
            differing += int(((a != b) & ~(np.isnan(d) & np.isnan(want))).sum()) if a.shape == b.shape else a.size
            compared += a.size
    differing = int(multi.sum_over_ranks(differing))
    compared = int(multi.sum_over_ranks(compared))
    return {"workload": name, "algs": list(algs), "grid_values_compared": compared, "differing_values": differing, "bit_identical": differing == 0,
            "against": "oracle/dense_oracle.c (pinned to the unmodified reference in tests/test_oracle.py)"}


def grid_digest(ctx, params, world):
    """sha256 over the blocks' density arrays in gid order (every rank hashes its own blocks; the per-block digests
    are combined on every rank).  Strong-scaling configs produce the same digest at every N."""
    import torch.distributed as dist
    res = ctx.download(params, want_grid=False)
    mine = [(int(g), hashlib.sha256(np.ascontiguousarray(d).tobytes()).hexdigest()) for g, d in zip(res.gids, res.block_density)]
    if world > 1:
        allp = [None] * world
        dist.all_gather_object(allp, mine)
        mine = [x for part in allp for x in part]
    h = hashlib.sha256()
    for g, dg in sorted(mine):
        h.update(f"{g}:{dg};".encode())
    return h.hexdigest(), len(mine)


# ---- our arm ----------------------------------------------------------------------------------------------
def pinned_copy(a):
    import torch
    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype).pin_memory()
    out = t.numpy()
    out[...] = a
    return t, out


def h2d_ceiling(local_rank, barrier, max_over_ranks, world):
    """Plain pinned host-to-device copies, every rank at the same time: the box's ceiling for the e2e leg."""
    import torch
    nbytes = 256 << 20
    src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dst = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{local_rank}")
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    reps = 8
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    return {"per_rank_GBps": nbytes * reps / dt / 1e9, "aggregate_GBps": world * nbytes * reps / dt / 1e9, "bytes_per_copy": nbytes, "ranks_at_once": world}


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
    n, cfg = world, args.config
    t_host0 = time.time()
    w = build_workload(cfg, n, rank, args.scale)
    blocks, layout, owner, dmin, dmax, gsize, ng = w["blocks"], w["layout"], w["owner"], w["dmin"], w["dmax"], w["gsize"], w["ng"]
    if rank == 0:
        log(f"workload ready in {time.time() - t_host0:.1f} s: {w['host']}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = tess2_b200.Context(local_rank)
    if world > 1:
        multi.init_comm(ctx, layout, owner)
    main_alg = tess2_b200.DENSE_CIC if cfg == 5 else tess2_b200.DENSE_TESS

    # ---- parity first ---------------------------------------------------------------------------------
    parity = {}
    if not args.no_parity:
        parity["small_clone"] = check_small_clone(ctx, cfg, world, rank, (0, 1))
        if world > 1:
            multi.set_layout(ctx, layout, owner)
        if rank == 0:
            log(f"parity (small clone): {parity['small_clone']}")

    keep = []
    for b in blocks:                      # pinned host buffers: the e2e leg copies from these
        for k in ("particles", "tets", "vert_to_tet"):
            t, b[k] = pinned_copy(np.ascontiguousarray(b[k]))
            keep.append(t)
    params = ctx.make_params(main_alg, ng, dmin, dmax, False, (0.0, 0.0, 1.0), 1.0, 1e-4, gsize)

    # ---- device-resident timing -------------------------------------------------------------------
    ctx.upload(blocks)
    sampler = ClockSampler(local_rank)
    sampler.start()             # sampled every 20 ms from the warm-up to the end of the e2e leg (both timed regions)
    for _ in range(args.warmup):
        ctx.run(params)
    barrier()
    t0 = time.perf_counter()
    skeys = [nme for nme, _ in tess2_b200.lib.DenseStats._fields_ if nme.startswith("ms_")]
    stage = {k: 0.0 for k in skeys}
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
    dev_ms_min = -multi.max_over_ranks(-stage["ms_total_device"] / args.steps)
    wall_ms = multi.max_over_ranks(1e3 * wall / args.steps)
    G_local = int(st.num_grid_pts)
    G_total = gsize[0] * gsize[1] * gsize[2]
    T_local = int(st.num_tets)
    T_total = int(multi.sum_over_ranks(T_local))
    P_local = sum(len(b["particles"]) for b in blocks)
    cells_local = int(st.num_cells)
    cells_total = int(multi.sum_over_ranks(cells_local))
    spans = int(st.num_spans)
    value = G_total / (dev_ms * 1e-3)
    # mass balance over all ranks (src/dense.cpp:1325-1326 "tot_mass ... should be"): deposits that crossed ranks are
    # in the receiver's grid and in the sender's cell count, so only the sums balance
    dep_total = multi.sum_over_ranks(int(st.num_deposit_cells))
    mass_total = multi.sum_over_ranks(float(st.tot_mass))
    parity["mass_total"] = mass_total
    parity["mass_expected"] = dep_total * 1.0
    parity["mass_rel_err"] = abs(mass_total - dep_total) / max(1.0, dep_total)
    parity["mass_ok"] = parity["mass_rel_err"] <= 1e-6
    if not args.no_parity:
        parity["grid_sha256"], parity["blocks_hashed"] = grid_digest(ctx, params, world)
        parity["grid_sha256_note"] = "sha256 of every block's density in gid order; strong-scaling configs: identical at every N"

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
    checksum = float(multi.sum_over_ranks(float(sum(float(o.astype(np.float64).sum()) for o in out_blocks))))
    h2d_total, d2h_total = int(multi.sum_over_ranks(h2d)), int(multi.sum_over_ranks(d2h))
    ceiling = h2d_ceiling(local_rank, barrier, multi.max_over_ranks, world)
    e2e_floor_ms = 1e3 * max(h2d, 1) / (ceiling["per_rank_GBps"] * 1e9)

    # ---- the other estimators on the same resident inputs: a few steps each, reported beside ----
    other = {}
    ctx.upload(blocks)
    names = {tess2_b200.DENSE_TESS: "DENSE_TESS", tess2_b200.DENSE_CIC: "DENSE_CIC", tess2_b200.DENSE_DTFE: "DENSE_DTFE (not in the reference)"}
    for alg in (tess2_b200.DENSE_TESS, tess2_b200.DENSE_CIC, tess2_b200.DENSE_DTFE):
        if alg == main_alg or (alg == tess2_b200.DENSE_DTFE and cfg != 2):
            continue
        pa = ctx.make_params(alg, ng, dmin, dmax, False, (0.0, 0.0, 1.0), 1.0, 1e-4, gsize)
        for _ in range(2):
            ctx.run(pa)
        barrier()
        ms = 0.0
        reps = 3
        for _ in range(reps):
            so = ctx.run(pa)
            ms += so.ms_total_device
        barrier()
        ms = multi.max_over_ranks(ms / reps)
        other[names[alg]] = {"ms_per_step": ms, "grid_points_per_sec": G_total / (ms * 1e-3), "tot_mass": multi.sum_over_ranks(float(so.tot_mass)),
                             "max_dense": multi.max_over_ranks(float(so.max_dense))}
    # the projected variant of the main estimator (`project` = 1 along z, the reference run scripts' default: TESS_DENSE_TEST,
    # DENSE_TEST): every z of a column lands on one 2-D grid point, so every deposit is shared and the whole step's span records
    # go through the ordered path (sort + k_rows).  Reported in box points/s of the same 3-D index box for comparison.
    try:
        pp = ctx.make_params(main_alg, ng, dmin, dmax, True, (0.0, 0.0, 1.0), 1.0, 1e-4, gsize)
        for _ in range(2):
            ctx.run(pp)
        barrier()
        ms = 0.0
        for _ in range(3):
            so = ctx.run(pp)
            ms += so.ms_total_device
        barrier()
        ms = multi.max_over_ranks(ms / 3)
        other[names[main_alg] + " projected along z"] = {
            "ms_per_step": ms, "box_points_per_sec": G_total / (ms * 1e-3), "grid_points_2d": gsize[0] * gsize[1],
            "tot_mass": multi.sum_over_ranks(float(so.tot_mass)), "ms_sort_and_rows": multi.max_over_ranks(float(so.ms_deposit)),
            "max_dense": multi.max_over_ranks(float(so.max_dense))}
    except Exception as e:      # an auxiliary line: its failure is reported in it and must not take the headline down
        other[names[main_alg] + " projected along z"] = {"error": repr(e)}
    # ---- K2 (per-site Voronoi volume + zero-order density, volume() of src/volume.cpp:13-54) on this rank's
    # largest block: device time of the kernels (circumcenters + star walk + fan sums), 60 T + 24 P algorithmic bytes ----
    k2 = None
    if not args.no_k2 and blocks:
        bb = max(blocks, key=lambda b: len(b["tets"]))
        best = None
        for _ in range(3):
            ctx.cell_volumes(int(bb["num_orig"]), bb["tets"], bb["particles"], bb["vert_to_tet"])
            ms = ctx.cell_volumes_ms()
            best = ms if best is None else min(best, ms)
        k2_bytes = 60 * len(bb["tets"]) + 24 * len(bb["particles"])
        k2 = {"block_gid": int(bb["gid"]), "sites": int(bb["num_orig"]), "tets": int(len(bb["tets"])), "ms": best, "algorithmic_bytes": k2_bytes,
              "algorithmic_GBps": k2_bytes / (best * 1e-3) / 1e9 if best else None, "sites_per_sec": bb["num_orig"] / (best * 1e-3) if best else None}
    ctx.upload(blocks)
    st_main = ctx.run(params)

    # ---- roofline (algorithmic bytes: DESIGN.md section 4) ----------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak = 6650.0; peak_src = "fallback (B200_PROFILING.md)"
    roofline = make_roofline(st, stage, args.steps, T_local, P_local, cells_local, G_local, spans, dev_ms, peak, peak_src, main_alg, cfg, n)
    if k2:
        k2["frac_of_peak"] = k2["algorithmic_GBps"] / peak if k2["algorithmic_GBps"] else None
        roofline["stages"]["K2 k_cell_volumes on one block (60 T + 24 P), not part of dense()"] = k2

    # ---- CPU baseline beside it (rank 0, N = 1 only) -------------------------------------------------------
    cpu = None
    if rank == 0 and n == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        max_cells = int(os.environ.get("TESSB200_CPU_SAMPLE_CELLS", "65536"))
        plain = plain_blocks(blocks)
        t_s, cells, kind, procs, maps = cpu_dense_sample(plain, gsize, (dmin, dmax), max_cells, cores, 1 if cfg == 5 else 0)
        cpu_value = G_total * (cells / cells_total) / t_s
        cpu = {"value": cpu_value, "unit": UNIT, "cores": procs, "kind": kind, "native_so": maps,
               "sample": f"{cells} of {cells_total} cells ({procs} windows of cells over {len(blocks)} blocks, one process per window, all at the same time "
                         f"on {cores} host cores), slowest dense() interval {t_s:.2f} s, throughput scaled by cells"}

    if rank == 0:
        host = dict(w["host"])
        tess_s = host.get("tess_seconds")
        host["tess_plus_dense_seconds"] = (tess_s + e2e_ms * 1e-3) if tess_s else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms, "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(cfg, n, args.scale), "alg": names[main_alg], "mass": 1.0, "eps": 1e-4,
                       "blocks_per_gpu": len(blocks), "particles": w["particles_total"],
                       "l2": "inputs larger than L2 (tets %.0f MB per GPU vs 126 MB L2), no flush" % (32 * T_local / 1e6),
                       "timing": "CUDA events on the library stream inside tessb200_dense_run, max over ranks"},
            "tets_per_sec": T_total / (dev_ms * 1e-3), "cells_per_sec": cells_total / (dev_ms * 1e-3),
            "ms_per_step_wall": wall_ms, "ms_per_step_fastest_rank": dev_ms_min,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_total, "d2h_bytes_per_step": d2h_total, "ms_per_step": e2e_ms,
                    "checksum_sum_density": checksum, "pinned_h2d_ceiling": ceiling, "ms_of_h2d_alone_at_ceiling": e2e_floor_ms},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "parity": parity,
            "other_algs": other,
            "host_tess": host,
            "stats": {"shared_deposits": int(st.num_shared_deposits), "cells": cells_local, "tets": T_local, "particles_with_ghosts": P_local,
                      "grid_points": G_local, "spans": spans, "faces": int(st.num_faces),
                      "deposit_cells": int(st.num_deposit_cells), "cic_fallback_cells": int(st.num_cic_fallback), "slow_cells": int(st.num_slow_cells),
                      "outside_cells": int(st.num_outside), "incomplete_cells": int(st.num_incomplete),
                      "tot_mass": float(st.tot_mass), "scope": "rank 0" if world > 1 else "all"},
        }
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    del st_main, keep, out_t


def make_roofline(st, stage, steps, T, P, P0, G, S, dev_ms, peak, peak_src, alg, cfg=3, n=1):
    """Per-stage device ms (CUDA events inside the library, same stream) against SURVEY 8(d)'s algorithmic bytes."""
    F = int(st.num_faces)
    ms = {k: v / steps for k, v in stage.items()}
    n_shared = int(st.num_shared_deposits)
    if alg == 1:
        # K4 (CIC) = 12 P + 4 G: particles in, grid out (weights, base cells, the particle sort and the few boundary
        # records are intermediates)
        stages = {"K4 cloud-in-cell: k_cic_prepare + particle sort + k_cic_gather + boundary records (12 P0 + 4 G)":
                  (ms["ms_circumcenters"] + ms["ms_cells"] + ms["ms_scan"] + ms["ms_slow_path"] + ms["ms_sort"] + ms["ms_deposit"], 12 * P0 + 4 * G)}
    else:
        stages = {
            "K1 k_circumcenters + cell order (28 T + 12 P)": (ms["ms_circumcenters"], 28 * T + 12 * P),
            "K3a cell set-up + scan (44 T + 16 P + S)": (ms["ms_cells"] + ms["ms_scan"] + ms["ms_slow_path"], 44 * T + 16 * P + 16 * S),
        }
        if n_shared >= 0:
            stages["K3b deposit: count + place + ordered shared points (S + 4 G)"] = (ms["ms_sort"] + ms["ms_deposit"], 16 * S + 4 * G)
        else:
            stages["K3b deposit: sort + k_rows (S + 4 G)"] = (ms["ms_sort"] + ms["ms_deposit"], 16 * S + 4 * G)
    stages["span exchange (NCCL)"] = (ms["ms_exchange"], 0)
    detail = {k: {"ms": v[0], "algorithmic_bytes": v[1], "algorithmic_GBps": (v[1] / (v[0] * 1e-3) / 1e9 if v[0] > 0 else None),
                  "frac_of_peak": (v[1] / (v[0] * 1e-3) / 1e9 / peak if v[0] > 0 else None)} for k, v in stages.items()}
    dom = max((k for k in stages if k.startswith("K")), key=lambda k: stages[k][0])
    ach = detail[dom]["algorithmic_GBps"] or 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        # DRAM bytes of the stage's kernels in one step, from the ncu launch list of profiles/run_step.py on one GPU
        # (profiles/launches_r02.py); strong-scaling configs at N GPUs move 1/N of it per GPU
        traffic = json.load(open(tpath)).get(f"c{cfg}_{dom.split(' ')[0]}")
        if traffic is not None and n > 1:
            traffic = traffic / n
    whole = 12 * P0 + 4 * G if alg == 1 else 32 * T + 16 * P + 4 * G          # K4 reads the block's own particles only
    sub = {k: ms[k] for k in ms if ms[k] > 0}
    return {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
            "peak_source": peak_src, "algorithmic_bytes": stages[dom][1], "kernel_ms": stages[dom][0],
            "note": "SURVEY 8(d) bytes (inputs once + outputs once, intermediates excluded) over the stage's device time; the cell kernels are "
                    "latency / issue bound, not HBM bound: DESIGN.md 3 and profiles/",
            "stages": detail, "device_ms": sub,
            "whole_stage": {"algorithmic_bytes": whole, "achieved": whole / (dev_ms * 1e-3) / 1e9, "frac": whole / (dev_ms * 1e-3) / 1e9 / peak}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=int(os.environ.get("TESSB200_BENCH_CONFIG", "3")), choices=[2, 3, 4, 5])
    ap.add_argument("--scale", type=int, default=int(os.environ.get("TESSB200_BENCH_SCALE", "1")), help="development: shrink configs 3-5 by this factor per axis")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-k2", action="store_true")
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
                   "--steps", str(args.steps), "--warmup", str(args.warmup), "--config", str(args.config), "--scale", str(args.scale)] + \
                  (["--no-cpu-baseline"] if args.no_cpu_baseline else []) + (["--no-parity"] if args.no_parity else [])
            raise SystemExit(subprocess.call(cmd))
        raise SystemExit(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
