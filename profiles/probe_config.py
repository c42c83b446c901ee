"""BASELINE.json configs 3, 4 and 5 on N GPUs of one box (one process per GPU, blocks split over the ranks in gid order):
clustered Gaussian-clump particles, kd-tree blocks, the repo's own host tess() on each rank's blocks,
multi-GPU dense with the NCCL span exchange.  Strong scaling: the same input whatever N is.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 \\
        profiles/probe_config.py --side 256 --blocks 8 --gsize 512 [--alg both] [--steps 3] [--dry-run]

  config 3: --side 256 --blocks 8  --gsize 512            (N = 1, 2, 4, 8)
  config 4: --side 512 --blocks 64 --gsize 1024           (N = 8)
  config 5: --side 256 --blocks 8  --gsize 512 --alg both (N = 8)

Rank 0 generates the particles and the decomposition once and hands them to the other ranks through /dev/shm (every
rank needs all particles to find its ghosts).  --dry-run stops after the host side (no GPU, gloo): sizes and times.
Rank 0 prints one JSON line: per-stage device ms (max over ranks), grid points/s, tets/s, mass check."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=64)
    ap.add_argument("--blocks", type=int, default=8)
    ap.add_argument("--gsize", type=int, default=0, help="grid points per axis (default 2 * side)")
    ap.add_argument("--alg", default="0", choices=["0", "1", "both"])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--threads", type=int, default=0, help="host threads per rank for tess() (default cores / ranks)")
    ap.add_argument("--dry-run", action="store_true")
    ap.add_argument("--limit-blocks", type=int, default=0, help="dry runs: tessellate only the first k of this rank's blocks")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from tess2_b200 import host_tess, multi
    from tess2_b200.harness import particles

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        if not args.dry_run:
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="gloo" if args.dry_run else "nccl")
    n, nb = args.side, args.blocks
    gs = (args.gsize or 2 * n,) * 3
    dom = (np.zeros(3, np.float32), np.full(3, n - 1, np.float32))
    shm = f"/dev/shm/tessb200_probe_{os.environ.get('MASTER_PORT', '0')}_{n}_{nb}"
    t0 = time.time()
    if rank == 0:
        p = particles.clustered_particles(n ** 3, *dom, seed=2024 + 3)
        bounds, owner_of_particle = host_tess.kdtree_blocks(p, *dom, nb)
        np.save(shm + "_p.npy", p)
        np.save(shm + "_o.npy", owner_of_particle)
        np.save(shm + "_b.npy", np.array([np.concatenate([mn, mx]) for mn, mx in bounds], np.float32))
    if world > 1:
        dist.barrier()
    p = np.load(shm + "_p.npy", mmap_mode="r")
    owner_of_particle = np.load(shm + "_o.npy", mmap_mode="r")
    b6 = np.load(shm + "_b.npy")
    bounds = [(b6[g, :3].copy(), b6[g, 3:].copy()) for g in range(nb)]
    t_gen = time.time() - t0
    owner = multi.assign_blocks(nb, world)
    my_gids = [g for g in range(nb) if owner[g] == rank]
    if args.dry_run and args.limit_blocks > 0:
        my_gids = my_gids[:args.limit_blocks]
    threads = args.threads or max(1, (os.cpu_count() or 1) // world)
    t0 = time.time()
    blocks = host_tess.tess(np.asarray(p), np.asarray(owner_of_particle), bounds, *dom, threads=threads, gids=my_gids)
    t_tess = time.time() - t0
    if world > 1:
        dist.barrier()
    if rank == 0:
        for s in ("_p.npy", "_o.npy", "_b.npy"):
            os.remove(shm + s)
    host = dict(particles=int(len(p)), blocks=nb, ranks=world, my_blocks=len(my_gids), threads_per_rank=threads,
                my_particles_with_ghosts=int(sum(len(b["particles"]) for b in blocks)), my_tets=int(sum(len(b["tets"]) for b in blocks)),
                generate_s=round(t_gen, 2), tess_s=round(t_tess, 2), max_rounds=int(max(b["rounds"] for b in blocks)))
    if args.dry_run:
        tets = multi.sum_over_ranks(host["my_tets"])
        if rank == 0:
            print(json.dumps(dict(host, dry_run=True, tets_total=int(tets))))
        return
    import tess2_b200
    layout = [(g, bounds[g][0], bounds[g][1]) for g in range(nb)]
    ctx = tess2_b200.Context(local_rank)
    if world > 1:
        multi.init_comm(ctx, layout, owner)
    ng = 3 if world > 1 else 0          # the data bounds of a rank's share are not the global ones: give the domain
    ctx.upload(blocks)
    out = dict(host, gsize=gs[0], config=f"{n}^3 clustered, kd-tree {nb} blocks, {gs[0]}^3 grid, {world} GPU(s)")
    keys = ("ms_circumcenters", "ms_bfs", "ms_nbrs", "ms_faces", "ms_scan", "ms_exchange", "ms_sort", "ms_deposit", "ms_slow_path", "ms_total_device")
    for alg in ([0, 1] if args.alg == "both" else [int(args.alg)]):
        params = ctx.make_params(alg, ng, dom[0], dom[1], False, (0.0, 0.0, 1.0), 1.0, 1e-4, gs)
        for _ in range(args.warmup):
            ctx.run(params)
        acc = {k: 0.0 for k in keys}
        for _ in range(args.steps):
            if world > 1:
                dist.barrier()
            st = ctx.run(params)
            for k in keys:
                acc[k] += getattr(st, k) / args.steps
        res = {k: round(multi.max_over_ranks(v), 3) for k, v in acc.items()}
        ms = res["ms_total_device"]
        mass = multi.sum_over_ranks(st.tot_mass)
        cells = multi.sum_over_ranks(st.num_deposit_cells)
        res.update(grid_points_per_sec=gs[0] ** 3 / (ms * 1e-3), tets_per_sec=multi.sum_over_ranks(st.num_tets) / (ms * 1e-3),
                   deposit_cells=int(cells), cic_fallback_cells=int(multi.sum_over_ranks(st.num_cic_fallback)),
                   slow_cells=int(multi.sum_over_ranks(st.num_slow_cells)), total_mass=mass,
                   mass_rel_err=abs(mass - (cells if alg == 0 else len(p))) / max(1.0, float(cells if alg == 0 else len(p))))
        out["DENSE_TESS" if alg == 0 else "DENSE_CIC"] = res
    if rank == 0:
        print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
