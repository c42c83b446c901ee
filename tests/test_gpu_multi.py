"""Two ranks / two GPUs: the NCCL span exchange (the replacement of master.exchange() + recvd_pts,
src/dense.cpp:98-102,166-202) against the oracle run on all blocks in one process.
Skipped on boxes with fewer than two GPUs."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, assert_same_bits

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, h, alg, project, q, mode="dense"):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    try:
        import torch
        import torch.distributed as dist
        import tess2_b200
        from tess2_b200 import multi
        from tess2_b200.harness import workloads
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        nblocks = 8 * world
        owner = multi.assign_blocks(nblocks, world)
        gids = [g for g in range(nblocks) if owner[g] == rank]
        blocks, layout, dmin, dmax = workloads.uniform_regular(h, (2, 2, 2 * world), gids=gids, cache=False, workers=1)
        ctx = tess2_b200.Context(rank)
        multi.init_comm(ctx, layout, owner)
        gs = (4 * h, 4 * h, 4 * h * world)
        if mode == "peer_fails":
            # rank 1 fails before the span exchange (nothing uploaded): rank 0 must come back with TESSB200_EPEER, not wait
            params = ctx.make_params(alg, 3, dmin, dmax, project, (0.0, 0.0, 1.0), 1.0, 1e-4, gs)
            if rank == 0:
                ctx.upload(blocks)
            codes = []
            for _ in range(2):          # the second round starts from a clean exchange state again
                try:
                    ctx.run(params)
                    codes.append(0)
                except tess2_b200.TessB200Error as e:
                    codes.append(e.code)
            gathered = [None] * world if rank == 0 else None
            dist.gather_object(codes, gathered, dst=0)
            if rank == 0:
                q.put(("ok", gathered, 0))
            ctx.close()
            dist.destroy_process_group()
            return
        # three runs: the first agrees on the segment capacities (exact counts, one host read-back), the others exchange
        # fixed-size sentinel-padded segments without reading anything back before the deposit
        for _ in range(3):
            res = ctx.dense(alg, 3, dmin, dmax, project, (0.0, 0.0, 1.0), 1.0, 1e-4, gs, blocks, want_grid=False)
        out = [(g, mn, num, np.ascontiguousarray(d)) for g, mn, num, d in zip(res.gids, res.block_min_idx, res.block_num_idx, res.block_density)]
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(out, gathered, dst=0)
        if rank == 0:
            q.put(("ok", [x for part in gathered for x in part], int(res.stats.num_spans)))
        ctx.close()
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        import traceback
        q.put(("err", f"rank {rank}: {e!r}\n{traceback.format_exc()}", 0))


@pytest.mark.parametrize("alg,project", [(0, False), (1, False), (0, True)])
def test_two_gpu_dense_equals_oracle(port, alg, project):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from tess2_b200.harness import workloads
    h, world = 8, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    tcp = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, tcp, h, alg, project, q)) for r in range(world)]
    for p in procs:
        p.start()
    status, payload, nspans = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
    assert status == "ok", payload
    # the oracle on all 16 blocks in one process, same grid bounds
    blocks, layout, dmin, dmax = workloads.uniform_regular(h, (2, 2, 2 * world), cache=False, workers=1)
    gs = (4 * h, 4 * h, 4 * h * world)
    o = port.dense(blocks, gs, alg=alg, project=project, given_bounds=(dmin, dmax))
    by_gid = {g: (mn, num, d) for g, mn, num, d in payload}
    assert sorted(by_gid) == list(range(16))
    for i, b in enumerate(blocks):
        mn, num, d = by_gid[b["gid"]]
        assert mn == o["block_min_idx"][i] and num == o["block_num_idx"][i]
        assert_same_bits(d, o["block_density"][i], f"alg{alg} proj{project} block gid {b['gid']}")


def test_a_failing_rank_does_not_hang_its_peers():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    tcp = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, tcp, 8, 0, False, q, "peer_fails")) for r in range(world)]
    for p in procs:
        p.start()
    status, codes, _ = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
    assert status == "ok", codes
    # rank 1: TESSB200_ESTATE (-7, run before upload); rank 0: TESSB200_EPEER (-9), both rounds
    assert codes[1] == [-7, -7] and codes[0] == [-9, -9], codes
