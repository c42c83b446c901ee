"""Host-side logic of the multi-GPU path (no GPU): block -> rank assignment, the layout handed to
tessb200_dense_set_layout, the broadcast of the NCCL unique id and the assembly of the global
grid from per-rank block sub-arrays -- exercised with world_size 2 over gloo."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_assign_blocks_is_contiguous_and_balanced():
    from tess2_b200 import multi
    for nblocks, nranks in [(8, 1), (8, 2), (16, 2), (64, 8), (10, 4)]:
        owner = multi.assign_blocks(nblocks, nranks)
        assert list(owner) == sorted(owner)                       # contiguous in gid order
        counts = np.bincount(owner, minlength=nranks)
        assert counts.max() - counts.min() <= 1 and counts.sum() == nblocks


def test_layout_arrays_sorted_by_gid():
    from tess2_b200 import multi
    layout = [(2, [2, 0, 0], [3, 1, 1]), (0, [0, 0, 0], [1, 1, 1]), (1, [1, 0, 0], [2, 1, 1])]
    gids, b6, own = multi.layout_arrays(layout, [1, 0, 0])
    assert list(gids) == [0, 1, 2]
    assert list(b6[0]) == [0, 0, 0, 1, 1, 1] and list(b6[2]) == [2, 0, 0, 3, 1, 1]
    assert list(own) == [0, 0, 1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from tess2_b200 import multi
    from tess2_b200.dense import DenseResult
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. the unique id travels from rank 0 to everyone
        uid = bytes(range(128)) if rank == 0 else bytes(128)
        got = multi.broadcast_bytes(uid, src=0)
        assert got == bytes(range(128))
        # 2. timings are reported as the slowest rank, counts as the sum
        assert multi.max_over_ranks(1.0 + rank) == float(world)
        assert multi.sum_over_ranks(3.0) == 3.0 * world
        # 3. each rank owns a z-slab of a 4 x 4 x (2*world) grid; rank 0 assembles the global grid
        gs = (4, 4, 2 * world)
        res = DenseResult()
        res.block_min_idx = [[0, 0, 2 * rank]]
        res.block_num_idx = [[4, 4, 2]]
        res.block_density = [np.full((2, 4, 4), float(rank + 1), np.float32)]
        grid = multi.gather_global_grid(res, gs, dst=0)
        if rank == 0:
            assert grid.shape == (2 * world, 4, 4)
            for r in range(world):
                assert np.all(grid[2 * r:2 * r + 2] == r + 1)
        else:
            assert grid is None
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_host_plumbing_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(out) == [(0, "ok"), (1, "ok")], out


def test_weak_scaling_workload_layout():
    # bench.py's N-rank workload: rank r owns gids [8r, 8r+8), a z-slab of the (2,2,2N) block grid
    from tess2_b200 import multi
    from tess2_b200.harness import workloads
    blocks, layout, dmin, dmax = workloads.uniform_regular(4, (2, 2, 4), gids=[8, 9, 10, 11, 12, 13, 14, 15], cache=False, workers=1)
    owner = multi.assign_blocks(16, 2)
    assert [b["gid"] for b in blocks] == list(range(8, 16)) and all(owner[b["gid"]] == 1 for b in blocks)
    zmin = min(float(b["bounds_min"][2]) for b in blocks)
    assert zmin == pytest.approx(float(dmax[2]) / 2)
    assert all(b["num_orig"] == 4 ** 3 for b in blocks)
    assert all("vert_to_tet" in b for b in blocks)
