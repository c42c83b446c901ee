"""One process per GPU: the host-side plumbing of a multi-GPU dense run.

The reference runs one MPI rank per group of DIY blocks and moves out-of-block grid points with
master.exchange() (src/dense.cpp:98).  Here torch.distributed carries the small control traffic
(NCCL unique id, timings) and the library itself exchanges the boundary span records with NCCL
over NVLink (tessb200_comm_init / exchange inside tessb200_dense_run).
"""
import ctypes as C
import os

import numpy as np


def assign_blocks(nblocks, nranks):
    """Contiguous block -> rank assignment in gid order (diy::ContiguousAssigner's rule): rank r owns
    gids [r*nblocks/nranks, (r+1)*nblocks/nranks).  The library requires contiguity."""
    owner = np.empty(nblocks, dtype=np.int32)
    for r in range(nranks):
        lo = (r * nblocks) // nranks
        hi = ((r + 1) * nblocks) // nranks
        owner[lo:hi] = r
    return owner


def layout_arrays(layout, owner):
    """(gids, bounds6, owner) arrays for tessb200_dense_set_layout from [(gid, min, max), ...]."""
    gids = np.array([g for g, _, _ in layout], dtype=np.int32)
    order = np.argsort(gids)
    gids = gids[order]
    b6 = np.array([np.concatenate([np.asarray(layout[i][1], np.float32), np.asarray(layout[i][2], np.float32)]) for i in order],
                  dtype=np.float32)
    return gids, np.ascontiguousarray(b6), np.ascontiguousarray(np.asarray(owner, np.int32)[order])


def broadcast_bytes(buf, src=0):
    """Broadcast a bytes object of known length from rank `src` with torch.distributed
    (works on gloo/CPU and nccl/GPU)."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().tolist())


def set_layout(ctx, layout, owner):
    """Install the decomposition (bounds and owner rank of every block) on this rank's Context."""
    from . import lib as _l
    lib = _l.load()
    gids, b6, own = layout_arrays(layout, owner)
    _l.check(lib.tessb200_dense_set_layout(ctx.handle, len(gids), gids.ctypes.data_as(_l.i32p), b6.ctypes.data_as(_l.f32p),
                                           own.ctypes.data_as(_l.i32p)))


def _preload_nccl():
    """The library binds NCCL with dlopen("libnccl.so.2"); a process that imports torch afterwards needs torch's own
    (newer) copy under that name, so map that one first when it exists."""
    import ctypes
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia")
        for base in (spec.submodule_search_locations if spec else []):
            p = os.path.join(base, "nccl", "lib", "libnccl.so.2")
            if os.path.exists(p):
                ctypes.CDLL(p, mode=ctypes.RTLD_GLOBAL)
                return p
    except Exception:
        pass
    return None


def init_comm(ctx, layout, owner):
    """Join this rank's Context to the job-wide NCCL communicator and install the block layout."""
    _preload_nccl()
    import torch.distributed as dist
    from . import lib as _l
    rank, nranks = dist.get_rank(), dist.get_world_size()
    lib = _l.load()
    uid = (C.c_ubyte * 128)()
    if rank == 0:
        _l.check(lib.tessb200_comm_unique_id(uid))
    raw = broadcast_bytes(bytes(uid), src=0)
    uid = (C.c_ubyte * 128).from_buffer_copy(raw)
    _l.check(lib.tessb200_comm_init(ctx.handle, nranks, rank, uid))
    set_layout(ctx, layout, owner)
    return rank, nranks


def max_over_ranks(value):
    """Max of a python float over all ranks (timings are reported as the slowest rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_global_grid(result, gsize, dst=0):
    """Assemble the global C-order grid on rank `dst` from every rank's block sub-arrays
    (what WriteGrid's MPI-IO subarray views do, src/dense.cpp:831-850)."""
    import torch.distributed as dist
    rank = dist.get_rank()
    payload = [(mn, num, np.ascontiguousarray(d)) for mn, num, d in zip(result.block_min_idx, result.block_num_idx, result.block_density)]
    gathered = [None] * dist.get_world_size() if rank == dst else None
    dist.gather_object(payload, gathered, dst=dst)
    if rank != dst:
        return None
    grid = np.zeros((gsize[2], gsize[1], gsize[0]), dtype=np.float32)
    for part in gathered:
        for mn, num, d in part:
            grid[mn[2]:mn[2] + num[2], mn[1]:mn[1] + num[1], mn[0]:mn[0] + num[0]] = d.reshape(num[2], num[1], num[0])
    return grid
