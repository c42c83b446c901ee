"""Host side of the tessellation stage: ctypes binding of libtess_b200_host.so (include/tess_b200_host.h).

`tess()` mirrors the role of tess2's tess() (src/tess.cpp:52-116) for one process that holds all
particles: blocks with ghosts, Delaunay tets, vert_to_tet -- the dict layout the rest of the package
(and the oracle) takes for a block.  CPU only; the Delaunay engine is the repo's own
(tess2_b200/host/delaunay3.hpp) because neither Qhull nor CGAL is installed in this image."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


class HostBlock(C.Structure):
    _fields_ = [
        ("gid", C.c_int), ("bounds_min", C.c_float * 3), ("bounds_max", C.c_float * 3),
        ("num_orig_particles", C.c_int), ("num_particles", C.c_int), ("num_tets", C.c_int),
        ("particles", C.POINTER(C.c_float)), ("tets", C.POINTER(C.c_int)), ("vert_to_tet", C.POINTER(C.c_int)),
        ("global_ids", C.POINTER(C.c_int)), ("ghost_margin", C.c_float), ("rounds", C.c_int), ("seconds", C.c_double),
        ("settled", C.c_int), ("reserved", C.c_int),
    ]


class HostDBlock(C.Structure):
    """tessb200_host_dblock: one block of the hand-off file (the fields of load_block_light, src/tess.cpp:223-259)."""
    _fields_ = [
        ("gid", C.c_int), ("bounds_min", C.c_float * 3), ("bounds_max", C.c_float * 3), ("box_min", C.c_float * 3), ("box_max", C.c_float * 3),
        ("data_min", C.c_float * 3), ("data_max", C.c_float * 3), ("num_orig_particles", C.c_int), ("num_particles", C.c_int),
        ("particles", C.POINTER(C.c_float)), ("rem_gids", C.POINTER(C.c_int)), ("rem_lids", C.POINTER(C.c_int)),
        ("num_grid_pts", C.c_int), ("density", C.POINTER(C.c_float)), ("complete", C.c_int), ("num_tets", C.c_int),
        ("tets", C.POINTER(C.c_int)), ("vert_to_tet", C.POINTER(C.c_int)),
    ]


BOUNDS_DYNAMIC, BOUNDS_STATIC4 = 0, 1

EXPORTS = ["tessb200_host_delaunay", "tessb200_host_tess", "tessb200_host_tess_periodic", "tessb200_host_free_block", "tessb200_host_free", "tessb200_host_last_error",
           "tessb200_host_regular_blocks", "tessb200_host_kdtree_blocks",
           "tessb200_host_write_blocks", "tessb200_host_read_blocks", "tessb200_host_free_dblocks"]


def load():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libtess_b200_host.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `python __graft_entry__.py` (or make -C tess2_b200/host)")
        lib = C.CDLL(path)
        lib.tessb200_host_delaunay.argtypes = [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_int))]
        lib.tessb200_host_tess.argtypes = [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                           C.c_int, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int), C.c_float, C.c_int, C.c_float,
                                           C.c_int, C.POINTER(HostBlock)]
        lib.tessb200_host_tess_periodic.argtypes = [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                                    C.c_int, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int), C.c_float, C.c_int, C.c_float,
                                                    C.c_int, C.c_int, C.POINTER(HostBlock)]
        lib.tessb200_host_free_block.argtypes = [C.POINTER(HostBlock)]
        lib.tessb200_host_regular_blocks.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float)]
        lib.tessb200_host_kdtree_blocks.argtypes = [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int,
                                                    C.POINTER(C.c_float), C.POINTER(C.c_int)]
        lib.tessb200_host_free.argtypes = [C.c_void_p]
        lib.tessb200_host_write_blocks.argtypes = [C.c_char_p, C.c_int, C.POINTER(HostDBlock), C.c_int, C.c_void_p, C.c_size_t]
        lib.tessb200_host_read_blocks.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.POINTER(HostDBlock)), C.POINTER(C.c_int)]
        lib.tessb200_host_free_dblocks.argtypes = [C.c_int, C.POINTER(HostDBlock)]
        lib.tessb200_host_last_error.restype = C.c_char_p
        _lib = lib
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def delaunay(points):
    """Delaunay tets of float32 points: int32 [T, 8] in tet_t layout (verts[4], tets[4])."""
    lib = load()
    p = np.ascontiguousarray(points, dtype=np.float32)
    nt = C.c_int()
    tp = C.POINTER(C.c_int)()
    rc = lib.tessb200_host_delaunay(len(p), _fp(p), C.byref(nt), C.byref(tp))
    if rc:
        raise RuntimeError(lib.tessb200_host_last_error().decode())
    out = np.ctypeslib.as_array(tp, (nt.value, 8)).copy() if nt.value else np.zeros((0, 8), np.int32)
    lib.tessb200_host_free(tp)
    return out


def tess(points, owner, bounds, domain_min, domain_max, margin0=0.0, threads=0, gids=None, max_rounds=3, max_growth=2.5, wrap=False):
    """Blocks of one process.  bounds: list of (min[3], max[3]) indexed by gid; owner: gid per
    particle or None (containment); gids: the blocks to tessellate (default all).  Returns the list
    of block dicts used throughout the package (gid, particles, num_orig, tets, vert_to_tet,
    bounds_min, bounds_max, margin, rounds, global_ids).  wrap: periodic domain (the reference drivers' `wrap` argument): ghosts
    include the images of particles shifted by whole domain extents; an image keeps its particle's global id."""
    lib = load()
    p = np.ascontiguousarray(points, dtype=np.float32)
    nb = len(bounds)
    bb = np.ascontiguousarray([list(np.asarray(mn, np.float32)) + list(np.asarray(mx, np.float32)) for mn, mx in bounds], dtype=np.float32)
    dmin = np.ascontiguousarray(domain_min, dtype=np.float32)
    dmax = np.ascontiguousarray(domain_max, dtype=np.float32)
    own = None if owner is None else np.ascontiguousarray(owner, dtype=np.int32)
    g = None if gids is None else np.ascontiguousarray(gids, dtype=np.int32)
    arr = (HostBlock * (nb if g is None else len(g)))()
    rc = lib.tessb200_host_tess_periodic(len(p), _fp(p), own.ctypes.data_as(C.POINTER(C.c_int)) if own is not None else None, _fp(dmin), _fp(dmax),
                                         nb, _fp(bb), 0 if g is None else len(g), g.ctypes.data_as(C.POINTER(C.c_int)) if g is not None else None,
                                         float(margin0), int(max_rounds), float(max_growth), int(threads), 1 if wrap else 0, arr)
    if rc:
        raise RuntimeError(lib.tessb200_host_last_error().decode())
    out = []
    for b in arr:
        n, t = b.num_particles, b.num_tets
        out.append(dict(
            gid=b.gid, num_orig=b.num_orig_particles,
            particles=np.ctypeslib.as_array(b.particles, (n, 3)).copy() if n else np.zeros((0, 3), np.float32),
            tets=np.ctypeslib.as_array(b.tets, (t, 8)).copy() if t else np.zeros((0, 8), np.int32),
            vert_to_tet=np.ctypeslib.as_array(b.vert_to_tet, (n,)).copy() if n else np.zeros(0, np.int32),
            global_ids=np.ctypeslib.as_array(b.global_ids, (n,)).copy() if n else np.zeros(0, np.int32),
            bounds_min=np.array(b.bounds_min, np.float32), bounds_max=np.array(b.bounds_max, np.float32),
            margin=b.ghost_margin, rounds=b.rounds, seconds=b.seconds, settled=bool(b.settled)))
        lib.tessb200_host_free_block(C.byref(b))
    return out


def _bounds_list(bb):
    return [(bb[g, :3].copy(), bb[g, 3:].copy()) for g in range(len(bb))]


def regular_blocks(domain_min, domain_max, nblocks):
    """[(min[3], max[3])] of a regular decomposition, gid x-fastest."""
    lib = load()
    dmin = np.ascontiguousarray(domain_min, dtype=np.float32)
    dmax = np.ascontiguousarray(domain_max, dtype=np.float32)
    bb = np.zeros((nblocks, 6), np.float32)
    if lib.tessb200_host_regular_blocks(_fp(dmin), _fp(dmax), nblocks, _fp(bb)):
        raise RuntimeError(lib.tessb200_host_last_error().decode())
    return _bounds_list(bb)


def kdtree_blocks(points, domain_min, domain_max, nblocks):
    """(bounds list, owner gid per particle) of a median kd-tree decomposition."""
    lib = load()
    p = np.ascontiguousarray(points, dtype=np.float32)
    dmin = np.ascontiguousarray(domain_min, dtype=np.float32)
    dmax = np.ascontiguousarray(domain_max, dtype=np.float32)
    bb = np.zeros((nblocks, 6), np.float32)
    owner = np.zeros(len(p), np.int32)
    if lib.tessb200_host_kdtree_blocks(len(p), _fp(p), _fp(dmin), _fp(dmax), nblocks, _fp(bb), owner.ctypes.data_as(C.POINTER(C.c_int))):
        raise RuntimeError(lib.tessb200_host_last_error().decode())
    return _bounds_list(bb), owner


def write_blocks(path, blocks, data_min, data_max, layout=BOUNDS_DYNAMIC, extra=b""):
    """tess_save (src/tess.cpp:126-137): the blocks (dicts as tess() returns them) into a DIY block file.
    Optional keys: box_min/box_max (default: the block bounds), rem_gids/rem_lids (default -1), density, complete."""
    lib = load()
    arr = (HostDBlock * max(len(blocks), 1))()
    keep = []

    def arr_of(a, dtype):
        a = np.ascontiguousarray(a, dtype=dtype)
        keep.append(a)
        return a

    for d, b in zip(arr, blocks):
        p = arr_of(b["particles"], np.float32).reshape(-1, 3)
        t = arr_of(b["tets"], np.int32).reshape(-1, 8)
        v = arr_of(b["vert_to_tet"], np.int32)
        d.gid, d.num_orig_particles, d.num_particles, d.num_tets = int(b["gid"]), int(b["num_orig"]), len(p), len(t)
        d.bounds_min[:] = [float(x) for x in b["bounds_min"]]
        d.bounds_max[:] = [float(x) for x in b["bounds_max"]]
        d.box_min[:] = [float(x) for x in b.get("box_min", b["bounds_min"])]
        d.box_max[:] = [float(x) for x in b.get("box_max", b["bounds_max"])]
        d.data_min[:] = [float(x) for x in data_min]
        d.data_max[:] = [float(x) for x in data_max]
        d.particles = _fp(p)
        d.tets = t.ctypes.data_as(C.POINTER(C.c_int))
        d.vert_to_tet = v.ctypes.data_as(C.POINTER(C.c_int))
        for key in ("rem_gids", "rem_lids"):
            if b.get(key) is not None:
                setattr(d, key, arr_of(b[key], np.int32).ctypes.data_as(C.POINTER(C.c_int)))
        if b.get("density") is not None:
            g = arr_of(b["density"], np.float32).reshape(-1)
            d.num_grid_pts, d.density = len(g), _fp(g)
        d.complete = int(b.get("complete", 0))
    rc = lib.tessb200_host_write_blocks(os.fsencode(path), len(blocks), arr, int(layout), extra if extra else None, len(extra))
    if rc:
        raise RuntimeError(lib.tessb200_host_last_error().decode())


def read_blocks(path):
    """tess_load (src/tess.cpp:139-152): (list of block dicts in gid order, data_min, data_max, bounds layout found)."""
    lib = load()
    n = C.c_int()
    layout = C.c_int()
    arr = C.POINTER(HostDBlock)()
    rc = lib.tessb200_host_read_blocks(os.fsencode(path), C.byref(n), C.byref(arr), C.byref(layout))
    if rc:
        raise RuntimeError(lib.tessb200_host_last_error().decode())

    def copy(ptr, shape, dtype):
        return np.ctypeslib.as_array(ptr, shape).copy() if shape[0] else np.zeros(shape, dtype)

    out = []
    data_min = data_max = None
    for i in range(n.value):
        b = arr[i]
        npart, ng = b.num_particles, b.num_particles - b.num_orig_particles
        out.append(dict(
            gid=b.gid, num_orig=b.num_orig_particles, particles=copy(b.particles, (npart, 3), np.float32),
            tets=copy(b.tets, (b.num_tets, 8), np.int32), vert_to_tet=copy(b.vert_to_tet, (npart,), np.int32),
            rem_gids=copy(b.rem_gids, (ng,), np.int32), rem_lids=copy(b.rem_lids, (ng,), np.int32),
            density=copy(b.density, (b.num_grid_pts,), np.float32), complete=b.complete,
            bounds_min=np.array(b.bounds_min, np.float32), bounds_max=np.array(b.bounds_max, np.float32),
            box_min=np.array(b.box_min, np.float32), box_max=np.array(b.box_max, np.float32)))
        data_min, data_max = np.array(b.data_min, np.float32), np.array(b.data_max, np.float32)
    lib.tessb200_host_free_dblocks(n.value, arr)
    return out, data_min, data_max, layout.value
