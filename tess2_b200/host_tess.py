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
    ]


EXPORTS = ["tessb200_host_delaunay", "tessb200_host_tess", "tessb200_host_free_block", "tessb200_host_free", "tessb200_host_last_error",
           "tessb200_host_regular_blocks", "tessb200_host_kdtree_blocks"]


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
        lib.tessb200_host_free_block.argtypes = [C.POINTER(HostBlock)]
        lib.tessb200_host_regular_blocks.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float)]
        lib.tessb200_host_kdtree_blocks.argtypes = [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int,
                                                    C.POINTER(C.c_float), C.POINTER(C.c_int)]
        lib.tessb200_host_free.argtypes = [C.c_void_p]
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


def tess(points, owner, bounds, domain_min, domain_max, margin0=0.0, threads=0, gids=None, max_rounds=3, max_growth=2.5):
    """Blocks of one process.  bounds: list of (min[3], max[3]) indexed by gid; owner: gid per
    particle or None (containment); gids: the blocks to tessellate (default all).  Returns the list
    of block dicts used throughout the package (gid, particles, num_orig, tets, vert_to_tet,
    bounds_min, bounds_max, margin, rounds, global_ids)."""
    lib = load()
    p = np.ascontiguousarray(points, dtype=np.float32)
    nb = len(bounds)
    bb = np.ascontiguousarray([list(np.asarray(mn, np.float32)) + list(np.asarray(mx, np.float32)) for mn, mx in bounds], dtype=np.float32)
    dmin = np.ascontiguousarray(domain_min, dtype=np.float32)
    dmax = np.ascontiguousarray(domain_max, dtype=np.float32)
    own = None if owner is None else np.ascontiguousarray(owner, dtype=np.int32)
    g = None if gids is None else np.ascontiguousarray(gids, dtype=np.int32)
    arr = (HostBlock * (nb if g is None else len(g)))()
    rc = lib.tessb200_host_tess(len(p), _fp(p), own.ctypes.data_as(C.POINTER(C.c_int)) if own is not None else None, _fp(dmin), _fp(dmax),
                                nb, _fp(bb), 0 if g is None else len(g), g.ctypes.data_as(C.POINTER(C.c_int)) if g is not None else None,
                                float(margin0), int(max_rounds), float(max_growth), int(threads), arr)
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
            margin=b.ghost_margin, rounds=b.rounds, seconds=b.seconds))
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
