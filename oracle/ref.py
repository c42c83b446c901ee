"""ctypes bindings for the two CPU checkers (TEST INFRASTRUCTURE ONLY).

  * oracle/_ref/libtess_ref.so    -- the unmodified reference sources (kind "reference")
  * oracle/_ref/libtess_oracle.so -- the plain-C restatement dense_oracle.c (kind "port")

Both export the same C entry points (ref_* / orc_*), so one wrapper class serves both.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product (tess2_b200/) never does.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int)


class Block(C.Structure):
    _fields_ = [
        ("gid", C.c_int), ("num_orig_particles", C.c_int), ("num_particles", C.c_int),
        ("particles", f32p), ("num_tets", C.c_int), ("tets", i32p), ("vert_to_tet", i32p),
        ("bounds_min", C.c_float * 3), ("bounds_max", C.c_float * 3),
        ("density", f32p), ("density_capacity", C.c_longlong),
        ("block_min_idx", C.c_int * 3), ("block_num_idx", C.c_int * 3), ("num_grid_pts", C.c_int),
    ]


class Params(C.Structure):
    _fields_ = [
        ("alg", C.c_int), ("num_given_bounds", C.c_int),
        ("given_mins", C.c_float * 3), ("given_maxs", C.c_float * 3),
        ("project", C.c_int), ("proj_plane", C.c_float * 3),
        ("mass", C.c_float), ("eps", C.c_float), ("glo_num_idx", C.c_int * 3),
        ("data_mins", C.c_float * 3), ("data_maxs", C.c_float * 3),
        ("grid_phys_mins", C.c_float * 3), ("grid_phys_maxs", C.c_float * 3),
        ("grid_step_size", C.c_float * 3), ("seconds", C.c_double),
    ]


def _fp(a):
    return a.ctypes.data_as(f32p)


def _ip(a):
    return a.ctypes.data_as(i32p)


class Checker:
    """One of the two CPU implementations behind a common Python interface."""

    def __init__(self, kind, path=None, prefix=None):
        self.kind = kind
        if path is None:
            assert kind in ("reference", "port", "dropin")
            name = {"reference": "libtess_ref.so", "port": "libtess_oracle.so", "dropin": "libtess_dropin.so"}[kind]
            prefix = {"reference": "ref_", "port": "orc_", "dropin": "drp_"}[kind]
            path = os.path.join(HERE, "_ref", name)
        self.prefix = prefix
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        self.lib = C.CDLL(path)
        self._f("dense").restype = C.c_int

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    # ---- per-array helpers -------------------------------------------------
    def fill_vert_to_tet(self, num_particles, tets):
        out = np.empty(num_particles, dtype=np.int32)
        self._f("fill_vert_to_tet")(C.c_int(num_particles), C.c_int(len(tets)), _ip(tets), _ip(out))
        return out

    def circumcenters(self, tets, particles):
        out = np.empty((len(tets), 3), dtype=np.float32)
        self._f("circumcenters")(C.c_int(len(tets)), _ip(tets), _fp(particles), _fp(out))
        return out

    def complete(self, num_verts, tets, vert_to_tet):
        out = np.empty(num_verts, dtype=np.int32)
        self._f("complete")(C.c_int(num_verts), C.c_int(len(tets)), _ip(tets), _ip(vert_to_tet), _ip(out))
        return out

    def volumes(self, num_verts, tets, particles, vert_to_tet):
        out = np.empty(num_verts, dtype=np.float32)
        self._f("volumes")(C.c_int(num_verts), C.c_int(len(tets)), _ip(tets), _fp(particles),
                           _ip(vert_to_tet), _fp(out))
        return out

    # ---- the dense stage ---------------------------------------------------
    def dense(self, blocks, gsize, alg=0, mass=1.0, eps=1e-4, project=False, proj_plane=(0.0, 0.0, 1.0),
              given_bounds=None, only_gid=-1, outfile=None, max_cells=-1, first_cell=0, assemble=True):
        """blocks: list of dicts (gid, particles, num_orig, tets, bounds_min, bounds_max[, vert_to_tet]).
        Returns dict(grid=global [gz,gy,gx] (or [gy,gx] when projected... per-block arrays only),
        block_density=[...], block_min_idx, block_num_idx, params)."""
        nb = len(blocks)
        arr = (Block * nb)()
        keep = []
        gs = [int(g) for g in gsize]
        cap = gs[0] * gs[1] * gs[2]
        for i, b in enumerate(blocks):
            pa = np.ascontiguousarray(b["particles"], dtype=np.float32)
            te = np.ascontiguousarray(b["tets"], dtype=np.int32)
            keep += [pa, te]
            arr[i].gid = int(b["gid"])
            arr[i].num_orig_particles = int(b["num_orig"])
            arr[i].num_particles = len(pa)
            arr[i].particles = _fp(pa)
            arr[i].num_tets = len(te)
            arr[i].tets = _ip(te)
            v2t = b.get("vert_to_tet")
            if v2t is not None:
                v2t = np.ascontiguousarray(v2t, dtype=np.int32)
                keep.append(v2t)
                arr[i].vert_to_tet = _ip(v2t)
            for d in range(3):
                arr[i].bounds_min[d] = float(b["bounds_min"][d])
                arr[i].bounds_max[d] = float(b["bounds_max"][d])
            dens = np.zeros(cap if nb * cap <= (1 << 24) else min(cap, self._block_cap(b, blocks, gs)), dtype=np.float32)
            keep.append(dens)
            arr[i].density = _fp(dens)
            arr[i].density_capacity = len(dens)
            b["_dens"] = dens
        p = Params()
        p.alg = alg
        p.num_given_bounds = 0
        if given_bounds is not None:
            p.num_given_bounds = len(given_bounds[0])     # 1..3 leading axes (src/dense.cpp:1725-1735)
            for d in range(p.num_given_bounds):
                p.given_mins[d] = float(given_bounds[0][d])
                p.given_maxs[d] = float(given_bounds[1][d])
        p.project = 1 if project else 0
        for d in range(3):
            p.proj_plane[d] = float(proj_plane[d])
            p.glo_num_idx[d] = gs[d]
        p.mass = mass
        p.eps = eps
        if first_cell:      # reference / port only: a window of cells [first_cell, max_cells) per block
            self._f("set_cell_window")(C.c_int(first_cell))
        rc = self._f("dense")(C.byref(p), C.c_int(nb), arr, C.c_int(only_gid),
                              outfile.encode() if outfile else None, C.c_int(max_cells))
        if rc != 0:
            raise RuntimeError(f"{self.prefix}dense failed: {rc}")
        out = dict(params=p, seconds=p.seconds, block_density=[], block_min_idx=[], block_num_idx=[])
        if self.kind == "port":
            # deposits that fell outside their block's sub-grid: the reference writes out of bounds there, so its
            # result is undefined for this input (0 on every input the parity tests use)
            f = self.lib.orc_out_of_range_deposits
            f.restype = C.c_longlong
            out["out_of_range"] = int(f())
        for i, b in enumerate(blocks):
            n = arr[i].num_grid_pts
            num = [arr[i].block_num_idx[d] for d in range(3)]
            out["block_min_idx"].append([arr[i].block_min_idx[d] for d in range(3)])
            out["block_num_idx"].append(num)
            dens = b.pop("_dens")[:n]
            shape = (num[1], num[0]) if project else (num[2], num[1], num[0])
            out["block_density"].append(dens.reshape(shape))
        out["step"] = np.array([p.grid_step_size[d] for d in range(3)], dtype=np.float32)
        out["grid_phys_mins"] = np.array([p.grid_phys_mins[d] for d in range(3)], dtype=np.float32)
        out["data_mins"] = np.array([p.data_mins[d] for d in range(3)], dtype=np.float32)
        out["data_maxs"] = np.array([p.data_maxs[d] for d in range(3)], dtype=np.float32)
        if not project and assemble:
            out["grid"] = assemble_grid(gs, out["block_min_idx"], out["block_num_idx"], out["block_density"])
        return out

    @staticmethod
    def _block_cap(b, blocks, gs):
        # generous bound on a block's sub-grid: fraction of the domain per axis, plus slack
        lo = np.min([bb["bounds_min"] for bb in blocks], axis=0).astype(np.float64)
        hi = np.max([bb["bounds_max"] for bb in blocks], axis=0).astype(np.float64)
        n = 1
        for d in range(3):
            frac = (float(b["bounds_max"][d]) - float(b["bounds_min"][d])) / max(float(hi[d] - lo[d]), 1e-30)
            n *= min(gs[d], int(frac * gs[d]) + 4)
        return n

    def dtfe_vertex_density(self, num_verts, tets, particles, vert_to_tet, mass=1.0):
        """port only: per-vertex density of the DTFE mode (oracle/dense_oracle.c, orc_dtfe_vertex_density)"""
        tets = np.ascontiguousarray(tets, dtype=np.int32)
        particles = np.ascontiguousarray(particles, dtype=np.float32)
        v2t = np.ascontiguousarray(vert_to_tet, dtype=np.int32)
        rho = np.empty(num_verts, dtype=np.float32)
        self._f("dtfe_vertex_density")(C.c_int(num_verts), C.c_int(len(tets)), _ip(tets), _fp(particles), _ip(v2t), C.c_float(mass), _fp(rho))
        return rho

    def cell_points(self, block, cell, data_mins, data_maxs, grid_phys_mins, step, mass=1.0, eps=1e-4, cap=1 << 16):
        arr = Block()
        pa = np.ascontiguousarray(block["particles"], dtype=np.float32)
        te = np.ascontiguousarray(block["tets"], dtype=np.int32)
        arr.gid = int(block["gid"]); arr.num_orig_particles = int(block["num_orig"])
        arr.num_particles = len(pa); arr.particles = _fp(pa); arr.num_tets = len(te); arr.tets = _ip(te)
        v2t = np.ascontiguousarray(block["vert_to_tet"], dtype=np.int32)
        arr.vert_to_tet = _ip(v2t)
        idx = np.zeros((cap, 3), dtype=np.int32)
        ms = np.zeros(cap, dtype=np.float32)
        cmin = np.zeros(3, np.float32); cmax = np.zeros(3, np.float32)
        nf = C.c_int(0)
        f = self._f("cell_points")
        f.restype = C.c_int
        n = f(C.byref(arr), C.c_int(cell), _fp(np.asarray(data_mins, np.float32)), _fp(np.asarray(data_maxs, np.float32)),
              _fp(np.asarray(grid_phys_mins, np.float32)), _fp(np.asarray(step, np.float32)), C.c_float(mass), C.c_float(eps),
              C.c_int(cap), _ip(idx), _fp(ms), _fp(cmin), _fp(cmax), C.byref(nf))
        m = max(0, min(n, cap))
        return n, idx[:m].copy(), ms[:m].copy(), cmin, cmax, nf.value


def assemble_grid(gs, block_min_idx, block_num_idx, block_density):
    """What the reference's WriteGrid does with MPI-IO subarrays (src/dense.cpp:831-850):
    every block's [nz][ny][nx] array lands at its (min_idx) offset of the global C-order grid."""
    grid = np.zeros((gs[2], gs[1], gs[0]), dtype=np.float32)
    for mn, num, d in zip(block_min_idx, block_num_idx, block_density):
        grid[mn[2]:mn[2] + num[2], mn[1]:mn[1] + num[1], mn[0]:mn[0] + num[0]] = d
    return grid
