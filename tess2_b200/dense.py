"""Host-side mirror of tess2's dense-stage interface over the C ABI (include/tess_b200.h).

Names and argument meaning follow the reference (include/tess/dense.hpp:33-98, 146-239;
include/tess/volume.h; src/tess.cpp:767-787):

    dense(alg_type, num_given_bounds, given_mins, given_maxs, project, proj_plane, mass,
          eps, glo_num_idx, blocks)                       <- dense(), src/dense.cpp:30-103
    WriteGrid(outfile, result)                            <- WriteGrid(), src/dense.cpp:751-870
    fill_circumcenters(tets, particles)                   <- src/volume.cpp:6-11
    volume(...) / complete(...)                           <- src/volume.cpp:13-54, src/tet.cpp:337-378
    fill_vert_to_tet(num_particles, tets)                 <- src/tess.cpp:767-787

A "block" is a dict with the dblock_t fields the stage reads (include/tess/delaunay.h:38-63):
gid, particles [n,3] f32 (originals first), num_orig, tets [T,8] i32 (verts[4], tets[4]),
bounds_min, bounds_max, and optionally vert_to_tet.  The out-parameters of the reference's
dense() (data_mins/maxs, grid_phys_mins/maxs, grid_step_size, DBlock::density) come back in a
DenseResult.  All compute happens on the GPU; this module only marshals pointers.
"""
import ctypes as C
import numpy as np
from . import lib as _l

DENSE_TESS = 0   # include/tess/dense.hpp:35
DENSE_CIC = 1    # include/tess/dense.hpp:36
DENSE_DTFE = 2   # not in the reference: first-order DTFE


def _fp(a):
    return a.ctypes.data_as(_l.f32p)


def _ip(a):
    return a.ctypes.data_as(_l.i32p)


class DenseResult:
    """What the reference's dense() leaves behind: the float[3] out-parameters, each block's
    density sub-array + BlockGridParams, and the stats dense_stats() prints."""

    def __init__(self):
        self.params = None
        self.data_mins = self.data_maxs = None
        self.grid_phys_mins = self.grid_phys_maxs = self.grid_step_size = None
        self.block_density = []
        self.block_min_idx = []
        self.block_num_idx = []
        self.gids = []
        self.grid = None
        self.stats = None
        self.project = False
        self._blocks_c = None
        self._keep = None

    @property
    def div(self):
        s = self.grid_step_size
        return np.float32(s[0]) * np.float32(s[1]) if self.project else np.float32(s[0]) * np.float32(s[1]) * np.float32(s[2])


class Context:
    """One GPU.  Wraps tessb200_create / tessb200_destroy."""

    def __init__(self, device=0):
        self.lib = _l.load()
        self.handle = C.c_void_p()
        _l.check(self.lib.tessb200_create(C.byref(self.handle), int(device)))
        self.device = device
        self._keep = None
        self._arr = None
        self._nb = 0
        self._blocks = None

    def close(self):
        if self.handle:
            self.lib.tessb200_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- marshalling -------------------------------------------------------------------
    def _marshal(self, blocks, with_density, gsize, project):
        nb = len(blocks)
        arr = (_l.Block * nb)()
        keep = []
        for i, b in enumerate(blocks):
            pa = b["particles"]
            if not (isinstance(pa, np.ndarray) and pa.dtype == np.float32 and pa.flags.c_contiguous):
                pa = np.ascontiguousarray(pa, dtype=np.float32)
            te = b["tets"]
            if not (isinstance(te, np.ndarray) and te.dtype == np.int32 and te.flags.c_contiguous):
                te = np.ascontiguousarray(te, dtype=np.int32)
            keep += [pa, te]
            arr[i].gid = int(b["gid"])
            arr[i].num_orig_particles = int(b["num_orig"])
            arr[i].num_particles = pa.shape[0]
            arr[i].particles = _fp(pa)
            arr[i].num_tets = te.shape[0]
            arr[i].tets = _ip(te)
            v2t = b.get("vert_to_tet")
            if v2t is not None:
                if not (isinstance(v2t, np.ndarray) and v2t.dtype == np.int32 and v2t.flags.c_contiguous):
                    v2t = np.ascontiguousarray(v2t, dtype=np.int32)
                keep.append(v2t)
                arr[i].vert_to_tet = _ip(v2t)
            for d in range(3):
                arr[i].bounds_min[d] = float(b["bounds_min"][d])
                arr[i].bounds_max[d] = float(b["bounds_max"][d])
        return arr, keep

    @staticmethod
    def make_params(alg_type, num_given_bounds, given_mins, given_maxs, project, proj_plane, mass, eps, glo_num_idx):
        p = _l.DenseParams()
        p.alg = int(alg_type)
        p.num_given_bounds = int(num_given_bounds)
        for d in range(3):
            p.given_mins[d] = float(given_mins[d]) if given_mins is not None and d < len(given_mins) else 0.0
            p.given_maxs[d] = float(given_maxs[d]) if given_maxs is not None and d < len(given_maxs) else 0.0
            p.proj_plane[d] = float(proj_plane[d]) if proj_plane is not None else (1.0 if d == 2 else 0.0)
            p.glo_num_idx[d] = int(glo_num_idx[d])
        p.project = 1 if project else 0
        p.mass = float(mass)
        p.eps = float(eps)
        return p

    # ---- three-step interface (inputs resident in HBM between calls) ------------------------
    def upload(self, blocks):
        arr, keep = self._marshal(blocks, False, None, False)
        _l.check(self.lib.tessb200_dense_upload(self.handle, len(blocks), arr))
        self._arr, self._keep, self._nb, self._blocks = arr, keep, len(blocks), blocks

    def run(self, params, want_stats=True):
        st = _l.DenseStats()
        _l.check(self.lib.tessb200_dense_run(self.handle, C.byref(params), C.byref(st) if want_stats else None))
        return st

    def geometry(self, params):
        _l.check(self.lib.tessb200_dense_geometry(self.handle, C.byref(params), self._nb, self._arr))
        return [([self._arr[i].block_min_idx[d] for d in range(3)], [self._arr[i].block_num_idx[d] for d in range(3)],
                 int(self._arr[i].num_grid_pts)) for i in range(self._nb)]

    def download(self, params, want_grid=True, out_blocks=None, out_grid=None):
        """D2H of every block's density (and the assembled global grid for 3-D runs)."""
        geo = self.geometry(params)
        res = DenseResult()
        res.project = bool(params.project)
        dens = []
        for i, (mn, num, npts) in enumerate(geo):
            d = out_blocks[i] if out_blocks is not None else np.empty(npts, dtype=np.float32)
            dens.append(d)
            self._arr[i].density = _fp(d)
            self._arr[i].density_capacity = d.size
        grid = None
        if want_grid and not params.project:
            gs = [params.glo_num_idx[d] for d in range(3)]
            grid = out_grid if out_grid is not None else np.zeros((gs[2], gs[1], gs[0]), dtype=np.float32)
        _l.check(self.lib.tessb200_dense_download(self.handle, self._nb, self._arr, _fp(grid) if grid is not None else None))
        self._fill_result(res, params, geo, dens, grid)
        return res

    def _fill_result(self, res, params, geo, dens, grid):
        res.params = params
        for name in ("data_mins", "data_maxs", "grid_phys_mins", "grid_phys_maxs", "grid_step_size"):
            setattr(res, name, np.array([getattr(params, name)[d] for d in range(3)], dtype=np.float32))
        for i, (mn, num, npts) in enumerate(geo):
            res.gids.append(int(self._arr[i].gid))
            res.block_min_idx.append(mn)
            res.block_num_idx.append(num)
            shape = (num[1], num[0]) if params.project else (num[2], num[1], num[0])
            res.block_density.append(dens[i][:npts].reshape(shape))
        res.grid = grid
        res._blocks_c = self._arr
        res._keep = (self._keep, dens)

    # ---- one-call interface == the reference's dense() ---------------------------------------
    def dense(self, alg_type, num_given_bounds, given_mins, given_maxs, project, proj_plane, mass, eps, glo_num_idx,
              blocks, want_grid=True, out_blocks=None, want_stats=True):
        """tessb200_dense(): host buffers in, host buffers out, copies pipelined with the kernels."""
        params = self.make_params(alg_type, num_given_bounds, given_mins, given_maxs, project, proj_plane, mass, eps, glo_num_idx)
        return self.dense_params(params, blocks, want_grid=want_grid, out_blocks=out_blocks, want_stats=want_stats)

    def dense_params(self, params, blocks, want_grid=True, out_blocks=None, want_stats=True):
        arr, keep = self._marshal(blocks, False, None, False)
        self._arr, self._keep, self._nb, self._blocks = arr, keep, len(blocks), blocks
        geo = self.geometry(params)                 # host-only: BlockGridParams of every block
        dens = []
        for i, (mn, num, npts) in enumerate(geo):
            d = out_blocks[i] if out_blocks is not None else np.empty(npts, dtype=np.float32)
            dens.append(d)
            arr[i].density = _fp(d)
            arr[i].density_capacity = d.size
        grid = None
        if want_grid and not params.project:
            gs = [params.glo_num_idx[d] for d in range(3)]
            grid = np.zeros((gs[2], gs[1], gs[0]), dtype=np.float32)
        st = _l.DenseStats()
        _l.check(self.lib.tessb200_dense(self.handle, C.byref(params), len(blocks), arr, _fp(grid) if grid is not None else None,
                                         C.byref(st) if want_stats else None))
        res = DenseResult()
        res.project = bool(params.project)
        self._fill_result(res, params, geo, dens, grid)
        res.stats = st
        return res

    # ---- per-tet / per-site -------------------------------------------------------------------
    def fill_vert_to_tet(self, num_particles, tets):
        tets = np.ascontiguousarray(tets, dtype=np.int32)
        out = np.empty(num_particles, dtype=np.int32)
        _l.check(self.lib.tessb200_fill_vert_to_tet(self.handle, num_particles, tets.shape[0], _ip(tets), _ip(out)))
        return out

    def fill_circumcenters(self, tets, particles):
        tets = np.ascontiguousarray(tets, dtype=np.int32)
        particles = np.ascontiguousarray(particles, dtype=np.float32)
        out = np.empty((tets.shape[0], 3), dtype=np.float32)
        _l.check(self.lib.tessb200_circumcenters(self.handle, particles.shape[0], _fp(particles), tets.shape[0], _ip(tets), _fp(out)))
        return out

    def cell_volumes(self, num_sites, tets, particles, vert_to_tet=None, mass=1.0):
        """(complete, volume, density) for sites [0, num_sites)."""
        tets = np.ascontiguousarray(tets, dtype=np.int32)
        particles = np.ascontiguousarray(particles, dtype=np.float32)
        v2t = None if vert_to_tet is None else np.ascontiguousarray(vert_to_tet, dtype=np.int32)
        comp = np.empty(num_sites, dtype=np.int32)
        vol = np.empty(num_sites, dtype=np.float32)
        den = np.empty(num_sites, dtype=np.float32)
        _l.check(self.lib.tessb200_cell_volumes(self.handle, num_sites, particles.shape[0], _fp(particles), tets.shape[0], _ip(tets),
                                                _ip(v2t) if v2t is not None else None, float(mass), _ip(comp), _fp(vol), _fp(den)))
        return comp, vol, den

    def cell_volumes_ms(self):
        """device ms of the kernels of the last cell_volumes() call (no copies)"""
        ms = C.c_float(0.0)
        _l.check(self.lib.tessb200_cell_volumes_ms(self.handle, C.byref(ms)))
        return float(ms.value)

    def dtfe_vertex_density(self, tets, particles, vert_to_tet=None, mass=1.0):
        """Per-particle density of the first-order DTFE mode (not in the reference): 4 m / (volume of the star), -1 where
        the star is infinite."""
        tets = np.ascontiguousarray(tets, dtype=np.int32)
        particles = np.ascontiguousarray(particles, dtype=np.float32)
        v2t = None if vert_to_tet is None else np.ascontiguousarray(vert_to_tet, dtype=np.int32)
        rho = np.empty(particles.shape[0], dtype=np.float32)
        _l.check(self.lib.tessb200_dtfe_vertex_density(self.handle, particles.shape[0], _fp(particles), tets.shape[0], _ip(tets),
                                                       _ip(v2t) if v2t is not None else None, float(mass), _fp(rho)))
        return rho


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def dense(alg_type, num_given_bounds, given_mins, given_maxs, project, proj_plane, mass, eps, glo_num_idx, blocks,
          ctx=None, want_grid=True):
    """The reference's dense() (include/tess/dense.hpp:75-89) for a list of blocks on one GPU."""
    ctx = ctx or default_context()
    return ctx.dense(alg_type, num_given_bounds, given_mins, given_maxs, project, proj_plane, mass, eps, glo_num_idx,
                     blocks, want_grid=want_grid)


def WriteGrid(outfile, result):
    """WriteGrid (src/dense.cpp:751-870): raw C-order float32, x fastest, no header; with
    projection the z-stacked blocks are summed first (ProjectGrid, src/dense.cpp:881-1023)."""
    lib = _l.load()
    arr = result._blocks_c
    for i, d in enumerate(result.block_density):
        arr[i].density = _fp(np.ascontiguousarray(d).reshape(-1))
    _l.check(lib.tessb200_write_grid(str(outfile).encode(), C.byref(result.params), len(result.block_density), arr))


def check_blocks(blocks, deep=False):
    """tessb200_check_block on every block (host code, no device): raises TessB200Error with the first finding."""
    lib = _l.load()
    arr, keep = Context._marshal(None, blocks, False, None, False)
    for i in range(len(blocks)):
        _l.check(lib.tessb200_check_block(C.byref(arr[i]), 1 if deep else 0))


def fill_vert_to_tet(num_particles, tets, ctx=None):
    return (ctx or default_context()).fill_vert_to_tet(num_particles, tets)


def fill_circumcenters(tets, particles, ctx=None):
    return (ctx or default_context()).fill_circumcenters(tets, particles)


def volume(num_sites, tets, particles, vert_to_tet=None, ctx=None):
    """volume() for every site in [0, num_sites): -1 infinite, -2 site in no tet."""
    return (ctx or default_context()).cell_volumes(num_sites, tets, particles, vert_to_tet)[1]


def complete(num_sites, tets, particles, vert_to_tet=None, ctx=None):
    return (ctx or default_context()).cell_volumes(num_sites, tets, particles, vert_to_tet)[0]
