"""ctypes binding of libtess_b200.so (the C ABI in include/tess_b200.h).

There is no CPU fallback: if the library has not been built, or no CUDA device is present,
every entry point fails loudly (ImportError / TessB200Error).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtess_b200.so")

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int)

# every symbol include/tess_b200.h declares (tests check that the library exports all of them)
EXPORTS = [
    "tessb200_create", "tessb200_destroy", "tessb200_last_error", "tessb200_version",
    "tessb200_dense", "tessb200_dense_upload", "tessb200_dense_run", "tessb200_dense_download",
    "tessb200_dense_geometry", "tessb200_dense_device_density",
    "tessb200_fill_vert_to_tet", "tessb200_circumcenters", "tessb200_cell_volumes", "tessb200_cell_volumes_ms",
    "tessb200_write_grid", "tessb200_check_block", "tessb200_dtfe_vertex_density",
    "tessb200_comm_unique_id", "tessb200_comm_init", "tessb200_comm_size", "tessb200_dense_set_layout",
]


class TessB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"tess_b200 error {code}: {msg}")
        self.code = code


class Block(C.Structure):
    """struct tessb200_block"""
    _fields_ = [
        ("gid", C.c_int), ("num_orig_particles", C.c_int), ("num_particles", C.c_int),
        ("particles", f32p), ("num_tets", C.c_int), ("tets", i32p), ("vert_to_tet", i32p),
        ("bounds_min", C.c_float * 3), ("bounds_max", C.c_float * 3),
        ("density", f32p), ("density_capacity", C.c_int64),
        ("block_min_idx", C.c_int * 3), ("block_num_idx", C.c_int * 3), ("num_grid_pts", C.c_int64),
    ]


class DenseParams(C.Structure):
    """struct tessb200_dense_params"""
    _fields_ = [
        ("alg", C.c_int), ("num_given_bounds", C.c_int),
        ("given_mins", C.c_float * 3), ("given_maxs", C.c_float * 3),
        ("project", C.c_int), ("proj_plane", C.c_float * 3),
        ("mass", C.c_float), ("eps", C.c_float), ("glo_num_idx", C.c_int * 3),
        ("data_mins", C.c_float * 3), ("data_maxs", C.c_float * 3),
        ("grid_phys_mins", C.c_float * 3), ("grid_phys_maxs", C.c_float * 3),
        ("grid_step_size", C.c_float * 3),
    ]


class DenseStats(C.Structure):
    """struct tessb200_dense_stats"""
    _fields_ = [
        ("num_cells", C.c_int64), ("num_no_tet", C.c_int64), ("num_incomplete", C.c_int64),
        ("num_outside", C.c_int64), ("num_deposit_cells", C.c_int64), ("num_cic_fallback", C.c_int64),
        ("num_slow_cells", C.c_int64), ("num_spans", C.c_int64), ("num_tets", C.c_int64),
        ("num_grid_pts", C.c_int64), ("num_kernel_launches", C.c_int64), ("tot_mass", C.c_double), ("max_dense", C.c_float),
        ("ms_upload", C.c_float), ("ms_circumcenters", C.c_float), ("ms_cells", C.c_float),
        ("ms_scan", C.c_float), ("ms_sort", C.c_float), ("ms_deposit", C.c_float),
        ("ms_exchange", C.c_float), ("ms_download", C.c_float), ("ms_total_device", C.c_float),
        ("ms_bfs", C.c_float), ("ms_nbrs", C.c_float), ("ms_faces", C.c_float),
        ("num_faces", C.c_int64), ("num_candidates", C.c_int64),
        ("ms_slow_path", C.c_float), ("ms_fused", C.c_float), ("num_shared_deposits", C.c_int64),
        ("ms_emit", C.c_float), ("ms_direct", C.c_float),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_lib = None


def load():
    """Loads libtess_b200.so (building is __graft_entry__.build()'s / `make -C tess2_b200/csrc`'s job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `make -C tess2_b200/csrc` "
                          "(tess2_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.tessb200_last_error.restype = C.c_char_p
    lib.tessb200_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    lib.tessb200_destroy.argtypes = [C.c_void_p]
    lib.tessb200_destroy.restype = None
    lib.tessb200_dense.argtypes = [C.c_void_p, C.POINTER(DenseParams), C.c_int, C.POINTER(Block), f32p, C.POINTER(DenseStats)]
    lib.tessb200_dense_upload.argtypes = [C.c_void_p, C.c_int, C.POINTER(Block)]
    lib.tessb200_dense_run.argtypes = [C.c_void_p, C.POINTER(DenseParams), C.POINTER(DenseStats)]
    lib.tessb200_dense_download.argtypes = [C.c_void_p, C.c_int, C.POINTER(Block), f32p]
    lib.tessb200_dense_geometry.argtypes = [C.c_void_p, C.POINTER(DenseParams), C.c_int, C.POINTER(Block)]
    lib.tessb200_dense_device_density.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    lib.tessb200_fill_vert_to_tet.argtypes = [C.c_void_p, C.c_int, C.c_int, i32p, i32p]
    lib.tessb200_circumcenters.argtypes = [C.c_void_p, C.c_int, f32p, C.c_int, i32p, f32p]
    lib.tessb200_cell_volumes.argtypes = [C.c_void_p, C.c_int, C.c_int, f32p, C.c_int, i32p, i32p, C.c_float, i32p, f32p, f32p]
    lib.tessb200_cell_volumes_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    lib.tessb200_write_grid.argtypes = [C.c_char_p, C.POINTER(DenseParams), C.c_int, C.POINTER(Block)]
    lib.tessb200_check_block.argtypes = [C.POINTER(Block), C.c_int]
    lib.tessb200_dtfe_vertex_density.argtypes = [C.c_void_p, C.c_int, f32p, C.c_int, i32p, i32p, C.c_float, f32p]
    lib.tessb200_comm_unique_id.argtypes = [C.c_void_p]
    lib.tessb200_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.tessb200_comm_size.argtypes = [C.c_void_p]
    lib.tessb200_dense_set_layout.argtypes = [C.c_void_p, C.c_int, i32p, f32p, i32p]
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise TessB200Error(rc, load().tessb200_last_error().decode(errors="replace"))
