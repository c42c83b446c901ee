import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _make(args, cwd):
    subprocess.run(["make", "--no-print-directory"] + args, cwd=cwd, check=True, stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def port():
    """the plain-C restatement (oracle/dense_oracle.c)"""
    from oracle import ref
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libtess_oracle.so")):
        _make(["port"], os.path.join(ROOT, "oracle"))
    return ref.Checker("port")


@pytest.fixture(scope="session")
def reference():
    """the unmodified reference sources; only buildable where /root/reference exists"""
    from oracle import ref
    path = os.path.join(ROOT, "oracle", "_ref", "libtess_ref.so")
    if not os.path.exists(path):
        if os.path.isdir("/root/reference/src"):
            _make(["ref"], os.path.join(ROOT, "oracle"))
        else:
            pytest.skip("oracle/_ref/libtess_ref.so not built and /root/reference absent")
    return ref.Checker("reference")


@pytest.fixture(scope="session")
def emul():
    """CPU single-stepping of the kernels' __host__ __device__ logic (tests/emul/emul.cpp)"""
    from oracle import ref
    d = os.path.join(ROOT, "tests", "emul")
    out = os.path.join(d, "_build", "libtess_emul.so")
    srcs = [os.path.join(d, "emul.cpp"), os.path.join(ROOT, "tess2_b200", "csrc", "cell_core.cuh"),
            os.path.join(ROOT, "tess2_b200", "csrc", "host_geom.hpp")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
                        "-o", out, srcs[0]], check=True)
    return ref.Checker("emul", out, "emu_")


# ---- datasets (tessellated once per session) ---------------------------------------------------
def _with_v2t(blocks):
    from tess2_b200.harness import delaunay
    for b in blocks:
        b["vert_to_tet"] = delaunay.fill_vert_to_tet(len(b["particles"]), b["tets"])
    return blocks


_CACHE = {}


def dataset(name):
    """Named inputs shared by the CPU and GPU tests."""
    if name in _CACHE:
        return _CACHE[name]
    from tess2_b200.harness import particles, decomp, delaunay
    if name == "c1":            # BASELINE config 1: 32^3 uniform, 1 block (gen_particles, srand(0))
        dom = ([0, 0, 0], [31, 31, 31])
        p = particles.gen_particles(0, *dom)
        b = decomp.regular_blocks(*dom, 1)
        blocks = delaunay.tessellate(p, decomp.assign_regular(p, b), b, *dom)
    elif name == "u16x8":       # 16^3 uniform in 8 regular blocks (gen_particles per block)
        dom = ([0, 0, 0], [15, 15, 15])
        b = decomp.regular_blocks(*dom, 8)
        ps = [particles.gen_particles(g, mn, mx) for g, (mn, mx) in enumerate(b)]
        allp = np.concatenate(ps)
        owner = np.concatenate([np.full(len(q), g, np.int32) for g, q in enumerate(ps)])
        blocks = delaunay.tessellate(allp, owner, b, *dom)
    elif name == "clump8":      # clustered, kd-tree 8 blocks
        dom = ([0, 0, 0], [31, 31, 31])
        p = particles.clustered_particles(20000, *dom, seed=2031, n_clumps=12)
        b, owner = decomp.kdtree_blocks(p, *dom, 8)
        blocks = delaunay.tessellate(p, owner, b, *dom)
    elif name == "aniso":       # non-cubic domain: the grid is padded on two axes
        dom = ([0, 0, 0], [24, 13, 7])
        p = particles.uniform_particles(3000, *dom, seed=5)
        b = decomp.regular_blocks(*dom, 4)
        owner = decomp.assign_regular(p, b)
        blocks = delaunay.tessellate(p, owner, b, *dom)
    elif name == "tiny":        # a handful of particles: most cells are infinite
        dom = ([0, 0, 0], [3, 3, 3])
        p = particles.uniform_particles(40, *dom, seed=11)
        b = decomp.regular_blocks(*dom, 1)
        blocks = delaunay.tessellate(p, decomp.assign_regular(p, b), b, *dom)
    else:
        raise KeyError(name)
    _CACHE[name] = _with_v2t(blocks)
    return _CACHE[name]


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def assert_same_bits(a, b, what=""):
    """Bit equality; NaNs compare equal to NaNs whatever their payload (x86 and the GPU produce different
    quiet-NaN patterns for the same invalid operation, e.g. inf/inf in the CIC weights at large offsets)."""
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    diff = bits(a) != bits(b)
    if a.dtype == np.float32 and b.dtype == np.float32:
        diff &= ~(np.isnan(a) & np.isnan(b))
    if diff.any():
        nd = int(diff.sum())
        with np.errstate(all="ignore"):
            rel = np.nanmax(np.abs(a.astype(np.float64) - b.astype(np.float64)) / np.maximum(np.abs(a.astype(np.float64)), 1e-30))
        raise AssertionError(f"{what}: {nd} of {a.size} values differ (max rel {rel:.3g})")


def write_grid_from(o, blocks, gs, project, path):
    """tessb200_write_grid (host code of libtess_b200.so, no device needed) over per-block densities that an
    oracle run produced."""
    import ctypes as C
    from tess2_b200 import lib as _l
    arr = (_l.Block * len(blocks))()
    keep = []
    for i, d in enumerate(o["block_density"]):
        d = np.ascontiguousarray(d, np.float32).reshape(-1)
        keep.append(d)
        arr[i].gid = int(blocks[i]["gid"])
        arr[i].density = d.ctypes.data_as(_l.f32p)
        arr[i].density_capacity = d.size
        arr[i].num_grid_pts = d.size
        for k in range(3):
            arr[i].block_min_idx[k] = o["block_min_idx"][i][k]
            arr[i].block_num_idx[k] = o["block_num_idx"][i][k]
    prm = _l.DenseParams()
    prm.project = 1 if project else 0
    for k in range(3):
        prm.glo_num_idx[k] = int(gs[k])
    _l.check(_l.load().tessb200_write_grid(str(path).encode(), C.byref(prm), len(blocks), arr))
