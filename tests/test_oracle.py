"""The CPU checkers against each other and against the golden vectors (no GPU).

  reference  = unmodified /root/reference sources (oracle/_ref/libtess_ref.so; only where built)
  port       = oracle/dense_oracle.c, the plain-C restatement
  golden     = tests/golden/*.npz, written by the reference (tests/golden/make_golden.py)
"""
import hashlib
import os
import tempfile

import numpy as np
import pytest

from conftest import dataset, assert_same_bits
from golden_util import load_small, load_c1


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def check_against_golden(impl):
    z, blocks, gs = load_small()
    for i, b in enumerate(blocks):
        assert np.array_equal(impl.fill_vert_to_tet(len(b["particles"]), b["tets"]), z[f"b{i}_v2t"])
        assert_same_bits(impl.circumcenters(b["tets"], b["particles"]), z[f"b{i}_cc"], "circumcenters")
        assert np.array_equal(impl.complete(len(b["particles"]), b["tets"], b["vert_to_tet"]), z[f"b{i}_complete"])
        assert_same_bits(impl.volumes(len(b["particles"]), b["tets"], b["particles"], b["vert_to_tet"]), z[f"b{i}_volume"], "volume")
    for alg in (0, 1):
        for proj in (0, 1):
            o = impl.dense(blocks, gs, alg=alg, project=bool(proj))
            for i in range(len(blocks)):
                assert o["block_min_idx"][i] == list(z[f"alg{alg}_proj{proj}_b{i}_min_idx"])
                assert_same_bits(o["block_density"][i], z[f"alg{alg}_proj{proj}_b{i}_density"], f"alg{alg} proj{proj} block {i}")
            assert_same_bits(o["step"], z[f"alg{alg}_proj{proj}_step"], "step")


def test_port_matches_golden(port):
    check_against_golden(port)


def test_port_matches_golden_config1(port):
    c1 = load_c1()
    blk = dataset("c1")[0]
    if sha(blk["tets"]) != str(c1["tets_sha"]) or sha(blk["particles"]) != str(c1["particles_sha"]):
        # gen_particles must always reproduce; the tets depend on the SciPy/Qhull build
        assert_same_bits(blk["particles"][:8], c1["particles_head"], "gen_particles head")
        pytest.skip("this SciPy/Qhull produced different tets than the fixture's")
    assert sha(port.circumcenters(blk["tets"], blk["particles"])) == str(c1["cc_sha"])
    assert sha(port.volumes(blk["num_orig"], blk["tets"], blk["particles"], blk["vert_to_tet"])) == str(c1["volume_sha"])
    for alg in (0, 1):
        g = port.dense([blk], (64, 64, 64), alg=alg)["grid"].reshape(-1)
        assert sha(g) == str(c1[f"alg{alg}_grid_sha"])
        assert_same_bits(g[c1["sample_idx"]], c1[f"alg{alg}_sample"], "sample")


def test_gen_particles_is_the_reference_sequence():
    # src/tess.cpp:281-293 with glibc rand(): first particles of block gid 0 in [0,31]^3
    c1 = load_c1()
    from tess2_b200.harness import particles
    p = particles.gen_particles(0, [0, 0, 0], [31, 31, 31])
    assert p.shape == (32768, 3)
    assert_same_bits(p[:8], c1["particles_head"], "gen_particles")


def test_reference_matches_golden(reference):
    check_against_golden(reference)


@pytest.mark.parametrize("name,gs", [("u16x8", (32, 32, 32)), ("clump8", (48, 48, 48)), ("aniso", (40, 28, 17)), ("tiny", (8, 8, 8))])
def test_port_matches_reference(port, reference, name, gs):
    blocks = dataset(name)
    b = blocks[0]
    assert np.array_equal(port.fill_vert_to_tet(len(b["particles"]), b["tets"]), reference.fill_vert_to_tet(len(b["particles"]), b["tets"]))
    assert_same_bits(port.circumcenters(b["tets"], b["particles"]), reference.circumcenters(b["tets"], b["particles"]), "cc")
    assert np.array_equal(port.complete(b["num_orig"], b["tets"], b["vert_to_tet"]), reference.complete(b["num_orig"], b["tets"], b["vert_to_tet"]))
    assert_same_bits(port.volumes(b["num_orig"], b["tets"], b["particles"], b["vert_to_tet"]),
                     reference.volumes(b["num_orig"], b["tets"], b["particles"], b["vert_to_tet"]), "volumes")
    for alg in (0, 1):
        for proj in (False, True):
            o1 = reference.dense(blocks, gs, alg=alg, project=proj)
            o2 = port.dense(blocks, gs, alg=alg, project=proj)
            assert o1["block_min_idx"] == o2["block_min_idx"] and o1["block_num_idx"] == o2["block_num_idx"]
            for i, (d1, d2) in enumerate(zip(o1["block_density"], o2["block_density"])):
                assert_same_bits(d1, d2, f"{name} alg{alg} proj{proj} block {i}")


def test_port_given_bounds_and_eps(port, reference):
    blocks = dataset("u16x8")
    # given bounds must enclose the data: the reference indexes out of bounds otherwise (src/dense.cpp:288-290)
    gb = ([-1.0, -2.0, -0.5], [17.0, 16.5, 15.5])
    for alg in (0, 1):
        o1 = reference.dense(blocks, (30, 26, 22), alg=alg, given_bounds=gb, eps=1e-3, mass=2.5)
        o2 = port.dense(blocks, (30, 26, 22), alg=alg, given_bounds=gb, eps=1e-3, mass=2.5)
        for d1, d2 in zip(o1["block_density"], o2["block_density"]):
            assert_same_bits(d1, d2, "given bounds")


def test_reference_writegrid_is_the_assembled_grid(reference):
    # the reference's own MPI-IO subarray writer (through the stub) == placing each block at its min_idx
    blocks = dataset("u16x8")
    with tempfile.TemporaryDirectory() as td:
        f = os.path.join(td, "dense.raw")
        o = reference.dense(blocks, (32, 32, 32), outfile=f)
        raw = np.fromfile(f, dtype=np.float32).reshape(32, 32, 32)
    assert_same_bits(raw, o["grid"], "WriteGrid")


def test_mass_conservation_port(port):
    # sum(grid) * div == number of depositing cells * mass, up to grid points lost
    # outside the bounds; for the tess algorithm on config 1 nothing is lost
    blk = dataset("c1")[0]
    o = port.dense([blk], (64, 64, 64))
    div = float(o["step"][0]) * float(o["step"][1]) * float(o["step"][2])
    tot = o["grid"].astype(np.float64).sum() * div
    assert abs(tot - round(tot)) < 1e-6 * tot
    assert 20000 < round(tot) < 32768


@pytest.mark.parametrize("gb", [([-1.5], [16.5]), ([-1.5, -1.5], [16.5, 16.5]), ([-1.5, -1.5, -2.0], [16.5, 16.5, 17.0]),
                                ([-1.5, -1.5, 2.5], [16.5, 16.5, 12.5])])
def test_port_given_bounds_match_reference(port, reference, gb):
    # 1, 2 or 3 given bounds (src/dense.cpp:1725-1735), 3-D and projected; the last case projects with a z range
    # narrower than the data: well defined in the reference because the projected index drops z
    blocks = dataset("u16x8")
    for proj in (False, True):
        if len(gb[0]) == 3 and gb[0][2] > 0 and not proj:
            continue
        for alg in (0, 1):
            o1 = reference.dense(blocks, (24, 24, 24), alg=alg, project=proj, given_bounds=gb)
            o2 = port.dense(blocks, (24, 24, 24), alg=alg, project=proj, given_bounds=gb)
            assert o1["block_min_idx"] == o2["block_min_idx"] and o1["block_num_idx"] == o2["block_num_idx"]
            assert_same_bits(o1["step"], o2["step"], "step")
            for d1, d2 in zip(o1["block_density"], o2["block_density"]):
                assert_same_bits(d1, d2, f"given {gb} alg{alg} proj{proj}")


def test_cell_windows_partition_the_work(port, reference):
    # bench.py's CPU legs cut a block's cells into windows so that every host core works: a window leaves the other
    # cells to the reference with vert_to_tet = -1 (skipped at src/dense.cpp:251).  Port == reference on every window,
    # and the windows add up to the whole run.
    blocks = dataset("u16x8")
    full = reference.dense(blocks, (32, 32, 32))["grid"].astype(np.float64)
    acc = np.zeros_like(full)
    for g in (0, 5):
        for first in range(0, 512, 200):
            o1 = reference.dense(blocks, (32, 32, 32), only_gid=g, first_cell=first, max_cells=first + 200)
            o2 = port.dense(blocks, (32, 32, 32), only_gid=g, first_cell=first, max_cells=first + 200)
            assert_same_bits(o1["grid"], o2["grid"], f"window {g}:{first}")
    for g in range(8):
        for first in range(0, 512, 200):
            acc += reference.dense(blocks, (32, 32, 32), only_gid=g, first_cell=first, max_cells=first + 200)["grid"]
    assert np.abs(acc - full).max() < 1e-5 * full.max()
    # the caller's vert_to_tet is not touched
    assert (blocks[0]["vert_to_tet"][:200] >= -1).all() and (blocks[0]["vert_to_tet"][:200] != -1).any()
