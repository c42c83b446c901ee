"""examples/drivers/: the reference's three example drivers (examples/tess, examples/dense,
examples/tess-dense) over the two C ABIs, taking the reference's own command lines (TESS_TEST,
DENSE_TEST, TESS_DENSE_TEST).  CPU: `tess` writes a block file that reads back as the blocks the
Python binding of the same host library produces; argument errors; no device -> a loud failure, no
CPU fallback.  GPU: tess -> del.out -> dense equals tess-dense byte for byte, and the oracle on the
blocks of the file."""
import os
import subprocess

import numpy as np
import pytest

from tess2_b200 import host_tess
from tess2_b200.harness import particles
from conftest import assert_same_bits, ROOT


@pytest.fixture(scope="module")
def drivers(tmp_path_factory):
    out = tmp_path_factory.mktemp("drivers")
    subprocess.run(["make", "--no-print-directory", "-C", os.path.join(ROOT, "examples", "drivers"), f"OUT={out}"], check=True,
                   stdout=subprocess.DEVNULL)
    return out


def python_blocks(side, tb):
    dom = ([0, 0, 0], [side - 1] * 3)
    bounds = host_tess.regular_blocks(*dom, tb)
    ps = [particles.gen_particles(g, mn, mx) for g, (mn, mx) in enumerate(bounds)]
    p = np.concatenate(ps)
    own = np.concatenate([np.full(len(q), g, np.int32) for g, q in enumerate(ps)])
    return host_tess.tess(p, own, bounds, *dom), ps


def test_tess_driver_writes_the_blocks(drivers, tmp_path):
    f = tmp_path / "del.out"
    r = subprocess.run([str(drivers / "tess"), "8", "-1", "12", "12", "12", "0", "-1", "-1", "0", "0", str(f)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    back, dmin, dmax, _ = host_tess.read_blocks(str(f))
    blocks, ps = python_blocks(12, 8)
    assert list(dmin) == [0, 0, 0] and list(dmax) == [11, 11, 11] and len(back) == 8
    for a, b in zip(blocks, back):
        assert a["gid"] == b["gid"] and a["num_orig"] == b["num_orig"] and b["complete"] == 1
        for k in ("particles", "tets", "vert_to_tet", "bounds_min", "bounds_max"):
            assert_same_bits(np.asarray(a[k]), b[k], k)
        # every ghost is the particle rem_lids of block rem_gids (src/tess.cpp:654-673)
        ghosts = b["particles"][b["num_orig"]:]
        src = np.stack([ps[g][l] for g, l in zip(b["rem_gids"], b["rem_lids"])]) if len(ghosts) else ghosts
        assert_same_bits(ghosts, src, "ghost provenance")
    # "!" = no output file
    r = subprocess.run([str(drivers / "tess"), "1", "-1", "6", "6", "6", "0", "-1", "-1", "0", "0", "!"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0 and sorted(os.listdir(tmp_path)) == ["del.out"]


def test_tess_driver_wrap_and_walls(drivers, tmp_path):
    # wrap = 1: periodic neighbours (images of particles as ghosts): every cell of every block is complete, where the open domain
    # leaves the cells at its boundary unbounded; walls = 1 is parsed and read by nothing in the reference: accepted, no effect
    from tess2_b200.harness import delaunay
    from oracle import ref
    port = ref.Checker("port")
    out = {}
    for wrap, walls in ((0, 0), (1, 0), (0, 1)):
        f = tmp_path / f"del_{wrap}{walls}.out"
        r = subprocess.run([str(drivers / "tess"), "8", "-1", "10", "10", "10", "0", "-1", "-1", str(wrap), str(walls), str(f)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        out[(wrap, walls)] = host_tess.read_blocks(str(f))[0]
    for a, b in zip(out[(0, 0)], out[(0, 1)]):
        for k in ("particles", "tets", "vert_to_tet"):
            assert_same_bits(np.asarray(a[k]), b[k], "walls changes nothing: " + k)
    done_open = sum(int(port.complete(b["num_orig"], b["tets"], b["vert_to_tet"]).sum()) for b in out[(0, 0)])
    done_wrap = sum(int(port.complete(b["num_orig"], b["tets"], b["vert_to_tet"]).sum()) for b in out[(1, 0)])
    total = sum(b["num_orig"] for b in out[(1, 0)])
    assert done_wrap == total and done_open < total
    for b in out[(1, 0)]:
        assert np.array_equal(b["vert_to_tet"], delaunay.fill_vert_to_tet(len(b["particles"]), b["tets"]))


def test_driver_argument_errors(drivers, tmp_path):
    assert subprocess.run([str(drivers / "tess"), "8"], capture_output=True).returncode == 2
    assert subprocess.run([str(drivers / "dense"), "a", "b", "0", "8", "8"], capture_output=True).returncode == 2
    assert subprocess.run([str(drivers / "tess-dense"), "0", "8", "8", "8", "8"], capture_output=True).returncode == 2
    # an unreadable block file
    r = subprocess.run([str(drivers / "dense"), str(tmp_path / "none.out"), "x.raw", "0", "8", "8", "8", "!", "1", "0"], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open" in r.stderr


def test_dense_driver_without_a_device_fails_loudly(drivers, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    f = tmp_path / "del.out"
    subprocess.run([str(drivers / "tess"), "1", "-1", "6", "6", "6", "0", "-1", "-1", "0", "0", str(f)], check=True, capture_output=True)
    r = subprocess.run([str(drivers / "dense"), str(f), str(tmp_path / "dense.raw"), "0", "8", "8", "8", "!", "1", "0"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr and not (tmp_path / "dense.raw").exists()


@pytest.mark.gpu
def test_drivers_end_to_end(drivers, tmp_path, port):
    f = tmp_path / "del.out"
    subprocess.run([str(drivers / "tess"), "8", "-1", "16", "16", "16", "0", "-1", "-1", "0", "0", str(f)], check=True, capture_output=True)
    blocks, _, _, _ = host_tess.read_blocks(str(f))
    # 3-D, both algorithms: the decoupled flow, the one-program flow and the oracle agree bit for bit
    for alg in ("0", "1"):
        raw1, raw2 = tmp_path / f"dense_{alg}.raw", tmp_path / f"tess_dense_{alg}.raw"
        r = subprocess.run([str(drivers / "dense"), str(f), str(raw1), alg, "32", "32", "32", "!", "1", "0"], capture_output=True, text=True)
        assert r.returncode == 0 and "total mass" in r.stderr, r.stderr
        r = subprocess.run([str(drivers / "tess-dense"), alg, "8", "16", "16", "16", "0", "-1", "-1", "0", "0", str(raw2), "32", "32", "32", "!", "1", "0"],
                           capture_output=True, text=True)
        assert r.returncode == 0 and "Overall time" in r.stderr, r.stderr
        a, b = np.fromfile(raw1, np.float32), np.fromfile(raw2, np.float32)
        assert_same_bits(a, b, "dense vs tess-dense")
        o = port.dense(blocks, (32, 32, 32), alg=int(alg))
        assert_same_bits(a.reshape(32, 32, 32), o["grid"], f"dense.raw alg {alg} vs oracle")
    # DENSE_TEST's own argument shape: projection onto xy with two given bounds
    raw1, raw2 = tmp_path / "p1.raw", tmp_path / "p2.raw"
    tail = ["24", "24", "24", "0.0", "0.0", "1.0", "1", "2", "-1.5", "-1.5", "16.5", "16.5"]
    subprocess.run([str(drivers / "dense"), str(f), str(raw1), "0"] + tail, check=True, capture_output=True)
    subprocess.run([str(drivers / "tess-dense"), "0", "8", "16", "16", "16", "0", "-1", "-1", "0", "0", str(raw2)] + tail, check=True, capture_output=True)
    a, b = np.fromfile(raw1, np.float32), np.fromfile(raw2, np.float32)
    assert a.size == 24 * 24 and np.isfinite(a).all() and a.sum() > 0
    assert_same_bits(a, b, "projected dense vs tess-dense")
    # against the oracle: its per-block projected densities, assembled by the library's own WriteGrid (host code,
    # itself checked against the reference's writer in test_abi.py)
    from conftest import write_grid_from
    o = port.dense(blocks, (24, 24, 24), alg=0, project=True, given_bounds=([-1.5, -1.5], [16.5, 16.5]))
    raw3 = tmp_path / "p3.raw"
    write_grid_from(o, blocks, (24, 24, 24), True, raw3)
    assert_same_bits(a, np.fromfile(raw3, np.float32), "projected dense.raw vs oracle")
