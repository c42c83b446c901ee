"""The hand-off file between tess() and dense() ("del.out"): tess_save / tess_load
(src/tess.cpp:118-152) with the payload order of save_block_light / load_block_light
(src/tess.cpp:198-259).  CPU tests: write -> read round trip; files packed here with `struct`
straight from the grammar (a different, longer link record in front of every payload, both Bounds
layouts, blocks out of gid order) -> the reader; corrupt files -> an error code, never a crash; the
dense oracle on blocks that went through the file is bit-identical to the oracle on the originals."""
import os
import struct

import numpy as np
import pytest

from tess2_b200 import host_tess
from conftest import dataset, assert_same_bits

KEYS = ("particles", "tets", "vert_to_tet", "bounds_min", "bounds_max")


def same_block(a, b):
    assert a["gid"] == b["gid"] and a["num_orig"] == b["num_orig"]
    for k in KEYS:
        assert_same_bits(np.asarray(a[k]), np.asarray(b[k]).reshape(np.asarray(a[k]).shape), k)


def pack_bounds(mn, mx, layout):
    if layout == host_tess.BOUNDS_DYNAMIC:
        return struct.pack("<Q3fQ3f", 3, *map(float, mn), 3, *map(float, mx))
    return struct.pack("<4f4f", *map(float, mn), 0.0, *map(float, mx), 0.0)


def pack_file(blocks, data_min, data_max, layout, link, extra=b"xyz"):
    """diy::io::write_blocks' layout restated independently of block_file.cpp."""
    body, toc = b"", []
    for b in blocks:
        p = np.ascontiguousarray(b["particles"], np.float32)
        t = np.ascontiguousarray(b["tets"], np.int32)
        v = np.ascontiguousarray(b["vert_to_tet"], np.int32)
        ng = len(p) - b["num_orig"]
        rec = link(b)
        rec += struct.pack("<i", b["gid"]) + pack_bounds(b["bounds_min"], b["bounds_max"], layout)
        rec += pack_bounds(b["bounds_min"], b["bounds_max"], layout) + pack_bounds(data_min, data_max, layout)
        rec += struct.pack("<ii", b["num_orig"], len(p)) + p.tobytes()
        rec += np.full(ng, 7, np.int32).tobytes() + np.arange(ng, dtype=np.int32).tobytes()
        dens = np.asarray(b.get("density", np.zeros(0, np.float32)), np.float32)
        rec += struct.pack("<i", len(dens)) + dens.tobytes() + struct.pack("<ii", 1, len(t)) + t.tobytes() + v.tobytes()
        toc.append((b["gid"], len(body), len(rec)))
        body += rec
    footer = struct.pack("<Q", len(toc))
    for gid, off, cnt in sorted(toc):
        footer += struct.pack("<i4xqq", gid, off, cnt)
    footer += struct.pack("<QQ", 0, len(extra)) + extra
    return body + footer + struct.pack("<Q", len(footer))


def regular_link(b):
    """something like a RegularLink<Bounds<float>>: type id, neighbours, dimension, direction map,
    core / bounds / per-neighbour bounds, wrap vectors -- deliberately full of ints equal to the gid"""
    gid = b["gid"]
    name = b"N3diy11RegularLinkINS_6BoundsIfEEEE"
    rec = struct.pack("<Q", len(name)) + name
    rec += struct.pack("<Q", 3) + struct.pack("<6i", gid, 0, gid + 1, 0, gid, 0)
    rec += struct.pack("<i", 3) + struct.pack("<Q", 3) + b"".join(struct.pack("<Q3ii", 3, gid, 0, -1, k) for k in range(3))
    rec += struct.pack("<Q3fQ3f", 3, 0.0, 0.0, 0.0, 3, 1.0, 1.0, 1.0) * 5
    rec += struct.pack("<Q", 3) + struct.pack("<Q3i", 3, gid, gid, gid) * 3
    return rec


def test_round_trip(tmp_path):
    blocks = dataset("u16x8")
    path = str(tmp_path / "del.out")
    host_tess.write_blocks(path, blocks, [0, 0, 0], [15, 15, 15], extra=b"times")
    back, dmin, dmax, layout = host_tess.read_blocks(path)
    assert layout == host_tess.BOUNDS_DYNAMIC and len(back) == len(blocks)
    assert list(dmin) == [0, 0, 0] and list(dmax) == [15, 15, 15]
    for a, b in zip(blocks, back):
        same_block(a, b)
        assert (b["rem_gids"] == -1).all() and len(b["rem_lids"]) == len(a["particles"]) - a["num_orig"]
        assert len(b["density"]) == 0
    size = sum(4 + 3 * 40 + 8 + 12 * len(b["particles"]) + 8 * (len(b["particles"]) - b["num_orig"]) + 4 + 8 + 32 * len(b["tets"])
               + 4 * len(b["particles"]) for b in blocks)
    assert os.path.getsize(path) > size      # payload bytes + link records + footer


@pytest.mark.parametrize("layout", [host_tess.BOUNDS_DYNAMIC, host_tess.BOUNDS_STATIC4])
def test_reads_files_packed_from_the_grammar(tmp_path, layout):
    blocks = [dict(b) for b in dataset("clump8")]
    blocks[3]["density"] = np.arange(24, dtype=np.float32)          # a block that already carries a density array
    order = [5, 0, 7, 2, 1, 6, 3, 4]                                # DIY writes rank by rank: not in gid order
    path = str(tmp_path / "del.out")
    with open(path, "wb") as f:
        f.write(pack_file([blocks[i] for i in order], [0, 0, 0], [31, 31, 31], layout, regular_link))
    back, dmin, dmax, found = host_tess.read_blocks(path)
    assert found == layout and [b["gid"] for b in back] == list(range(8))
    for a, b in zip(blocks, back):
        same_block(a, b)
        assert b["complete"] == 1 and (b["rem_gids"] == 7).all()
        assert (b["rem_lids"] == np.arange(len(b["rem_lids"]))).all()
    assert_same_bits(back[3]["density"], blocks[3]["density"], "density")
    # written again by the library in the other layout and read back: same blocks
    other = 1 - layout
    host_tess.write_blocks(path, back, dmin, dmax, layout=other)
    again, _, _, found = host_tess.read_blocks(path)
    assert found == other
    for a, b in zip(blocks, again):
        same_block(a, b)
    assert_same_bits(again[3]["density"], blocks[3]["density"], "density")


def test_empty_and_particle_free_blocks(tmp_path):
    path = str(tmp_path / "del.out")
    host_tess.write_blocks(path, [], [0, 0, 0], [1, 1, 1])
    back, _, _, _ = host_tess.read_blocks(path)
    assert back == []
    empty = dict(gid=4, num_orig=0, particles=np.zeros((0, 3), np.float32), tets=np.zeros((0, 8), np.int32),
                 vert_to_tet=np.zeros(0, np.int32), bounds_min=[0, 0, 0], bounds_max=[1, 1, 1])
    host_tess.write_blocks(path, [empty], [0, 0, 0], [1, 1, 1])
    back, _, _, _ = host_tess.read_blocks(path)
    assert len(back) == 1 and back[0]["gid"] == 4 and len(back[0]["particles"]) == 0 and len(back[0]["tets"]) == 0


def test_corrupt_files_are_errors(tmp_path):
    blocks = dataset("tiny")
    good = pack_file(blocks, [0, 0, 0], [3, 3, 3], host_tess.BOUNDS_DYNAMIC, regular_link)
    path = str(tmp_path / "bad.out")
    cases = {
        "empty": b"",
        "short": good[:11],
        "truncated": good[:-9],
        "footer size": good[:-8] + struct.pack("<Q", 1 << 40),
        "payload count": good[:200] + good[204:],
        "not a block file": os.urandom(4096),
    }
    for what, data in cases.items():
        with open(path, "wb") as f:
            f.write(data)
        with pytest.raises(RuntimeError):
            host_tess.read_blocks(path)
    with pytest.raises(RuntimeError):
        host_tess.read_blocks(str(tmp_path / "missing.out"))
    # a wrong num_particles in an otherwise intact file: the payload no longer ends at the block's end
    i = good.index(struct.pack("<ii", blocks[0]["num_orig"], len(blocks[0]["particles"])))
    with open(path, "wb") as f:
        f.write(good[:i + 4] + struct.pack("<i", len(blocks[0]["particles"]) - 1) + good[i + 8:])
    with pytest.raises(RuntimeError):
        host_tess.read_blocks(path)


def test_dense_oracle_through_the_file(tmp_path, port):
    blocks = dataset("u16x8")
    path = str(tmp_path / "del.out")
    host_tess.write_blocks(path, blocks, [0, 0, 0], [15, 15, 15])
    back, _, _, _ = host_tess.read_blocks(path)
    a = port.dense(blocks, (32, 32, 32), alg=0)
    b = port.dense(back, (32, 32, 32), alg=0)
    assert_same_bits(a["grid"], b["grid"], "grid")
