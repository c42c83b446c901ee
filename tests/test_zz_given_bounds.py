"""GPU parity for the grid-bounds arguments of the run scripts: 1-3 given bounds, 3-D and projected, including a
projection whose given z range is narrower than the data (points of any z index deposit: the projected index
drops z, src/dense.cpp:1047-1090).  The CPU side of the same cases: test_oracle.py (port vs the unmodified
reference) and test_emul.py (device logic vs port)."""
import numpy as np
import pytest

from conftest import dataset, assert_same_bits
from test_emul import GIVEN

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("gb", GIVEN)
def test_given_bounds_match_oracle(port, gb):
    import tess2_b200
    blocks = dataset("u16x8")
    ctx = tess2_b200.Context(0)
    try:
        for proj in (False, True):
            if len(gb[0]) == 3 and gb[0][2] > 0 and not proj:
                continue
            for alg in (0, 1):
                o = port.dense(blocks, (24, 24, 24), alg=alg, project=proj, given_bounds=gb)
                res = ctx.dense(alg, len(gb[0]), gb[0], gb[1], proj, (0.0, 0.0, 1.0), 1.0, 1e-4, (24, 24, 24), blocks)
                assert res.block_min_idx == o["block_min_idx"] and res.block_num_idx == o["block_num_idx"]
                for i, (d1, d2) in enumerate(zip(res.block_density, o["block_density"])):
                    assert_same_bits(d1, d2, f"given {gb} alg{alg} proj{proj} block {i}")
    finally:
        ctx.close()


def test_random_configurations_match_oracle(port):
    # tests/fuzz_logic.py's generator (distributions, offsets, 1-16 regular / kd-tree blocks, grids of 4-90 points,
    # both algorithms, 3-D / projected, eps, 0-3 given bounds wider or narrower than the data) through the CUDA path.
    # The bits must agree with the port in every case: where the port reports deposits outside a block's sub-grid
    # (the reference is undefined there; port and device logic both skip them) and where the reference's own
    # arithmetic is non-finite (seed 2025 case 11: offset 1e4, CIC -- the 2*eps nudge vanishes in fp32, vol = 0,
    # w = v0/vol, src/dense.cpp:1837-1841).  NaNs compare equal to NaNs (assert_same_bits).
    import tess2_b200
    import fuzz_logic
    rng = np.random.default_rng(2025)
    ctx = tess2_b200.Context(0)
    compared = 0
    try:
        for case in range(60):
            try:
                blocks, gs, args, desc = fuzz_logic.random_case(rng)
            except RuntimeError:
                continue
            gb = args["given_bounds"]
            try:
                o = port.dense(blocks, gs, **args)
            except RuntimeError:
                continue            # a block without grid points: rejected by the oracle (and an error in the library)
            res = ctx.dense(args["alg"], 0 if gb is None else len(gb[0]), None if gb is None else gb[0], None if gb is None else gb[1],
                            args["project"], (0.0, 0.0, 1.0), args["mass"], args["eps"], gs, blocks)
            assert res.block_min_idx == o["block_min_idx"] and res.block_num_idx == o["block_num_idx"], desc
            for i, (d1, d2) in enumerate(zip(res.block_density, o["block_density"])):
                assert_same_bits(d1, d2, f"case {case} block {i}: {desc}")
            compared += 1
    finally:
        ctx.close()
    assert compared >= 25


@pytest.mark.parametrize("name", ["c1", "clump8", "tiny"])
def test_dtfe_vertex_density_export(port, name):
    # per-site density of the DTFE mode (not in the reference: compared with the repo's CPU statement)
    import tess2_b200
    ctx = tess2_b200.Context(0)
    try:
        for b in dataset(name)[:2]:
            n = len(b["particles"])
            want = port.dtfe_vertex_density(n, b["tets"], b["particles"], b["vert_to_tet"], mass=0.5)
            got = ctx.dtfe_vertex_density(b["tets"], b["particles"], b["vert_to_tet"], mass=0.5)
            assert_same_bits(got, want, f"{name} vertex densities")
            assert_same_bits(ctx.dtfe_vertex_density(b["tets"], b["particles"], None, mass=0.5), want, f"{name} vertex densities, NULL vert_to_tet")
            assert (want[want > 0] > 0).any() and (want == -1).any()       # hull vertices have no finite star
    finally:
        ctx.close()
