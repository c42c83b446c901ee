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
