"""The kernels' __host__ __device__ logic (tess2_b200/csrc/cell_core.cuh, host_geom.hpp),
single-stepped on the CPU by tests/emul/emul.cpp, against the oracle.  This is what can be
checked without a GPU: star walk, edge circulation, Newell normals, plane tests, the scan-line
state machine, CIC weights, span emission + key order, the accumulate step, BlockGridParams."""
import numpy as np
import pytest

from conftest import dataset, assert_same_bits
from golden_util import load_small


@pytest.mark.parametrize("name,gs", [("c1", (64, 64, 64)), ("u16x8", (32, 32, 32)), ("clump8", (48, 48, 48)),
                                     ("aniso", (40, 28, 17)), ("tiny", (8, 8, 8)), ("clump8", (20000, 6, 6))])
def test_emul_matches_port(port, emul, name, gs):
    blocks = dataset(name)
    for b in blocks[:2]:
        assert np.array_equal(emul.fill_vert_to_tet(len(b["particles"]), b["tets"]), b["vert_to_tet"])
        assert_same_bits(emul.circumcenters(b["tets"], b["particles"]), port.circumcenters(b["tets"], b["particles"]), "cc")
        assert np.array_equal(emul.complete(b["num_orig"], b["tets"], b["vert_to_tet"]), port.complete(b["num_orig"], b["tets"], b["vert_to_tet"]))
        assert_same_bits(emul.volumes(b["num_orig"], b["tets"], b["particles"], b["vert_to_tet"]),
                         port.volumes(b["num_orig"], b["tets"], b["particles"], b["vert_to_tet"]), "volumes")
    for alg in (0, 1):
        for proj in (False, True):
            o1 = port.dense(blocks, gs, alg=alg, project=proj)
            o2 = emul.dense(blocks, gs, alg=alg, project=proj)
            assert o1["block_min_idx"] == o2["block_min_idx"] and o1["block_num_idx"] == o2["block_num_idx"]
            for i, (d1, d2) in enumerate(zip(o1["block_density"], o2["block_density"])):
                assert_same_bits(d1, d2, f"{name} alg{alg} proj{proj} block {i}")


def test_emul_matches_golden(emul):
    z, blocks, gs = load_small()
    for alg in (0, 1):
        for proj in (0, 1):
            o = emul.dense(blocks, gs, alg=alg, project=bool(proj))
            for i in range(len(blocks)):
                assert_same_bits(o["block_density"][i], z[f"alg{alg}_proj{proj}_b{i}_density"], f"alg{alg} proj{proj} block {i}")


@pytest.mark.parametrize("name,gs", [("c1", (64, 64, 64)), ("u16x8", (32, 32, 32)), ("clump8", (48, 48, 48)), ("aniso", (40, 28, 17))])
def test_emul_dtfe_matches_port(port, emul, name, gs):
    # alg 2 (first-order DTFE) has no reference implementation: the comparison is between
    # the device logic and the repo's own CPU statement of it ("parity unpinned")
    blocks = dataset(name)
    o1 = port.dense(blocks, gs, alg=2)
    o2 = emul.dense(blocks, gs, alg=2)
    covered = 0
    for i, (d1, d2) in enumerate(zip(o1["block_density"], o2["block_density"])):
        assert_same_bits(d1, d2, f"{name} dtfe block {i}")
        covered += int((d1 > 0).sum())
    # most of the grid is covered by valid tets (not the non-cubic case, whose grid is mostly padding)
    assert covered > (0.0 if name == "aniso" else 0.6) * gs[0] * gs[1] * gs[2]


# given grid bounds (the run scripts pass two, with a projection: DENSE_TEST, TESS_DENSE_TEST).  Bounds wider
# than the data on x and y -- narrower ones make the reference itself write out of bounds (test_oracle.py) --
# and a z range NARROWER than the data under projection: the projected index drops z, so those points deposit.
GIVEN = [([-1.5], [16.5]), ([-1.5, -1.5], [16.5, 16.5]), ([-1.5, -1.5, -2.0], [16.5, 16.5, 17.0]), ([-1.5, -1.5, 2.5], [16.5, 16.5, 12.5])]


@pytest.mark.parametrize("gb", GIVEN)
def test_emul_given_bounds(port, emul, gb):
    blocks = dataset("u16x8")
    narrow_z = len(gb[0]) == 3 and gb[0][2] > 0
    for proj in (False, True):
        if narrow_z and not proj:
            continue          # 3-D with a narrower z range: the reference writes out of bounds
        for alg in (0, 1):
            o1 = port.dense(blocks, (24, 24, 24), alg=alg, project=proj, given_bounds=gb)
            o2 = emul.dense(blocks, (24, 24, 24), alg=alg, project=proj, given_bounds=gb)
            assert o1["block_min_idx"] == o2["block_min_idx"] and o1["block_num_idx"] == o2["block_num_idx"]
            for i, (d1, d2) in enumerate(zip(o1["block_density"], o2["block_density"])):
                assert_same_bits(d1, d2, f"given {gb} alg{alg} proj{proj} block {i}")
            if narrow_z:      # the points beyond the z range did deposit: about the mass of the run whose z range holds the data
                wide = port.dense(blocks, (24, 24, 24), alg=alg, project=True, given_bounds=(gb[0][:2], gb[1][:2]))
                tot = lambda o: sum(float(d.astype(np.float64).sum()) for d in o["block_density"])
                assert abs(tot(o1) - tot(wide)) < 0.05 * tot(wide)


def test_fuzz_reference_port_device_logic(reference, port, emul):
    # a fixed slice of tests/fuzz_logic.py (run that script for longer sweeps)
    import fuzz_logic
    n, failures = fuzz_logic.run(seed=12345, cases=120, checkers=(reference, port, emul), log=lambda *a: None)
    assert n == 120 and not failures, failures[:3]


def test_cic_gather_merge_adds_in_particle_order(emul):
    """k_cic_gather's sum at a grid point (cic_merge_sum, cell_core.cuh): eight ascending id lists merged by id, float adds in
    that order (src/dense.cpp:539) -- against a plain loop over the sorted ids; a step past the end of the lists yields -0.0f,
    which must not change a bit (NaN and infinities included)."""
    import ctypes
    lib = emul.lib
    rng = np.random.default_rng(31)
    for case in range(300):
        lens = rng.integers(0, [1, 3, 9, 40][case % 4] + 1, size=8)
        if case % 7 == 0:
            lens[rng.integers(0, 8)] = rng.integers(50, 400)
        n_part = int(lens.sum()) + 5
        ids = rng.permutation(n_part)[:int(lens.sum())].astype(np.uint32)
        # the sorted array: some foreign entries in front, then list after list, each ascending
        front = rng.integers(0, 4)
        sorted_ids = np.concatenate([np.full(front, 0, np.uint32)] + [np.sort(ids[int(lens[:n].sum()):int(lens[:n + 1].sum())]) for n in range(8)] + [np.zeros(1, np.uint32)])
        pos = (front + np.concatenate([[0], np.cumsum(lens)[:-1]])).astype(np.uint32)
        end = (pos + lens).astype(np.uint32)
        vals = (rng.random((n_part, 8), dtype=np.float32) * np.float32(10.0) ** rng.integers(-6, 6, size=(n_part, 8)).astype(np.float32)).astype(np.float32)
        if case % 11 == 0 and n_part:
            vals[rng.integers(0, n_part), :] = np.float32(np.inf) if case % 2 else np.float32(np.nan)
        if case % 13 == 0:
            vals[:] = 0.0
        out = np.zeros(2, np.float32)
        vals_sorted = np.ascontiguousarray(vals[sorted_ids])          # k_cic_permute: the weights in the order of the sorted ids
        lib.emu_cic_point(sorted_ids.ctypes.data_as(ctypes.c_void_p), vals_sorted.ctypes.data_as(ctypes.c_void_p), pos.ctypes.data_as(ctypes.c_void_p),
                          end.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
        want = np.float32(0.0)
        order = sorted((int(sorted_ids[q]), n) for n in range(8) for q in range(int(pos[n]), int(end[n])))
        with np.errstate(all="ignore"):
            for pid, n in order:
                want = np.float32(want + vals[pid, n])
        assert_same_bits(out[0:1], np.array([want], np.float32), f"case {case}")
        assert_same_bits(out[1:2], np.array([want], np.float32), f"case {case} one step past the end")
