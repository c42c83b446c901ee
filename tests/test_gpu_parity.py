"""Parity of the CUDA path (through the C ABI, include/tess_b200.h) with the oracle on the same
inputs.  The bar: bit-exact for indices / flags; for fp32 results the tolerance north_star states
is 1e-5 relative -- the arithmetic is kept in the reference's order without FMA, so the tests ask
for bit equality and report the relative error if that ever fails."""
import hashlib
import os
import tempfile

import numpy as np
import pytest

from conftest import dataset, assert_same_bits
from golden_util import load_small, load_c1

pytestmark = pytest.mark.gpu

TOL = 1e-5  # north_star: cell volumes, densities and grid values within 1e-5 relative (fp32)


@pytest.fixture(scope="module")
def ctx():
    import tess2_b200
    c = tess2_b200.Context(0)
    yield c
    c.close()


def run_gpu(ctx, blocks, gs, alg=0, project=False, given_bounds=None, eps=1e-4, mass=1.0):
    ng = 0 if given_bounds is None else 3
    gmin = given_bounds[0] if given_bounds else None
    gmax = given_bounds[1] if given_bounds else None
    return ctx.dense(alg, ng, gmin, gmax, project, (0.0, 0.0, 1.0), mass, eps, gs, blocks)


def compare_dense(res, o, what):
    assert res.block_min_idx == o["block_min_idx"], what
    assert res.block_num_idx == o["block_num_idx"], what
    assert_same_bits(res.grid_step_size, o["step"], what + " step")
    assert_same_bits(res.grid_phys_mins, o["grid_phys_mins"], what + " grid_phys_mins")
    for i, (d1, d2) in enumerate(zip(res.block_density, o["block_density"])):
        assert_same_bits(d1, d2, f"{what} block {i}")
    if res.grid is not None:
        assert_same_bits(res.grid, o["grid"], what + " global grid")


# ---- per-tet / per-site ------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c1", "clump8", "tiny"])
def test_vert_to_tet_circumcenters_volumes(ctx, port, name):
    for b in dataset(name)[:2]:
        n = len(b["particles"])
        assert np.array_equal(ctx.fill_vert_to_tet(n, b["tets"]), b["vert_to_tet"])
        assert_same_bits(ctx.fill_circumcenters(b["tets"], b["particles"]), port.circumcenters(b["tets"], b["particles"]), "circumcenters")
        comp, vol, den = ctx.cell_volumes(n, b["tets"], b["particles"], b["vert_to_tet"])
        assert np.array_equal(comp, port.complete(n, b["tets"], b["vert_to_tet"]))
        ref_vol = port.volumes(n, b["tets"], b["particles"], b["vert_to_tet"])
        assert_same_bits(vol, ref_vol, "volumes")
        fin = ref_vol > 0
        assert np.allclose(den[fin], 1.0 / ref_vol[fin], rtol=TOL, atol=0)
        # vert_to_tet recomputed on the device when NULL
        comp2, vol2, _ = ctx.cell_volumes(n, b["tets"], b["particles"], None)
        assert np.array_equal(comp2, comp) and np.array_equal(vol2.view(np.uint32), vol.view(np.uint32))


# ---- the dense stage -----------------------------------------------------------------------------
@pytest.mark.parametrize("name,gs", [("c1", (64, 64, 64)), ("u16x8", (32, 32, 32)), ("clump8", (48, 48, 48)),
                                     ("aniso", (40, 28, 17)), ("tiny", (8, 8, 8))])
@pytest.mark.parametrize("alg", [0, 1])
@pytest.mark.parametrize("project", [False, True])
def test_dense_matches_oracle(ctx, port, name, gs, alg, project):
    blocks = dataset(name)
    o = port.dense(blocks, gs, alg=alg, project=project)
    res = run_gpu(ctx, blocks, gs, alg=alg, project=project)
    compare_dense(res, o, f"{name} alg{alg} proj{project}")


def test_dense_matches_golden_small(ctx):
    z, blocks, gs = load_small()
    for alg in (0, 1):
        for proj in (0, 1):
            res = run_gpu(ctx, blocks, gs, alg=alg, project=bool(proj))
            for i in range(len(blocks)):
                assert res.block_min_idx[i] == list(z[f"alg{alg}_proj{proj}_b{i}_min_idx"])
                assert_same_bits(res.block_density[i], z[f"alg{alg}_proj{proj}_b{i}_density"], f"golden alg{alg} proj{proj} block {i}")
            # WriteGrid: same bytes as the reference's MPI-IO writer produced
            with tempfile.TemporaryDirectory() as td:
                import tess2_b200
                f = os.path.join(td, "dense.raw")
                tess2_b200.WriteGrid(f, res)
                raw = np.fromfile(f, dtype=np.float32)
            assert_same_bits(raw, z[f"alg{alg}_proj{proj}_raw"], f"dense.raw alg{alg} proj{proj}")


def test_dense_matches_golden_config1(ctx):
    c1 = load_c1()
    blk = dataset("c1")[0]
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    if sha(blk["tets"]) != str(c1["tets_sha"]):
        pytest.skip("this SciPy/Qhull produced different tets than the fixture's")
    assert sha(ctx.fill_circumcenters(blk["tets"], blk["particles"])) == str(c1["cc_sha"])
    assert sha(ctx.cell_volumes(blk["num_orig"], blk["tets"], blk["particles"], blk["vert_to_tet"])[1]) == str(c1["volume_sha"])
    for alg in (0, 1):
        res = run_gpu(ctx, [blk], (64, 64, 64), alg=alg)
        assert sha(res.grid.reshape(-1)) == str(c1[f"alg{alg}_grid_sha"])


def test_vert_to_tet_null_and_given_bounds(ctx, port):
    blocks = [dict(b) for b in dataset("u16x8")]
    for b in blocks:
        b.pop("vert_to_tet")
    gb = ([-1.0, -2.0, -0.5], [17.0, 16.5, 15.5])
    for alg in (0, 1):
        o = port.dense(blocks, (30, 26, 22), alg=alg, given_bounds=gb, eps=1e-3, mass=2.5)
        res = run_gpu(ctx, blocks, (30, 26, 22), alg=alg, given_bounds=gb, eps=1e-3, mass=2.5)
        compare_dense(res, o, f"given bounds alg{alg}")


def test_stats_and_mass_conservation(ctx, port):
    blk = dataset("c1")[0]
    res = run_gpu(ctx, [blk], (64, 64, 64))
    st = res.stats
    comp = port.complete(blk["num_orig"], blk["tets"], blk["vert_to_tet"])
    assert st.num_cells == blk["num_orig"]
    assert st.num_incomplete == int((comp == 0).sum())
    assert st.num_no_tet == int((comp == -1).sum())
    assert st.num_deposit_cells + st.num_outside + st.num_incomplete + st.num_no_tet == st.num_cells
    # total deposited mass == number of depositing cells (mass 1), to 1e-6 (north_star)
    assert abs(st.tot_mass - st.num_deposit_cells) <= 1e-6 * st.num_deposit_cells
    tot = res.grid.astype(np.float64).sum() * float(res.div)
    assert abs(tot - st.num_deposit_cells) <= 1e-6 * st.num_deposit_cells
    assert st.max_dense == res.grid.max()
    assert st.num_tets == len(blk["tets"])


def test_big_cells_and_many_faces(ctx, port):
    # sparse particles on a fine grid: every index box exceeds the shared-memory scan kernel's
    # 2048-point limit, so the large-cell kernel does the work
    from tess2_b200.harness import particles, decomp, delaunay
    dom = ([0, 0, 0], [7, 7, 7])
    p = particles.uniform_particles(300, *dom, seed=3)
    b = decomp.regular_blocks(*dom, 1)
    blocks = delaunay.tessellate(p, decomp.assign_regular(p, b), b, *dom, workers=1)
    o = port.dense(blocks, (96, 96, 96))
    res = run_gpu(ctx, blocks, (96, 96, 96))
    assert res.stats.num_slow_cells > 0
    compare_dense(res, o, "big cells")
    # a site with a very large star: points on a sphere around a centre (degenerate, many faces)
    rng = np.random.default_rng(9)
    v = rng.standard_normal((400, 3))
    v /= np.linalg.norm(v, axis=1)[:, None]
    shell = (4.0 + 1.5 * v * (1 + 1e-3 * rng.standard_normal((400, 1)))).astype(np.float32)
    outer = (4.0 + 3.4 * v[:200]).astype(np.float32)
    pts = np.concatenate([[[4.0, 4.0, 4.0]], shell, outer]).astype(np.float32)
    blocks = delaunay.tessellate(pts, np.zeros(len(pts), np.int32), [(np.zeros(3, np.float32), np.full(3, 8, np.float32))],
                                 [0, 0, 0], [8, 8, 8], workers=1)
    o = port.dense(blocks, (40, 40, 40))
    res = run_gpu(ctx, blocks, (40, 40, 40))
    compare_dense(res, o, "large star")
    comp, vol, _ = ctx.cell_volumes(len(pts), blocks[0]["tets"], blocks[0]["particles"], None)
    assert_same_bits(vol, port.volumes(len(pts), blocks[0]["tets"], blocks[0]["particles"],
                                        port.fill_vert_to_tet(len(pts), blocks[0]["tets"])), "large star volume")


def test_long_rows(ctx, port):
    # a grid 20000 points wide: the deposit kernel's per-warp row buffer no longer fits four warps per
    # CTA, index boxes are wider than 32 points (generic scan-line walk), spans are long
    gs = (20000, 6, 6)
    for name in ("tiny", "clump8"):
        blocks = dataset(name)
        for alg in (0, 1):
            o = port.dense(blocks, gs, alg=alg)
            res = run_gpu(ctx, blocks, gs, alg=alg)
            compare_dense(res, o, f"long rows {name} alg {alg}")


def test_dense_on_native_tess_blocks(ctx, port):
    # blocks from the C++ tess() driver and its own Delaunay engine (include/tess_b200_host.h): the GPU
    # and the oracle take the same tets, so the bar is bit equality as everywhere else
    from tess2_b200 import host_tess
    from tess2_b200.harness import particles, decomp
    dom = ([0, 0, 0], [19, 19, 19])
    p = particles.clustered_particles(20 ** 3, *dom, seed=12)
    b, own = decomp.kdtree_blocks(p, *dom, 4)
    blocks = host_tess.tess(p, own, b, *dom)
    for alg in (0, 1):
        compare_dense(run_gpu(ctx, blocks, (40, 40, 40), alg=alg), port.dense(blocks, (40, 40, 40), alg=alg), f"native tess alg {alg}")


def test_dense_on_periodic_blocks(ctx, port):
    # wrap = 1 of the reference drivers: ghosts include the images of particles beyond the domain (tessb200_host_tess_periodic);
    # every cell is complete, and the cells that reach across the domain boundary are the ones dense() drops by its data-bounds test
    from tess2_b200 import host_tess
    from tess2_b200.harness import particles
    dom = (np.zeros(3, np.float32), np.full(3, 15, np.float32))
    p = particles.uniform_particles(16 ** 3, *dom, seed=21)
    bounds = host_tess.regular_blocks(*dom, 8)
    blocks = host_tess.tess(p, None, bounds, *dom, wrap=True, max_rounds=6, max_growth=8.0)
    assert all(b["settled"] for b in blocks)
    for alg in (0, 1):
        for proj in (False, True):
            compare_dense(run_gpu(ctx, blocks, (32, 32, 32), alg=alg, project=proj), port.dense(blocks, (32, 32, 32), alg=alg, project=proj),
                          f"periodic blocks alg {alg} proj {proj}")


@pytest.mark.parametrize("kd", [0, 1])
def test_c_example_end_to_end(ctx, port, tmp_path, kd):
    # examples/tess_dense.c: particles -> decomposition -> tess -> dense -> raw file through the two C ABIs only
    import subprocess
    from tess2_b200 import host_tess
    from tess2_b200.harness import particles
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "tess_dense"
    subprocess.run(["gcc", "-O2", "-I", os.path.join(root, "include"), "-o", str(exe), os.path.join(root, "examples", "tess_dense.c"),
                    "-L", os.path.join(root, "tess2_b200"), "-ltess_b200", "-ltess_b200_host", "-Wl,-rpath," + os.path.join(root, "tess2_b200")], check=True)
    raw = tmp_path / "dense.raw"
    out = subprocess.run([str(exe), "12", "8", "40", str(raw), "0", "0", str(kd)], check=True, capture_output=True, text=True).stdout
    assert "total mass" in out
    # the same pipeline from Python, checked against the oracle on the same blocks
    dom = ([0, 0, 0], [23, 23, 23])
    bounds = host_tess.regular_blocks(*dom, 8)
    ps = [particles.gen_particles(g, mn, mx) for g, (mn, mx) in enumerate(bounds)]
    p = np.concatenate(ps)
    own = np.concatenate([np.full(len(q), g, np.int32) for g, q in enumerate(ps)])
    if kd:
        bounds, own = host_tess.kdtree_blocks(p, *dom, 8)
    blocks = host_tess.tess(p, own, bounds, *dom)
    o = port.dense(blocks, (40, 40, 40))
    got = np.fromfile(raw, dtype=np.float32).reshape(40, 40, 40)
    assert_same_bits(got, o["grid"], "C example dense.raw")


def test_many_small_blocks(ctx, port):
    # 128 blocks through the one-call path: more groups than the per-block event rings hold (64), blocks
    # with a handful of cells, most deposits crossing block faces
    from tess2_b200 import host_tess
    from tess2_b200.harness import particles
    dom = ([0, 0, 0], [15, 15, 15])
    p = particles.uniform_particles(16 ** 3, *dom, seed=21)
    bounds = host_tess.regular_blocks(*dom, 128)
    blocks = host_tess.tess(p, None, bounds, *dom)
    assert len(blocks) == 128
    o = port.dense(blocks, (32, 32, 32))
    compare_dense(run_gpu(ctx, blocks, (32, 32, 32)), o, "128 blocks, one call")
    ctx.upload(blocks)
    params = ctx.make_params(0, 0, None, None, False, (0.0, 0.0, 1.0), 1.0, 1e-4, (32, 32, 32))
    ctx.run(params)
    res = ctx.download(params)
    assert_same_bits(res.grid, o["grid"], "128 blocks, resident")


def test_edge_cases(ctx):
    import tess2_b200
    # a block with particles but no tets, next to a normal block: nothing deposits from it
    blocks = [dict(b) for b in dataset("u16x8")]
    blocks[3] = dict(blocks[3], tets=np.zeros((0, 8), np.int32), vert_to_tet=None)
    res = run_gpu(ctx, blocks, (32, 32, 32))
    assert res.stats.num_no_tet >= blocks[3]["num_orig"]
    assert np.isfinite(res.grid).all()
    # bad arguments come back as error codes, not crashes
    with pytest.raises(tess2_b200.TessB200Error):
        run_gpu(ctx, blocks, (1, 32, 32))
    with pytest.raises(tess2_b200.TessB200Error):
        ctx.dense(7, 0, None, None, False, None, 1.0, 1e-4, (32, 32, 32), blocks)
    with pytest.raises(tess2_b200.TessB200Error):
        ctx.dense(0, 0, None, None, True, (1.0, 0.0, 0.0), 1.0, 1e-4, (32, 32, 32), blocks)
    # a layout whose owner ranks do not ascend with the gid would route records to the wrong rank: refused
    from tess2_b200 import multi
    layout = [(b["gid"], b["bounds_min"], b["bounds_max"]) for b in blocks]
    with pytest.raises(tess2_b200.TessB200Error):
        multi.set_layout(ctx, layout, [1, 1, 1, 1, 0, 0, 0, 0])
    with pytest.raises(tess2_b200.TessB200Error):
        multi.set_layout(ctx, layout, [0, 0, 0, 0, 0, 0, 0, -1])
    multi.set_layout(ctx, [], [])        # back to the single-rank state


def test_three_step_interface_equals_one_call(ctx, port):
    # upload -> run -> download (inputs resident, one group of all blocks) vs tessb200_dense()
    # (per-block groups pipelined against the copies): same bytes, and both equal the oracle
    blocks = dataset("clump8")
    for alg in (0, 1):
        params = ctx.make_params(alg, 0, None, None, False, (0.0, 0.0, 1.0), 1.0, 1e-4, (48, 48, 48))
        ctx.upload(blocks)
        st = ctx.run(params)
        res3 = ctx.download(params)
        res1 = run_gpu(ctx, blocks, (48, 48, 48), alg=alg)
        assert_same_bits(res3.grid, res1.grid, "three-step vs one-call")
        for d3, d1 in zip(res3.block_density, res1.block_density):
            assert_same_bits(d3, d1, "three-step vs one-call block")
        assert st.num_deposit_cells == res1.stats.num_deposit_cells and st.num_spans == res1.stats.num_spans
        compare_dense(res3, port.dense(blocks, (48, 48, 48), alg=alg), f"three-step alg{alg}")


def test_repeatability(ctx):
    # two runs on the same inputs give the same bytes (no atomics on values, ordered accumulation)
    blocks = dataset("clump8")
    a = run_gpu(ctx, blocks, (48, 48, 48))
    b = run_gpu(ctx, blocks, (48, 48, 48))
    assert_same_bits(a.grid, b.grid, "repeatability")


def test_full_size_properties(ctx, port):
    """BASELINE config 2 shape at a size the CPU can still tessellate in the test budget
    (64^3 particles in 8 regular blocks -> 128^3 grid): size-independent properties + a sampled
    comparison with the oracle on one block."""
    from tess2_b200.harness import particles, decomp, delaunay
    n = int(os.environ.get("TESSB200_TEST_N", "64"))
    dom = ([0, 0, 0], [n - 1] * 3)
    b = decomp.regular_blocks(*dom, 8)
    ps = [particles.gen_particles(g, mn, mx) for g, (mn, mx) in enumerate(b)]
    allp = np.concatenate(ps)
    owner = np.concatenate([np.full(len(q), g, np.int32) for g, q in enumerate(ps)])
    blocks = delaunay.tessellate(allp, owner, b, *dom)
    gs = (2 * n, 2 * n, 2 * n)
    res = run_gpu(ctx, blocks, gs)
    st = res.stats
    assert abs(st.tot_mass - st.num_deposit_cells) <= 1e-6 * st.num_deposit_cells
    assert st.num_deposit_cells > 0.9 * st.num_cells
    # CIC conserves all mass that lands inside the grid
    cic = run_gpu(ctx, blocks, gs, alg=1)
    assert abs(cic.stats.tot_mass - st.num_cells) <= 1e-4 * st.num_cells
    # linearity in the particle mass: exactly 2x for a power-of-two factor
    res2 = run_gpu(ctx, blocks, gs, mass=2.0)
    assert_same_bits(res2.grid, (res.grid * np.float32(2.0)).astype(np.float32), "linearity in mass")
    # the oracle on the same inputs (a few seconds per block): bit equality of the whole grid
    o = port.dense(blocks, gs)
    compare_dense(res, o, "config-2 shape")


# ---- alg 2: first-order DTFE (not in the reference; checked against the repo's own CPU statement) ----
@pytest.mark.parametrize("name,gs", [("c1", (64, 64, 64)), ("u16x8", (32, 32, 32)), ("clump8", (48, 48, 48)), ("aniso", (40, 28, 17))])
def test_dtfe_matches_cpu_statement(ctx, port, name, gs):
    blocks = dataset(name)
    o = port.dense(blocks, gs, alg=2)
    res = run_gpu(ctx, blocks, gs, alg=2)
    assert res.block_min_idx == o["block_min_idx"] and res.block_num_idx == o["block_num_idx"]
    for i, (d1, d2) in enumerate(zip(res.block_density, o["block_density"])):
        assert_same_bits(d1, d2, f"{name} dtfe block {i}")
    # three-step (resident) interface gives the same bytes
    params = ctx.make_params(2, 0, None, None, False, (0.0, 0.0, 1.0), 1.0, 1e-4, gs)
    ctx.upload(blocks)
    ctx.run(params)
    res3 = ctx.download(params)
    assert_same_bits(res3.grid, res.grid, "dtfe three-step vs one-call")


def test_dtfe_properties(ctx):
    """Properties that do not need an oracle: the DTFE field integrates to the mass of the covered
    particles (each tet contributes V/4 * sum of its vertex densities), it is linear in the particle
    mass, and it is continuous (no zero holes strictly inside the covered region)."""
    blk = dataset("c1")[0]
    gs = (64, 64, 64)
    res = run_gpu(ctx, [blk], gs, alg=2)
    g = res.grid.astype(np.float64)
    div = float(res.div)
    tot = g.sum() * div
    n_int = int((g[8:-8, 8:-8, 8:-8] > 0).sum())
    assert n_int == g[8:-8, 8:-8, 8:-8].size                      # the interior is fully covered
    assert 0.85 * blk["num_orig"] < tot < 1.05 * blk["num_orig"]  # grid quadrature of a mass-conserving field
    res2 = run_gpu(ctx, [blk], gs, alg=2, mass=2.0)
    assert_same_bits(res2.grid, (res.grid * np.float32(2.0)).astype(np.float32), "dtfe linear in mass")
    with pytest.raises(Exception):
        run_gpu(ctx, [blk], gs, alg=2, project=True)


@pytest.mark.skipif(not os.environ.get("TESSB200_BIG_TESTS"), reason="opt-in (about two minutes): set TESSB200_BIG_TESTS=1")
def test_clustered_kdtree_at_scale(ctx):
    """BASELINE configs 3-5 shape at a size the host can tessellate in a test: 128^3 clustered
    particles, kd-tree 16 blocks, 256^3 grid.  The oracle runs one process per block on the host cores
    (each with every block's bounds, so forwarded points land where the reference sends them); the
    assembled global grid must be bit-identical."""
    import multiprocessing as mp
    from tess2_b200.harness import particles, decomp, delaunay
    n, nb = int(os.environ.get("TESSB200_BIG_N", "128")), 16
    dom = ([0, 0, 0], [n - 1] * 3)
    p = particles.clustered_particles(n ** 3, *dom, seed=2027)
    bounds, owner = decomp.kdtree_blocks(p, *dom, nb)
    blocks = delaunay.tessellate(p, owner, bounds, *dom)
    for b in blocks:
        b["vert_to_tet"] = delaunay.fill_vert_to_tet(len(b["particles"]), b["tets"])
    gs = (2 * n,) * 3
    res = run_gpu(ctx, blocks, gs)
    st = res.stats
    assert abs(st.tot_mass - st.num_deposit_cells) <= 1e-6 * st.num_deposit_cells
    # oracle: the contributions of block g's cells to every block, one process per g; the chain order of
    # the global grid is (owner's own cells, then sources by gid), which a per-source split cannot
    # reproduce bit for bit, so the whole-run oracle is used on a sub-sample of blocks and the rest of
    # the grid is checked through mass conservation above
    from oracle import ref
    port = ref.Checker("port")
    o = port.dense(blocks, gs)
    compare_dense(res, o, "clustered kd-tree 16 blocks")
