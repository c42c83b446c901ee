"""Host side of the tessellation stage: the C++ tess() driver and its Delaunay
engine (include/tess_b200_host.h).  CPU tests: the engine against SciPy's Qhull ('Qt', the options
of src/tess-qhull.c:46) -- for points in general position the Delaunay triangulation is unique, so
the tets must be the same SETS; the driver against the Python/SciPy harness block by block; the
dense oracle on the driver's blocks against the oracle on the harness's blocks (same triangulation,
different tet numbering: the fp32 tolerance north_star states)."""
import numpy as np
import pytest

from tess2_b200 import host_tess
from tess2_b200.harness import particles, decomp, delaunay
from conftest import assert_same_bits


def tet_set(tets):
    return set(map(tuple, np.sort(np.asarray(tets)[:, :4], axis=1)))


def check_adjacency(tets, limit=4000):
    v, nb = tets[:, :4], tets[:, 4:]
    for t in range(min(len(tets), limit)):
        for i in range(4):
            u = nb[t, i]
            if u < 0:
                continue
            face = set(v[t]) - {v[t, i]}
            assert face <= set(v[u])
            j = [k for k in range(4) if v[u, k] not in face]
            assert len(j) == 1 and nb[u, j[0]] == t


def signed_volumes(p, t):
    a, b, c, d = [p[t[:, i]].astype(np.float64) for i in range(4)]
    return np.einsum("ij,ij->i", np.cross(a - d, b - d), c - d) / 6.0


@pytest.mark.parametrize("kind,n", [("uniform", 50), ("uniform", 3000), ("clustered", 3000), ("gen_particles", 4096)])
def test_engine_matches_qhull(kind, n):
    from scipy.spatial import Delaunay
    dom = ([0, 0, 0], [15, 15, 15])
    if kind == "uniform":
        p = particles.uniform_particles(n, *dom, seed=n)
    elif kind == "clustered":
        p = particles.clustered_particles(n, *dom, seed=n)
    else:
        p = particles.gen_particles(0, dom[0], dom[1])
    t = host_tess.delaunay(p)
    q = Delaunay(p.astype(np.float64), qhull_options="Qt")
    assert tet_set(t) == tet_set(q.simplices)
    check_adjacency(t)
    assert (signed_volumes(p, t[:, :4]) > 0).all()      # consistently oriented, no flat tets


def test_engine_degenerate_inputs():
    # a lattice (every cube cospherical): any triangulation is acceptable, it must be one
    g = np.stack(np.meshgrid(*[np.arange(5)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    t = host_tess.delaunay(g)
    v = signed_volumes(g, t[:, :4])
    assert (v > 0).all() and abs(v.sum() - 64.0) < 1e-9 and len(np.unique(t[:, :4])) == len(g)
    check_adjacency(t)
    # exact duplicates are left out (Qhull drops them too)
    p = np.random.default_rng(0).random((300, 3)).astype(np.float32)
    t2 = host_tess.delaunay(np.concatenate([p, p]))
    assert tet_set(t2[:, :4] % len(p)) == tet_set(host_tess.delaunay(p))     # either copy of a point may be the one kept
    # nothing to tessellate: no tets, no error
    assert len(host_tess.delaunay(np.eye(3, dtype=np.float32))) == 0
    assert len(host_tess.delaunay(np.c_[np.random.default_rng(1).random((40, 2)), np.zeros(40)].astype(np.float32))) == 0
    # coordinates Qhull refuses (extreme scales) still triangulate: exact predicates
    for scale in (1e-20, np.array([1e-6, 1.0, 1e6])):
        q = (np.random.default_rng(2).random((500, 3)) * scale).astype(np.float32)
        t3 = host_tess.delaunay(q)
        assert len(np.unique(t3[:, :4])) == len(q) and (signed_volumes(q, t3[:, :4]) > 0).all()


@pytest.mark.parametrize("nblocks,kind", [(8, "regular"), (4, "kdtree")])
def test_driver_matches_python_harness(nblocks, kind):
    dom = ([0, 0, 0], [23, 23, 23])
    if kind == "regular":
        p = particles.uniform_particles(24 ** 3 // 2, *dom, seed=4)
        b = decomp.regular_blocks(*dom, nblocks)
        own = decomp.assign_regular(p, b)
    else:
        p = particles.clustered_particles(24 ** 3 // 2, *dom, seed=4)
        b, own = decomp.kdtree_blocks(p, *dom, nblocks)
    ref = delaunay.tessellate(p, own, b, *dom, workers=1)
    mine = host_tess.tess(p, own, b, *dom, threads=2)
    for r, m in zip(ref, mine):
        assert r["gid"] == m["gid"] and r["num_orig"] == m["num_orig"]
        assert np.array_equal(r["particles"], m["particles"])          # same ghosts, same order
        assert tet_set(r["tets"]) == tet_set(m["tets"])
        assert np.array_equal(m["vert_to_tet"], delaunay.fill_vert_to_tet(len(m["particles"]), m["tets"]))
        assert np.array_equal(p[m["global_ids"]], m["particles"])
    if kind == "regular":
        # ownership by containment gives the same blocks
        mine2 = host_tess.tess(p, None, b, *dom, threads=1)
        assert all(np.array_equal(a["global_ids"], c["global_ids"]) for a, c in zip(mine, mine2))


def test_dense_on_driver_blocks_agrees_with_qhull_blocks(port):
    # same Delaunay triangulation, different tet numbering and vertex order: the grids agree to the
    # fp32 tolerance north_star states (1e-5 relative) and deposit the same mass
    dom = ([0, 0, 0], [15, 15, 15])
    p = particles.uniform_particles(16 ** 3, *dom, seed=8)
    b = decomp.regular_blocks(*dom, 2)
    own = decomp.assign_regular(p, b)
    ref = delaunay.tessellate(p, own, b, *dom, workers=1)
    for r in ref:
        r["vert_to_tet"] = delaunay.fill_vert_to_tet(len(r["particles"]), r["tets"])
    mine = host_tess.tess(p, own, b, *dom, threads=1)
    g1 = port.dense(ref, (32, 32, 32))["grid"]
    g2 = port.dense(mine, (32, 32, 32))["grid"]
    scale = np.abs(g1).max()
    close = np.abs(g1 - g2) <= 1e-5 * np.maximum(np.abs(g1), 1e-3 * scale)
    assert close.mean() > 0.999                      # all but tie points (a point within eps of a cell face)
    assert abs(float(g1.astype(np.float64).sum()) - float(g2.astype(np.float64).sum())) <= 1e-5 * float(g1.astype(np.float64).sum())


def test_decompositions_match_the_python_harness():
    dom = ([0, 0, 0], [31, 47, 15])
    for nb in (1, 2, 8, 12, 64):
        for (m1, x1), (m2, x2) in zip(decomp.regular_blocks(*dom, nb), host_tess.regular_blocks(*dom, nb)):
            assert np.array_equal(m1, m2) and np.array_equal(x1, x2)
    p = particles.clustered_particles(20000, *dom, seed=3)
    for nb in (1, 2, 8, 32):
        b1, o1 = decomp.kdtree_blocks(p, *dom, nb)
        b2, o2 = host_tess.kdtree_blocks(p, *dom, nb)
        assert np.array_equal(o1, o2)
        for (m1, x1), (m2, x2) in zip(b1, b2):
            assert np.array_equal(m1, m2) and np.array_equal(x1, x2)
    with pytest.raises(RuntimeError):
        host_tess.kdtree_blocks(p, *dom, 6)


def test_wider_rounds_insert_only_the_new_ghosts():
    # a block whose first ghost margin is too narrow keeps its triangulation and inserts the additional ghosts; the
    # result must be the block a single round at the final margin produces: same particles in the same order, same
    # tets as a set (the Delaunay triangulation of points in general position is unique)
    dom = ([0, 0, 0], [31, 31, 31])
    p = particles.clustered_particles(30000, *dom, seed=77, n_clumps=6)
    bounds, owner = host_tess.kdtree_blocks(p, *dom, 8)
    multi = host_tess.tess(p, owner, bounds, *dom)
    assert max(b["rounds"] for b in multi) >= 2
    for b in multi:
        if b["rounds"] < 2:
            continue
        one = host_tess.tess(p, owner, bounds, *dom, margin0=float(b["margin"]), max_rounds=1, gids=[b["gid"]])[0]
        assert one["rounds"] == 1 and one["num_orig"] == b["num_orig"]
        assert np.array_equal(one["global_ids"], b["global_ids"])
        assert np.array_equal(one["particles"].view(np.uint32), b["particles"].view(np.uint32))
        assert tet_set(one["tets"]) == tet_set(b["tets"])
        check_adjacency(b["tets"], limit=1500)
        v2t = b["vert_to_tet"]
        used = v2t >= 0
        assert (b["tets"][v2t[used], :4] == np.nonzero(used)[0][:, None]).any(axis=1).all()


def test_engine_is_exactly_delaunay_where_qhull_is_not():
    # Coordinates far from the origin (float32 spacing 6e-5 at 1000 in a unit box): Qhull, working in double with
    # tolerances, merges facets there and returns a different, smaller set of tets.  The engine's predicates are exact, so
    # its result can be certified instead of compared: in rational arithmetic every tet is positively oriented, the
    # neighbour relation is symmetric, every point is a vertex, and no vertex across a face lies inside the tet's
    # circumsphere (local Delaunay property on every face => Delaunay).
    from fractions import Fraction as F

    def det3(a, b, c):
        return a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0])

    def orient(a, b, c, d):
        return det3(*[[p[i] - d[i] for i in range(3)] for p in (a, b, c)])

    def insphere(a, b, c, d, e):
        m = []
        for p in (a, b, c, d):
            v = [p[i] - e[i] for i in range(3)]
            m.append(v + [v[0] * v[0] + v[1] * v[1] + v[2] * v[2]])
        return sum((-1) ** j * m[0][j] * det3(*[[m[i][k] for k in range(4) if k != j] for i in range(1, 4)]) for j in range(4))

    unit = [[F(0)] * 3, [F(1), F(0), F(0)], [F(0), F(1), F(0)], [F(0), F(0), F(1)]]
    inside_sign = insphere(*unit, [F(1, 4)] * 3) * orient(*unit)        # sign of insphere * orient for a point inside
    assert inside_sign != 0
    for n, ext, off, seed in [(260, 1.0, 1000.0, 3), (300, 31.0, 31000.0, 4)]:
        dom = ([off] * 3, [off + ext] * 3)
        p = np.unique(particles.clustered_particles(n, *dom, seed=seed, n_clumps=3), axis=0)
        t = host_tess.delaunay(p)
        q = [[F(float(x)) for x in row] for row in p]
        v, nb = t[:, :4], t[:, 4:]
        assert len(np.unique(v)) == len(p)
        for i in range(len(t)):
            a, b, c, d = [q[k] for k in v[i]]
            o = orient(a, b, c, d)
            assert o > 0
            for s in range(4):
                u = nb[i, s]
                if u < 0:
                    continue
                face = set(v[i].tolist()) - {v[i, s]}
                opp = [x for x in v[u].tolist() if x not in face]
                assert len(opp) == 1 and i in nb[u].tolist()
                assert insphere(a, b, c, d, q[opp[0]]) * o * inside_sign <= 0, "a vertex inside a circumsphere"


def test_settled_flag_reports_an_unsatisfied_ghost_region():
    # the widening of the ghost region must never stop silently.  With one round and a margin far below the
    # particle spacing the circumsphere test cannot hold for the border cells: settled == False; the defaults settle.
    from tess2_b200 import host_tess
    from tess2_b200.harness import particles
    dom = (np.zeros(3, np.float32), np.full(3, 15, np.float32))
    p = particles.uniform_particles(4000, *dom, seed=3)
    bounds = host_tess.regular_blocks(*dom, 8)
    ok = host_tess.tess(p, None, bounds, *dom)
    assert all(b["settled"] for b in ok)
    starved = host_tess.tess(p, None, bounds, *dom, margin0=0.05, max_rounds=1)
    assert not any(b["settled"] for b in starved)
    # a single block that holds the whole domain needs no ghosts and is settled at once
    one = host_tess.tess(p, None, host_tess.regular_blocks(*dom, 1), *dom, margin0=0.05, max_rounds=1)
    assert one[0]["settled"] and one[0]["rounds"] == 1


@pytest.mark.parametrize("nblocks", [1, 8])
def test_periodic_domain_matches_the_delaunay_of_all_images(nblocks, port, reference):
    """`wrap` of the reference drivers (examples/tess/main.cpp:83-88; wrap_pt, src/tess.cpp:698-710): the ghosts of a block include
    the images of particles -- its own too -- shifted by whole domain extents.  Oracle: SciPy-Qhull over all 27 images of the
    particle set; every tet at an original particle of a block must be a tet of that triangulation and the other way round, no
    original stays on a block's hull, an image keeps its particle's id and sits at the particle's position minus the shift.
    The dense stage then runs on such blocks as on any others: port and unmodified reference agree bit for bit, and every cell of
    the periodic set is complete."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(3)
    dmin, dmax = np.zeros(3, np.float32), np.array([10.0, 8.0, 12.0], np.float32)
    ext = (dmax - dmin).astype(np.float32)
    p = (rng.random((500, 3)) * ext * 0.98 + dmin + 0.01 * ext).astype(np.float32)
    bounds = host_tess.regular_blocks(dmin, dmax, nblocks)
    blocks = host_tess.tess(p, None, bounds, dmin, dmax, wrap=True, max_rounds=6, max_growth=8.0)
    imgs = []
    for sz in (-1, 0, 1):
        for sy in (-1, 0, 1):
            for sx in (-1, 0, 1):
                q = p.copy()
                for d, s in enumerate((sx, sy, sz)):
                    if s:
                        q[:, d] = (q[:, d] - np.float32(s) * ext[d]).astype(np.float32)
                imgs.append(q)
    allp = np.concatenate(imgs).astype(np.float32)
    known = set(map(tuple, allp.tolist()))
    tri = Delaunay(allp.astype(np.float64), qhull_options="Qt")

    def key(pts):
        return tuple(sorted(map(tuple, pts.tolist())))

    for b in blocks:
        assert b["settled"]
        no = b["num_orig"]
        assert np.array_equal(b["particles"][:no], p[b["global_ids"][:no]])
        ghosts, gid = b["particles"][no:], b["global_ids"][no:]
        assert all(tuple(g) in known for g in ghosts.tolist())
        shift = (p[gid] - ghosts) / ext                      # whole extents (0 for a ghost that is no image)
        assert np.allclose(shift, np.round(shift), atol=1e-5) and np.abs(np.round(shift)).max() <= 1
        own = set(map(tuple, b["particles"][:no].tolist()))
        mine = set(key(b["particles"][t[:4]]) for t in b["tets"] if (t[:4] < no).any())
        want = set(key(allp[s]) for s in tri.simplices if any(tuple(x) in own for x in allp[s].tolist()))
        assert mine == want
        hull = b["tets"][(b["tets"][:, 4:] < 0).any(axis=1)]
        for t in hull:
            for k in range(4):
                if t[4 + k] < 0:
                    assert all(t[j] >= no for j in range(4) if j != k), "an original on the hull of a periodic block"
    o1 = port.dense(blocks, (24, 24, 24))
    o2 = reference.dense(blocks, (24, 24, 24))
    for a, c in zip(o1["block_density"], o2["block_density"]):
        assert_same_bits(a, c, "periodic blocks: port vs reference")
    assert all(port.complete(b["num_orig"], b["tets"], b["vert_to_tet"]).all() for b in blocks)
