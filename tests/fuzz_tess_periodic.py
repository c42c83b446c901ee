"""Randomised periodic domains through the host tess() driver (`wrap`, tessb200_host_tess_periodic) against SciPy-Qhull over all
27 images of the particle set: every tet at an original particle of a block must be a tet of that triangulation and the other
way round.  Test infrastructure only.

    python tests/fuzz_tess_periodic.py [seed] [seconds]

40-700 particles (uniform or three clumps), boxes of extent 0.4 .. 64 at the origin, 1-8 regular or kd-tree blocks.  Round 2:
seeds 1 and 2, 540 s, 316 cases: no difference at the origin; two 5-tet differences in boxes of extent 0.5 at offsets -20 / 100,
where Qhull's double-precision tolerances merge nearly cospherical facets and stop being a yardstick (DESIGN.md section 10,
tests/test_host_tess.py::test_engine_is_exactly_delaunay_where_qhull_is_not)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tess2_b200 import host_tess
from tess2_b200.harness import particles
from scipy.spatial import Delaunay
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
rng = np.random.default_rng(seed)
t0 = time.time(); cases = 0; bad = 0; unsettled = 0
key = lambda pts: tuple(sorted(map(tuple, pts.tolist())))
while time.time() - t0 < secs:
    n = int(rng.integers(40, 700))
    off = 0.0
    ext = (np.array(rng.choice([1.0, 10.0, 64.0]) * rng.uniform(0.4, 1.0, 3))).astype(np.float32)
    dmin = np.full(3, off, np.float32); dmax = (dmin + ext).astype(np.float32)
    extf = (dmax - dmin).astype(np.float32)
    if rng.random() < 0.5:
        p = (rng.random((n, 3)) * extf * 0.98 + dmin + 0.01 * extf).astype(np.float32)
    else:
        p = particles.clustered_particles(n, dmin, dmax, seed=int(rng.integers(1 << 30)), n_clumps=3)
    p = np.unique(p, axis=0)
    nb = int(rng.choice([1, 2, 4, 8]))
    if rng.random() < 0.5:
        bounds = host_tess.regular_blocks(dmin, dmax, nb); owner = None
    else:
        bounds, owner = host_tess.kdtree_blocks(p, dmin, dmax, nb)
    blocks = host_tess.tess(p, owner, bounds, dmin, dmax, wrap=True, max_rounds=8, max_growth=30.0)
    imgs = []
    for sz in (-1, 0, 1):
        for sy in (-1, 0, 1):
            for sx in (-1, 0, 1):
                q = p.copy()
                for d, s in enumerate((sx, sy, sz)):
                    if s: q[:, d] = (q[:, d] - np.float32(s) * extf[d]).astype(np.float32)
                imgs.append(q)
    allp = np.concatenate(imgs).astype(np.float32)
    if len(np.unique(allp, axis=0)) != len(allp):
        continue
    tri = Delaunay(allp.astype(np.float64), qhull_options="Qt")
    cases += 1
    for b in blocks:
        if not b["settled"]:
            unsettled += 1
            continue
        no = b["num_orig"]
        own = set(map(tuple, b["particles"][:no].tolist()))
        mine = set(key(b["particles"][t[:4]]) for t in b["tets"] if (t[:4] < no).any())
        want = set(key(allp[s]) for s in tri.simplices if any(tuple(x) in own for x in allp[s].tolist()))
        if mine != want:
            bad += 1
            print("DIFF seed", seed, "case", cases, "n", len(p), "nb", nb, "gid", b["gid"], len(mine), len(want), len(mine ^ want), "margin", b["margin"], "ext", extf)
print(cases, "cases", bad, "differences", unsettled, "unsettled blocks")
