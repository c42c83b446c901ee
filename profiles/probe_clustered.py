"""Stage times on a clustered (Gaussian-clump) input: BASELINE configs 3-5 are clustered, and the
large-cell kernels (index boxes > 2048 points) only matter there.  Usage (GPU box):
    python profiles/probe_clustered.py [n_side=64] [nblocks=8]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tess2_b200  # noqa: E402
from tess2_b200.harness import particles, decomp, delaunay  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dom = ([0, 0, 0], [n - 1] * 3)
t0 = time.time()
p = particles.clustered_particles(n ** 3, *dom, seed=2024 + 3)
bounds, owner = decomp.kdtree_blocks(p, *dom, nb)
blocks = delaunay.tessellate(p, owner, bounds, *dom)
for b in blocks:
    b["vert_to_tet"] = delaunay.fill_vert_to_tet(len(b["particles"]), b["tets"])
print(f"tessellated {len(p)} clustered particles in {nb} kd-tree blocks: {time.time() - t0:.1f} s, "
      f"{sum(len(b['tets']) for b in blocks)} tets", file=sys.stderr)
ctx = tess2_b200.Context(0)
gs = (2 * n,) * 3
params = ctx.make_params(0, 0, None, None, False, (0, 0, 1), 1.0, 1e-4, gs)
ctx.upload(blocks)
for _ in range(3):
    st = ctx.run(params)
out = {k: getattr(st, k) for k in ("num_cells", "num_tets", "num_deposit_cells", "num_cic_fallback", "num_slow_cells", "num_spans", "num_faces",
                                   "num_incomplete", "num_outside", "ms_circumcenters", "ms_bfs", "ms_nbrs", "ms_faces", "ms_cells", "ms_scan",
                                   "ms_sort", "ms_deposit", "ms_total_device", "tot_mass")}
out["grid_points_per_sec"] = gs[0] ** 3 / (st.ms_total_device * 1e-3)
print(json.dumps(out))
