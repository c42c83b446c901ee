"""Run one estimator (alg 0 / 1 / 2) on the bench workload with resident inputs; used under
`ncu --metrics gpu__time_duration.sum` to get a launch list for that estimator.
    python profiles/probe_alg.py <alg> [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tess2_b200  # noqa: E402
from tess2_b200.harness import workloads  # noqa: E402

alg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
H = int(os.environ.get("TESSB200_BENCH_H", "64"))
blocks, layout, dmin, dmax = workloads.uniform_regular(H, (2, 2, 2))
ctx = tess2_b200.Context(0)
params = ctx.make_params(alg, 0, None, None, False, (0, 0, 1), 1.0, 1e-4, (4 * H,) * 3)
ctx.upload(blocks)
for _ in range(steps):
    st = ctx.run(params)
print({k: getattr(st, k) for k in ("ms_total_device", "num_slow_cells", "num_kernel_launches", "tot_mass")})
