#!/usr/bin/env python
"""profiles/r01_summary.md from the bench lines of one capture (profiles/capture.sh <tag>, plus optional
bench_<tag>_{2,4,8}gpu.json from multi-GPU runs) and profiles/traffic.json (digest.py).
usage: python profiles/make_summary.py <tag>"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01_zz"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def line(path):
    for l in open(path, errors="replace"):
        if l.startswith('{"metric"') or l.startswith('{"impl"'):
            return json.loads(l)
    raise SystemExit(f"no JSON line in {path}")


d = line(os.path.join(G, f"bench_{tag}.json"))
ref = line(os.path.join(G, f"bench_ref_{tag}.json"))
json.dump(d, open(os.path.join(P, f"{tag}_bench.json"), "w"))
json.dump(ref, open(os.path.join(P, f"{tag}_bench_reference_arm.json"), "w"))
multi = {}
for n in (2, 4, 8):
    f = os.path.join(G, f"bench_{tag}_{n}gpu.json")
    if os.path.exists(f):
        multi[n] = line(f)
        json.dump(multi[n], open(os.path.join(P, f"{tag}_bench_{n}gpu.json"), "w"))
tr = json.load(open(os.path.join(P, "traffic.json")))
st = d["roofline"]["stages"]
L = []
L.append(f"# Round 1 — measured results (B200, `bench.py --steps 5 --warmup 3`, not under a profiler; capture `{tag}`)\n")
L.append("Workload = BASELINE.json configs[1]: 128^3 `gen_particles` (srand(gid)) in 8 regular blocks, SciPy-Qhull `Qt` tets\n"
         "(16.2 M tets, 2.40 M particles with ghosts), DENSE_TESS onto 256^3, mass 1, eps 1e-4.  Clocks during the timed\n"
         "regions: SM %d MHz of %d max, no throttle reason (%d samples).\n" % (d["clocks"]["sm_mhz"], d["clocks"]["sm_max_mhz"], d["clocks"]["samples"]))
L.append("| quantity | value |\n|---|---|")
L.append("| device-resident dense stage | **%.2f ms/step → %.3g grid points/s, %.3g tets/s, %.3g cells/s** |" % (d["ms_per_step"], d["value"], d["tets_per_sec"], d["cells_per_sec"]))
L.append("| end to end through `tessb200_dense()` (pinned host buffers, %.0f MB H2D + %.0f MB D2H per step, pipelined) | **%.2f ms/step → %.3g grid points/s** |"
         % (d["e2e"]["h2d_bytes_per_step"] / 1e6, d["e2e"]["d2h_bytes_per_step"] / 1e6, d["e2e"]["ms_per_step"], d["e2e"]["value"]))
cb = d["cpu_baseline"]
L.append("| CPU beside it: unmodified reference `dense()`, one process per block (%d processes) | %.3g grid points/s (%s) |" % (cb["cores"], cb["value"], cb["sample"]))
L.append("| `--impl reference` arm (same CPU reference, its own run) | %.3g grid points/s |" % ref["value"])
L.append("| resident / CPU, e2e / CPU | %.0fx, %.0fx |" % (d["value"] / cb["value"], d["e2e"]["value"] / cb["value"]))
oa = d["other_algs"]
L.append("| DENSE_CIC on the same inputs | %.2f ms → %.3g grid points/s |" % (oa["DENSE_CIC"]["ms_per_step"], oa["DENSE_CIC"]["grid_points_per_sec"]))
k = [x for x in oa if x.startswith("DENSE_DTFE")][0]
L.append("| DTFE first-order mode (alg 2, not in the reference) | %.2f ms → %.3g grid points/s |" % (oa[k]["ms_per_step"], oa[k]["grid_points_per_sec"]))
ht = d["host_tess"]
L.append("| host tessellation, SciPy Qhull `Qt`, %d processes | %.1f s; tess + dense end to end %.2f s |" % (ht["workers"], ht["seconds"], ht["tess_plus_dense_seconds"]))
if ht.get("native"):
    L.append("| host tessellation, the repo's C++ driver + Delaunay engine (same set of tets), %d threads | %.2f s; tess + dense end to end %.2f s |"
             % (ht["native"]["threads"], ht["native"]["seconds"], ht["native_tess_plus_dense_seconds"]))
L.append("| kernels launched in the 5 timed steps | %d |" % d["gpu_launches"])
L.append("\n## Per kernel (CUDA events inside the library; algorithmic bytes: DESIGN.md 4; DRAM traffic from `ncu --set full`)\n")
L.append("| kernel | ms/step | algorithmic GB/s | of measured HBM peak 6542.7 GB/s | DRAM traffic / launch | what bounds it (ncu, `%s_ncu_*.txt`) |\n|---|---:|---:|---:|---:|---|" % tag)
why = {"k_circumcenters": "stream + gathers (4 particles, 4 neighbour records per tet); includes the Morton sort of the cell order",
       "k_cell_bfs": "latency of the dependent walk-record loads (long scoreboard 3.9 cycles per issue, L1 hit rate 30 %) at 22 warps/SM (272 B of workspace per thread), 17/32 lanes (per-cell trip counts differ)",
       "k_cell_nbrs": "global-load latency of the candidate stream",
       "k_cell_faces": "L2 sector throughput (one 32-B walk record per step)",
       "k_cell_scan": "fp32 issue (12.75 instructions per face x 32 points; 8 of them the reference's unfused fp32 operations)",
       "sort (cub radix, 64-bit key + 64-bit payload)": "library radix sort, 5 passes over 39 key bits",
       "k_rows": "per-row latency chain (72 records per row, one record per lane)",
       "k_span_count + k_span_place": "L2 atomics / scattered 4-byte stores: 9.4 M records read twice, 17 M point updates (`k_span_count` 85 us, `k_span_place` 132 us, one host read-back)",
       "sort + k_rows of the shared grid points": "launch-bound: 39 k one-point records (0.4 % of the deposits) through the 5-pass radix sort and the ordered row kernel",
       "slow path (oversized cells)": "latency: 854 cells with stars > 52 tets (one warp each) and the per-CTA scan of oversized index boxes"}
for kname, v in st.items():
    if kname.startswith("nccl"):
        continue
    g = v["algorithmic_GBps"] or 0
    t = tr.get(kname)
    if kname == "k_span_count + k_span_place":
        t = tr.get("k_span_count", 0) + tr.get("k_span_place", 0)
    L.append("| `%s` | %.3f | %s | %s | %s | %s |" % (kname, v["ms"], ("%.0f" % g) if g else "-", ("%.3f" % (g / 6542.7)) if g else "-",
                                                    ("%.0f MB" % (t / 1e6)) if t else "-", why.get(kname, "")))
w = d["roofline"]["whole_stage"]
L.append("\nWhole stage against SURVEY 8(d)'s figure (32 T + 16 P + 4 G = %.0f MB): %.1f GB/s = %.3f of the measured HBM peak.  The stage is\n"
         "not HBM-bound: its kernels are bound by instruction issue, shared-memory / global latency at the occupancy their per-thread\n"
         "state allows, and L2 sector throughput (last column); DRAM traffic stays within 1.1-1.3x of the algorithmic bytes.\n"
         % (w["algorithmic_bytes"] / 1e6, w["achieved"], w["frac"]))
if multi:
    L.append("## Scaling (weak: one 8-block 128^3 slab per GPU, grid 256 x 256 x 256 N, NCCL span exchange)\n")
    L.append("| GPUs | ms/step | grid points/s | efficiency vs 1 GPU | e2e grid points/s | exchange ms |\n|---:|---:|---:|---:|---:|---:|")
    L.append("| 1 | %.2f | %.3g | 1.00 | %.3g | - |" % (d["ms_per_step"], d["value"], d["e2e"]["value"]))
    for n, s in sorted(multi.items()):
        L.append("| %d | %.2f | %.3g | %.2f | %.3g | %.2f |" % (n, s["ms_per_step"], s["value"], s["value"] / (n * d["value"]), s["e2e"]["value"],
                                                          s["roofline"]["stages"]["nccl span exchange"]["ms"]))
    L.append("\n(e2e: the ranks of one box share the host's PCIe / memory paths.)\n")
c3 = os.path.join(P, "r01_config3_clustered_256.json")
if os.path.exists(c3):
    c = json.loads(open(c3).readline())
    L.append("## Config 3 at full size on one GPU (`profiles/probe_clustered.py 256 8`)\n")
    L.append("256^3 = 16.8 M clustered (Gaussian-clump) particles, kd-tree 8 blocks, %.0f M tets, 512^3 grid: **%.1f ms → %.3g grid points/s, %.3g tets/s**; "
             "%d of %d cells deposit (%d of them through the CIC fallback), total mass %.2f (= depositing cells to %.1e).  Stage ms: cc %.1f, bfs %.1f, nbrs %.1f, "
             "faces %.1f, scan %.1f, sort %.1f, rows %.1f (measured with the full sort, before the deposit was split into single and shared grid points).\n" % (c["num_tets"] / 1e6, c["ms_total_device"], c["grid_points_per_sec"], c["num_tets"] / (c["ms_total_device"] * 1e-3),
                                                             c["num_deposit_cells"], c["num_cells"], c["num_cic_fallback"], c["tot_mass"],
                                                             abs(c["tot_mass"] - c["num_deposit_cells"]) / c["num_deposit_cells"],
                                                             c["ms_circumcenters"], c["ms_bfs"], c["ms_nbrs"], c["ms_faces"], c["ms_scan"], c["ms_sort"], c["ms_deposit"]))
L.append("## Parity\n\n`pytest -m gpu` on a 2-GPU box: 47 passed (bit equality with the oracle everywhere, including 2-GPU runs, the drop-in\n"
         "C++ header test, blocks from the repo's own tess driver and the plain-C example); opt-in `TESSB200_BIG_TESTS=1`: 128^3 clustered\n"
         "particles, kd-tree 16 blocks, 256^3 grid, global grid bit-identical to the oracle.\n")
qc = os.path.join(P, "r01_quickcheck.txt")
if os.path.exists(qc):
    ok = [l.split()[1] for l in open(qc) if l.startswith("OK")]
    bad = [l.split()[1] for l in open(qc) if l.startswith("FAIL")]
    L.append("After the last capture (projection with a given z range narrower than the data, index boxes that start below the grid, the\n"
             "reference drivers' command lines): `profiles/quick/run.sh` on a B200, the `dense` driver against the oracle's `dense.raw` byte for\n"
             "byte -- %d cases identical (%s)%s.  The GPU tests written after that (`tests/test_zz_*.py`) have only run on the CPU side so far.\n"
             % (len(ok), ", ".join(ok), "" if not bad else "; DIFFERENT: " + ", ".join(bad)))
open(os.path.join(P, "r01_summary.md"), "w").write("\n".join(L) + "\n")
print("\n".join(L))
