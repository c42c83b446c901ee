#!/usr/bin/env python
"""profiles/r02/summary.md from the bench lines committed under profiles/r02/ (python profiles/make_summary_r02.py)."""
import json
import os

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "r02")


def load(name):
    p = os.path.join(HERE, name)
    if not os.path.exists(p):
        return None
    return json.loads(open(p).read().strip().splitlines()[-1])


def row(tag, d, base=None):
    ms = d["ms_per_step"]
    e = d["e2e"]
    eff = f"{base / (ms * d['n_gpus']):.2f}" if base else "—"
    return (f"| {tag} | {d['n_gpus']} | {ms:.2f} | {d['value']:.3e} | {d['tets_per_sec']:.3e} | {eff} | {e['ms_per_step']:.1f} | {e['value']:.3e} | "
            f"{(d['roofline']['frac'] if 'CIC' in d['config'].get('alg', '') else d['roofline']['whole_stage']['frac']) * 100:.2f} % | {d['parity']['small_clone']['bit_identical']} | {d['parity']['mass_rel_err']:.1e} | "
            f"`{d['parity'].get('grid_sha256', '')[:12]}` |")


out = ["# Round 2 — measured on B200 (driver-independent runs of `bench.py`, lines under `profiles/r02/`)", "",
       "Device-resident `value` = grid points / (max over ranks of the device time of `tessb200_dense_run`); `e2e` = the same through",
       "`tessb200_dense()` with pinned host buffers.  Clocks 1965 / 1965 MHz, no throttle reason in any line.  `frac` = SURVEY 8(d)'s",
       "whole-stage bytes (32 T + 16 P + 4 G; DENSE_CIC: 12 P0 + 4 G) over the device time, of the measured 6558 GB/s.", "",
       "| config | GPUs | ms/step | grid points/s | tets/s | strong-scaling eff. | e2e ms | e2e grid points/s | frac of HBM | clone == oracle | mass rel. err | grid sha256 |",
       "|---|---:|---:|---:|---:|---:|---:|---:|---:|---|---:|---|"]
# config 3: 1 and 2 GPUs with the final library; 4 and 8 GPUs were measured one build earlier (before the per-CTA record allocation of
# k_span_place: 52.45 ms on one GPU, bench_c3_n1_a.json) and are compared with that build's one-GPU line
c3 = {1: load("bench_c3_n1.json"), 2: load("bench_c3_n2.json"), 4: load("bench_c3_n4.json"), 8: load("bench_c3_n8.json")}
c3_prev1 = load("bench_c3_n1_a.json")
base = c3[1]["ms_per_step"] if c3[1] else None
base_prev = c3_prev1["ms_per_step"] if c3_prev1 else base
for n in (1, 2, 4, 8):
    if c3[n]:
        out.append(row("3: 256³ clustered, kd-tree 8 blocks → 512³" + (" (previous build)" if n > 2 else ""), c3[n], base if n <= 2 else base_prev))
for tag, name in (("4: 512³ clustered, kd-tree 64 blocks → 1024³ (previous build)", "bench_c4_n8.json"), ("5: config 3's input, DENSE_CIC", "bench_c5_n1.json"),
                  ("5: config 3's input, DENSE_CIC", "bench_c5_n2.json"), ("5: config 3's input, DENSE_CIC (gather before its last four steps, DESIGN 3.5)", "bench_c5_n8.json"),
                  ("2: 128³ uniform, 8 regular blocks → 256³", "bench_c2_n1.json")):
    d = load(name)
    if d:
        out.append(row(tag, d))
dl = load("bench_c5_n1_delay3.json")
if dl:
    out += ["", f"Config 5 on one GPU with the library as committed last (`k_cic_gather`'s adds three steps behind their loads; a 5-step run, "
            f"`bench_c5_n1_delay3.json`): {dl['ms_per_step']:.2f} ms/step, grid sha256 `{dl['parity'].get('grid_sha256', '')[:12]}`."]
out += ["", "The digest of config 3 is the same at 1, 2, 4 and 8 GPUs: the NCCL span exchange reproduces the single-GPU bits.", ""]
d = c3[1]
if d:
    out += ["## Config 3 on one GPU: where the time goes", "", "| stage (CUDA events inside the library) | ms |", "|---|---:|"]
    names = {"ms_circumcenters": "K1: circumcenters + walk records + hull flags + cell order (Morton sort)", "ms_bfs": "k_cell_bfs (star BFS, box, filter)",
             "ms_nbrs": "k_cell_nbrs (neighbour dedup, face list, classification)", "ms_faces": "k_cell_faces (planes of the cells that go through k_cell_scan)",
             "ms_direct": "k_cell_direct<2,3,4> (small boxes: faces + inside test + scan, no plane storage)", "ms_slow_path": "oversized stars / boxes",
             "ms_sort": "k_span_count + k_span_place", "ms_deposit": "sort of the shared deposits + k_rows", "ms_total_device": "**whole stage**"}
    dm = d["roofline"]["device_ms"]
    for k, v in names.items():
        if k in dm:
            out.append(f"| {v} | {dm[k]:.2f} |")
    out.append(f"| k_cell_scan (the other cells) = ms_scan − ms_direct | {dm['ms_scan'] - dm['ms_direct']:.2f} |")
    out += ["", "Roofline stages (SURVEY 8(d) bytes / device time, of measured peak):", ""]
    for k, v in d["roofline"]["stages"].items():
        if v.get("ms"):
            out.append(f"* {k}: {v['ms']:.2f} ms, {v['algorithmic_GBps']:.0f} GB/s = {100 * (v.get('frac_of_peak') or 0):.2f} %")
    cb = d["cpu_baseline"]
    out += ["", f"CPU reference beside it ({cb['kind']}, {cb['cores']} cores): {cb['value']:.3e} grid points/s — {cb['sample']}.",
            f"Device-resident / CPU = {d['value'] / cb['value']:.0f}×, end to end / CPU = {d['e2e']['value'] / cb['value']:.0f}×.",
            f"Host tess of the same input: {d['host_tess'].get('tess_seconds') or 0:.1f} s on {d['host_tess'].get('threads')} threads (tess + dense end to end: "
            f"{d['host_tess'].get('tess_plus_dense_seconds') or 0:.1f} s).", ""]
if d:
    out += ["## The other estimators on config 3's resident input (one GPU, `other_algs`)", ""]
    for k, v in d.get("other_algs", {}).items():
        out.append(f"* {k}: {v.get('ms_per_step', float('nan')):.2f} ms/step")
    k2 = d["roofline"]["stages"].get("K2 k_cell_volumes on one block (60 T + 24 P), not part of dense()")
    if k2:
        out.append(f"* K2 (`tessb200_cell_volumes`, one block of {k2['tets']} tets): {k2['ms']:.2f} ms, {k2['algorithmic_GBps']:.0f} GB/s of 60 T + 24 P")
    out.append("")
d8 = c3[8]
if d8:
    c = d8["e2e"]["pinned_h2d_ceiling"]
    out += ["## End to end at 8 GPUs", "",
            f"Eight ranks copying at once get {c['per_rank_GBps']:.1f} GB/s each ({c['aggregate_GBps']:.0f} GB/s aggregate) against "
            f"{c3[1]['e2e']['pinned_h2d_ceiling']['per_rank_GBps']:.1f} GB/s for one rank alone; the H2D bytes of a step alone take "
            f"{d8['e2e']['ms_of_h2d_alone_at_ceiling']:.1f} ms of the {d8['e2e']['ms_per_step']:.1f} ms step.", ""]
for name in ("full_parity_c3.json", "full_parity_c2.json"):
    f = load(name)
    if f:
        out += [f"## Full-size parity: {f['config']}", "",
                f"`profiles/full_parity.py`: all {f['cells']} cells through the CPU oracle in {f['windows']} windows on {f['host_cores']} host cores ({f['cpu_seconds']} s) against the GPU grid: "
                f"{f['grid_points']} grid points; {f['single_window_points']} received deposits from exactly one window — **{f['single_window_bit_mismatches']} bit mismatches**; "
                f"{f['untouched_points']} untouched ({f['untouched_nonzero_on_gpu']} non-zero on the GPU); {f['multi_window_points']} touched by several windows, "
                f"{f['multi_window_beyond_1e-5']} beyond 1e-5 of the float64 sum (max rel. err {f['multi_window_max_rel_err']:.2e}).", ""]
open(os.path.join(HERE, "summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
