"""Host side of bench.py (no GPU): the workload builders of the BASELINE configs at reduced size, the windows of the CPU
reference arm, and `--impl reference` end to end (the reference's own dense() on the host cores)."""
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT

sys.path.insert(0, ROOT)


def test_workload_names_and_specs():
    import bench
    assert "256^3 clustered" in bench.workload_name(3, 1) and "kd-tree 8 blocks" in bench.workload_name(3, 8) and "512^3" in bench.workload_name(3, 1)
    assert "512^3 clustered" in bench.workload_name(4, 8) and "1024^3" in bench.workload_name(4, 8) and "64 blocks" in bench.workload_name(4, 8)
    assert "DENSE_CIC" in bench.workload_name(5, 8)
    assert "128^3 uniform" in bench.workload_name(2, 1)
    assert "REDUCED" in bench.workload_name(3, 1, 4) and "REDUCED" not in bench.workload_name(3, 1, 1)      # a development run says so
    s = bench.clustered_spec(3, 8)
    assert s["side"] == 32 and s["g"] == 64 and s["nb"] == 8


def test_cpu_windows_cover_the_sample_once():
    import bench
    blocks = [dict(gid=g, num_orig=n) for g, n in enumerate([70000, 65536, 1000, 40000])]
    jobs = bench.cpu_jobs(blocks, 65536, "port", 16)
    assert len(jobs) == 16                        # one window per core
    for bi, b in enumerate(blocks):
        w = sorted((lo, hi) for j, lo, hi, _, _ in jobs if j == bi)
        assert w[0][0] == 0 and w[-1][1] == min(b["num_orig"], 65536)
        assert all(a[1] == c[0] for a, c in zip(w, w[1:]))      # contiguous, no cell twice
    assert len(bench.cpu_jobs(blocks, 65536, "port", 2)) == 2   # never more processes than cores


def test_clustered_workload_and_small_clone(tmp_path, monkeypatch):
    import bench
    monkeypatch.setattr(bench, "CACHE_DIR", str(tmp_path))
    bench.build_workload(3, 2, 0, scale=8)                      # rank 0 generates the particles and the decomposition once per box
    w = bench.build_workload(3, 2, 1, scale=8)                  # rank 1 of 2: blocks 4..7 of the 32^3 stand-in
    assert [b["gid"] for b in w["blocks"]] == [4, 5, 6, 7] and w["scaling"] == "strong" and w["ng"] == 0
    assert len(w["layout"]) == 8 and list(w["owner"]) == [0, 0, 0, 0, 1, 1, 1, 1]
    assert w["host"]["all_blocks_settled"] and w["host"]["blocks_tessellated_now"] == 4
    again = bench.build_workload(3, 2, 1, scale=8)             # the second arm on the same box reuses the blocks
    assert again["host"]["blocks_from_cache"] == 4 and again["host"]["tess_seconds"] > 0
    for a, b in zip(w["blocks"], again["blocks"]):
        assert np.array_equal(a["tets"], b["tets"]) and np.array_equal(a["particles"], b["particles"])
    name, blocks, layout, gs = bench.small_clone(3)
    assert len(blocks) == 8 and gs == (96, 96, 96) and sum(b["num_orig"] for b in blocks) > 100000


def test_reference_arm_line(tmp_path):
    env = dict(os.environ, TESSB200_CACHE=str(tmp_path), TESSB200_CPU_SAMPLE_CELLS="2048")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--scale", "8", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "dense_grid_points_per_sec" and line["value"] > 0
    assert line["scaling"] == "strong" and line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and "libtess_" in cb["native_so"]
