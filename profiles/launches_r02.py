#!/usr/bin/env python
"""ncu launch list (gpu__time_duration + dram bytes per launch, csv) of profiles/run_step.py -> per-kernel table of the LAST
step and the DRAM traffic of the stages bench.py's roofline names.

    python profiles/launches_r02.py gpurun_out/launches_c3.csv <launches per step> [profiles/traffic.json key prefix]"""
import collections
import csv
import json
import os
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, ni, vi, ui, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
launch = collections.OrderedDict()
for r in rows[1:]:
    d = launch.setdefault(int(r[ii]), {"name": re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("tb::", "")[:70]})
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    if r[ni] == "gpu__time_duration.sum":
        d["us"] = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else (v * 1e6 if u == "s" else v))
    else:
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        d[r[ni]] = v * mult
ids = sorted(launch)
per_step = int(sys.argv[2]) if len(sys.argv) > 2 else len(ids)
last = ids[-per_step:]
agg = collections.OrderedDict()
for i in last:
    d = launch[i]
    a = agg.setdefault(d["name"], [0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("us", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
totb = sum(a[2] for a in agg.values())
print(f"{len(last)} launches in the step, {tot / 1e3:.3f} ms under ncu (serialised, cold caches: compare shares), {totb / 1e9:.3f} GB of DRAM traffic")
print(f"{'kernel':70s} {'n':>4s} {'ms':>9s} {'share':>6s} {'DRAM MB':>10s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:70s} {a[0]:4d} {a[1] / 1e3:9.3f} {100 * a[1] / tot:5.1f}% {a[2] / 1e6:10.1f}")


def stage(name):
    if name.startswith(("k_circumcenters", "k_morton", "k_pack", "k_vert_to_tet", "k_fill")):
        return "K1"
    if name.startswith(("k_cell_", "k_advance")):
        return "K3a"
    if name.startswith(("k_span", "k_row", "k_point")):
        return "K3b"
    if name.startswith("k_cic"):
        return "K4"
    return "library (cub)"


st = collections.OrderedDict()
for k, a in agg.items():
    s = st.setdefault(stage(k), [0.0, 0.0])
    s[0] += a[1]
    s[1] += a[2]
print("\nby stage:")
for k, s in st.items():
    print(f"  {k:16s} {s[0] / 1e3:9.3f} ms {100 * s[0] / tot:5.1f}%  {s[1] / 1e6:10.1f} MB")
if len(sys.argv) > 3:
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
    t = json.load(open(path)) if os.path.exists(path) else {}
    for k, s in st.items():
        t[f"{sys.argv[3]}{k.split(' ')[0]}"] = s[1]
    json.dump(t, open(path, "w"), indent=1, sort_keys=True)
