"""Turns gpurun_out/{launches,prof_*}_<tag>.* into the committed summaries under profiles/:
  <tag>_launches.csv / _launch_summary.txt   per-launch device times of the bench command
  <tag>_ncu_<kernel>.txt                      key counters of one `ncu --set full` capture per kernel
  traffic.json                                dram__bytes_read + dram__bytes_write per launch, per kernel
Usage (in the development container): python profiles/digest.py <tag>"""
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01_final"
out = os.path.join(ROOT, "profiles")
src = os.path.join(ROOT, "gpurun_out")

shutil.copy(os.path.join(src, f"launches_{tag}.csv"), os.path.join(out, f"{tag}_launches.csv"))
txt = subprocess.run([sys.executable, os.path.join(out, "launchsum.py"), os.path.join(out, f"{tag}_launches.csv")], capture_output=True, text=True).stdout
open(os.path.join(out, f"{tag}_launch_summary.txt"), "w").write(txt)

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
traffic = {}
for f in sorted(os.listdir(src)):
    m = re.match(rf"prof_(k_\w+)_{tag}\.ncu-rep$", f)
    if not m:
        continue
    kern = m.group(1)
    raw = subprocess.run(["ncu", "-i", os.path.join(src, f), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    lines = [f"kernel: {vals[hdr.index('Kernel Name')]}"]
    rec = {}
    for i, h in enumerate(hdr):
        if h in WANT or re.search(r"average_warps_issue_stalled_(long|short|wait|barrier|branch|math|mio|lg|not_sel|no_inst)\w*_per_issue_active", h):
            lines.append(f"  {h}: {vals[i]} {units[i]}")
            rec[h] = (vals[i], units[i])

    def to_bytes(name):
        v, u = rec[name]
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    traffic[kern] = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
    lines.append(f"  dram traffic per launch: {traffic[kern] / 1e6:.1f} MB")
    open(os.path.join(out, f"{tag}_ncu_{kern}.txt"), "w").write("\n".join(lines) + "\n")
json.dump(traffic, open(os.path.join(out, "traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
