#!/usr/bin/env python
"""Top source lines of an `ncu --set full --import-source on` capture by executed warp instructions.

usage: python profiles/srclines.py gpurun_out/prof_<kernel>_<tag>.ncu-rep [N]
"""
import csv, io, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, agg = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    d = dict(zip(hdr, r))
    try:
        ins = int(d["Instructions Executed"] or 0)
    except ValueError:
        continue
    if r[2] != "-":      # SASS rows carry an address; the CUDA row above them carries the line total
        continue
    thr = int(d["Thread Instructions Executed"] or 0)
    smp = int(d["# Samples"] or 0)
    k = (fname, int(r[0]))
    a = agg.setdefault(k, [0, 0, 0, r[1]])
    a[0] += ins; a[1] += thr; a[2] += smp
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[2] for a in agg.values()) or 1
print("total warp instructions %d, samples %d" % (tot, tots))
print("%-22s %6s %7s %6s %6s  %s" % ("file:line", "inst%", "Minst", "lanes", "smpl%", "source"))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-22s %6.2f %7.1f %6.1f %6.2f  %s" % ("%s:%d" % (f, ln), 100.0 * a[0] / tot, a[0] / 1e6,
                                                 a[1] / max(a[0], 1), 100.0 * a[2] / tots, a[3].strip()[:110]))
