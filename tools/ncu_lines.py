#!/usr/bin/env python3
"""Aggregate an ncu source page (cuda,sass view) by CUDA source line.
usage: ncu_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, agg = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        ie = hdr.index("Instructions Executed")
        it = hdr.index("Thread Instructions Executed")
        ism = hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= it:
        continue
    try:
        key = (fname, int(r[0]))
        e, t, s = int(r[ie] or 0), int(r[it] or 0), int(r[ism] or 0)
    except ValueError:
        continue
    a = agg.setdefault(key, [0, 0, 0, r[1]])
    a[0] += e
    a[1] += t
    a[2] += s
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[2] for a in agg.values()) or 1
print("total warp-inst %d  thread-inst %d  avg lanes %.2f  samples %d" % (tot, sum(a[1] for a in agg.values()), sum(a[1] for a in agg.values()) / tot, tots))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.2f%% inst  %5.2f%% samp  lanes %4.1f  %s:%d  %s" % (100 * a[0] / tot, 100 * a[2] / tots, a[1] / max(a[0], 1), f, ln, a[3].strip()[:90]))
