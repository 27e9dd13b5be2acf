#!/usr/bin/env python3
"""Source lines of a kernel ranked by one stall reason (ncu source page).
usage: ncu_stalls.py report.ncu-rep kernel_regex stall_long_sb [top_n]"""
import csv
import io
import subprocess
import sys

rep, kern, col = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
fname, hdr, agg = None, None, {}
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ic, isamp = hdr.index(col), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= ic or r[0] in ("Function Name",):
        continue
    try:
        key = (fname, int(r[0]))
        v, sm = int(r[ic] or 0), int(r[isamp] or 0)
    except ValueError:
        continue
    a = agg.setdefault(key, [0, 0, r[1]])
    a[0] += v
    a[1] += sm
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[1] for a in agg.values()) or 1
print("%s: %d of %d samples (%.1f%%)" % (col, tot, tots, 100.0 * tot / tots))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.2f%%  %s:%d  %s" % (100 * a[0] / tot, f, ln, a[2].strip()[:110]))
