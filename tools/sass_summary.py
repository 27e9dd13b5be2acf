#!/usr/bin/env python3
"""Per-kernel SASS evidence of a built library: size, registers / stack / spills (cuobjdump -res-usage) and how often the
instructions that matter here occur (global / shared / local loads and stores, cp.async = LDGSTS, TMA bulk copy = UBLKCP,
mbarrier = SYNCS, shuffles, 64-bit multiplies ...).  usage: sass_summary.py lib.so [out.txt]"""
import collections
import re
import subprocess
import sys

so = sys.argv[1]
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
usage = {}
cur = None
for ln in res.splitlines():
    m = re.search(r"Function (\S+):", ln)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", ln)
    if m and cur:
        usage[cur] = tuple(int(x) for x in m.groups())
KEYS = ["LDG", "STG", "LDS", "STS", "LDL", "STL", "LDGSTS", "UBLKCP", "UTMALDG", "SYNCS", "SHFL", "IMAD.WIDE", "IMAD.HI", "BAR", "WARPSYNC", "ATOM", "RED", "UTCHMMA", "HMMA"]
kern, cnt = None, None
rows = []
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        if kern:
            rows.append((kern, cnt))
        kern, cnt = m.group(1), collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and kern:
        op = m.group(1)
        cnt["_total"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                cnt[k] += 1
if kern:
    rows.append((kern, cnt))
lines = ["SASS summary of %s (cuobjdump -sass / -res-usage, sm_100a)" % so.split("/")[-1],
         "%-34s %6s %4s %6s %6s | %s" % ("kernel", "instr", "regs", "stack", "smem", " ".join("%s" % k for k in KEYS))]
for k, c in sorted(rows, key=lambda r: demangle(r[0])):
    u = usage.get(k, (0, 0, 0, 0))
    name = demangle(k).replace("opusb200::", "").replace("void ", "")
    lines.append("%-34s %6d %4d %6d %6d | %s" % (name[:34], c["_total"], u[0], u[1], u[2], " ".join("%*d" % (len(x), c[x]) for x in KEYS)))
txt = "\n".join(lines) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "a").write(txt)
print(txt)
