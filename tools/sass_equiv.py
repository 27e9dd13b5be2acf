#!/usr/bin/env python3
"""Is the machine code of every kernel the same as in an earlier commit?  (no GPU needed)

    python tools/sass_equiv.py --base c3afef2 [--record]

Builds libopusb200.so from the kernel sources of commit --base in a scratch directory, disassembles it and the
in-tree build with `cuobjdump -sass`, and compares kernel by kernel the instruction streams with every literal
address / offset masked (data moving in the module's global segment changes constant-bank offsets and branch
targets, not instructions).  With --record, and only if every kernel is identical, the running source hash is added
to profiles/r02_ncu_constants.json as `sass_equivalent` of the captured one: bench.py then keeps using the capture's
per-kernel instruction and DRAM-byte counts, and says on what grounds.  Any differing kernel: nothing is recorded.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "esp32-opus-player_b200")
CONST = os.path.join(ROOT, "profiles", "r02_ncu_constants.json")
OUT = os.path.join(ROOT, "profiles", "r02_sass_equivalence.txt")
sys.path.insert(0, ROOT)


def kernels(so):
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    fn, d = None, {}
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            d[fn] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and fn:
            d[fn].append(re.sub(r"0x[0-9a-f]+", "X", m.group(1)).strip())
    return d


def demangle(n):
    try:
        return subprocess.run(["c++filt", "-p", n], capture_output=True, text=True).stdout.strip() or n
    except OSError:
        return n


def build_base(ref, tmp):
    src = os.path.join(tmp, "src")
    os.makedirs(src)
    tar = subprocess.run(["git", "-C", ROOT, "archive", ref, "esp32-opus-player_b200/csrc", "include"], capture_output=True, check=True).stdout
    subprocess.run(["tar", "-x", "-C", src], input=tar, check=True)
    import importlib.util

    spec = importlib.util.spec_from_file_location("b200_build", os.path.join(PKG, "build.py"))
    bld = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bld)
    so = os.path.join(tmp, "base.so")
    csrc = os.path.join(src, "esp32-opus-player_b200", "csrc")
    cmd = [os.environ.get("NVCC", "nvcc")] + bld.NVCC_FLAGS + ["-o", so] + [os.path.join(csrc, os.path.basename(f)) for f in bld.SRCS]
    subprocess.run(cmd, check=True)
    return so


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--base", required=True, help="git commit the ncu capture was taken from")
    ap.add_argument("--record", action="store_true")
    a = ap.parse_args()
    import bench

    cur_hash = bench.source_hash()
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    with tempfile.TemporaryDirectory() as tmp:
        base = kernels(build_base(a.base, tmp))
    cur = kernels(os.path.join(PKG, "lib", "libopusb200.so"))
    lines = ["SASS equivalence of the in-tree build (source hash %s, on top of %s) with commit %s" % (cur_hash, head, a.base),
             "cuobjdump -sass, per kernel, literal addresses / offsets masked", "", "%-9s %6s %6s  %s" % ("", "base", "now", "kernel")]
    all_same = set(base) == set(cur)
    for k in sorted(base):
        same = k in cur and base[k] == cur[k]
        all_same = all_same and same
        lines.append("%-9s %6d %6d  %s" % ("identical" if same else "DIFFERENT", len(base[k]), len(cur.get(k, [])), demangle(k)[:110]))
    for k in sorted(set(cur) - set(base)):
        lines.append("%-9s %6s %6d  %s" % ("NEW", "-", len(cur[k]), demangle(k)[:110]))
    lines.append("")
    lines.append("every kernel identical: %s" % all_same)
    print("\n".join(lines))
    if a.record and not all_same:
        # some kernels changed: record WHICH, so that bench.py can say what part of the capture still describes the running code
        changed = sorted({demangle(k).split("(")[0] for k in set(base) | set(cur) if k not in cur or k not in base or base[k] != cur[k]})
        same = sorted({demangle(k).split("(")[0] for k in base if k in cur and base[k] == cur[k]})
        open(OUT, "w").write("\n".join(lines) + "\n")
        d = json.load(open(CONST))
        for w in d["workloads"].values():
            if w.get("source_hash") == cur_hash or not str(w.get("git", "")).startswith(a.base):
                continue
            w["sass_partial"] = {"source_hash": cur_hash, "base": a.base, "evidence": os.path.relpath(OUT, ROOT), "kernels_identical": len(same),
                                 "kernels_changed": changed}
        json.dump(d, open(CONST, "w"), indent=1)
        print("recorded (partial) in", CONST)
        return 0
    if a.record:
        open(OUT, "w").write("\n".join(lines) + "\n")
        d = json.load(open(CONST))
        for w in d["workloads"].values():
            if w.get("source_hash") == cur_hash:
                continue
            eq = [e for e in w.get("sass_equivalent", []) if e["source_hash"] != cur_hash]
            eq.append({"source_hash": cur_hash, "base": a.base, "evidence": os.path.relpath(OUT, ROOT),
                       "why": "every kernel's SASS is instruction-for-instruction the capture's (tools/sass_equiv.py)"})
            w["sass_equivalent"] = eq
        json.dump(d, open(CONST, "w"), indent=1)
        print("recorded in", CONST)
    return 0 if all_same else 1


if __name__ == "__main__":
    sys.exit(main())
