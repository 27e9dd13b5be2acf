#!/usr/bin/env python3
"""Turn an ncu capture of ONE bench step into the constants bench.py reports (profiles/r02_ncu_constants.json).

  ncu_capture.py cmd --workload celt
        prints the command to run under gpurun: the plain device-only bench run (its JSON line, which carries the
        source hash, goes to gpurun_out/ncu_<workload>_plain.json) followed by the same command under
        `ncu --set full`, restricted to the decode kernels of the one timed step behind the three warm-up steps.
  ncu_capture.py parse gpurun_out/ncu_celt.ncu-rep --workload celt
        reads the report here (no GPU needed), sums instructions / DRAM bytes / device time per kernel, writes
        profiles/r02_ncu_constants.json (merging other workloads already in it) and a readable summary beside it.

Nothing in bench.py is typed in from a profile: it loads this file and says whether its source hash is the running one.
"""
import argparse
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r02_ncu_constants.json")
KERNEL_RE = "^k_(entropy|sequence|leaf|bands|imdct|postfilter|synth|deemph|silk)"
# launches of one step (decode kernels only), per workload part (see launches_per_step)
STAGES = {"entropy": ["k_entropy"], "sequence": ["k_sequence"], "leaves": ["k_leaf_count", "k_leaf_place", "k_leaf_expand"], "bands": ["k_bands"],
          "imdct": ["k_imdct_long", "k_imdct"], "postfilter": ["k_postfilter", "k_synth"], "deemph": ["k_deemph", "k_deemph_fix"],
          "silk": ["k_silk_params", "k_silk_fix", "k_silk_core", "k_silk_unmix", "k_silk_up2", "k_silk_fir"]}
METRICS = {"warp_inst": "smsp__inst_executed.sum", "lanes": "smsp__thread_inst_executed_per_inst_executed.ratio", "dram_read": "dram__bytes_read.sum",
           "dram_write": "dram__bytes_write.sum", "time_ns": "gpu__time_duration.sum", "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "registers": "launch__registers_per_thread",
           "smem_bank_conflicts": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "local_load_bytes": "smsp__inst_executed_op_local_ld.sum",
           "local_store_bytes": "smsp__inst_executed_op_local_st.sum"}
STALLS = "smsp__average_warps_issue_stalled_"


def short_name(k):
    k = k.strip()
    if k.startswith("void "):
        k = k[5:]
    k = k.split("(")[0]
    k = re.sub(r"<.*>$", "", k)
    return k.split("::")[-1]


def launches_per_step(workload):
    sys.path.insert(0, ROOT)
    import bench

    comps = bench.WORKLOADS[workload][0]
    keys = []
    for c in comps:
        if c.key not in keys:
            keys.append(c.key)
    n = 0
    for key in keys:
        # raw CELT context: 9 kernels (+ k_imdct_long for 20 ms frames); Opus-mode context: + k_imdct_long, the SILK
        # instantiation of k_entropy and 6 SILK kernels (they return at once without SILK frames)
        # 20 ms frames: + k_imdct_long and k_deemph_fix (the de-emphasis chains run as two segments)
        n += (9 + (2 if key[2] == 960 else 0)) if key[0] else 18
    return n


def cmd(args):
    n = launches_per_step(args.workload)
    base = "python bench.py --device-only --no-overlap --steps 1 --warmup 3 --workload %s" % args.workload
    if args.streams:
        base += " --streams %d" % args.streams
    tag = "ncu_" + args.workload
    print("%s > gpurun_out/%s_plain.json 2> gpurun_out/%s_plain.err && ncu --set full --clock-control none --import-source on -k regex:'%s' -s %d -c %d "
          "-f -o gpurun_out/%s %s > gpurun_out/%s.log 2>&1" % (base, tag, tag, KERNEL_RE, 3 * n, n, tag, base, tag))


def parse(args):
    out = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    ki = hdr.index("Kernel Name")
    col = {k: (hdr.index(m) if m in hdr else None) for k, m in METRICS.items()}
    stall_cols = [(h[len(STALLS):].replace("_per_issue_active.ratio", ""), i) for i, h in enumerate(hdr)
                  if h.startswith(STALLS) and h.endswith("per_issue_active.ratio") and "not_issued" not in h]
    kernels, order = {}, []

    def num(r, i):
        if i is None:
            return 0.0
        try:
            return float(r[i].replace(",", ""))
        except ValueError:
            return 0.0

    for r in rows[2:]:
        if len(r) <= ki:
            continue
        name = short_name(r[ki])
        k = kernels.setdefault(name, {"launches": 0, "warp_inst": 0.0, "thread_inst": 0.0, "dram_bytes": 0.0, "dram_read": 0.0, "dram_write": 0.0,
                                      "time_us": 0.0, "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")]})
        if name not in order:
            order.append(name)
        k["launches"] += 1
        k["warp_inst"] += num(r, col["warp_inst"])
        k["thread_inst"] += num(r, col["lanes"]) * num(r, col["warp_inst"])
        k["dram_read"] += num(r, col["dram_read"])
        k["dram_write"] += num(r, col["dram_write"])
        k["dram_bytes"] = k["dram_read"] + k["dram_write"]
        t_us = num(r, col["time_ns"]) / 1e3
        k["time_us"] += t_us
        k["smem_bank_conflicts"] = k.get("smem_bank_conflicts", 0.0) + num(r, col["smem_bank_conflicts"])
        if t_us >= k.get("_longest", -1.0):  # rates and stalls: those of the kernel's longest launch (template instantiations share a name)
            k["_longest"] = t_us
            k["registers"] = num(r, col["registers"])
            k["issue_active_pct"] = num(r, col["issue_active_pct"])
            k["warps_active_pct"] = num(r, col["warps_active_pct"])
            k["stalls_per_issue"] = {n: round(num(r, i), 3) for n, i in stall_cols if num(r, i) >= 0.05}
    for k in kernels.values():
        k.pop("_longest", None)
    plain = os.path.join(os.path.dirname(os.path.abspath(args.report)), "ncu_%s_plain.json" % args.workload)
    src_hash, frames, value = None, args.frames_per_step, None
    if os.path.exists(plain):
        for line in open(plain):
            line = line.strip()
            if line.startswith("{"):
                d = json.loads(line)
                src_hash = d.get("source_hash")
                frames = frames or d.get("frames_per_step_per_gpu")
                value = d.get("value")
    if not frames:
        raise SystemExit("frames per step unknown: pass --frames-per-step or keep %s beside the report" % plain)
    git = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    dirty = bool(subprocess.run(["git", "-C", ROOT, "status", "--porcelain", "--", "esp32-opus-player_b200/csrc"], capture_output=True, text=True).stdout.strip())
    rec = {"git": git + ("+dirty" if dirty else ""), "source_hash": src_hash, "frames_per_step": frames, "report": os.path.basename(args.report),
           "plain_run_value": value, "how": "ncu --set full --clock-control none, the decode kernels of one bench step (bench.py --device-only) behind 3 warm-up steps",
           "kernels": kernels, "stage_kernels": {s: [k for k in ks if k in kernels] for s, ks in STAGES.items()}}
    data = {"workloads": {}}
    if os.path.exists(args.out):
        data = json.load(open(args.out))
    data["workloads"][args.workload] = rec
    json.dump(data, open(args.out, "w"), indent=1)
    # readable summary
    tot_t = sum(k["time_us"] for k in kernels.values()) or 1.0
    tot_i = sum(k["warp_inst"] for k in kernels.values())
    tot_d = sum(k["dram_bytes"] for k in kernels.values())
    lines = ["ncu summary, workload %s, %d frames per step, git %s, source %s" % (args.workload, frames, rec["git"], src_hash),
             "%-16s %9s %6s %12s %6s %9s %7s %7s %5s  top stalls per issue" % ("kernel", "us", "share", "warp inst", "lanes", "DRAM MB", "issue%", "warps%", "regs")]
    for name in order:
        k = kernels[name]
        st = sorted(k.get("stalls_per_issue", {}).items(), key=lambda kv: -kv[1])[:4]
        lines.append("%-16s %9.1f %5.1f%% %12.4g %6.2f %9.1f %7.1f %7.1f %5d  %s" % (
            name, k["time_us"], 100 * k["time_us"] / tot_t, k["warp_inst"], k["thread_inst"] / max(k["warp_inst"], 1), k["dram_bytes"] / 1e6,
            k["issue_active_pct"], k["warps_active_pct"], int(k["registers"]), ", ".join("%s %.2f" % kv for kv in st)))
    lines.append("%-16s %9.1f        %12.4g %6.2f %9.1f   (per frame: %.0f warp inst, %.0f DRAM bytes)" % (
        "step", tot_t, tot_i, sum(k["thread_inst"] for k in kernels.values()) / max(tot_i, 1), tot_d / 1e6, tot_i / frames, tot_d / frames))
    txt = "\n".join(lines) + "\n"
    summ = args.summary or os.path.join(ROOT, "profiles", "r02_ncu_summary_%s.txt" % args.workload)
    open(summ, "w").write(txt)
    print(txt)


def main():
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="what", required=True)
    c = sub.add_parser("cmd")
    c.add_argument("--workload", default="celt")
    c.add_argument("--streams", type=int, default=0)
    p = sub.add_parser("parse")
    p.add_argument("report")
    p.add_argument("--workload", default="celt")
    p.add_argument("--frames-per-step", type=int, default=0)
    p.add_argument("--out", default=OUT)
    p.add_argument("--summary", default="")
    a = ap.parse_args()
    cmd(a) if a.what == "cmd" else parse(a)


if __name__ == "__main__":
    main()
