#!/usr/bin/env python3
"""bench.py -- decoded audio-seconds per second (x realtime) of the batched Opus decode path.

Workload (BASELINE.json configs[1]): 4096 concurrent CELT-only fullband stereo 20 ms streams per GPU,
160-byte packets (TOC 0xFC + 159 xorshift payload bytes, stream i seeded 0x9E3779B97F4A7C15 ^ i).
A "step" decodes `--frames` (default 50 = 1 s of audio) consecutive packets of every stream; decoder
state carries from step to step, so W+K steps are W+K consecutive seconds of every stream.

  value      kernel path, inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e        decode_many_async() through the C ABI with pinned HOST buffers: H2D packets + kernels + D2H PCM,
             next to the measured bare D2H ceiling of the box (e2e.d2h_ceiling_gbs)
  parity     every stream's PCM over all W+K steps, hashed on host threads (FNV-1a) and compared with the
             reference decoder's hashes of the same streams (the CPU leg); at N > 1 the hashes of all ranks are gathered
             by global stream id first, so sharded == unsharded == reference is checked on the GPUs
  roofline   bound = SM issue slots: algorithmic integer ops / step time vs SMs x 4 x f x 32; the HBM view of
             the dominant kernel is the secondary object roofline.hbm; instruction and DRAM-byte counts come from the ncu
             capture recorded in profiles/r02_ncu_constants.json (tools/ncu_capture.py), never typed in
  cpu_baseline  the unmodified reference (oracle/_ref) forked one worker per host core on a bounded sample

`--impl reference` times that CPU reference alone and prints the same JSON line with "impl": "reference".
Multi-GPU: one process per GPU (torchrun), streams sharded round-robin, no collective on the data path.
`--scaling strong` splits --streams over the GPUs instead of giving every GPU that many.
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "decoded audio-sec/sec (x realtime), 4096 CELT FB streams, 1/2/4/8 B200"
UNIT = "audio-s/s"
OPS_PER_FRAME = 5.35e5  # reference x86-64 instructions per CELT FB stereo 20 ms frame
NCU_CONSTANTS = os.path.join(ROOT, "profiles", "r02_ncu_constants.json")


class Comp:
    """One kind of stream: Opus packets behind a TOC (20 ms frames, player-level semantics), or raw CELT frames for
    celt_decode_with_ec (the short-frame oracle is the CELT layer).  ops: integer ops per audio-second
    relative to CELT FB stereo (the short-frame figure is not in the survey and is taken as 1.0)."""

    def __init__(self, name, toc, payload, channels, frame_size=960, ops=1.0):
        self.name, self.toc, self.payload, self.channels, self.frame_size, self.ops = name, toc, payload, channels, frame_size, ops
        self.raw = toc is None

    @property
    def key(self):  # streams of one decoder context: same output channel count, same call mode / frame size
        return (self.raw, self.channels, self.frame_size)


CELT_FB = Comp("CELT FB stereo 20 ms 160 B", 0xFC, 159, 2, ops=1.0)
SILK_WB = Comp("SILK WB mono 20 ms 60 B", 0x48, 59, 1, ops=0.4)
HYB_FB = Comp("hybrid FB stereo 20 ms 160 B", 0x7C, 159, 2, ops=1.1)
HYB_SWB = Comp("hybrid SWB mono 20 ms 100 B", 0x68, 99, 1, ops=0.6)
CELT_10 = Comp("CELT 10 ms stereo 80 B (CELT layer)", None, 80, 2, 480)
CELT_5 = Comp("CELT 5 ms stereo 40 B (CELT layer)", None, 40, 2, 240)
CELT_2_5 = Comp("CELT 2.5 ms stereo 30 B (CELT layer)", None, 30, 2, 120)
# name -> (components, default streams per GPU, description)
WORKLOADS = {
    "celt": ([CELT_FB], 4096, "configs[1]: %d concurrent CELT-only FB stereo 20 ms streams per GPU, 160 B packets"),
    "silk": ([SILK_WB], 16384, "configs[2]: %d concurrent SILK-only WB mono 20 ms streams per GPU, 60 B packets"),
    "hybrid": ([HYB_FB], 4096, "configs[3]: %d concurrent hybrid FB stereo 20 ms streams per GPU, 160 B packets"),
    "hybrid_swb": ([HYB_SWB], 8192, "configs[3]: %d concurrent hybrid SWB mono 20 ms streams per GPU, 100 B packets"),
    "celt10": ([CELT_10], 4096, "configs[3]: %d concurrent CELT 10 ms stereo streams per GPU, 80 B frames, short-block IMDCT (CELT-layer call, SURVEY F2)"),
    "celt5": ([CELT_5], 4096, "configs[3]: %d concurrent CELT 5 ms stereo streams per GPU, 40 B frames (CELT-layer call, SURVEY F2)"),
    "celt2_5": ([CELT_2_5], 4096, "configs[3]: %d concurrent CELT 2.5 ms stereo streams per GPU, 30 B frames (CELT-layer call, SURVEY F2)"),
    # configs[4] / SURVEY 8d config 5: a quarter each of configs 2 / 3 / hybrid FB / CELT 10 ms; 65 536 streams over 8 GPUs = 8192 per GPU
    "mix": ([CELT_FB, SILK_WB, HYB_FB, CELT_10], 8192, "configs[4]: %d concurrent mixed-mode streams per GPU (25%% each CELT FB stereo 160 B, SILK WB mono 60 B, "
            "hybrid FB stereo 160 B, CELT 10 ms stereo 80 B at the CELT layer; one decoder context per channel count and call mode)"),
    "mix20": ([CELT_FB, SILK_WB, HYB_FB, HYB_SWB], 8192, "%d concurrent mixed-mode 20 ms streams per GPU (25%% each CELT FB stereo, SILK WB mono, hybrid FB stereo, "
              "hybrid SWB mono; player-level semantics only)"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def source_hash():
    """Hash of the kernel sources: ties an ncu capture to the code it was taken from."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "esp32-opus-player_b200", "csrc")
    for fn in sorted(os.listdir(d)):
        if fn.endswith((".cu", ".cuh")):
            h.update(fn.encode())
            h.update(open(os.path.join(d, fn), "rb").read())
    return h.hexdigest()[:16]


def ncu_constants(workload):
    """Per-kernel instruction / DRAM-byte counts of one full bench step, written by tools/ncu_capture.py from an ncu
    capture; None if there is no record for this workload."""
    if not os.path.exists(NCU_CONSTANTS):
        return None
    d = json.load(open(NCU_CONSTANTS))
    return d.get("workloads", {}).get(workload)


def stage_traffic_view(nc, scale, stage_ms, hbm_peak):
    """Measured DRAM traffic (ncu capture) over the live device time of every stage: which stage moves how many GB/s, as a
    fraction of the measured HBM peak.  The de-emphasis is the one stage of the path that is HBM bound."""
    if not nc or not stage_ms or not hbm_peak:
        return None
    kern, sk = nc.get("kernels", {}), nc.get("stage_kernels", {})
    out = {}
    for stage, ms in stage_ms.items():
        b = sum(kern[k]["dram_bytes"] for k in sk.get(stage, []) if k in kern) * (scale or 1.0)
        if b > 0 and ms and ms > 0:
            gbs = b / (ms * 1e-3) / 1e9
            out[stage] = {"GBps": round(gbs, 1), "frac_of_hbm_peak": round(gbs / hbm_peak, 4)}
    return out or None


def ncu_provenance(nc):
    """Which capture the instruction / DRAM-byte counts come from, and whether it is of the running kernels: either the
    kernel sources hash to the capture's, or tools/sass_equiv.py recorded that every kernel's SASS is the capture's
    instruction for instruction (a change that moved data or touched host code only)."""
    if not nc:
        return None
    cur = source_hash()
    out = {"file": os.path.relpath(NCU_CONSTANTS, ROOT), "git": nc.get("git"), "source_hash": nc.get("source_hash"),
           "matches_this_source": nc.get("source_hash") == cur}
    if not out["matches_this_source"]:
        sp = nc.get("sass_partial")
        if sp and sp.get("source_hash") == cur:  # some kernels changed since the capture: say which (the others are SASS-identical)
            out.update(kernels_identical_to_capture=sp.get("kernels_identical"), kernels_changed_since_capture=sp.get("kernels_changed"),
                       evidence=sp.get("evidence"))
        for e in nc.get("sass_equivalent", []):
            if e.get("source_hash") == cur:
                out.update(matches_this_source=True, matched_by="sass_equivalent: " + e.get("why", ""), evidence=e.get("evidence"))
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).  nvidia-smi takes up to a
    second to produce its first line, so the sampler is started early (start) and only the lines read between begin() and
    end() -- the timed region -- are used; if the region was too short to catch one, the lines of the warm-up steps
    just before it count instead (the GPU is under the same load) and the result says so."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.t_load = self.t_begin = self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def load_begins(self):
        self.t_load = time.perf_counter()

    def begin(self):
        self.t_begin = time.perf_counter()

    def end(self):
        self.t_end = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def collect(t0, t1):
            sm, mx, reasons = [], [], set()
            for ts, r in self.rows:
                if ts < t0 or ts > t1:
                    continue
                try:
                    sm.append(float(r[0]))
                    mx.append(float(r[1]))
                    for k, nm in enumerate(names):
                        if r[3 + k].lower().startswith("active"):
                            reasons.add(nm)
                except Exception:
                    pass
            return sm, mx, reasons

        window = "timed region"
        sm, mx, reasons = collect(self.t_begin, self.t_end + 0.02)  # a line is read up to one period after it was sampled
        if not sm and self.t_load is not None:
            window = "warm-up steps and timed region"
            sm, mx, reasons = collect(self.t_load, self.t_end + 0.02)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


def layout(comps, S, order):
    """Component of every local stream index: contiguous groups (what the host scheduler does) or index mod n."""
    nm = len(comps)
    return [(s % nm) if order == "interleaved" else min(nm - 1, s * nm // S) for s in range(S)]


def cpu_reference(comps, mode_of, rank_world, seconds_of_audio, target_s, jobs=None, max_streams=None):
    """The reference's own CPU decoder (oracle/_ref) on all host cores over a bounded sample: for every component the
    first n GLOBAL streams of that kind (n sized for target_s of wall time in total), `seconds_of_audio` seconds each.
    Returns the timing dict plus {global stream id: hash}."""
    from oracle import ref

    world = rank_world
    cores = jobs or os.cpu_count() or 1
    S = len(mode_of)
    wall_t, audio_t, streams_t, hashes = 0.0, 0.0, 0, {}
    for ci, c in enumerate(comps):
        local = [s for s in range(S) if mode_of[s] == ci]
        gids = np.array([r + world * s for s in local for r in range(world)], np.int64)  # in (s, r) order: low ids first
        n_frames = int(round(seconds_of_audio * 48000 / c.frame_size))
        probe = min(n_frames, 250 * 960 // c.frame_size)
        t0 = time.perf_counter()
        if c.raw:
            ref.decode_synth_celt(c.frame_size, c.channels, c.payload, probe, want_pcm=False)
        else:
            ref.decode_synth_stream(c.toc, c.payload + 1, c.channels, probe, want_pcm=False)
        rate1 = probe * c.frame_size / 48000.0 / max(time.perf_counter() - t0, 1e-6)  # x realtime on one core
        n = max(cores, int(target_s / len(comps) * rate1 * cores / seconds_of_audio))
        n = min(n, gids.size, max_streams or gids.size)
        wall, hs = ref.bench_fork_ids(c.toc or 0, c.payload + (0 if c.raw else 1), c.channels, gids[:n], n_frames,
                                      frame_size=c.frame_size if c.raw else 0, jobs=cores)
        wall_t += wall
        audio_t += n * seconds_of_audio
        streams_t += n
        hashes.update({int(g): int(h) for g, h in zip(gids[:n], hs)})
    return {"value": audio_t / wall_t, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": "%d streams x %.0f s of audio (%.0f audio-s) of the same corpus, one forked reference decoder per core, %.1f s wall"
                      % (streams_t, seconds_of_audio, audio_t, wall_t), "wall_s": wall_t, "audio_s": audio_t}, hashes


def make_config(args, wl_desc, S):
    F = args.frames
    cfg = {"workload": (wl_desc % S) + ", %d frames (%.1f s) per stream per step" % (F, F * 0.02), "streams_per_gpu": S, "frames_per_step": F,
           "sharding": "stream i -> GPU i mod N, no collective"}
    if len(WORKLOADS[args.workload][0]) > 1:
        cfg["mix_order"] = args.mix_order
    if args.no_overlap:
        cfg["device_leg"] = "dm_decode_device, calls not overlapped"
    cfg["l2"] = "no flush needed: every step decodes new packets, and a step's inputs plus intermediates exceed the 126 MB L2 many times over"
    return cfg


def run_reference(args, wl_desc, S):
    """--impl reference: the reference CPU decoder alone, same metric / config; a step = a bounded sample of the workload."""
    comps = WORKLOADS[args.workload][0]
    mode_of = layout(comps, S, args.mix_order)
    secs = args.frames * 0.02 * 10  # 10 steps' worth of every sampled stream per call: decoder set-up is amortised as in a long run
    for _ in range(args.warmup):
        cpu_reference(comps, mode_of, 1, secs, 1.0)
    t_total, a_total, last = 0.0, 0.0, None
    for _ in range(args.steps):
        last, _ = cpu_reference(comps, mode_of, 1, secs, 3.0)
        t_total += last["wall_s"]
        a_total += last["audio_s"]
    v = a_total / t_total
    cfg = make_config(args, wl_desc, S)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_total / max(args.steps, 1), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "int16/int32 fixed-point", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": last["cores"], "kind": "reference",
                             "sample": "each step: " + last["sample"] + " -- a sample of the config's streams (throughput of independent streams)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="celt", choices=sorted(WORKLOADS), help="celt = BASELINE.json's metric config (default)")
    ap.add_argument("--streams", type=int, default=0, help="streams per GPU (default: the workload's); with --scaling strong: streams in total")
    ap.add_argument("--frames", type=int, default=50, help="20 ms frames per stream per step (50 = 1 s of audio; short-frame streams get the same duration)")
    ap.add_argument("--payload", type=int, default=0, help="payload bytes per packet behind the TOC (single-component workloads; SURVEY 8d length sensitivity)")
    ap.add_argument("--mix-order", default="grouped", choices=["grouped", "interleaved"],
                    help="mixed workloads: streams of one kind contiguous in the decoder context (what the host scheduler does) or kind = index mod 4")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="device leg through dm_decode_device (one stream-ordered call after the other) instead of "
                    "dm_decode_device_async (the tail of step k beside the head of step k+1)")
    ap.add_argument("--device-only", action="store_true", help="the device leg alone (what tools/ncu_capture.py profiles)")
    args = ap.parse_args()
    global METRIC
    comps, def_streams, wl_desc = WORKLOADS[args.workload]
    if args.payload and len(comps) == 1:
        c = comps[0]
        comps = [Comp(c.name, c.toc, args.payload, c.channels, c.frame_size, c.ops)]
        wl_desc = wl_desc.rsplit(",", 1)[0] + ", %d B packets" % (args.payload + (0 if c.raw else 1))
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    total = args.streams or def_streams
    S = total // world if args.scaling == "strong" else total
    if args.workload != "celt" or (args.payload and args.payload != 159):
        METRIC = "decoded audio-sec/sec (x realtime), %s workload (not the BASELINE.json headline)" % args.workload
    if args.device_only:
        args.no_e2e = args.no_parity = args.no_cpu_baseline = True
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        if rank == 0:
            run_reference(args, wl_desc, S)
        return

    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    pkg = ge.load_package()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the decode path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    D = dist if world > 1 else None
    F, K, W = args.frames, args.steps, args.warmup
    n_steps = W + K
    pkg.lib()
    sampler = ClockSampler(local_rank)
    sampler.start()  # early: its first line takes a while, the corpus below takes longer

    # ---- synthetic corpus: global stream g = rank + world * s (round-robin shard), n_steps seconds*F/50 of audio each.
    # A decoder context has ONE output channel count and one call mode, so a mixed workload becomes one part per
    # (channel count, mode) -- what the player layer does with mono and stereo files (opus_player.cpp RoundScheduler).
    mode_of = layout(comps, S, args.mix_order)
    ops_per_step = sum(OPS_PER_FRAME * F * comps[m].ops for m in mode_of)  # algorithmic integer ops of one step on this GPU
    audio_per_step = S * F * 0.02

    class Part:
        """The streams of one decoder context: pinned packets per step, device copies, a decoder."""

        def __init__(self, key, ids):
            self.raw, self.cc, self.fs = key
            self.ids, self.n = ids, len(ids)
            self.F = F * 960 // self.fs  # frames per step: the same audio duration whatever the frame size
            cs = [comps[mode_of[s]] for s in ids]
            self.pay = np.array([c.payload for c in cs], np.int64)
            self.toc = np.array([c.toc or 0 for c in cs], np.uint8)
            base = np.concatenate([[0], np.cumsum(self.pay * self.F)])  # byte offset of every stream's packets inside one step
            self.nb_step = int(base[-1])
            self.h_bytes = [pkg.api.PinnedArray(self.nb_step, np.uint8) for _ in range(n_steps)]
            self.h_frames = [pkg.api.PinnedArray((self.n, self.F), pkg.FRAME_DTYPE) for _ in range(n_steps)]
            for j, s in enumerate(ids):  # payload from the SURVEY generator seeded by the GLOBAL stream id
                pl = int(self.pay[j])
                b1, _ = pkg.api.synth_corpus(1, n_steps * self.F, pl, None if self.raw else int(self.toc[j]), first_stream=pkg.shard.global_id(s, rank, world))
                b1 = b1.reshape(n_steps, self.F * pl)
                for k in range(n_steps):
                    self.h_bytes[k].array[base[j]:base[j + 1]] = b1[k]
            for k in range(n_steps):
                fr = self.h_frames[k].array
                fr["offset"] = (base[:-1, None] + np.arange(self.F)[None, :] * self.pay[:, None]).astype(np.uint32)
                fr["len"] = self.pay[:, None].astype(np.uint16)
                fr["toc"] = self.toc[:, None]
                fr["flags"] = 0
            self.kw = dict(mode=pkg.DM_MODE_CELT_RAW if self.raw else pkg.DM_MODE_OPUS, frame_size=self.fs, coded_channels=self.cc, start_band=0)
            self.dec = self.new_decoder()
            self.d_bytes = [torch.from_numpy(self.h_bytes[k].array).cuda() for k in range(n_steps)]
            self.d_frames = [torch.from_numpy(self.h_frames[k].array.view(np.uint8).reshape(-1)).cuda() for k in range(n_steps)]
            self.d_pcm = torch.empty(self.n * self.F * self.fs * self.cc, dtype=torch.int16, device="cuda")
            self.d_status = torch.empty(self.n * self.F, dtype=torch.int32, device="cuda")
            self.pcm_bytes_step = self.n * self.F * self.fs * self.cc * 2

        def new_decoder(self):
            return pkg.BatchDecoder(self.n, self.F, channels=self.cc, device=local_rank)

        def step_device(self, k, stream, ready=None):
            """One step with device-resident packets.  Steps overlap (dm_decode_device_async: the serial tail of step k runs
            beside the parse and spectral stages of step k + 1) unless --no-overlap; the first timed step starts behind
            `ready` (the start event of the timed region), finish_device() makes the timing stream wait for the last."""
            if args.no_overlap:
                rc = self.dec.decode_device(self.d_bytes[k].data_ptr(), self.nb_step, self.d_frames[k].data_ptr(), self.F, self.d_pcm.data_ptr(),
                                            self.d_status.data_ptr(), stream, **self.kw)
            else:
                rc, _ = self.dec.decode_device_async(self.d_bytes[k].data_ptr(), self.nb_step, self.d_frames[k].data_ptr(), self.F, self.d_pcm.data_ptr(),
                                                     self.d_status.data_ptr(), ready, **self.kw)
            if rc:
                raise RuntimeError("dm_decode_device rc=%d" % rc)

        def finish_device(self, stream):
            if not args.no_overlap:
                assert self.dec.stream_wait(stream) == 0

        def outs(self):
            return (pkg.api.PinnedArray((self.n, self.F, self.fs, self.cc), np.int16).array, pkg.api.PinnedArray((self.n, self.F), np.int32).array)

    keys = []
    for m in mode_of:
        if comps[m].key not in keys:
            keys.append(comps[m].key)
    parts = [Part(key, [s for s in range(S) if comps[mode_of[s]].key == key]) for key in keys]
    nb_step = sum(p.nb_step for p in parts)
    n_frames_step = sum(p.n * p.F for p in parts)
    pcm_bytes_step = sum(p.pcm_bytes_step for p in parts)
    toc_bytes_step = sum(p.n * p.F for p in parts if not p.raw)
    algo_bytes_step = nb_step + toc_bytes_step + pcm_bytes_step  # packets in (payload + TOC) + int16 PCM out
    stream = torch.cuda.current_stream().cuda_stream

    def step_device(k, ready=None):
        for p in parts:
            p.step_device(k, stream, ready)

    def finish_device():
        for p in parts:
            p.finish_device(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident, CUDA events
    sampler.load_begins()
    for k in range(W):
        step_device(k)
    finish_device()
    barrier()
    launches0 = sum(p.dec.launch_count for p in parts)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.begin()
    ev0.record()
    for k in range(W, n_steps):
        step_device(k, ev0.cuda_event if k == W else None)  # the timed steps start behind ev0 ...
    finish_device()                                          # ... and ev1 is behind the last kernel of the last step
    ev1.record()
    barrier()
    sampler.end()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = sum(p.dec.launch_count for p in parts) - launches0
    for p in parts:
        assert (p.d_status.cpu().numpy() == p.fs).all(), "decode errors in the timed region"
    device_bytes = sum(p.dec.device_bytes for p in parts)
    ms_max = pkg.shard.max_over_ranks(ms, D, "cuda")
    value = pkg.shard.aggregate_rate(K * audio_per_step, world, ms_max * 1e-3)
    last_pcm_device = [p.d_pcm.cpu().numpy() for p in parts]
    for p in parts:
        p.dec.close()
    # per-stage times: the timed steps overlap (the tail of one beside the head of the next), where a stage has no time
    # of its own; so once more, untimed, one stream-ordered call after the other (dm_decode_device) on fresh decoders:
    # the last of W + 1 steps, CUDA events recorded inside the library on the streams the kernels run on, summed over the parts
    stage_acc = {}
    for p in parts:
        d = p.new_decoder()
        for k in range(W + 1):
            rc = d.decode_device(p.d_bytes[k].data_ptr(), p.nb_step, p.d_frames[k].data_ptr(), p.F, p.d_pcm.data_ptr(), p.d_status.data_ptr(), stream, **p.kw)
            assert rc == 0
        torch.cuda.synchronize()
        for name, v in d.stage_times().items():
            stage_acc[name] = stage_acc.get(name, 0.0) + v
        d.close()
    staged_step_ms = stage_acc.pop("pipeline", None)

    # ---- e2e: C ABI with pinned host buffers (fresh decoders so that the same packets are legal input)
    e2e = None
    if not args.no_e2e:
        # three calls in flight per context: the kernels of step k+1 overlap the PCM read-back of step k, and step k+2 is
        # already queued when the host wakes up for step k (decode_many_async/dm_wait); every step's packets come from
        # pinned host memory and its PCM lands in pinned host memory
        DEPTH = 3
        decs = [p.new_decoder() for p in parts]
        outs = [[p.outs() for _ in range(DEPTH)] for p in parts]

        def submit(k):
            t = []
            for i, p in enumerate(parts):
                o = outs[i][k % DEPTH]
                rc, ticket = decs[i].decode_many_async(p.h_bytes[k].array, p.h_frames[k].array, o[0], o[1], **p.kw)
                assert rc == 0
                t.append(ticket)
            return t

        def wait(tickets):
            for i, t in enumerate(tickets):
                assert decs[i].wait(t) == 0

        for k in range(W):  # warm-up through the same pipelined call (its staging buffers are allocated on first use)
            wait(submit(k))
        barrier()
        t0 = time.perf_counter()
        pending = []
        ok = True
        for k in range(W, n_steps):
            pending.append((k, submit(k)))
            if len(pending) == DEPTH:
                kk, tk = pending.pop(0)
                wait(tk)
                for i, p in enumerate(parts):  # the step's result is read on the host
                    st = outs[i][kk % DEPTH][1]
                    ok = ok and bool(st[0, 0] == p.fs and st[-1, -1] == p.fs)
        last_k = None
        for kk, tk in pending:
            wait(tk)
            last_k = kk
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        same = True
        for i, p in enumerate(parts):
            lp, ls = outs[i][last_k % DEPTH]
            assert ok and (ls == p.fs).all()
            # the two legs decode the same stream positions: their last step must agree bit for bit
            same = same and bool(np.array_equal(lp.reshape(-1), last_pcm_device[i]))
        dt_max = pkg.shard.max_over_ranks(dt, D, "cuda")
        kernel_ms_last = sum(d.stage_times()["pipeline"] for d in decs)  # (before the probe below re-uses the library's events)
        # the bare read-back of one step's PCM, all ranks at once: what PCIe and the host memory system allow
        barrier()
        d2h_ms = sum(decs[i].probe_d2h(outs[i][0][0], p.pcm_bytes_step, reps=5) for i, p in enumerate(parts))
        d2h_ms_max = pkg.shard.max_over_ranks(d2h_ms, D, "cuda")
        ceiling = world * pcm_bytes_step / (d2h_ms_max * 1e-3) / 1e9
        achieved_d2h = world * pcm_bytes_step * K / dt_max / 1e9
        e2e = {"value": pkg.shard.aggregate_rate(K * audio_per_step, world, dt_max), "unit": UNIT,
               "h2d_bytes_per_step": int(nb_step + n_frames_step * 8), "d2h_bytes_per_step": int(pcm_bytes_step + n_frames_step * 4),
               "ms_per_step": 1e3 * dt_max / K, "matches_device_leg": same,
               "api": "decode_many_async + dm_wait, %d calls in flight, pinned host buffers" % DEPTH,
               "kernel_ms_last_call": kernel_ms_last,
               "d2h_gbs": achieved_d2h, "d2h_ceiling_gbs": ceiling, "frac_of_ceiling": achieved_d2h / ceiling,
               "d2h_ceiling_how": "bare cudaMemcpyAsync of one step's PCM per GPU (dm_probe_d2h), %d GPU(s) at once, aggregate" % world}
        for d in decs:
            d.close()
        del outs

    # ---- parity: every stream over all W+K steps against the reference's hashes (untimed)
    parity, cpu_line = None, None
    if not args.no_parity:
        decs = [p.new_decoder() for p in parts]
        bufs = [p.outs() for p in parts]
        hashes = np.full(S, pkg.api.FNV_START, np.uint64)
        part_hashes = [np.full(p.n, pkg.api.FNV_START, np.uint64) for p in parts]
        bad_status = 0
        for k in range(n_steps):
            for i, p in enumerate(parts):
                rc, _, _ = decs[i].decode_many(p.h_bytes[k].array, p.h_frames[k].array, pcm=bufs[i][0], status=bufs[i][1], **p.kw)
                assert rc == 0
                bad_status += int((bufs[i][1] != p.fs).sum())
                pkg.api.fnv_update(part_hashes[i], bufs[i][0], p.cc, dup_mono=not p.raw)
        for i, p in enumerate(parts):
            hashes[np.array(p.ids)] = part_hashes[i]
            decs[i].close()
        by_gid = pkg.shard.gather_by_global_id(hashes, rank, world, D, "cuda")  # every rank's streams, by global id
        if rank == 0:
            secs = n_steps * F * 0.02
            cb, want = cpu_reference(comps, mode_of, world, secs, 10.0 if world == 1 and not args.no_cpu_baseline else 4.0)
            mism = [g for g, h in want.items() if int(by_gid[g]) != h]
            parity = {"streams_checked": len(want), "of_streams": int(world * S), "seconds_per_stream": secs, "mismatches": len(mism),
                      "bad_status_frames": bad_status, "ranks_covered": sorted({g % world for g in want}),
                      "how": "FNV-1a (SURVEY 8c) of every checked stream's PCM over all %d steps through decode_many(), hashed on host threads, "
                             "gathered by global stream id, == the unmodified reference decoder's hash of the same stream" % n_steps}
            if mism:
                parity["first_mismatches"] = mism[:8]
            if world == 1 and not args.no_cpu_baseline:
                cpu_line = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        hbm_peak, peak_src, sm_max = peaks()
        top = max(stage_acc, key=lambda n: stage_acc[n]) if stage_acc else None
        top_ms = stage_acc.get(top, 0.0) if top else 0.0
        sm_clock = (clocks.get("sm_mhz") or sm_max) * 1e6
        n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
        issue_peak = n_sm * 4 * sm_max * 1e6 * 32  # lane-ops/s at max clock
        step_s = ms_max * 1e-3 / K
        nc = ncu_constants(args.workload) if not args.payload and args.mix_order == "grouped" else None
        scale = (n_frames_step / nc["frames_per_step"]) if nc else None  # the capture's step size may differ (--streams / --frames)
        kern = (nc or {}).get("kernels", {})
        stage_kernels = (nc or {}).get("stage_kernels", {})

        def stage_sum(stage, field):
            ks = stage_kernels.get(stage, [])
            return sum(kern[k][field] for k in ks if k in kern) * scale if nc and ks else None

        inst_step = sum(v["warp_inst"] for v in kern.values()) * scale if nc else None
        dram_step = sum(v["dram_bytes"] for v in kern.values()) * scale if nc else None
        achieved_hbm = algo_bytes_step / (top_ms * 1e-3) / 1e9 if top_ms > 0 else None
        roof = {"bound": "issue", "achieved": ops_per_step / step_s, "peak": issue_peak, "unit": "lane-op/s", "frac": ops_per_step / step_s / issue_peak,
                "traffic": dram_step,
                "how": "algorithmic integer ops (5.35e5 per CELT FB stereo frame x the workload's relative cost, SURVEY 8d) per step / step time, "
                       "vs SMs x 4 schedulers x f_max x 32 lanes; traffic = DRAM bytes of all kernels of one step (ncu)",
                "frac_at_clock_under_load": ops_per_step / step_s / (n_sm * 4 * sm_clock * 32), "sm_clock_hz_under_load": sm_clock,
                # SURVEY 8d issue_utilisation: executed warp instructions (ncu) / (t x SMs x 4 x f_max)
                "issue_utilisation": (inst_step / step_s / (issue_peak / 32)) if inst_step else None,
                "lanes_per_instruction": (sum(v["thread_inst"] for v in kern.values()) / sum(v["warp_inst"] for v in kern.values())) if nc else None,
                "kernel": "k_" + str(top), "kernel_ms": top_ms, "stage_ms": stage_acc,
                "stage_ms_how": "one untimed step on its own (dm_decode_device); the timed steps overlap the tail of one step with the head of the next",
                "staged_step_ms": staged_step_ms,
                "hbm": {"bound": "hbm", "kernel": "k_" + str(top), "achieved": achieved_hbm, "peak": hbm_peak, "unit": "GB/s",
                        "frac": (achieved_hbm / hbm_peak) if achieved_hbm else None, "traffic": stage_sum(top, "dram_bytes"),
                        "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes_step,
                        "step_traffic_gbs": (dram_step / step_s / 1e9) if dram_step else None,
                        "step_traffic_frac_of_peak": (dram_step / step_s / 1e9 / hbm_peak) if dram_step else None},
                "ncu_constants": ncu_provenance(nc)}
        try:  # (an extra view; nothing in the line depends on it)
            roof["hbm"]["by_stage"] = stage_traffic_view(nc, scale, stage_acc, hbm_peak)
        except Exception:
            roof["hbm"]["by_stage"] = None
        cfg = make_config(args, wl_desc, S)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K,
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "int16/int32 fixed-point", "data": "synthetic",
                "config": cfg, "clocks": clocks, "roofline": roof, "gpu_launches": int(launches), "source_hash": source_hash(),
                "frames_per_step_per_gpu": int(n_frames_step), "device_bytes_per_gpu": int(device_bytes)}
        if e2e:
            line["e2e"] = e2e
        if parity:
            line["parity"] = parity
        if cpu_line:
            line["cpu_baseline"] = cpu_line
        print(json.dumps(line))
        if parity and (parity["mismatches"] or parity["bad_status_frames"]):
            raise SystemExit("PARITY FAILURE: %d of %d checked streams differ from the reference" % (parity["mismatches"], parity["streams_checked"]))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
