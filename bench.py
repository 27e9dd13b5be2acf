#!/usr/bin/env python3
"""bench.py -- decoded audio-seconds per second (x realtime) of the batched Opus decode path.

Workload (BASELINE.json configs[1]): 4096 concurrent CELT-only fullband stereo 20 ms streams per GPU,
160-byte packets (TOC 0xFC + 159 xorshift payload bytes, stream i seeded 0x9E3779B97F4A7C15 ^ i).
A "step" decodes `--frames` (default 50 = 1 s of audio) consecutive packets of every stream; decoder
state carries from step to step, so W+K steps are W+K consecutive seconds of every stream.

  value      kernel path, inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e        decode_many() through the C ABI with pinned HOST buffers: H2D packets + kernels + D2H PCM
  roofline   dominant kernel, algorithmic bytes (4000 B/frame) / measured kernel time vs measured HBM peak
             (the path is issue bound, not HBM bound: see roofline.issue and DESIGN.md)
  cpu_baseline  the unmodified reference (oracle/_ref) forked one worker per host core on a bounded sample

`--impl reference` times that CPU reference alone and prints the same JSON line with "impl": "reference".
Multi-GPU: one process per GPU (torchrun), streams sharded round-robin, no collective on the data path.
"""
import argparse
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
OPS_PER_FRAME = 5.35e5            # reference x86-64 instructions per CELT FB stereo frame
# name -> (TOC, payload bytes, channels, default streams per GPU, ops per frame relative to CELT FB stereo, description)
WORKLOADS = {
    "celt": (0xFC, 159, 2, 4096, 1.0, "configs[1]: %d concurrent CELT-only FB stereo 20 ms streams per GPU, 160 B packets"),
    "silk": (0x48, 59, 1, 16384, 0.4, "configs[2]: %d concurrent SILK-only WB mono 20 ms streams per GPU, 60 B packets"),
    "hybrid": (0x7C, 159, 2, 4096, 1.1, "configs[3]: %d concurrent hybrid FB stereo 20 ms streams per GPU, 160 B packets"),
    "hybrid_swb": (0x68, 99, 1, 8192, 0.6, "configs[3]: %d concurrent hybrid SWB mono 20 ms streams per GPU, 100 B packets"),
    # configs[4]: a quarter each of CELT FB stereo / SILK WB mono / hybrid FB stereo / hybrid SWB mono, the stereo and the mono streams in a decoder context each
    # (the reference plays 20 ms frames only, SURVEY F1, so the short-CELT quarter of SURVEY 8d config 5 is hybrid SWB here);
    # 65 536 streams over 8 GPUs = 8192 per GPU.  TOC/payload/ops below are placeholders: MIX gives the components.
    "mix": (0xFC, 159, 2, 8192, 0.775, "configs[4]: %d concurrent mixed-mode streams per GPU (25%% each CELT FB stereo 160 B, SILK WB mono 60 B, "
            "hybrid FB stereo 160 B, hybrid SWB mono 100 B; one decoder context per channel count)"),
}
MIX = [(0xFC, 159, 1.0), (0x48, 59, 0.4), (0x7C, 159, 1.1), (0x68, 99, 0.6)]  # (TOC, payload bytes, relative ops per frame)
TOC, PAYLOAD, CHANNELS = 0xFC, 159, 2
COMPONENTS = None  # mixed workload: [(toc, payload, ops_rel)]
# dram__bytes_read.sum + dram__bytes_write.sum per decoded CELT FB stereo frame, from the ncu --set full capture of one
# full bench step (204 800 frames per launch, profiles/r01_v17_ncu_summary_celt.txt); a launch moves this times its frame count
NCU_DRAM_BYTES_PER_FRAME = {"entropy": (0.688528e9 + 1.861747e9) / 204800, "bands": (1.223216e9 + 0.746794e9) / 204800,
                            "leaves": (0.193186e9 + 0.003928e9 + 0.357798e9 + 0.025795e9 + 1.342981e9 + 0.446985e9) / 204800,
                            "imdct": (0.852076e9 + 1.519578e9) / 204800, "postfilter": (1.667189e9 + 1.568961e9) / 204800,
                            "deemph": (1.626948e9 + 0.761595e9) / 204800, "sequence": (0.084038e9 + 0.014776e9) / 204800}
# smsp__inst_executed.sum of the same capture, all nine CELT kernels of one step: warp instructions per decoded frame
NCU_WARP_INST_PER_FRAME = (1.978627e9 + 0.010017e9 + 0.023592e9 + 0.761288e9 + 0.080024e9 + 2.368702e9 + 1.784372e9 + 1.414198e9
                           + 0.184346e9) / 204800


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


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


def cpu_reference(n_frames, target_s, jobs=None):
    """The reference's own CPU decoder on all host cores, bounded sample. Returns dict."""
    from oracle import ref

    cores = jobs or os.cpu_count() or 1
    parts = COMPONENTS or [(TOC, PAYLOAD, 1.0)]
    wall_t, audio_t, streams_t = 0.0, 0.0, 0
    for toc, payload, _ in parts:  # a mixed workload is timed one mode after the other, an equal number of streams each
        t0 = time.perf_counter()
        ref.decode_synth_stream(toc, payload + 1, 2 if toc & 4 else 1, 250, want_pcm=False)
        rate1 = 250 * 0.02 / max(time.perf_counter() - t0, 1e-6)  # x realtime on one core
        if len(parts) > 1:
            rate1 = 400.0  # a mixed workload: the same stream count for every mode
        per_core = max(1, int(round(target_s / len(parts) * rate1 / (n_frames * 0.02))))
        n_streams = cores * per_core
        wall, hashes = ref.bench_fork(toc, payload + 1, 2 if toc & 4 else 1, n_streams, n_frames, jobs=cores)
        wall_t += wall
        audio_t += n_streams * n_frames * 0.02
        streams_t += n_streams
    return {"value": audio_t / wall_t, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": "%d streams x %d frames (%.0f audio-s) of the same corpus, one forked reference decoder per core, %.1f s wall"
                      % (streams_t, n_frames, audio_t, wall_t), "wall_s": wall_t, "audio_s": audio_t}


def run_reference(args, rank, world):
    if rank != 0:
        return
    steps = []
    for _ in range(args.warmup):
        cpu_reference(args.frames * 10, 1.0)
    t_total, a_total, last = 0.0, 0.0, None
    for _ in range(args.steps):
        last = cpu_reference(args.frames * 10, 3.0)
        t_total += last["wall_s"]
        a_total += last["audio_s"]
        steps.append(last["wall_s"])
    v = a_total / t_total
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_total / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16/int32 fixed-point",
            "data": "synthetic", "config": {"workload": (WORKLOADS[args.workload][5] % 0).replace("0 concurrent ", "") + ", reference CPU decoder (oracle/_ref) forked per core",
                                            "frames_per_stream_per_step": args.frames * 10},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": last["cores"], "kind": "reference", "sample": last["sample"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="celt", choices=sorted(WORKLOADS), help="celt = BASELINE.json's metric config (default)")
    ap.add_argument("--streams", type=int, default=0, help="streams per GPU (default: the workload's)")
    ap.add_argument("--frames", type=int, default=50, help="frames per stream per step (50 = 1 s of audio)")
    ap.add_argument("--payload", type=int, default=0, help="payload bytes per packet behind the TOC (default: the workload's; SURVEY 8d length sensitivity)")
    ap.add_argument("--mix-order", default="grouped", choices=["grouped", "interleaved"],
                    help="mix workload: streams of one mode contiguous in the decoder context (what the host scheduler does) or mode = index mod 4")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    global TOC, PAYLOAD, CHANNELS, METRIC, COMPONENTS
    TOC, PAYLOAD, CHANNELS, def_streams, ops_rel, wl_desc = WORKLOADS[args.workload]
    COMPONENTS = MIX if args.workload == "mix" else None
    if args.payload and not COMPONENTS:
        PAYLOAD = args.payload
        wl_desc = wl_desc.rsplit(",", 1)[0] + ", %d B packets" % (PAYLOAD + 1)
    args.streams = args.streams or def_streams
    if args.workload != "celt" or (args.payload and args.payload != 159):
        METRIC = "decoded audio-sec/sec (x realtime), %s workload (not the BASELINE.json headline)" % args.workload
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
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
    S, F, K, W = args.streams, args.frames, args.steps, args.warmup
    n_steps = W + K
    lib = pkg.lib()
    sampler = ClockSampler(local_rank)
    sampler.start()  # early: its first line takes a while, the corpus below takes longer

    # ---- synthetic corpus: global stream g = rank + world * s (round-robin shard), n_steps*F frames each.
    # A decoder context has ONE output channel count (opus_decoder_init's `channels`), so a mixed workload becomes one
    # part per channel count -- what the player layer does with mono and stereo files (opus_player.cpp RoundScheduler).
    if COMPONENTS:
        nm = len(COMPONENTS)
        mode_of = [(s % nm) if args.mix_order == "interleaved" else min(nm - 1, s * nm // S) for s in range(S)]
    else:
        mode_of = [0] * S
    comps = COMPONENTS or [(TOC, PAYLOAD, ops_rel)]
    ops_rel = float(np.mean([comps[m][2] for m in mode_of]))
    ops_per_frame = OPS_PER_FRAME * ops_rel

    class Part:
        """The streams of one output channel count: pinned packets per step, device copies, a decoder context."""

        def __init__(self, cc, ids):
            self.cc, self.ids, self.n = cc, ids, len(ids)
            self.pay = np.array([comps[mode_of[s]][1] for s in ids], np.int64)
            self.toc = np.array([comps[mode_of[s]][0] for s in ids], np.uint8)
            base = np.concatenate([[0], np.cumsum(self.pay * F)])  # byte offset of every stream's F packets inside one step
            self.nb_step = int(base[-1])
            self.h_bytes = [pkg.api.PinnedArray(self.nb_step, np.uint8) for _ in range(n_steps)]
            self.h_frames = [pkg.api.PinnedArray((self.n, F), pkg.FRAME_DTYPE) for _ in range(n_steps)]
            if len(set(self.pay.tolist())) == 1 and len(set(self.toc.tolist())) == 1 and ids == list(range(S)):
                full, _ = pkg.shard.rank_corpus(pkg.api, rank, world, S, n_steps * F, int(self.pay[0]), int(self.toc[0]))
                full = full.reshape(S, n_steps, F * int(self.pay[0]))
                for k in range(n_steps):
                    self.h_bytes[k].array[:] = full[:, k, :].reshape(-1)
                del full
            else:
                for j, s in enumerate(ids):  # payload from the SURVEY generator seeded by the GLOBAL stream id
                    pl = int(self.pay[j])
                    b1, _ = pkg.api.synth_corpus(1, n_steps * F, pl, int(self.toc[j]), first_stream=pkg.shard.global_id(s, rank, world))
                    b1 = b1.reshape(n_steps, F * pl)
                    for k in range(n_steps):
                        self.h_bytes[k].array[base[j]:base[j + 1]] = b1[k]
            for k in range(n_steps):
                fr = self.h_frames[k].array
                fr["offset"] = (base[:-1, None] + np.arange(F)[None, :] * self.pay[:, None]).astype(np.uint32)
                fr["len"] = self.pay[:, None].astype(np.uint16)
                fr["toc"] = self.toc[:, None]
                fr["flags"] = 0
            self.dec = pkg.BatchDecoder(self.n, F, channels=cc, device=local_rank)
            self.d_bytes = [torch.from_numpy(self.h_bytes[k].array).cuda() for k in range(n_steps)]
            self.d_frames = [torch.from_numpy(self.h_frames[k].array.view(np.uint8).reshape(-1)).cuda() for k in range(n_steps)]
            self.d_pcm = torch.empty(self.n * F * 960 * cc, dtype=torch.int16, device="cuda")
            self.d_status = torch.empty(self.n * F, dtype=torch.int32, device="cuda")

        def step_device(self, k, stream):
            rc = self.dec.decode_device(self.d_bytes[k].data_ptr(), self.nb_step, self.d_frames[k].data_ptr(), F, self.d_pcm.data_ptr(),
                                        self.d_status.data_ptr(), stream)
            if rc:
                raise RuntimeError("dm_decode_device rc=%d" % rc)

        def outs(self):
            return (pkg.api.PinnedArray((self.n, F, 960, self.cc), np.int16).array, pkg.api.PinnedArray((self.n, F), np.int32).array)

    def out_channels(toc):
        return 2 if toc & 4 else 1

    parts = []
    for cc in (2, 1):
        ids = [s for s in range(S) if out_channels(comps[mode_of[s]][0]) == cc]
        if ids:
            parts.append(Part(cc, ids))
    nb_step = sum(p.nb_step for p in parts)
    pcm_bytes_step = sum(p.n * F * 960 * p.cc * 2 for p in parts)
    bytes_per_frame_algo = (nb_step + S * F + pcm_bytes_step) / (S * F)  # packet in (payload + TOC) + int16 PCM out
    stream = torch.cuda.current_stream().cuda_stream

    def step_device(k):
        for p in parts:
            p.step_device(k, stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident, CUDA events
    sampler.load_begins()
    for k in range(W):
        step_device(k)
    barrier()
    launches0 = sum(p.dec.launch_count for p in parts)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.begin()
    ev0.record()
    for k in range(W, n_steps):
        step_device(k)
    ev1.record()
    barrier()
    sampler.end()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = sum(p.dec.launch_count for p in parts) - launches0
    for p in parts:
        assert (p.d_status.cpu().numpy() == 960).all(), "decode errors in the timed region"
    # per-stage times of the last step (events recorded inside the library on the same stream), summed over the parts
    stage_acc = {}
    for p in parts:
        for name, v in p.dec.stage_times().items():
            stage_acc[name] = stage_acc.get(name, 0.0) + v
    device_bytes = sum(p.dec.device_bytes for p in parts)
    ms_max = pkg.shard.max_over_ranks(ms, dist if world > 1 else None, "cuda")
    audio_per_step = S * F * 0.02
    value = pkg.shard.aggregate_rate(K * audio_per_step, world, ms_max * 1e-3)

    # ---- e2e: C ABI with pinned host buffers (fresh decoders so that the same packets are legal input)
    e2e = None
    if not args.no_e2e:
        # three calls in flight per context: the kernels of step k+1 overlap the PCM read-back of step k, and step k+2 is
        # already queued when the host wakes up for step k (decode_many_async/dm_wait); every step's packets come from
        # pinned host memory and its PCM lands in pinned host memory
        DEPTH = 3
        decs = [pkg.BatchDecoder(p.n, F, channels=p.cc, device=local_rank) for p in parts]
        outs = [[p.outs() for _ in range(DEPTH)] for p in parts]

        def submit(k):
            t = []
            for i, p in enumerate(parts):
                o = outs[i][k % DEPTH]
                rc, ticket = decs[i].decode_many_async(p.h_bytes[k].array, p.h_frames[k].array, o[0], o[1])
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
                for i in range(len(parts)):  # the step's result is read on the host
                    st = outs[i][kk % DEPTH][1]
                    ok = ok and bool(st[0, 0] == 960 and st[-1, -1] == 960)
        last_k = None
        for kk, tk in pending:
            wait(tk)
            last_k = kk
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        same = True
        for i, p in enumerate(parts):
            lp, ls = outs[i][last_k % DEPTH]
            assert ok and (ls == 960).all()
            # the two legs decode the same stream positions: their last step must agree bit for bit
            same = same and bool(np.array_equal(lp.reshape(-1), p.d_pcm.cpu().numpy()))
        dt_max = pkg.shard.max_over_ranks(dt, dist if world > 1 else None, "cuda")
        e2e = {"value": pkg.shard.aggregate_rate(K * audio_per_step, world, dt_max), "unit": UNIT,
               "h2d_bytes_per_step": int(nb_step + S * F * 8), "d2h_bytes_per_step": int(pcm_bytes_step + S * F * 4),
               "ms_per_step": 1e3 * dt_max / K, "matches_device_leg": same,
               "api": "decode_many_async + dm_wait, %d calls in flight, pinned host buffers" % DEPTH,
               "kernel_ms_last_call": sum(sum(d.stage_times().values()) for d in decs)}
        for d in decs:
            d.close()

    if rank == 0:
        hbm_peak, peak_src, sm_max = peaks()
        top = max(stage_acc, key=lambda n: stage_acc[n]) if stage_acc else None
        top_ms = stage_acc.get(top, 0.0) if top else 0.0
        algo_bytes = S * F * bytes_per_frame_algo
        achieved = algo_bytes / (top_ms * 1e-3) / 1e9 if top_ms > 0 else None
        sm_clock = (clocks.get("sm_mhz") or sm_max) * 1e6
        n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
        issue_peak = n_sm * 4 * sm_max * 1e6 * 32  # lane-ops/s at max clock
        step_s = ms_max * 1e-3 / K
        roof = {"bound": "hbm", "kernel": "k_" + str(top), "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": (achieved / hbm_peak) if achieved else None,
                "traffic": (NCU_DRAM_BYTES_PER_FRAME[top] * S * F) if (args.workload == "celt" and top in NCU_DRAM_BYTES_PER_FRAME) else None,
                "traffic_source": "ncu dram bytes per frame of a full 204 800-frame launch (profiles/r01_v17_ncu_summary_celt.txt) x frames per launch", "peak_source": peak_src,
                "kernel_ms": top_ms, "stage_ms": stage_acc,
                "issue": {"note": "the path is issue bound (SURVEY 8d): algorithmic integer ops vs SM issue peak over the whole step",
                          "achieved_lane_ops_per_s": S * F * ops_per_frame / step_s, "peak_lane_ops_per_s": issue_peak,
                          "frac": S * F * ops_per_frame / step_s / issue_peak, "sm_clock_hz_under_load": sm_clock,
                          # SURVEY 8d issue_utilisation: executed warp instructions (ncu) / (t x SMs x 4 x f_max)
                          "issue_utilisation": (NCU_WARP_INST_PER_FRAME * S * F / step_s / (issue_peak / 32)) if args.workload == "celt" and not args.payload else None,
                          "issue_utilisation_source": "smsp__inst_executed.sum per frame from profiles/r01_v17_ncu_summary_celt.txt x frames per step"}}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16/int32 fixed-point", "data": "synthetic",
                "config": {"workload": (wl_desc % S) + ", %d frames (%.1f s) per stream per step" % (F, F * 0.02), "streams_per_gpu": S, "frames_per_step": F, "sharding": "stream i -> GPU i mod N, no collective",
                           **({"mix_order": args.mix_order} if COMPONENTS else {}),
                           "l2": "inputs and intermediates per step (%.1f GB) exceed L2; every step decodes new packets" % (device_bytes / 1e9)},
                "clocks": clocks, "roofline": roof, "gpu_launches": int(launches)}
        if e2e:
            line["e2e"] = e2e
        if not args.no_cpu_baseline and world == 1:  # the CPU baseline is reported at N = 1 only
            cb = cpu_reference(3000 if F * n_steps >= 3000 else F * n_steps, 10.0)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    for p in parts:
        p.dec.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
