#!/usr/bin/env python3
"""Player-level throughput: N synthetic Ogg Opus files -> host-thread demux -> batched GPU decode ->
op_read_stereo drain.  Prints one JSON line.  usage: player_bench.py [--files 4096] [--packets 500] [--threads 0]"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=4096)
    ap.add_argument("--packets", type=int, default=500)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--frames-per-round", type=int, default=50)
    ap.add_argument("--drain", type=int, default=1, help="1: read every file with op_read_stereo semantics")
    a = ap.parse_args()
    pkg = ge.load_package()
    t0 = time.perf_counter()
    blobs = [pkg.synth_ogg(2, 0xFC, 160, a.packets, per_page=4, seed=0x9E3779B97F4A7C15 ^ i, serial=1 + i) for i in range(a.files)]
    t_gen = time.perf_counter() - t0
    ogg_bytes = sum(b.size for b in blobs)
    t0 = time.perf_counter()
    st = pkg.OggOpusFileSet(blobs, frames_per_round=a.frames_per_round, n_threads=a.threads)
    t_open = time.perf_counter() - t0  # includes the demux of every file (opfb_set_prepare) and the context creation
    L = pkg.lib()
    buf = np.zeros(2048, np.int16)
    t_dec = t_drain = 0.0
    frames_out = 0
    done = np.zeros(a.files, bool)
    while True:
        t1 = time.perf_counter()
        more = st.decode_round()
        t_dec += time.perf_counter() - t1
        assert more >= 0
        if a.drain:
            t1 = time.perf_counter()
            frames_out += st.drain(2048)
            t_drain += time.perf_counter() - t1
        if not more:
            break
    stats = st.stats()
    audio = a.files * a.packets * 0.02
    out = {"files": a.files, "packets_per_file": a.packets, "ogg_MB": ogg_bytes / 1e6, "threads": a.threads or os.cpu_count(),
           "demux_s": stats["demux_s"], "demux_MBps": ogg_bytes / 1e6 / max(stats["demux_s"], 1e-9),
           "demux_x_realtime": audio / max(stats["demux_s"], 1e-9),
           "decode_rounds_s": t_dec, "drain_s": t_drain,
           "player_x_realtime": audio / max(stats["demux_s"] + t_dec + t_drain, 1e-9), "frames_read": frames_out, "expected_frames": a.files * a.packets * 960 if a.drain else 0,
           "host_phases_s": {k: stats[k] for k in ("keep_s", "plan_s", "reset_s", "stage_s", "gpu_wait_s", "finish_s")},
           "open_s_incl_demux_and_ctx": t_open, "gen_s": t_gen, "launches": stats["launches"]}
    print(json.dumps(out))
    st.close()


if __name__ == "__main__":
    main()
