#!/usr/bin/env python3
"""Host ceiling of the many-files player, CPU only: the whole host side (demux of every file, then rounds of plan / stage /
finish and the op_read_stereo consumer) over the mock device of tests/mockdev (which "decodes" by writing a checksum per
sample: a memory-bound stand-in, so gpu_wait_s here is the mock's own time and is reported separately).

    python tools/host_player_bench.py [--files 4096] [--packets 1500] [--threads 0] [--frames-per-round 50]

Prints one JSON line: seconds per host phase and the audio-seconds per second the HOST side alone sustains -- the number
the GPU's rate (bench.py) has to be fed at."""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def load_mock():
    d = os.path.join(ROOT, "tests", "mockdev")
    so = os.path.join(d, "libplayer_mock.so")
    csrc = os.path.join(ROOT, "esp32-opus-player_b200", "csrc")
    srcs = [os.path.join(d, "mock_device.cpp"), os.path.join(csrc, "opus_player.cpp"), os.path.join(csrc, "host_api.cpp")]
    deps = srcs + [os.path.join(csrc, "ogg_demux.hpp")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wl,-Bsymbolic", "-o", so] + srcs + ["-lpthread"], check=True)
    L = ctypes.CDLL(so)
    P, I = ctypes.c_void_p, ctypes.c_int
    L.opfb_set_create.argtypes = [ctypes.POINTER(P), I, I, I, I]
    L.opfb_set_destroy.argtypes = [P]
    L.opfb_set_open_memory.argtypes = [P, I, P, ctypes.c_size_t, I]
    L.opfb_set_prepare.argtypes = [P]
    L.opfb_set_decode_round.argtypes = [P]
    L.opfb_set_drain.argtypes = [P, I]
    L.opfb_set_drain.restype = ctypes.c_int64
    L.opfb_set_stats.argtypes = [P, P, I]
    L.opfb_synth_ogg.restype = ctypes.c_size_t
    L.opfb_synth_ogg.argtypes = [P, ctypes.c_size_t, I, I, I, I, I, I, ctypes.c_uint64, I]
    return L


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=4096)
    ap.add_argument("--packets", type=int, default=1500)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--frames-per-round", type=int, default=50)
    a = ap.parse_args()
    os.environ["MOCK_DEVICE_FAST"] = "1"  # the mock device reports every frame decoded and writes no samples
    L = load_mock()
    n = L.opfb_synth_ogg(None, 0, 2, 0, 0xFC, 160, a.packets, 4, ctypes.c_uint64(1), 1)
    blobs = []
    for i in range(a.files):
        b = np.empty(n, np.uint8)
        L.opfb_synth_ogg(b.ctypes.data, n, 2, 0, 0xFC, 160, a.packets, 4, ctypes.c_uint64(0x9E3779B97F4A7C15 ^ i), 1 + i)
        blobs.append(b)
    st = ctypes.c_void_p()
    assert L.opfb_set_create(ctypes.byref(st), a.files, 0, a.frames_per_round, a.threads) == 0
    for i, b in enumerate(blobs):
        assert L.opfb_set_open_memory(st, i, b.ctypes.data, b.size, 0) == 0
    assert L.opfb_set_prepare(st) == 0
    t_dec = t_drain = 0.0
    frames = 0
    while True:
        t0 = time.perf_counter()
        more = L.opfb_set_decode_round(st)
        t_dec += time.perf_counter() - t0
        assert more >= 0
        t0 = time.perf_counter()
        frames += L.opfb_set_drain(st, 2048)
        t_drain += time.perf_counter() - t0
        if not more:
            break
    v = (ctypes.c_double * 12)()
    L.opfb_set_stats(st, v, 12)
    L.opfb_set_destroy(st)
    audio = a.files * a.packets * 0.02
    ph = dict(zip(("keep_s", "plan_s", "reset_s", "stage_s", "mock_device_s", "finish_s"), [float(x) for x in v[6:12]]))
    # stage_s includes the call into the mock device (it does its work inside decode_many_async): take that out
    host = v[0] + ph["keep_s"] + ph["plan_s"] + ph["reset_s"] + ph["finish_s"] + t_drain
    print(json.dumps({"what": "host side of the many-files player over a mock device, CPU only", "files": a.files, "packets_per_file": a.packets,
                      "threads": a.threads or os.cpu_count(), "audio_s": audio, "demux_s": float(v[0]), "rounds_s": t_dec, "drain_s": t_drain,
                      "phases_s": ph, "frames_read_ok": bool(frames == a.files * a.packets * 960),
                      "note": "stage_s contains the mock device's own work (it fills the PCM inside decode_many_async)",
                      "x_realtime_demux": audio / max(float(v[0]), 1e-9), "x_realtime_drain": audio / max(t_drain, 1e-9)}))


if __name__ == "__main__":
    main()
