#!/usr/bin/env python3
"""Host demux stage alone, CPU only: Ogg pages -> CRC check -> packets -> time stamps / play windows of a
synthetic Ogg Opus file (30 s of 160-byte CELT packets, 10 packets per page), `--files` times.

  ours       esp32-opus-player_b200 csrc/ogg_demux.hpp + the event generator of csrc/opus_player.cpp (opfb_debug_events, which
             also copies every packet into the file's arena as the batched player does), called from --threads Python
             threads (ctypes releases the GIL; the per-call Python overhead counts against this arm)
  reference  the unmodified ogg.cpp + opusfile.cpp (opus_init_decoder, then op_fetch_and_process_page page after page with the
             buffered packets marked consumed: oracle/ref_api.cpp ref_ogg_demux_only), one forked process per core -- its
             demux state is a process-wide singleton.  oracle/ is used here as the thing timed BESIDE the product, as in
             bench.py's reference arm; it is not on any product path.

Both arms see the same packets (count and bytes are compared).  Prints one JSON line."""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=4096)
    ap.add_argument("--seconds", type=int, default=30, help="audio per file")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    a = ap.parse_args()
    import __graft_entry__ as ge

    pkg = ge.load_package()
    from oracle import ref

    R = ref.lib()
    R.ref_demux_fork.restype = ctypes.c_double
    R.ref_demux_fork.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    n_packets = a.seconds * 50
    blob = pkg.synth_ogg(2, 0xFC, 160, n_packets, per_page=10)
    L = pkg.api._plib()
    res = ctypes.c_int(0)
    cap = n_packets + 16

    def ours_one(out):
        return L.opfb_debug_events(blob.ctypes.data, blob.size, 0, out.ctypes.data, cap, ctypes.byref(ctypes.c_int(0)))

    out0 = np.zeros((cap, 6), np.int32)
    n_ev = ours_one(out0)
    ours_packets = int((out0[:n_ev, 0] == 0).sum())
    ours_bytes = int(out0[:n_ev][out0[:n_ev, 0] == 0][:, 2].sum())

    def worker(lo, hi):
        out = np.zeros((cap, 6), np.int32)
        for _ in range(lo, hi):
            ours_one(out)

    T = max(1, min(a.threads, a.files))
    th = [threading.Thread(target=worker, args=(a.files * j // T, a.files * (j + 1) // T)) for j in range(T)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    ours_s = time.perf_counter() - t0

    pk = ctypes.c_long(0)
    ref_s = R.ref_demux_fork(blob.ctypes.data, blob.size, a.files, T, ctypes.byref(pk))
    audio = a.files * a.seconds
    mb = a.files * blob.size / 1e6
    line = {"what": "Ogg Opus demux only, CPU", "files": a.files, "file_bytes": int(blob.size), "audio_s_per_file": a.seconds, "threads": T,
            "same_packets": bool(ours_packets == pk.value == n_packets), "packets_per_file": ours_packets, "packet_bytes_per_file": ours_bytes,
            "ours": {"seconds": ours_s, "MB_per_s": mb / ours_s, "files_per_s": a.files / ours_s, "x_realtime": audio / ours_s},
            "reference": {"seconds": ref_s, "MB_per_s": mb / ref_s, "files_per_s": a.files / ref_s, "x_realtime": audio / ref_s},
            "ratio": ref_s / ours_s}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
