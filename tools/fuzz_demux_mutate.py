"""Memory-safety fuzzing of the host demux stage on CORRUPTED input (the demuxer parses untrusted files): valid random Ogg
Opus layouts from tools/fuzz_demux.py with bit flips, byte runs, truncations, duplicated and re-ordered spans, and header
fields set to extremes, pushed through opfb_debug_events (page framer, packetizer, the open / link / timestamp logic and
the packet parser).  No oracle: the reference asserts or reads out of bounds on much of this.  Meant to run against an
ASan + UBSan build of the host sources:

    cd esp32-opus-player_b200 && g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fPIC -shared -o /tmp/libhost_asan.so \\
        csrc/opus_player.cpp csrc/host_api.cpp -Llib -lopusb200 -Wl,-rpath,$PWD/lib -lpthread
    OPUSB_LIB=/tmp/libhost_asan.so LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 \\
        python tools/fuzz_demux_mutate.py --cases 20000

Checks: no sanitizer report / crash, the event list is well formed (play windows inside the packet, durations of whole
20 ms frames, sizes within the Opus packet limit x 48 frames), and the walk ends.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import fuzz_demux as fd  # noqa: E402


def mutate(rng, blob):
    b = bytearray(blob)
    for _ in range(int(rng.integers(1, 6))):
        if not b:
            break
        k = int(rng.integers(0, 8))
        i = int(rng.integers(0, len(b)))
        if k == 0:
            b[i] ^= 1 << int(rng.integers(0, 8))
        elif k == 1:
            n = int(rng.integers(1, 64))
            b[i:i + n] = bytes([int(rng.choice([0, 255, int(rng.integers(0, 256))]))]) * min(n, len(b) - i)
        elif k == 2:
            del b[i:i + int(rng.integers(1, 400))]
        elif k == 3:
            j = int(rng.integers(0, len(b)))
            b[i:i] = b[j:j + int(rng.integers(1, 600))]
        elif k == 4:
            b = b[:i]
        elif k == 5:  # aim at a page header: segment count / lacing / granule / flags
            p = bytes(b).find(b"OggS", i)
            if p >= 0 and p + 27 < len(b):
                f = int(rng.choice([5, 6, 13, 14, 18, 26, 27, 28]))
                b[p + f] = int(rng.choice([0, 1, 127, 128, 254, 255]))
        elif k == 6:  # aim at OpusHead
            p = bytes(b).find(b"OpusHead")
            if p >= 0 and p + 19 < len(b):
                b[p + 8 + int(rng.integers(0, 11))] = int(rng.choice([0, 1, 2, 3, 127, 128, 255]))
        else:  # a TOC / frame-count byte of some packet
            p = bytes(b).find(b"OggS", i)
            if p >= 0 and p + 27 < len(b):
                q = p + 27 + b[p + 26]
                if q + 1 < len(b):
                    b[q] = int(rng.integers(0, 256))
                    b[q + 1] = int(rng.choice([0, 1, 48, 49, 63, 64, 128, 191, 192, 255]))
    return bytes(b)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=5000)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    import __graft_entry__ as ge

    pkg = ge.load_package()
    bad, opened, events = [], 0, 0
    for k in range(a.cases):
        seed = (a.seed << 24) + k
        rng = np.random.default_rng(seed ^ 0x5EED)
        blob, _ = fd.make_file(seed, bool(k & 1))
        data = np.frombuffer(mutate(rng, blob), np.uint8)
        if data.size == 0:
            continue
        for clean_eof in (False, True):
            res, ev = pkg.demux_events(data, clean_eof=clean_eof)
            opened += res == 0
            events += len(ev)
            for kind, code, nbytes, dur, pos, size in ev.tolist():
                if kind == 0 and not (0 <= pos <= size <= dur and dur % 960 == 0 and 0 < dur <= 5760 and 0 < nbytes <= 1275 * 48 + 100):
                    bad.append({"seed": seed, "event": [kind, code, nbytes, dur, pos, size]})
                    break
            if len(ev) and ev[-1][0] != 1 and res == 0:
                bad.append({"seed": seed, "what": "walk does not end in a return code"})
    print(json.dumps({"cases": a.cases, "opened": int(opened), "events": int(events), "malformed": bad[:10], "malformed_n": len(bad), "seed": a.seed}))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
