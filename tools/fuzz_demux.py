"""Differential fuzzing of the host demux stage (csrc/ogg_demux.hpp + the event generator of csrc/opus_player.cpp) against
the reference's own player (opus_init_decoder + op_read_stereo over an in-memory SD_read, oracle/_ref): random Ogg Opus
layouts -- links, pre-skip, end trim, packets per page, multi-frame packets (code 0/1/2/3, padding), packets spanning
pages, lacing edges (255 / 510 bytes), junk between pages, multiplexed foreign streams, corrupt last pages, missing EOS,
odd granule positions, random op_read_stereo buffer sizes.  CPU only; TEST INFRASTRUCTURE (checker only).

    python tools/fuzz_demux.py --cases 3000 [--seed 1] [--wild]

The sequence of op_read_stereo return values is fully determined by the demux stage (every packet is a valid 20 ms CELT
packet, which always decodes), so it is compared call by call with what our event list predicts.  The reference runs in
a forked child: where it dies on one of its own asserts (SURVEY F-list: holes in unseekable streams, failed opens) the
case is counted as `ref_died` and skipped.  --wild adds granule positions that run backwards / jump, where the
reference's behaviour is mostly error paths.
"""
import argparse
import json
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oggutil as og  # noqa: E402


def rbytes(rng, n):
    return rng.integers(0, 256, n, dtype=np.uint8).tobytes()


def frame_len(rng):
    k = int(rng.integers(0, 10))
    if k == 0:
        return int(rng.choice([253, 254, 255, 256, 509, 510, 511]))
    if k == 1:
        return int(rng.integers(600, 1200))
    return int(rng.integers(8, 240))


def make_packet(rng, toc):
    """-> (bytes, frames).  20 ms frames only (the reference's Opus layer cannot decode others)."""
    code = int(rng.choice([0, 0, 0, 1, 2, 3]))
    if code == 0:
        return bytes([toc]) + rbytes(rng, frame_len(rng) - 1), 1
    if code == 1:
        n = int(rng.integers(4, 200))
        return bytes([toc | 1]) + rbytes(rng, 2 * n), 2
    if code == 2:
        a, b = int(rng.integers(2, 500)), int(rng.integers(2, 300))  # frames of < 2 bytes fail in the decoder, not in the demux
        hdr = bytes([a]) if a < 252 else bytes([252 + (a & 3), (a - (252 + (a & 3))) >> 2])
        return bytes([toc | 2]) + hdr + rbytes(rng, a) + rbytes(rng, b), 2
    m = int(rng.integers(1, 5))
    vbr = bool(rng.integers(0, 2))
    pad = int(rng.choice([0, 0, 3, 254, 255, 300])) if rng.integers(0, 2) else 0
    head = bytes([toc | 3, m | (0x80 if vbr else 0) | (0x40 if pad else 0)])
    if pad:
        p, pb = pad, b""
        while p >= 255:
            pb += bytes([255])
            p -= 254
        head += pb + bytes([p])
    body = b""
    if vbr:
        lens = [int(rng.integers(2, 240)) for _ in range(m)]
        for L in lens[:-1]:
            head += bytes([L])
        for L in lens:
            body += rbytes(rng, L)
    else:
        L = int(rng.integers(2, 200))
        body = rbytes(rng, L * m)
    return head + body + bytes(pad), m


def make_link(rng, w, channels, wild, foreign):
    pre_skip = int(rng.choice([0, 0, 312, 960, 3840])) if rng.integers(0, 2) else int(rng.integers(0, 5000))
    w.write_page([og.opus_head(channels, pre_skip, gain=int(rng.integers(-500, 500)) if rng.integers(0, 4) == 0 else 0)], 0, bos=True)
    if foreign is not None and rng.integers(0, 2):
        foreign.out = []
        foreign.write_page([b"\x80theora-ish" + rbytes(rng, 20)], 0, bos=True)
        w.out += foreign.out
    w.write_page([og.opus_tags(rbytes(rng, int(rng.integers(0, 600))))], 0)
    n = int(rng.integers(1, 40))
    toc_base = 0xF8
    gp = 0
    i = 0
    while i < n:
        per_page = int(rng.integers(1, 9))
        pk, dur = [], 0
        for _ in range(min(per_page, n - i)):
            toc = toc_base | (4 if rng.integers(0, 2) else 0)
            p, fr = make_packet(rng, toc)
            pk.append(p)
            dur += 960 * fr
            i += 1
        last = i >= n
        if rng.integers(0, 12) == 0:
            w.raw(rbytes(rng, int(rng.integers(1, 300))) + (b"OggS" if rng.integers(0, 3) == 0 else b"Og"))
        if foreign is not None and rng.integers(0, 3) == 0:
            foreign.out = []
            foreign.write_page([rbytes(rng, int(rng.integers(1, 400)))], int(rng.integers(0, 1 << 20)))
            w.out += foreign.out
        gp += dur
        g = gp
        if last and rng.integers(0, 2):
            g = gp - int(rng.integers(0, min(dur + 2000, gp) + 1))  # end trim, possibly more than the page holds
        elif wild and rng.integers(0, 10) == 0:
            g = gp + int(rng.integers(-3000, 3000))
            gp = max(g, 0)
        # a packet spanning two pages now and then (the first page completes no packet: granule -1)
        if len(pk) == 1 and len(pk[0]) > 255 and rng.integers(0, 2):
            w.write_spanning(pk[0], g, 255 * int(rng.integers(1, (len(pk[0]) - 1) // 255 + 1)), eos=last and bool(rng.integers(0, 4)))
        else:
            w.write_page(pk, g, eos=last and bool(rng.integers(0, 4)), corrupt=last and rng.integers(0, 10) == 0)


def make_file(seed, wild):
    rng = np.random.default_rng(seed)
    channels = int(rng.integers(1, 3))
    foreign = og.OggWriter(0x99) if rng.integers(0, 5) == 0 else None
    out = b""
    if rng.integers(0, 8) == 0:
        out += rbytes(rng, int(rng.integers(1, 2000)))
    for link in range(int(rng.choice([1, 1, 1, 2, 3]))):
        w = og.OggWriter(0x1000 + link)
        make_link(rng, w, channels, wild, foreign)
        out += w.blob()
    if rng.integers(0, 10) == 0:
        out += rbytes(rng, int(rng.integers(1, 500)))
    if rng.integers(0, 15) == 0:
        out = out[: len(out) - int(rng.integers(1, min(len(out), 400)))]  # truncated file
    buf = int(rng.choice([2048, 2048, 1920, 600, 4096, 11520, 64, 2]))
    return out, buf


def ref_in_child(blob, buf):
    """(rets, died): the reference player's op_read_stereo return sequence, from a forked child."""
    r, wfd = os.pipe()
    pid = os.fork()
    if pid == 0:
        try:
            os.close(r)
            os.dup2(os.open(os.devnull, os.O_WRONLY), 2)
            os.dup2(os.open(os.devnull, os.O_WRONLY), 1)
            import faulthandler

            faulthandler.disable()
            from oracle import ref

            R = ref.lib()
            data = np.frombuffer(blob, np.uint8).copy()
            rc = R.ref_ogg_open(data.ctypes.data, data.size)
            rets = None
            if rc == 0:
                pcm = np.zeros(max(buf, 2), np.int16)
                rets = []
                for _ in range(200000):
                    v = R.ref_ogg_read_stereo(pcm.ctypes.data, buf)
                    rets.append(v)
                    if v == 0 or v == -128:
                        break
            os.write(wfd, pickle.dumps((rc, rets)))
        finally:
            os._exit(0)
    os.close(wfd)
    chunks = []
    while True:
        b = os.read(r, 1 << 16)
        if not b:
            break
        chunks.append(b)
    os.close(r)
    os.waitpid(pid, 0)
    if not chunks:
        return None, None, True
    rc, rets = pickle.loads(b"".join(chunks))
    return rc, rets, False


def predicted_returns(events, buf):
    out = []
    for kind, code, nbytes, dur, pos, size in events.tolist():
        if kind == 1:
            out.append(code)
        elif kind == 0:
            n = size - pos
            while n > 0:
                c = min(n, buf // 2)
                if c <= 0:
                    return out + ["stuck"]
                out.append(c)
                n -= c
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=1000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--wild", action="store_true")
    ap.add_argument("--replay", type=int, default=None)
    args = ap.parse_args()
    import __graft_entry__ as ge

    pkg = ge.load_package()
    from oracle import ref

    ref.lib()  # loaded before the forks
    stats = {"cases": 0, "compared": 0, "ref_died": 0, "open_failed_both": 0, "failures": []}
    seeds = [args.replay] if args.replay is not None else [(args.seed << 24) + k for k in range(args.cases)]
    for seed in seeds:
        blob, buf = make_file(seed, args.wild)
        stats["cases"] += 1
        rc, want, died = ref_in_child(blob, buf)
        res, ev = pkg.demux_events(np.frombuffer(blob, np.uint8))
        if died:
            stats["ref_died"] += 1
            continue
        if rc != 0:
            if res == 0:
                stats["failures"].append({"seed": seed, "what": "reference refuses the file (%d), we open it" % rc})
            else:
                stats["open_failed_both"] += 1
            continue
        if res != 0:
            stats["failures"].append({"seed": seed, "what": "we refuse the file (%d), the reference opens it" % res})
            continue
        got = predicted_returns(ev, buf)
        stats["compared"] += 1
        if got != want:
            k = next((i for i, (a, b) in enumerate(zip(got, want)) if a != b), min(len(got), len(want)))
            stats["failures"].append({"seed": seed, "buf": buf, "what": "call %d: %s vs reference %s (lengths %d / %d)" % (
                k, got[k:k + 4], want[k:k + 4], len(got), len(want))})
        if args.replay is not None:
            print("want", want[:60], "\ngot ", got[:60], "\nevents", ev[:40].tolist())
    stats["failures_n"] = len(stats["failures"])
    stats["failures"] = stats["failures"][:20]
    stats.update(seed=args.seed, wild=args.wild)
    print(json.dumps(stats))
    return 1 if stats["failures_n"] else 0


if __name__ == "__main__":
    sys.exit(main())
