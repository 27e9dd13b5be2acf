"""Host side of the many-files player (csrc/opus_player.cpp: RoundScheduler + Track::read) against a sequential model, on the
CPU: the device behind the dm_* ABI is tests/mockdev/mock_device.cpp, whose "PCM" of a frame is a checksum of what the
scheduler handed over for it (bytes, TOC, frames decoded on the stream since its last reset).  So the samples
opfb_set_read_stereo returns tell whether every frame of every file reached the device exactly once, in order, with the
right bytes, behind the right decoder resets, and came back through the right play window (pre-skip, end trim) -- with
rounds in flight, files of different lengths, multi-frame packets, chained links, failing packets, slow readers whose
samples have to be copied aside, and odd read sizes.  The GPU tests (test_gpu_player.py) check the same path end to end
against the reference player with the real decoder."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oggutil as og

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PENDING = -2100


@pytest.fixture(scope="module")
def mock():
    d = os.path.join(HERE, "mockdev")
    so = os.path.join(d, "libplayer_mock.so")
    csrc = os.path.join(ROOT, "esp32-opus-player_b200", "csrc")
    srcs = [os.path.join(d, "mock_device.cpp"), os.path.join(csrc, "opus_player.cpp"), os.path.join(csrc, "host_api.cpp")]
    deps = srcs + [os.path.join(csrc, "ogg_demux.hpp"), os.path.join(ROOT, "include", "opus_b200.h"), os.path.join(ROOT, "include", "opusfile_b200.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wl,-Bsymbolic", "-o", so] + srcs + ["-lpthread"], check=True)
    L = ctypes.CDLL(so)
    P, I = ctypes.c_void_p, ctypes.c_int
    L.opfb_set_create.argtypes = [ctypes.POINTER(P), I, I, I, I]
    L.opfb_set_destroy.argtypes = [P]
    L.opfb_set_open_memory.argtypes = [P, I, P, ctypes.c_size_t, I]
    L.opfb_set_prepare.argtypes = [P]
    L.opfb_set_open_result.argtypes = [P, I]
    L.opfb_set_decode_round.argtypes = [P]
    L.opfb_set_read_stereo.argtypes = [P, I, P, I]
    L.opfb_debug_events.argtypes = [P, ctypes.c_size_t, I, P, I, P]
    L.opusb_packet_parse.argtypes = [P, I, P, P, P, P]
    return L


def fnv_frame(seq, toc, payload):
    h = 2166136261
    for v in (seq & 255, (seq >> 8) & 255, (seq >> 16) & 255, toc):
        h = ((h ^ v) * 16777619) & 0xFFFFFFFF
    for v in payload:
        h = ((h ^ v) * 16777619) & 0xFFFFFFFF
    return h


def frame_pcm(h, ch):
    i = np.arange(960, dtype=np.uint32)[:, None]
    c = np.arange(ch, dtype=np.uint32)[None, :]
    return ((np.uint32(h) + np.uint32(131) * i + np.uint32(7) * c) & np.uint32(0xFFFF)).astype(np.uint16).view(np.int16)


def parse_packet(L, pkt):
    a = np.frombuffer(pkt, np.uint8).copy()
    toc = ctypes.c_uint8(0)
    sizes = (ctypes.c_int16 * 48)()
    off = ctypes.c_int(0)
    n = L.opusb_packet_parse(a.ctypes.data, a.size, ctypes.cast(ctypes.byref(toc), ctypes.c_void_p), None, ctypes.cast(sizes, ctypes.c_void_p),
                             ctypes.cast(ctypes.byref(off), ctypes.c_void_p))
    frames, pos = [], off.value
    for k in range(max(n, 0)):
        frames.append(pkt[pos:pos + sizes[k]])
        pos += sizes[k]
    return n, toc.value, frames


def model(L, blob, packets, ch):
    """What op_read_stereo must deliver for one file: (stereo samples int16[n][2], return codes in order)."""
    data = np.frombuffer(blob, np.uint8).copy()
    ev = np.zeros((len(packets) + 64, 6), np.int32)
    res = ctypes.c_int(0)
    n = L.opfb_debug_events(data.ctypes.data, data.size, 0, ev.ctypes.data, ev.shape[0], ctypes.byref(res))
    assert res.value == 0 and n <= ev.shape[0]
    out, codes = [], []
    seq, k = 0, 0
    for kind, code, nbytes, dur, pos, size in ev[:n].tolist():
        if kind == 3:
            seq = 0
        elif kind == 2:
            seq = 1 << 20
        elif kind == 1:
            codes.append(code)
        else:
            pkt = packets[k]
            k += 1
            assert len(pkt) == nbytes
            cnt, toc, frames = parse_packet(L, pkt)
            assert cnt * 960 == dur
            ok, parts = True, []
            for fr in frames:
                if not ok:
                    break
                if len(fr) and fr[0] == 0xEE:
                    ok = False
                    break
                parts.append(frame_pcm(fnv_frame(seq, toc, fr), ch))
                seq += 1
            if not ok:
                codes.append(-136)
                continue
            pcm = np.concatenate(parts)[pos:size]
            out.append(np.repeat(pcm, 2, axis=1) if ch == 1 else pcm)
    return (np.concatenate(out) if out else np.zeros((0, 2), np.int16)), codes


def make_file(rng, idx, allow_fail):
    """-> (blob, packets in stream order, channels)"""
    ch = int(rng.integers(1, 3))
    toc = 0xF8 | (4 if ch == 2 else 0)
    blob, packets = b"", []
    for link in range(int(rng.choice([1, 1, 2, 3]))):
        n = int(rng.integers(1, 90))
        pk, dur = [], []
        for _ in range(n):
            code = int(rng.choice([0, 0, 0, 1, 3]))
            r = lambda m: rng.integers(0, 256, m, dtype=np.uint8).tobytes()
            if code == 0:
                p, f = bytes([toc]) + r(int(rng.integers(3, 120))), 1
            elif code == 1:
                p, f = bytes([toc | 1]) + r(2 * int(rng.integers(3, 60))), 2
            else:
                m = int(rng.integers(1, 4))
                p, f = bytes([toc | 3, m]) + r(m * int(rng.integers(3, 50))), m
            if allow_fail and rng.integers(0, 25) == 0:  # a frame the mock device refuses
                q = bytearray(p)
                q[1 if code < 3 else 2] = 0xEE
                p = bytes(q)
            pk.append(p)
            dur.append(960 * f)
        pre = 0 if allow_fail else int(rng.choice([0, 0, 312, 960, 2500]))
        total = sum(dur)
        trim = int(rng.integers(0, min(total - pre, 900) + 1)) if total - pre > 0 and rng.integers(0, 2) else 0
        if total - trim < pre:
            pre = 0
        w = og.OggWriter(0x4000 + 16 * idx + link)
        w.headers(ch, pre)
        i, gp = 0, 0
        while i < n:
            k = int(rng.integers(1, 9))
            gp += sum(dur[i:i + k])
            last = i + k >= n
            w.write_page(pk[i:i + k], gp - (trim if last else 0), eos=last)
            i += k
        blob += w.blob()
        packets += pk
    return blob, packets, ch


@pytest.mark.parametrize("seed,n_files,frames_per_round,allow_fail", [(1, 40, 12, False), (2, 90, 50, False), (3, 33, 7, True), (4, 150, 24, True)])
def test_scheduler_delivers_every_frame_once_in_order(mock, seed, n_files, frames_per_round, allow_fail):
    L = mock
    rng = np.random.default_rng(seed)
    files = [make_file(rng, i, allow_fail) for i in range(n_files)]
    blobs = [np.frombuffer(b, np.uint8).copy() for b, _, _ in files]
    want = [model(L, b, pk, ch) for b, pk, ch in files]
    st = ctypes.c_void_p()
    assert L.opfb_set_create(ctypes.byref(st), n_files, 0, frames_per_round, 4) == 0
    try:
        for i, b in enumerate(blobs):
            assert L.opfb_set_open_memory(st, i, b.ctypes.data, b.size, 0) == 0
        assert L.opfb_set_prepare(st) == 0
        got = [[] for _ in files]
        codes = [[] for _ in files]
        done = [False] * n_files
        # readers of different temper: some read everything after every round, some only now and then (their samples have
        # to be kept aside when the round buffer comes round again), with their own buffer sizes
        lazy = rng.integers(0, 4, n_files)
        bufsz = [int(rng.choice([2048, 1920, 600, 4096, 64, 11520])) for _ in files]
        buf = np.zeros(11520, np.int16)
        rounds = 0
        while not all(done):
            more = L.opfb_set_decode_round(st)
            assert more >= 0
            rounds += 1
            assert rounds < 5000
            for i in range(n_files):
                if done[i] or (lazy[i] and more and rounds % (2 + int(lazy[i])) != 0):
                    continue
                for _ in range(100000):
                    r = L.opfb_set_read_stereo(st, i, buf.ctypes.data, bufsz[i])
                    if r == PENDING:
                        break
                    if r > 0:
                        got[i].append(buf[: 2 * r].copy())
                    else:
                        codes[i].append(r)
                        if r in (0, -128):
                            done[i] = True
                            break
        for i in range(n_files):
            g = np.concatenate(got[i]).reshape(-1, 2) if got[i] else np.zeros((0, 2), np.int16)
            assert codes[i] == want[i][1], "file %d: return codes" % i
            assert g.shape == want[i][0].shape and np.array_equal(g, want[i][0]), "file %d: samples" % i
    finally:
        L.opfb_set_destroy(st)


@pytest.mark.parametrize("seed,allow_fail", [(11, False), (12, True)])
def test_single_stream_player_over_a_read_callback(mock, seed, allow_fail):
    """opus_init_decoder() + op_read_stereo() (opusb_init_decoder / opusb_read_stereo) over a read callback that hands out the
    file in odd chunk sizes: the framer's own buffer (not the in-place path of memory images), one track per scheduler, the
    decode-ahead loop inside the read call -- same samples and return codes as the model."""
    L = mock
    P, I = ctypes.c_void_p, ctypes.c_int
    READ = ctypes.CFUNCTYPE(I, P, P, I)
    L.opusb_init_decoder.restype = P
    L.opusb_init_decoder.argtypes = [READ, P, I, ctypes.POINTER(I)]
    L.opusb_read_stereo.argtypes = [P, P, I]
    L.opusb_free_decoder.argtypes = [P]
    rng = np.random.default_rng(seed)
    for k in range(12):
        blob, packets, ch = make_file(rng, k, allow_fail)
        want_pcm, want_codes = model(L, blob, packets, ch)
        pos = [0]
        short = int(rng.integers(0, 2))

        def cb(user, buf, n, blob=blob, pos=pos, short=short):
            if n == 0:
                return 0
            m = min(n, len(blob) - pos[0])
            if m <= 0:
                return -1
            if short:  # a source that returns less than asked for
                m = max(1, min(m, 1 + (pos[0] * 7919) % 700))
            ctypes.memmove(buf, blob[pos[0]:pos[0] + m], m)
            pos[0] += m
            return m

        fn = READ(cb)
        err = I(0)
        of = L.opusb_init_decoder(fn, None, 0, ctypes.byref(err))
        assert of and err.value == 0
        buf = np.zeros(11520, np.int16)
        bs = int(rng.choice([2048, 600, 11520, 64]))
        got, codes = [], []
        for _ in range(200000):
            r = L.opusb_read_stereo(of, buf.ctypes.data, bs)
            if r > 0:
                got.append(buf[: 2 * r].copy())
            else:
                codes.append(r)
                if r in (0, -128):
                    break
        L.opusb_free_decoder(of)
        g = np.concatenate(got).reshape(-1, 2) if got else np.zeros((0, 2), np.int16)
        assert codes == want_codes, "file %d" % k
        assert g.shape == want_pcm.shape and np.array_equal(g, want_pcm), "file %d" % k
