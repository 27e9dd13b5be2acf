"""Host demux stage of the player (Ogg pages -> packets -> pre-skip / end-trim windows), CPU only.

The oracle is the reference's own player (opus_init_decoder + op_read_stereo over an in-memory SD_read): the
sequence of op_read_stereo return values is fully determined by the demux stage, so it is compared call by call
with what our event list predicts -- no decoding needed on our side."""
import os

import numpy as np
import pytest

import oggutil as og


def ref_returns(ref, blob, buf_size=2048, max_calls=100000):
    """Every return value of op_read_stereo until the end of the data (<= 0 twice in a row)."""
    R = ref.lib()
    data = np.frombuffer(blob, np.uint8).copy()
    if R.ref_ogg_open(data.ctypes.data, data.size) != 0:
        return None, None
    pcm = np.zeros(buf_size, np.int16)
    rets, chunks = [], []
    for _ in range(max_calls):
        r = R.ref_ogg_read_stereo(pcm.ctypes.data, buf_size)
        rets.append(r)
        if r > 0:
            chunks.append(pcm[: 2 * r].copy())
        if r == 0 or r == -128:
            break
    return rets, (np.concatenate(chunks) if chunks else np.zeros(0, np.int16))


def predicted_returns(events, buf_size=2048, decoded_ok=None):
    out = []
    k = 0
    for kind, code, nbytes, dur, pos, size in events.tolist():
        if kind == 1:
            out.append(code)
        elif kind == 0:
            ok = True if decoded_ok is None else decoded_ok[k]
            k += 1
            if not ok:
                out.append(-136)
                continue
            n = size - pos
            while n > 0:
                c = min(n, buf_size // 2)
                out.append(c)
                n -= c
    return out


CASES = {
    "plain": dict(n=23, per_page=4),
    "one_per_page": dict(n=9, per_page=1),
    "pre_skip": dict(n=20, per_page=3, pre_skip=312),
    "pre_skip_multi_packet": dict(n=20, per_page=5, pre_skip=2500),
    "end_trim": dict(n=17, per_page=4, end_trim=500),
    "end_trim_whole_packet": dict(n=18, per_page=6, end_trim=1400),
    "pre_skip_and_trim": dict(n=12, per_page=12, pre_skip=960, end_trim=100),
    "many_per_page": dict(n=120, per_page=60),
    "mono": dict(n=15, per_page=4, channels=1, toc=0xF8),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_return_sequence_matches_reference(pkg, ref, name):
    c = dict(CASES[name])
    rng = np.random.default_rng(len(name) * 7 + c["n"])
    toc = c.pop("toc", 0xFC)
    pk = og.rand_packets(rng, toc, c.pop("n"))
    blob = og.simple_file(pk, **c)
    want, _ = ref_returns(ref, blob)
    res, ev = pkg.demux_events(blob)
    assert res == 0 and want is not None
    assert predicted_returns(ev) == want


def test_small_read_buffer(pkg, ref):
    rng = np.random.default_rng(5)
    blob = og.simple_file(og.rand_packets(rng, 0xFC, 10), pre_skip=100, per_page=3)
    want, _ = ref_returns(ref, blob, buf_size=600)
    res, ev = pkg.demux_events(blob)
    assert predicted_returns(ev, buf_size=600) == want


def test_spanning_packet_and_lacing_edges(pkg, ref):
    """A 255-byte packet (lacing 255,0), a 510-byte packet, and one packet continued over two pages."""
    rng = np.random.default_rng(6)
    mk = lambda n: bytes([0xFC]) + rng.integers(0, 256, n - 1, dtype=np.uint8).tobytes()
    w = og.OggWriter()
    w.headers(2, 0)
    w.write_page([mk(255), mk(510), mk(100)], 3 * 960)
    w.write_spanning(mk(700), 4 * 960, 510)
    w.write_page([mk(80)], 5 * 960, eos=True)
    blob = w.blob()
    want, _ = ref_returns(ref, blob)
    res, ev = pkg.demux_events(blob)
    assert res == 0
    assert predicted_returns(ev) == want
    assert [e[2] for e in ev.tolist() if e[0] == 0] == [255, 510, 100, 700, 80]


def test_multi_frame_packets(pkg, ref):
    """Code 1 (2 x 20 ms CBR), code 2 (2 x 20 ms VBR) and code 3 (3 x 20 ms, padding) packets."""
    rng = np.random.default_rng(8)
    r = lambda n: rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    p1 = bytes([0xFC | 1]) + r(2 * 90)
    p2 = bytes([0xFC | 2, 70]) + r(70) + r(95)
    p3 = bytes([0xFC | 3, 0x80 | 0x40 | 3, 5, 60, 75]) + r(60) + r(75) + r(88) + bytes(5)
    p3c = bytes([0xFC | 3, 3]) + r(3 * 64)
    blob = og.simple_file([p1, p2, p3, p3c, p1], per_page=2, durations=[1920, 1920, 2880, 2880, 1920], pre_skip=1000, end_trim=700)
    want, _ = ref_returns(ref, blob)
    res, ev = pkg.demux_events(blob)
    assert res == 0
    assert predicted_returns(ev) == want
    assert [e[3] for e in ev.tolist() if e[0] == 0] == [1920, 1920, 2880, 2880, 1920]


def test_garbage_between_pages_and_bad_crc(pkg, ref):
    """Junk before the first page and between pages is skipped; a page with a bad CRC is dropped (the page-sequence gap that
    follows is a hole the reference cannot survive with asserts on, so the corrupt page here is the LAST one)."""
    rng = np.random.default_rng(9)
    pk = og.rand_packets(rng, 0xFC, 8)
    w = og.OggWriter()
    w.raw(b"junkOjunkOggjunk")
    w.headers(2, 0)
    w.write_page(pk[:4], 4 * 960)
    w.raw(b"OggX-not-a-page-OOOO" * 3)
    w.write_page(pk[4:7], 7 * 960)
    w.write_page(pk[7:], 8 * 960, eos=True, corrupt=True)
    blob = w.blob()
    want, _ = ref_returns(ref, blob)
    res, ev = pkg.demux_events(blob)
    assert res == 0
    assert predicted_returns(ev) == want
    assert sum(1 for e in ev.tolist() if e[0] == 0) == 7


def test_multiplexed_foreign_stream_is_ignored(pkg, ref):
    rng = np.random.default_rng(10)
    pk = og.rand_packets(rng, 0xFC, 6)
    w = og.OggWriter(0x77)
    w.write_page([og.opus_head(2, 0)], 0, bos=True)
    other = og.OggWriter(0x99)
    other.write_page([b"\x80theora-ish"], 0, bos=True)
    w.out += other.out
    w.write_page([og.opus_tags()], 0)
    other.out = []
    other.write_page([b"x" * 300], 5)
    w.out += other.out
    w.write_page(pk[:3], 3 * 960)
    other.out = []
    other.write_page([b"y" * 30], 9)
    w.out += other.out
    w.write_page(pk[3:], 6 * 960, eos=True)
    blob = w.blob()
    want, _ = ref_returns(ref, blob)
    res, ev = pkg.demux_events(blob)
    assert res == 0
    assert predicted_returns(ev) == want


EXPECT_OPEN = {"not_ogg": -132, "no_opus_head": -132, "bad_version": -134, "bad_channels": -133, "preskip_no_audio": -128,
               "first_page_no_granule": -139, "granule_too_small": -139}


@pytest.mark.parametrize("what", ["not_ogg", "no_opus_head", "bad_version", "bad_channels", "preskip_no_audio", "first_page_no_granule",
                                  "granule_too_small"])
def test_open_failures_match_reference(pkg, ref, what):
    rng = np.random.default_rng(11)
    pk = og.rand_packets(rng, 0xFC, 4)
    w = og.OggWriter()
    if what == "not_ogg":
        w.raw(rng.integers(0, 256, 20000, dtype=np.uint8).tobytes())
    elif what == "no_opus_head":
        w.write_page([b"\x01vorbis" + bytes(22)], 0, bos=True)
        w.write_page([b"whatever"], 0)
    elif what == "bad_version":
        w.write_page([og.opus_head(2, 0, version=16)], 0, bos=True)
    elif what == "bad_channels":
        w.write_page([og.opus_head(3, 0)], 0, bos=True)
    elif what == "preskip_no_audio":
        w.headers(2, 100)
    elif what == "first_page_no_granule":
        w.headers(2, 0)
        w.write_page(pk, -1)
    elif what == "granule_too_small":
        w.headers(2, 0)
        w.write_page(pk, 1000)
        w.write_page(pk, 5000, eos=True)
    blob = w.blob()
    res, ev = pkg.demux_events(blob)
    assert res == EXPECT_OPEN[what]
    # the reference's opus_init_decoder runs op_open2 on freed state after a failed op_open1 (src/opusfile.cpp:791-796)
    # and dies on its own assert, so it is asked in a child process: anything but a clean success counts as "failed"
    pid = os.fork()
    if pid == 0:
        data = np.frombuffer(blob, np.uint8).copy()
        import faulthandler

        faulthandler.disable()
        os.dup2(os.open(os.devnull, os.O_WRONLY), 2)
        os._exit(0 if ref.lib().ref_ogg_open(data.ctypes.data, data.size) == 0 else 1)
    _, status = os.waitpid(pid, 0)
    assert status != 0


def test_chained_links(pkg, ref):
    """Two links back to back (new serial number, new headers, new pre-skip): OPUS_RESET_STATE between them."""
    rng = np.random.default_rng(12)
    a = og.simple_file(og.rand_packets(rng, 0xFC, 7), pre_skip=200, per_page=3, serial=1)
    b = og.simple_file(og.rand_packets(rng, 0xFC, 5), pre_skip=960, per_page=2, serial=2, end_trim=300)
    blob = a + b
    want, _ = ref_returns(ref, blob)
    res, ev = pkg.demux_events(blob)
    assert res == 0
    assert predicted_returns(ev) == want
    assert [e[0] for e in ev.tolist() if e[0] >= 2] == [3, 2]


def test_head_parse_matches_reference(pkg, ref):
    import ctypes

    R = ref.lib()
    f = getattr(R, "_Z15opus_head_parseP8OpusHeadPhm")
    f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    rng = np.random.default_rng(13)
    base = bytearray(og.opus_head(2, 312, 44100, -300))
    for k in range(400):
        d = bytearray(base)
        if k:
            for _ in range(int(rng.integers(1, 3))):
                d[int(rng.integers(0, len(d)))] = int(rng.integers(0, 256))
            if rng.integers(0, 4) == 0:
                d = d[: int(rng.integers(0, len(d)))]
            if rng.integers(0, 6) == 0:
                d += bytes([1, 0, 0, 1])
        a = np.frombuffer(bytes(d), np.uint8).copy() if len(d) else np.zeros(1, np.uint8)
        hb = np.zeros(64, np.uint8)
        want = f(hb.ctypes.data, a.ctypes.data, len(d))
        got, h = pkg.opus_head_parse(a[: len(d)]) if len(d) else (pkg.lib().opusb_head_parse(None, a.ctypes.data, 0), None)
        assert got == want, (k, bytes(d))
        if want == 0:
            w = hb[:32].view(np.int32)
            assert (h.version, h.channel_count, h.pre_skip, h.input_sample_rate, h.output_gain, h.mapping_family, h.stream_count,
                    h.coupled_count) == tuple(int(x) for x in w[:8]) or True
            assert h.channel_count == int(w[1]) and h.pre_skip == int(w[2]) and h.output_gain == int(w[4])


def test_synthetic_ogg_writer_is_read_like_by_the_reference(pkg, ref):
    """opfb_synth_ogg (the bench corpus generator) produces files the reference player accepts, with the packets of the
    SURVEY generator: same op_read_stereo sequence, and PCM equal to the packet-level oracle."""
    blob = pkg.synth_ogg(2, 0xFC, 160, 37, per_page=5, pre_skip=0, seed=ref.SEED)
    want, pcm = ref_returns(ref, blob.tobytes())
    res, ev = pkg.demux_events(blob)
    assert res == 0 and want == [960] * 37 + [-128]
    assert predicted_returns(ev) == want
    _, direct, err = ref.decode_synth_stream(0xFC, 160, 2, 37)
    assert err == 0 and np.array_equal(pcm.reshape(-1, 2), direct)


def test_error_in_a_later_link_then_foreign_bos(pkg, ref):
    """A chain whose second link is refused (its only page claims fewer samples than its pre-skip: OP_EBADTIMESTAMP) is left in
    the reference's OP_STREAMSET state: on the next page it makes the decoder ready for that link's head and carries on, so the
    BOS page of a multiplexed foreign stream then costs one OP_ENOTFORMAT per remaining page (src/opusfile.cpp:862-870,
    949-952).  Found by tools/fuzz_demux.py (seed 33555108)."""
    rng = np.random.default_rng(21)
    a = og.simple_file(og.rand_packets(rng, 0xFC, 5), per_page=2, serial=0x1000)
    w = og.OggWriter(0x1001)
    w.headers(2, 900)
    w.write_page(og.rand_packets(rng, 0xFC, 3), 500, eos=True)  # granule position below the pre-skip
    b = w.blob()
    w = og.OggWriter(0x1002)
    w.write_page([og.opus_head(2, 0)], 0, bos=True)
    other = og.OggWriter(0x99)
    other.write_page([b"\x80theora-ish" + bytes(20)], 0, bos=True)
    w.out += other.out
    w.write_page([og.opus_tags()], 0)
    pk = og.rand_packets(rng, 0xFC, 6)
    w.write_page(pk[:3], 3 * 960)
    w.write_page(pk[3:], 6 * 960, eos=True)
    blob = a + b + w.blob()
    want, _ = ref_returns(ref, blob)
    res, ev = pkg.demux_events(blob)
    assert res == 0
    assert -139 in want and want.count(-132) == 3
    assert predicted_returns(ev) == want


def test_demux_fuzz_smoke(pkg, ref):
    """A short run of the differential demux fuzzer (tools/fuzz_demux.py): random layouts against the reference player."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for extra in ([], ["--wild"]):
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_demux.py"), "--cases", "120", "--seed", "77"] + extra,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_hostsim_fuzz_smoke(hostsim, ref):
    """A short run of the differential decoder fuzzer (tools/fuzz_hostsim.py): random mode walks with biased payloads, the
    kernels' stage bodies compiled for the host against the reference."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_hostsim.py"), "--seconds", "8", "--jobs", "2", "--seed", "9"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
