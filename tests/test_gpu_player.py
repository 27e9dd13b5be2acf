"""GPU parity tests of the player layer (SURVEY 8b player API, 8f rows 1-3): Ogg Opus files through
opus_init_decoder()/op_read_stereo()-shaped calls, single stream and many files at once, against the reference's own
player (oracle: opus_init_decoder + op_read_stereo over an in-memory SD_read).  Same samples, same return value per call."""
import numpy as np
import pytest

import oggutil as og
from test_player_demux import ref_returns

pytestmark = pytest.mark.gpu


def ours_single(pkg, blob, buf_size=2048, max_calls=100000):
    f = pkg.OggOpusFile(blob)
    rets, chunks = [], []
    for _ in range(max_calls):
        r, pcm = f.read_stereo(buf_size)
        rets.append(r)
        if r > 0:
            chunks.append(pcm.reshape(-1).copy())
        if r == 0 or r == -128:
            break
    f.close()
    return rets, (np.concatenate(chunks) if chunks else np.zeros(0, np.int16))


def ours_set(pkg, blobs, buf_size=2048, frames_per_round=50):
    st = pkg.OggOpusFileSet(blobs, frames_per_round=frames_per_round, n_threads=4)
    n = len(blobs)
    rets = [[] for _ in range(n)]
    chunks = [[] for _ in range(n)]
    done = [st.open_result(i) != 0 for i in range(n)]
    for _ in range(10000):
        more = st.decode_round()
        assert more >= 0
        for i in range(n):
            while not done[i]:
                r, pcm = st.read_stereo(i, buf_size)
                if r == pkg.api.OPFB_PENDING:
                    break
                rets[i].append(r)
                if r > 0:
                    chunks[i].append(pcm.reshape(-1).copy())
                if r == 0 or r == -128:
                    done[i] = True
        if all(done):
            break
    stats = st.stats()
    st.close()
    return rets, [np.concatenate(c) if c else np.zeros(0, np.int16) for c in chunks], stats


def mixed_packets(rng, stereo, n):
    s = 4 if stereo else 0
    tocs = [0xF8 | s, 0x48 | s, 0x78 | s, 0x08 | s, 0x68 | s, 0xD8 | s]
    out = []
    while len(out) < n:
        t = int(rng.choice(tocs))
        out += og.rand_packets(rng, t, int(rng.integers(1, 5)), 20, 250)
    return out[:n]


@pytest.mark.parametrize("case", ["celt", "pre_skip_trim", "mono_silk", "mixed_modes", "small_buf"])
def test_single_stream_matches_reference_player(pkg, ref, case):
    rng = np.random.default_rng(len(case))
    buf = 2048
    if case == "celt":
        blob = og.simple_file(og.rand_packets(rng, 0xFC, 70), per_page=5)
    elif case == "pre_skip_trim":
        blob = og.simple_file(og.rand_packets(rng, 0xFC, 30), per_page=4, pre_skip=3840, end_trim=777)
    elif case == "mono_silk":
        blob = og.simple_file(og.rand_packets(rng, 0x48, 60, 30, 90), channels=1, per_page=7, pre_skip=312)
    elif case == "mixed_modes":
        blob = og.simple_file(mixed_packets(rng, True, 90), per_page=6, pre_skip=400, end_trim=200)
    else:
        blob = og.simple_file(og.rand_packets(rng, 0xFC, 12), per_page=3, pre_skip=100)
        buf = 500
    want_r, want_pcm = ref_returns(ref, blob, buf)
    got_r, got_pcm = ours_single(pkg, blob, buf)
    assert got_r == want_r
    assert np.array_equal(got_pcm, want_pcm)


def test_multi_frame_and_failing_packets(pkg, ref):
    """Code 1/2/3 packets; a packet whose second frame is 1 byte long (CELT refuses it: OP_EBADPACKET, the frames behind
    it in the packet are not decoded, the stream carries on); a packet with a broken frame-count byte."""
    rng = np.random.default_rng(21)
    r = lambda n: rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    pk = og.rand_packets(rng, 0xFC, 3)
    pk.append(bytes([0xFC | 1]) + r(2 * 90))
    pk.append(bytes([0xFC | 2, 70]) + r(70) + r(95))
    pk.append(bytes([0xFC | 3, 0x80 | 3, 60, 1]) + r(60) + r(1) + r(88))    # frame 2 undecodable
    pk += og.rand_packets(rng, 0xFC, 2)
    pk.append(bytes([0xFC | 3, 0x80 | 0x40 | 3, 5, 60, 75]) + r(60) + r(75) + r(88) + bytes(5))
    pk.append(bytes([0xFC | 1]) + r(2 * 80 + 1))                             # code 1 with an odd length: invalid
    pk += og.rand_packets(rng, 0xFC, 3)
    dur = [960] * 3 + [1920, 1920, 2880, 960, 960, 2880, 1920] + [960] * 3
    blob = og.simple_file(pk, per_page=3, durations=dur, pre_skip=500, end_trim=300)
    want_r, want_pcm = ref_returns(ref, blob)
    got_r, got_pcm = ours_single(pkg, blob)
    assert -136 in want_r
    assert got_r == want_r
    assert np.array_equal(got_pcm, want_pcm)


def test_chained_links_reset_state(pkg, ref):
    rng = np.random.default_rng(22)
    a = og.simple_file(mixed_packets(rng, True, 25), pre_skip=200, per_page=4, serial=1)
    b = og.simple_file(mixed_packets(rng, True, 25), pre_skip=960, per_page=3, serial=2, end_trim=300)
    c = og.simple_file(og.rand_packets(rng, 0xFC, 10), pre_skip=0, per_page=5, serial=3)
    blob = a + b + c
    want_r, want_pcm = ref_returns(ref, blob)
    got_r, got_pcm = ours_single(pkg, blob)
    assert got_r == want_r
    assert np.array_equal(got_pcm, want_pcm)


def test_many_files_at_once(pkg, ref):
    """96 files of different lengths, channel counts, modes, pre-skips and page layouts decoded together."""
    rng = np.random.default_rng(23)
    blobs = []
    for i in range(96):
        stereo = bool(i % 3)
        n = int(rng.integers(5, 130))
        pk = mixed_packets(rng, stereo, n) if i % 2 else og.rand_packets(rng, 0xFC if stereo else 0xF8, n)
        blobs.append(og.simple_file(pk, channels=2 if stereo else 1, per_page=int(rng.integers(1, 9)), pre_skip=int(rng.integers(0, 2000)),
                                    end_trim=int(rng.integers(0, 900)), serial=1000 + i))
    blobs.append(b"this is not an ogg file" * 100)
    got_r, got_pcm, stats = ours_set(pkg, blobs, frames_per_round=32)
    assert stats["launches"] > 0 and stats["packets"] > 96 * 5
    for i, blob in enumerate(blobs[:-1]):
        want_r, want_pcm = ref_returns(ref, blob)
        assert got_r[i] == want_r, i
        assert np.array_equal(got_pcm[i], want_pcm), i
    assert got_r[-1] == []


def test_config1_plumbing_3000_packets(pkg, ref):
    """BASELINE configs[0]: one 48 kHz stereo CELT FB 20 ms stream, 3000 x 160 B packets in Ogg, through the player API;
    equals the packet-level oracle on the same packets."""
    n = 3000
    g = ref.lib()
    st = np.array([ref.SEED], np.uint64)
    pk = []
    buf = np.zeros(160, np.uint8)
    for _ in range(n):
        g.ref_gen_packet(st.ctypes.data, 0xFC, 160, buf.ctypes.data)
        pk.append(buf.tobytes())
    w = og.OggWriter()
    w.headers(2, 0)
    for i in range(n):
        w.write_page([pk[i]], 960 * (i + 1), eos=i == n - 1)
    blob = w.blob()
    got_r, got_pcm = ours_single(pkg, blob)
    assert got_r[:-1] == [960] * n and got_r[-1] == -128
    _, want, err = ref.decode_synth_stream(0xFC, 160, 2, n)
    assert err == 0
    assert np.array_equal(got_pcm.reshape(-1, 2), want)
    assert "%016x" % ref.fnv_hash(got_pcm[: 500 * 960 * 2], 2) == "6dffccd132b87ac1"
