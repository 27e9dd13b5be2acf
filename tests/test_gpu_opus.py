"""GPU parity tests of the Opus frame layer: SILK-only, hybrid and the mode transitions, through the C ABI
(decode_many) against the oracle (the unmodified reference, packet level).  SURVEY 8a rows a13-a20, configs 3-5."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "survey_hashes.json")))


def _pack(pkg, plans, seed):
    """plans[s] = list of (toc, payload_len) per frame; random payloads. Returns (bytes, frames, packets[s][f])."""
    rng = np.random.default_rng(seed)
    n_streams, n_frames = len(plans), len(plans[0])
    frames = pkg.make_frames(n_streams, n_frames)
    chunks, pos, pk = [], 0, []
    for s, plan in enumerate(plans):
        row = []
        for f, (toc, ln) in enumerate(plan):
            L = ln if ln > 0 else int(rng.integers(3, 400))
            payload = rng.integers(0, 256, L, dtype=np.uint8)
            frames[s, f] = (pos, L, toc, 0)
            chunks.append(payload)
            pos += L
            row.append(np.concatenate([np.array([toc], np.uint8), payload]))
        pk.append(row)
    return np.concatenate(chunks), frames, pk


def _check(pkg, ref, ch, plans, seed, calls=1):
    data, frames, pk = _pack(pkg, plans, seed)
    n_streams, n_frames = frames.shape
    per = n_frames // calls
    dec = pkg.BatchDecoder(n_streams, per, channels=ch)
    pcm = np.zeros((n_streams, n_frames, 960, ch), np.int16)
    status = np.zeros((n_streams, n_frames), np.int32)
    for k in range(calls):
        rc, p, st = dec.decode_many(data, np.ascontiguousarray(frames[:, k * per:(k + 1) * per]))
        assert rc == 0
        pcm[:, k * per:(k + 1) * per] = p
        status[:, k * per:(k + 1) * per] = st
    dec.close()
    for s in range(n_streams):
        o = ref.PacketDecoder(ch)
        try:
            for f in range(n_frames):
                r, want = o.decode(pk[s][f])
                assert status[s, f] == r, "stream %d frame %d: status %d vs %d" % (s, f, status[s, f], r)
                if r > 0:
                    assert np.array_equal(pcm[s, f], want), "stream %d frame %d toc %#x" % (s, f, pk[s][f][0])
        finally:
            o.close()


@pytest.mark.parametrize("toc,ch,ln", [(0x08, 1, 40), (0x28, 1, 50), (0x48, 1, 60), (0x4C, 2, 80), (0x0C, 2, 60), (0x2C, 2, 70),
                                       (0x68, 1, 100), (0x6C, 2, 120), (0x78, 1, 110), (0x7C, 2, 160), (0x48, 1, 0), (0x4C, 2, 0),
                                       (0x7C, 2, 0)])
def test_single_mode_streams_bit_exact(pkg, ref, toc, ch, ln):
    _check(pkg, ref, ch, [[(toc, ln)] * 40 for _ in range(37)], toc * 31 + ln)


@pytest.mark.parametrize("ch", [1, 2])
def test_mode_switching_bit_exact(pkg, ref, ch):
    """Every stream walks its own random sequence of SILK NB/MB/WB, hybrid SWB/FB and CELT frames; decoded in
    3 calls so that transitions also straddle call boundaries (prev_mode, SILK/CELT state carried in HBM)."""
    s = 4 if ch == 2 else 0
    tocs = [0x08, 0x28, 0x48, 0x68, 0x78, 0xF8, 0xD8, 0x98]
    rng = np.random.default_rng(5 + ch)
    plans = []
    for _ in range(48):
        plan = []
        while len(plan) < 60:
            t = int(rng.choice(tocs)) | s
            plan += [(t, 0)] * int(rng.integers(1, 6))
        plans.append(plan[:60])
    _check(pkg, ref, ch, plans, 900 + ch, calls=3)


@pytest.mark.parametrize("ch", [1, 2])
def test_config5_mixed_modes_side_by_side(pkg, ref, ch):
    """BASELINE configs[4] shape at oracle-sized scale: the streams of one batch alternate CELT FB, SILK WB, hybrid FB
    and hybrid SWB (mode = stream index mod 4, so every warp of the entropy kernel holds all four), decoded in two
    calls; one decoder context per channel count, as the player layer groups its files."""
    s = 4 if ch == 2 else 0
    comp = [(0xF8 | s, 159), (0x48 | s, 59), (0x78 | s, 159), (0x68 | s, 99)]
    _check(pkg, ref, ch, [[comp[k % 4]] * 12 for k in range(132)], 77 + ch, calls=2)


@pytest.mark.parametrize("idx", [1, 2, 3, 4])
def test_survey_hashes_silk_hybrid(pkg, ref, idx):
    """SURVEY 8c golden hashes over 500 frames: SILK WB mono/stereo, hybrid FB stereo, hybrid SWB mono."""
    e = GOLD["packet_level"][idx]
    ch = 2 if e["toc"] & 4 else 1
    data, frames = pkg.synth_corpus(1, 500, e["len"] - 1, e["toc"])
    dec = pkg.BatchDecoder(1, 500, channels=ch)
    rc, pcm, status = dec.decode_many(data, frames)
    assert rc == 0 and (status == 960).all()
    assert "%016x" % ref.fnv_hash(pcm, ch) == e["hash"], e
    dec.close()


def test_config3_silk_wb_mono_streams(pkg, ref):
    """BASELINE config 3 shape at oracle-sized scale: 256 SILK WB mono streams x 50 frames, 60 B packets."""
    n_streams, n_frames = 256, 50
    data, frames = pkg.synth_corpus(n_streams, n_frames, 59, 0x48)
    dec = pkg.BatchDecoder(n_streams, n_frames, channels=1)
    rc, pcm, status = dec.decode_many(data, frames)
    assert rc == 0 and (status == 960).all()
    for s in range(0, n_streams, 3):
        _, want, err = ref.decode_synth_stream(0x48, 60, 1, n_frames, seed=ref.SEED ^ s)
        assert err == 0
        assert np.array_equal(pcm[s].reshape(-1, 1), want), "stream %d" % s
    dec.close()


def test_config3_full_size_properties(pkg, ref):
    """BASELINE configs[2] at its full per-GPU size (16 384 SILK WB mono streams x 50 frames): every status is 960, a
    sample of streams spread over the batch is bit-exact against the oracle, and those streams decoded alone in a small
    context give the same PCM (a stream's output does not depend on its neighbours in the batch)."""
    n_streams, n_frames = 16384, 50
    data, frames = pkg.synth_corpus(n_streams, n_frames, 59, 0x48)
    dec = pkg.BatchDecoder(n_streams, n_frames, channels=1)
    rc, pcm, status = dec.decode_many(data, frames)
    assert rc == 0 and (status == 960).all()
    rows = list(range(0, n_streams, 512)) + [n_streams - 1]
    got = pcm[rows].copy()
    del pcm
    dec.close()
    for j, s in enumerate(rows[::4]):
        _, want, err = ref.decode_synth_stream(0x48, 60, 1, n_frames, seed=ref.SEED ^ s)
        assert err == 0
        assert np.array_equal(got[j * 4].reshape(-1, 1), want), "stream %d" % s
    small = pkg.BatchDecoder(len(rows), n_frames, channels=1)
    rc, pcm2, status2 = small.decode_many(data, np.ascontiguousarray(frames[rows]))
    assert rc == 0 and (status2 == 960).all()
    assert np.array_equal(pcm2, got)
    small.close()


def test_single_stream_dropin_silk_and_hybrid(pkg, ref):
    """opus_decode()-shaped single-stream API over mixed packets (ref src/opus_decoder.cpp:351)."""
    rng = np.random.default_rng(3)
    d = pkg.OpusDecoder(48000, 2)
    o = ref.PacketDecoder(2)
    for toc in [0x4C] * 4 + [0x7C] * 4 + [0xFC] * 3 + [0x6C] * 3 + [0x0C] * 3:
        pkt = np.concatenate([np.array([toc], np.uint8), rng.integers(0, 256, int(rng.integers(20, 200)), dtype=np.uint8)])
        r, want = o.decode(pkt)
        n, got = d.decode(pkt)
        assert n == r
        assert np.array_equal(got, want)
    o.close()
    d.close()


@pytest.mark.parametrize("ch", [1, 2])
def test_tiny_and_maximal_frames_every_mode(pkg, ref, ch):
    """Payloads of 0, 1, 2, 3 bytes (the reference fails or decodes silence-like frames), ordinary ones and the 1275-byte
    maximum, for SILK, hybrid and CELT TOCs, interleaved so that failures sit between good frames."""
    s = 4 if ch == 2 else 0
    rng = np.random.default_rng(40 + ch)
    lens = [0, 1, 2, 3, 5, 8, 60, 1275, 2, 100, 1, 0, 40, 1275, 3, 200]
    plans = []
    for toc in (0x08, 0x48, 0x68, 0x78, 0xF8, 0xD8):
        plan = []
        for k, L in enumerate(lens):
            plan.append((toc | s, -L if L == 0 else L))
        plans.append(plan)
    # _pack treats ln <= 0 as "random length": build by hand instead
    n_streams, n_frames = len(plans), len(lens)
    frames = pkg.make_frames(n_streams, n_frames)
    chunks, pos, pk = [], 0, []
    for i, plan in enumerate(plans):
        row = []
        for f, (toc, _) in enumerate(plan):
            L = lens[f]
            payload = rng.integers(0, 256, L, dtype=np.uint8)
            frames[i, f] = (pos, L, toc, 0)
            chunks.append(payload)
            pos += L
            row.append(np.concatenate([np.array([toc], np.uint8), payload]))
        pk.append(row)
    data = np.concatenate(chunks)
    dec = pkg.BatchDecoder(n_streams, n_frames, channels=ch)
    rc, pcm, status = dec.decode_many(data, frames)
    assert rc == 0
    dec.close()
    for i in range(n_streams):
        o = ref.PacketDecoder(ch)
        try:
            for f in range(n_frames):
                r, want = o.decode(pk[i][f])
                assert status[i, f] == r, (i, f, status[i, f], r)
                if r > 0:
                    assert np.array_equal(pcm[i, f], want), (i, f, hex(pk[i][f][0]), lens[f])
        finally:
            o.close()


def test_filler_and_continuation_frames(pkg, ref):
    """DMF_NULL fillers behind a stream's last frame change nothing (the next call carries on as if they were not there);
    a DMF_CONT frame behind a failed frame of its packet is not decoded (opus_decode_native returns at the first failing
    frame, src/opus_decoder.cpp:334-338)."""
    rng = np.random.default_rng(77)
    toc = 0xFC
    pay = [rng.integers(0, 256, 120, dtype=np.uint8) for _ in range(8)]
    data = np.concatenate(pay + [np.zeros(1, np.uint8)])
    off = np.cumsum([0] + [len(p) for p in pay])
    DMF_CONT, DMF_NULL = 1, 2
    # call 1: frames 0,1,2 real, 3 = a 1-byte frame that CELT refuses, 4 = continuation of that packet (must be skipped), 5 = filler
    fr1 = pkg.make_frames(1, 6)
    for f in range(3):
        fr1[0, f] = (off[f], 120, toc, 0)
    fr1[0, 3] = (off[8], 1, toc, 0)
    fr1[0, 4] = (off[3], 120, toc, DMF_CONT)
    fr1[0, 5] = (0, 0, toc, DMF_NULL)
    # call 2: frames 4..7 real
    fr2 = pkg.make_frames(1, 6)
    for k, f in enumerate(range(4, 8)):
        fr2[0, k] = (off[f], 120, toc, 0)
    fr2[0, 4] = (0, 0, toc, DMF_NULL)
    fr2[0, 5] = (0, 0, toc, DMF_NULL)
    dec = pkg.BatchDecoder(1, 6, channels=2)
    rc, p1, s1 = dec.decode_many(data, fr1)
    rc2, p2, s2 = dec.decode_many(data, fr2)
    dec.close()
    assert rc == 0 and rc2 == 0
    assert list(s1[0]) == [960, 960, 960, -18, -2004, -2004]
    assert list(s2[0]) == [960, 960, 960, 960, -2004, -2004]
    assert not p1[0, 4].any() and not p1[0, 5].any()
    o = ref.PacketDecoder(2)
    want = []
    for f in (0, 1, 2):
        want.append(o.decode(np.concatenate([[toc], pay[f]]).astype(np.uint8))[1])
    r, _ = o.decode(np.array([toc | 1, 0, 0], np.uint8))  # code 1, two 1-byte frames: the first fails, the second is never decoded
    assert r == -18
    for f in (4, 5, 6, 7):
        want.append(o.decode(np.concatenate([[toc], pay[f]]).astype(np.uint8))[1])
    o.close()
    got = [p1[0, 0], p1[0, 1], p1[0, 2], p2[0, 0], p2[0, 1], p2[0, 2], p2[0, 3]]
    for k in range(7):
        assert np.array_equal(got[k], want[k]), k


def test_decoder_ctl_and_multistream_match_reference(pkg, ref):
    """opus_decoder_ctl requests and the one-stream multistream wrapper the player creates (src/opusfile.cpp:693),
    against the reference's own functions (variadic C++ symbols called through ctypes)."""
    import ctypes

    R = ref.lib()
    rctl = getattr(R, "_Z16opus_decoder_ctlP11OpusDecoderiz")
    ms_create = getattr(R, "_Z31opus_multistream_decoder_createiiiiPKhPi")
    ms_decode = getattr(R, "_Z23opus_multistream_decodeP13OpusMSDecoderPhiPsi")
    ms_create.restype = ctypes.c_void_p
    ms_create.argtypes = [ctypes.c_int32, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    ms_decode.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int]
    rng = np.random.default_rng(91)
    GET_BW, GET_RANGE, RESET, GET_FS, SET_GAIN, GET_GAIN, GET_DUR = 4009, 4031, 4028, 4029, 4034, 4045, 4039

    def ref_get(st, req):
        v = ctypes.c_int32(-77)
        rc = rctl(ctypes.c_void_p(st), ctypes.c_int(req), ctypes.byref(v))
        return rc, v.value

    o = ref.PacketDecoder(2)
    d = pkg.OpusDecoder(48000, 2)
    for toc in (0xFC, 0xDC, 0x4C, 0x7C | 1, 0x9C):
        body = rng.integers(0, 256, 2 * int(rng.integers(20, 90)), dtype=np.uint8)
        pkt = np.concatenate([np.array([toc], np.uint8), body])
        r, want = o.decode(pkt)
        n, got = d.decode(pkt)
        assert n == r and np.array_equal(got, want)
        for req in (GET_BW, GET_RANGE, GET_FS, GET_GAIN, GET_DUR):
            assert d.ctl(req) == ref_get(o.st, req), req
    assert d.ctl(SET_GAIN, 1234)[0] == rctl(ctypes.c_void_p(o.st), ctypes.c_int(SET_GAIN), ctypes.c_int32(1234)) == 0
    assert d.ctl(GET_GAIN) == ref_get(o.st, GET_GAIN) == (0, 1234)
    assert d.ctl(SET_GAIN, 40000)[0] == -1
    # OPUS_RESET_STATE, then the same packets again: same PCM as the reference after its own reset
    assert d.ctl(RESET)[0] == rctl(ctypes.c_void_p(o.st), ctypes.c_int(RESET)) == 0
    for toc in (0x4C, 0xFC, 0x7C):
        pkt = np.concatenate([np.array([toc], np.uint8), rng.integers(0, 256, 100, dtype=np.uint8)])
        r, want = o.decode(pkt)
        n, got = d.decode(pkt)
        assert n == r and np.array_equal(got, want)
    o.close()
    d.close()
    # multistream: a coupled stream played swapped, a mono stream fanned out to two channels with a muted third
    for channels, coupled, mapping, toc in ((2, 1, [1, 0], 0xFC), (3, 0, [0, 255, 0], 0x48), (1, 0, [0], 0xF8)):
        m = np.array(mapping, np.uint8)
        err = ctypes.c_int(0)
        rst = ms_create(48000, channels, 1, coupled, m.ctypes.data, ctypes.addressof(err))
        assert rst and err.value == 0
        ours = pkg.OpusMSDecoder(48000, channels, 1, coupled, mapping)
        for _ in range(4):
            pkt = np.concatenate([np.array([toc], np.uint8), rng.integers(0, 256, 80, dtype=np.uint8)])
            want = np.zeros((960, channels), np.int16)
            r = ms_decode(rst, pkt.ctypes.data, pkt.size, want.ctypes.data, 960)
            n, got = ours.decode(pkt, 960)
            assert n == r == 960
            assert np.array_equal(got, want)
        ours.close()
        ctypes.CDLL(None).free(ctypes.c_void_p(rst))


def test_packets_with_a_toc_frame_size_other_than_20ms_are_refused_without_touching_the_buffer(pkg):
    """Every frame is decoded as 960 samples whatever the TOC says; the reference asserts / overruns on
    2.5 / 5 / 10 / 40 / 60 ms TOCs.  The packet API refuses them (OPUS_UNIMPLEMENTED) or reports OPUS_BUFFER_TOO_SMALL
    by the reference's own rule, and never writes past frame_size samples."""
    rng = np.random.default_rng(77)
    r = lambda n: rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    dec = pkg.OpusDecoder(48000, 2)
    L = pkg.lib()
    guard = 0x5A5A

    def call(packet, frame_size):
        p = np.frombuffer(packet, np.uint8).copy()
        pcm = np.full(8 * 5760 * 2, guard, np.int16)
        n = L.opusb_decode(dec._st, p.ctypes.data, p.size, pcm.ctypes.data, frame_size)
        return n, pcm

    for toc in (0xE4, 0xEC, 0xF4, 0x04, 0x54, 0x5C, 0x64):  # CELT 2.5/5/10 ms, SILK 10 ms, SILK 40/60 ms, hybrid 10 ms
        spf = pkg.opus_packet_get_samples_per_frame(bytes([toc]), 48000)
        assert spf != 960
        n, pcm = call(bytes([toc]) + r(60), 5760)
        assert n == -5 and (pcm == guard).all()
        n, pcm = call(bytes([toc]) + r(60), spf)       # the advisor's case: one short frame, frame_size = its TOC size
        assert n in (-5, -2) and (pcm == guard).all()
    # code 3, 48 x 2.5 ms frames, frame_size 5760
    n, pcm = call(bytes([0xE4 | 3, 48]) + r(48 * 10), 5760)
    assert n == -5 and (pcm == guard).all()
    # 20 ms TOC, 3 frames, buffer for 2: the reference's OPUS_BUFFER_TOO_SMALL (src/opus_decoder.cpp:320)
    n, pcm = call(bytes([0xFC | 3, 3]) + r(3 * 80), 1920)
    assert n == -2 and (pcm == guard).all()
    n, pcm = call(bytes([0xFC | 3, 3]) + r(3 * 80), 2880)
    assert n == 2880 and (pcm[2880 * 2:] == guard).all() and not (pcm[: 2880 * 2] == guard).all()
    # multistream wrapper: same rules (its scratch holds 5760 samples per channel)
    ms = pkg.OpusMSDecoder(48000, 2, 1, 1, [0, 1])
    n, _ = ms.decode(bytes([0xE4 | 3, 48]) + r(48 * 10))
    assert n == -5
    ms.close()
    dec.close()


@pytest.mark.parametrize("kind", ["silk_wb_mono", "silk_stereo", "hybrid_fb", "hybrid_swb_mono", "mode_walk_stereo"])
def test_three_async_calls_in_flight_silk_hybrid_and_mode_walks(pkg, ref, kind):
    """decode_many_async / dm_wait with THREE calls in flight (the side-stream SILK fork/join, double-buffered staging)
    on SILK-only, hybrid and mode-switching streams: every frame equals the reference, and equals the blocking call."""
    rng = np.random.default_rng(len(kind))
    S, F, steps = 48, 6, 7
    ch = 1 if "mono" in kind else 2
    s4 = 4 if ch == 2 else 0
    if kind == "mode_walk_stereo":
        tocs = [0xFC, 0x4C, 0x7C, 0x0C, 0x6C, 0xDC]
        plans = [[(int(rng.choice(tocs)), 0) for _ in range(F * steps)] for _ in range(S)]
    else:
        toc = {"silk_wb_mono": 0x48, "silk_stereo": 0x4C, "hybrid_fb": 0x7C, "hybrid_swb_mono": 0x68}[kind]
        plans = [[(toc, 0) for _ in range(F * steps)] for _ in range(S)]
    data, frames, pk = _pack(pkg, plans, 1000 + len(kind))
    dec = pkg.BatchDecoder(S, F, channels=ch)
    hb = pkg.api.PinnedArray(data.size, np.uint8)
    hb.array[:] = data
    hf = [pkg.api.PinnedArray((S, F), pkg.FRAME_DTYPE) for _ in range(steps)]
    DEPTH = 3
    outs = [(pkg.api.PinnedArray((S, F, 960, ch), np.int16), pkg.api.PinnedArray((S, F), np.int32)) for _ in range(DEPTH)]
    got_pcm = np.zeros((S, F * steps, 960, ch), np.int16)
    got_st = np.zeros((S, F * steps), np.int32)
    pending = []

    def retire():
        k, ticket = pending.pop(0)
        assert dec.wait(ticket) == 0
        got_pcm[:, k * F:(k + 1) * F] = outs[k % DEPTH][0].array
        got_st[:, k * F:(k + 1) * F] = outs[k % DEPTH][1].array

    for k in range(steps):
        hf[k].array[:] = frames[:, k * F:(k + 1) * F]
        if len(pending) == DEPTH:
            retire()
        o = outs[k % DEPTH]
        rc, ticket = dec.decode_many_async(hb.array, hf[k].array, o[0].array, o[1].array)
        assert rc == 0
        pending.append((k, ticket))
    while pending:
        retire()
    dec.close()
    for s in range(S):
        o = ref.PacketDecoder(ch)
        try:
            for f in range(F * steps):
                r, want = o.decode(pk[s][f])
                assert got_st[s, f] == r, (s, f)
                if r > 0:
                    assert np.array_equal(got_pcm[s, f], want), "stream %d frame %d toc %#x" % (s, f, pk[s][f][0])
        finally:
            o.close()


@pytest.mark.parametrize("kind", ["celt", "mode_walk"])
def test_device_calls_overlapping_head_and_tail(pkg, kind):
    """dm_decode_device_async: consecutive calls with explicit events, the tail of call k beside the head of call k + 1
    (two alternating sets of intermediates): same PCM as the blocking host-buffer calls, for CELT and for streams that
    walk through SILK / hybrid / CELT (SILK stage on its side stream)."""
    import torch

    rng = np.random.default_rng(5 + len(kind))
    S, F, steps = 192, 8, 7
    tocs = [0xFC] if kind == "celt" else [0xFC, 0x4C, 0x7C, 0x0C, 0x6C, 0xDC]
    plans = [[(int(rng.choice(tocs)), 0) for _ in range(F * steps)] for _ in range(S)]
    data, frames, _ = _pack(pkg, plans, 31)
    a = pkg.BatchDecoder(S, F, channels=2)
    want = []
    for k in range(steps):
        rc, pcm, st = a.decode_many(data, np.ascontiguousarray(frames[:, k * F:(k + 1) * F]))
        assert rc == 0
        want.append((pcm.copy(), st.copy()))
    a.close()
    b = pkg.BatchDecoder(S, F, channels=2)
    d_bytes = torch.from_numpy(data).cuda()
    d_frames = [torch.from_numpy(np.ascontiguousarray(frames[:, k * F:(k + 1) * F]).view(np.uint8).reshape(-1)).cuda() for k in range(steps)]
    d_pcm = [torch.empty(S * F * 960 * 2, dtype=torch.int16, device="cuda") for _ in range(steps)]
    d_st = [torch.empty(S * F, dtype=torch.int32, device="cuda") for _ in range(steps)]
    torch.cuda.synchronize()
    for k in range(steps):
        rc, done = b.decode_device_async(d_bytes.data_ptr(), data.size, d_frames[k].data_ptr(), F, d_pcm[k].data_ptr(), d_st[k].data_ptr())
        assert rc == 0 and done
    assert b.stream_wait(torch.cuda.current_stream().cuda_stream) == 0
    torch.cuda.synchronize()
    for k in range(steps):
        assert np.array_equal(d_st[k].cpu().numpy().reshape(S, F), want[k][1]), k
        assert np.array_equal(d_pcm[k].cpu().numpy().reshape(S, F, 960, 2), want[k][0]), k
    b.close()
