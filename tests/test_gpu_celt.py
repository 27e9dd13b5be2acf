"""GPU parity tests: the CUDA path through the C ABI against the oracle (the unmodified reference)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "survey_hashes.json")))


def xorshift_bytes(seed, n):
    s, out, M = seed, np.empty(n, np.uint8), (1 << 64) - 1
    for i in range(n):
        s ^= (s << 13) & M
        s ^= s >> 7
        s ^= (s << 17) & M
        out[i] = (s >> 11) & 0xFF
    return out


def test_golden_fixture_through_c_abi(pkg):
    fx = np.load(os.path.join(HERE, "golden", "celt_fb_stereo.npz"))
    data = fx["payload"].reshape(-1).copy()
    frames = pkg.make_frames(1, 4)
    for f in range(4):
        frames[0, f] = (f * 159, 159, int(fx["toc"]), 0)
    dec = pkg.BatchDecoder(1, 4, channels=2)
    rc, pcm, status = dec.decode_many(data, frames)
    assert rc == 0 and (status == 960).all()
    assert np.array_equal(pcm.reshape(-1, 2), fx["pcm"])
    dec.close()


def test_golden_short_frames(pkg):
    fx = np.load(os.path.join(HERE, "golden", "celt_short.npz"))
    for fs, ch in [(120, 1), (120, 2), (240, 2), (480, 1), (480, 2)]:
        fr = fx["f_%d_%d" % (fs, ch)]
        n, ln = fr.shape
        frames = pkg.make_frames(1, n)
        for f in range(n):
            frames[0, f] = (f * ln, ln, 0, 0)
        dec = pkg.BatchDecoder(1, n, channels=ch)
        rc, pcm, status = dec.decode_many(fr.reshape(-1).copy(), frames, mode=pkg.DM_MODE_CELT_RAW, frame_size=fs, coded_channels=ch)
        assert rc == 0 and (status == fs).all()
        assert np.array_equal(pcm.reshape(-1, ch), fx["p_%d_%d" % (fs, ch)]), (fs, ch)
        dec.close()


def test_config2_streams_bit_exact(pkg, ref):
    """BASELINE config 2 shape at a size the oracle finishes in seconds: 96 streams x 50 frames, 160 B packets."""
    n_streams, n_frames = 96, 50
    data, frames = pkg.synth_corpus(n_streams, n_frames, 159, 0xFC)
    dec = pkg.BatchDecoder(n_streams, n_frames, channels=2)
    rc, pcm, status = dec.decode_many(data, frames)
    assert rc == 0 and (status == 960).all()
    for s in range(n_streams):
        h, want, err = ref.decode_synth_stream(0xFC, 160, 2, n_frames, seed=ref.SEED ^ s)
        assert err == 0
        assert np.array_equal(pcm[s].reshape(-1, 2), want), "stream %d" % s
    assert dec.launch_count > 0
    assert dec.leaf_overflow() == 0  # the sorted leaf path ran, not the fallback
    dec.close()


def test_config2_full_size_properties(pkg, ref):
    """BASELINE configs[1] at its full per-GPU size (4096 streams x 50 frames, two consecutive calls), checked through
    properties that do not need the oracle on every stream: (1) a sample of streams spread over the batch is bit-exact
    against the oracle across both calls, (2) stream i's PCM does not depend on the batch it is decoded in -- a decoder
    holding only every 64th stream reproduces those rows exactly, (3) every status is 960."""
    n_streams, n_frames, calls = 4096, 50, 2
    data, frames = pkg.synth_corpus(n_streams, n_frames * calls, 159, 0xFC)
    dec = pkg.BatchDecoder(n_streams, n_frames, channels=2)
    rows = list(range(0, n_streams, 64)) + [n_streams - 1]
    got = np.zeros((len(rows), n_frames * calls, 960, 2), np.int16)
    for k in range(calls):
        rc, pcm, status = dec.decode_many(data, np.ascontiguousarray(frames[:, k * n_frames:(k + 1) * n_frames]))
        assert rc == 0 and (status == 960).all()
        got[:, k * n_frames:(k + 1) * n_frames] = pcm[rows]
    dec.close()
    for j, s in enumerate(rows[::8]):
        _, want, err = ref.decode_synth_stream(0xFC, 160, 2, n_frames * calls, seed=ref.SEED ^ s)
        assert err == 0
        assert np.array_equal(got[j * 8].reshape(-1, 2), want), "stream %d" % s
    small = pkg.BatchDecoder(len(rows), n_frames * calls, channels=2)
    sub = np.ascontiguousarray(frames[rows])
    rc, pcm2, status2 = small.decode_many(data, sub)
    assert rc == 0 and (status2 == 960).all()
    assert np.array_equal(pcm2, got)
    small.close()


def test_leaf_list_overflow_falls_back(pkg, ref, monkeypatch):
    """The sorted leaf list has a fixed capacity per frame; a step with more coded leaves than it holds must take the
    band kernel's own leaf expansion.  OPUSB_LEAF_SLOTS=1 forces that path; the result is the same PCM."""
    monkeypatch.setenv("OPUSB_LEAF_SLOTS", "1")
    n_streams, n_frames = 40, 12
    data, frames = pkg.synth_corpus(n_streams, n_frames, 159, 0xFC)
    dec = pkg.BatchDecoder(n_streams, n_frames, channels=2)
    rc, pcm, status = dec.decode_many(data, frames)
    assert dec.leaf_overflow() == 1
    dec.close()
    assert rc == 0 and (status == 960).all()
    for s in range(0, n_streams, 5):
        _, want, err = ref.decode_synth_stream(0xFC, 160, 2, n_frames, seed=ref.SEED ^ s)
        assert err == 0
        assert np.array_equal(pcm[s].reshape(-1, 2), want), "stream %d" % s


def test_survey_hash_celt_fb_stereo_500_frames(pkg, ref):
    e = GOLD["packet_level"][0]
    data, frames = pkg.synth_corpus(1, 500, e["len"] - 1, e["toc"])
    dec = pkg.BatchDecoder(1, 500, channels=2)
    rc, pcm, status = dec.decode_many(data, frames)
    assert rc == 0
    assert "%016x" % ref.fnv_hash(pcm, 2) == e["hash"]
    dec.close()


def test_state_carries_across_calls(pkg, ref):
    """3000-frame stream decoded as 60 calls of 50 frames == the reference's single pass (state carry)."""
    n_calls, per = 12, 50
    data, frames = pkg.synth_corpus(2, n_calls * per, 159, 0xFC)
    dec = pkg.BatchDecoder(2, per, channels=2)
    out = []
    for k in range(n_calls):
        fr = np.ascontiguousarray(frames[:, k * per:(k + 1) * per])
        rc, pcm, status = dec.decode_many(data, fr)
        assert rc == 0 and (status == 960).all()
        out.append(pcm.copy())
    got = np.concatenate(out, axis=1)
    for s in range(2):
        _, want, _ = ref.decode_synth_stream(0xFC, 160, 2, n_calls * per, seed=ref.SEED ^ s)
        assert np.array_equal(got[s].reshape(-1, 2), want)
    dec.close()


@pytest.mark.parametrize("fs,ch,C,ln", [(960, 2, 2, 160), (960, 1, 1, 100), (480, 2, 2, 80), (240, 2, 2, 40), (120, 1, 1, 20),
                                        (120, 2, 2, 2), (960, 2, 2, 8), (960, 2, 2, 1275), (960, 2, 1, 90), (960, 1, 2, 90),
                                        (240, 2, 1, 50), (480, 1, 2, 60), (960, 2, 2, 0)])
def test_celt_layer_all_sizes(pkg, ref, fs, ch, C, ln):
    """CELT-layer parity: every frame size, mono/stereo and mono<->stereo coded/decoded mixes,
    smallest and largest frames, ragged lengths (ln == 0)."""
    n_streams, n_frames = 5, 30
    rng = np.random.default_rng(fs + 10 * ch + C + ln)
    lens = np.full((n_streams, n_frames), ln) if ln else rng.integers(2, 600, (n_streams, n_frames))
    frames = pkg.make_frames(n_streams, n_frames)
    chunks, pos = [], 0
    for s in range(n_streams):
        for f in range(n_frames):
            b = rng.integers(0, 256, int(lens[s, f]), dtype=np.uint8)
            chunks.append(b)
            frames[s, f] = (pos, len(b), 0, 0)
            pos += len(b)
    data = np.concatenate(chunks)
    dec = pkg.BatchDecoder(n_streams, n_frames, channels=ch)
    rc, pcm, status = dec.decode_many(data, frames, mode=pkg.DM_MODE_CELT_RAW, frame_size=fs, coded_channels=C)
    assert rc == 0
    for s in range(n_streams):
        layer = ref.CeltLayer(ch, C)
        for f in range(n_frames):
            fr = frames[s, f]
            n, want = layer.decode(data[fr["offset"]:fr["offset"] + fr["len"]], fs)
            assert status[s, f] == n
            assert np.array_equal(pcm[s, f], want), (s, f)
    dec.close()


def test_bad_frames_do_not_poison_the_batch(pkg, ref):
    """len <= 1 fails like the reference (ERR_OPUS_CELT_BAD_ARG, src/celt.cpp:2225) and leaves the stream's
    state untouched; neighbouring streams are unaffected."""
    n_streams, n_frames = 4, 6
    data, frames = pkg.synth_corpus(n_streams, n_frames, 159, 0xFC)
    frames[1, 2]["len"] = 1
    frames[2, 0]["len"] = 0
    dec = pkg.BatchDecoder(n_streams, n_frames, channels=2)
    rc, pcm, status = dec.decode_many(data, frames, mode=pkg.DM_MODE_CELT_RAW, frame_size=960, coded_channels=2)
    assert rc == 0
    assert status[1, 2] == -18 and status[2, 0] == -18
    for s in range(n_streams):
        layer = ref.CeltLayer(2, 2)
        for f in range(n_frames):
            fr = frames[s, f]
            n, want = layer.decode(data[fr["offset"]:fr["offset"] + fr["len"]], 960)
            assert status[s, f] == n
            if n > 0:
                assert np.array_equal(pcm[s, f], want), (s, f)
    dec.close()


def test_chunked_pipeline_and_determinism(pkg, ref):
    """>= 256 streams takes the multi-stream (4 CUDA streams) host pipeline; run twice, identical, and
    spot-check against the oracle."""
    n_streams, n_frames = 320, 8
    data, frames = pkg.synth_corpus(n_streams, n_frames, 159, 0xFC)
    outs = []
    for _ in range(2):
        dec = pkg.BatchDecoder(n_streams, n_frames, channels=2)
        rc, pcm, status = dec.decode_many(data, frames)
        assert rc == 0 and (status == 960).all()
        outs.append(pcm.copy())
        dec.close()
    assert np.array_equal(outs[0], outs[1])
    for s in (0, 79, 80, 159, 160, 239, 240, 319):
        _, want, _ = ref.decode_synth_stream(0xFC, 160, 2, n_frames, seed=ref.SEED ^ s)
        assert np.array_equal(outs[0][s].reshape(-1, 2), want), s


def test_reset_subset(pkg, ref):
    data, frames = pkg.synth_corpus(3, 5, 159, 0xFC)
    dec = pkg.BatchDecoder(3, 5, channels=2)
    rc, first, _ = dec.decode_many(data, frames)
    dec.reset([1])
    rc, second, _ = dec.decode_many(data, frames)
    assert np.array_equal(second[1], first[1])          # stream 1 restarted from a fresh decoder
    assert not np.array_equal(second[0], first[0])      # stream 0 kept its state
    dec.close()


def test_single_stream_drop_in(pkg, ref):
    """opus_decode() mirror: packet in, 960 samples out, same PCM as the reference's opus_decode."""
    dec = pkg.OpusDecoder(48000, 2)
    rdec = ref.PacketDecoder(2)
    payload = xorshift_bytes(ref.SEED, 6 * 159).reshape(6, 159)
    for f in range(6):
        pkt = np.concatenate([[0xFC], payload[f]]).astype(np.uint8)
        n, pcm = dec.decode(pkt)
        rn, rpcm = rdec.decode(pkt)
        assert n == rn == 960
        assert np.array_equal(pcm, rpcm)
    assert dec.decode(np.zeros(0, np.uint8))[0] == -1
    dec.close()
    rdec.close()


def test_device_resident_path(pkg, ref):
    import torch

    n_streams, n_frames = 16, 10
    data, frames = pkg.synth_corpus(n_streams, n_frames, 159, 0xFC)
    dec = pkg.BatchDecoder(n_streams, n_frames, channels=2)
    d_bytes = torch.from_numpy(data).cuda()
    d_frames = torch.from_numpy(frames.view(np.uint8).reshape(-1)).cuda()
    d_pcm = torch.zeros(n_streams * n_frames * 960 * 2, dtype=torch.int16, device="cuda")
    d_status = torch.zeros(n_streams * n_frames, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    rc = dec.decode_device(d_bytes.data_ptr(), data.size, d_frames.data_ptr(), n_frames, d_pcm.data_ptr(), d_status.data_ptr(), stream)
    assert rc == 0
    torch.cuda.synchronize()
    pcm = d_pcm.cpu().numpy().reshape(n_streams, n_frames * 960, 2)
    assert (d_status.cpu().numpy() == 960).all()
    for s in range(n_streams):
        _, want, _ = ref.decode_synth_stream(0xFC, 160, 2, n_frames, seed=ref.SEED ^ s)
        assert np.array_equal(pcm[s], want)
    times = dec.stage_times()
    assert set(times) >= {"entropy", "bands", "imdct", "postfilter", "deemph"} and all(v >= 0 for v in times.values())
    dec.close()


def test_async_calls_in_flight_match_blocking_calls(pkg):
    """decode_many_async/dm_wait with two calls in flight == the same steps through blocking decode_many."""
    S, F, steps = 256, 10, 6
    data, frames = pkg.synth_corpus(S, F * steps, 159, 0xFC)
    a = pkg.BatchDecoder(S, F, channels=2)
    want = []
    for k in range(steps):
        rc, pcm, st = a.decode_many(data, np.ascontiguousarray(frames[:, k * F:(k + 1) * F]))
        assert rc == 0 and (st == 960).all()
        want.append(pcm.copy())
    a.close()
    b = pkg.BatchDecoder(S, F, channels=2)
    hb = pkg.api.PinnedArray(data.size, np.uint8)
    hb.array[:] = data
    hf = [pkg.api.PinnedArray((S, F), pkg.FRAME_DTYPE) for _ in range(steps)]
    outs = [(pkg.api.PinnedArray((S, F, 960, 2), np.int16), pkg.api.PinnedArray((S, F), np.int32)) for _ in range(2)]
    pending = None
    got = []
    for k in range(steps):
        hf[k].array[:] = frames[:, k * F:(k + 1) * F]
        o = outs[k & 1]
        rc, ticket = b.decode_many_async(hb.array, hf[k].array, o[0].array, o[1].array)
        assert rc == 0
        if pending is not None:
            assert b.wait(pending[0]) == 0
            got.append(pending[1][0].array.copy())
        pending = (ticket, o)
    assert b.wait(pending[0]) == 0
    got.append(pending[1][0].array.copy())
    for k in range(steps):
        assert np.array_equal(got[k], want[k]), k
    b.close()


def test_script_capacity_is_a_true_worst_case(pkg, ref):
    """OPUSB_SCRIPT_OVERFLOW (-1000) is a status the reference never returns: the per-frame script (kMaxOps entries) must
    hold the worst case.  Maximum-size frames (1275 B: the most bits per band, the deepest splits) at every frame size,
    mono and stereo, random payloads (transient 1/8, so short blocks with B = 8 occur), several hundred frames each:
    no frame overflows and every one equals the reference."""
    rng = np.random.default_rng(99)
    worst = 0
    for fs, ch, ln in [(960, 2, 1275), (960, 1, 1275), (480, 2, 1275), (240, 2, 1275), (120, 2, 1275), (120, 1, 1275), (960, 2, 900), (480, 1, 1000)]:
        S, F = 24, 12
        data = rng.integers(0, 256, S * F * ln, dtype=np.uint8)
        frames = pkg.make_frames(S, F)
        frames["offset"] = (np.arange(S * F, dtype=np.uint32) * ln).reshape(S, F)
        frames["len"] = ln
        dec = pkg.BatchDecoder(S, F, channels=ch)
        rc, pcm, status = dec.decode_many(data, frames, mode=pkg.DM_MODE_CELT_RAW, frame_size=fs, coded_channels=ch)
        assert rc == 0
        assert not (status == -1000).any(), (fs, ch, ln)
        dec.close()
        for s in range(0, S, 5):
            o = ref.CeltLayer(ch)
            for f in range(F):
                n, want = o.decode(data[(s * F + f) * ln:(s * F + f + 1) * ln], fs)
                assert status[s, f] == n
                if n > 0:
                    assert np.array_equal(pcm[s, f], want), (fs, ch, ln, s, f)


@pytest.mark.parametrize("segs,warm", [(4, 8), (1, 8), (8, 8), (2, 3), (4, 1), (8, 1)])
def test_deemphasis_segments(pkg, ref, monkeypatch, segs, warm):
    """The de-emphasis chains of a 20 ms call run as `segs` runs of frames side by side, each warmed up on `warm` x 32 samples
    of the frame in front of it and checked against the serial order afterwards (k_deemph / k_deemph_fix).  Bit-exact PCM
    and carried state (two calls) whatever the cut; with the full warm-up (8 tiles) no segment of a stream of full CELT
    frames is done twice, with a 32-sample warm-up the memories have not met and the redo path does the work."""
    monkeypatch.setenv("OPUSB_DEEMPH_SEGS", str(segs))
    monkeypatch.setenv("OPUSB_DEEMPH_WARM", str(warm))
    n_streams, n_frames, calls = 70, 19, 2   # 140 chains: the last warp of a segment is ragged; 19 frames: uneven segments
    data, frames = pkg.synth_corpus(n_streams, n_frames * calls, 159, 0xFC)
    dec = pkg.BatchDecoder(n_streams, n_frames, channels=2)
    got = []
    for k in range(calls):
        rc, pcm, status = dec.decode_many(data, np.ascontiguousarray(frames[:, k * n_frames:(k + 1) * n_frames]))
        assert rc == 0 and (status == 960).all()
        got.append(pcm.copy())
    redone = dec.deemph_redone()
    dec.close()
    got = np.concatenate(got, axis=1)
    for s in range(n_streams):
        _, want, err = ref.decode_synth_stream(0xFC, 160, 2, n_frames * calls, seed=ref.SEED ^ s)
        assert err == 0
        assert np.array_equal(got[s].reshape(-1, 2), want), "stream %d" % s
    if segs == 1 or warm == 8:
        assert redone == 0
    if segs > 1 and warm == 1:
        assert redone > 0


def test_deemphasis_segments_behind_frames_without_celt_output(pkg, ref):
    """Segments that begin behind a SILK-only frame (no CELT samples to warm up on) or a hybrid -> SILK fade-out frame (120
    samples) are done in serial order by k_deemph_fix: mono and stereo mode walks, 24 frames per call."""
    for ch in (1, 2):
        s4 = 4 if ch == 2 else 0
        rng = np.random.default_rng(40 + ch)
        tocs = [0xF8 | s4, 0x48 | s4, 0x78 | s4, 0xF8 | s4]
        n_streams, n_frames = 40, 24
        plans = [[int(rng.choice(tocs)) for _ in range(n_frames)] for _ in range(n_streams)]
        frames = pkg.make_frames(n_streams, n_frames)
        chunks, pos, pk = [], 0, []
        for s in range(n_streams):
            row = []
            for f in range(n_frames):
                L = int(rng.integers(20, 200))
                payload = rng.integers(0, 256, L, dtype=np.uint8)
                frames[s, f] = (pos, L, plans[s][f], 0)
                chunks.append(payload)
                pos += L
                row.append(np.concatenate([np.array([plans[s][f]], np.uint8), payload]))
            pk.append(row)
        dec = pkg.BatchDecoder(n_streams, n_frames, channels=ch)
        rc, pcm, status = dec.decode_many(np.concatenate(chunks), frames)
        assert rc == 0
        assert dec.deemph_redone() > 0
        dec.close()
        for s in range(n_streams):
            o = ref.PacketDecoder(ch)
            try:
                for f in range(n_frames):
                    r, want = o.decode(pk[s][f])
                    assert status[s, f] == r
                    if r > 0:
                        assert np.array_equal(pcm[s, f], want), (ch, s, f)
            finally:
                o.close()
