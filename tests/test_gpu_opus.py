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
