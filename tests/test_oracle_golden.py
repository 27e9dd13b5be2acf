"""The oracle (oracle/_ref = the unmodified reference built for x86) against the survey's golden hashes
and the committed fixtures.  CPU only."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "survey_hashes.json")))


def test_packet_level_hashes(ref):
    for e in GOLD["packet_level"]:
        n = min(e["frames"], 500)
        if n != e["frames"]:
            continue  # the 6000-frame case is covered by make_golden.py / bench, keep the CPU suite short
        h, _, err = ref.decode_synth_stream(e["toc"], e["len"], e["channels"], n, want_pcm=False)
        assert err == 0
        assert "%016x" % h == e["hash"], e["what"]


def test_celt_layer_hashes(ref):
    for e in GOLD["celt_layer"]:
        h, _, rng = ref.decode_synth_celt(e["frame_size"], e["channels"], e["len"], e["frames"], want_pcm=False)
        assert "%016x" % h == e["hash"], e
        assert "%08x" % rng == e["rng"], e


def test_fixture_matches_oracle(ref):
    fx = np.load(os.path.join(HERE, "golden", "celt_fb_stereo.npz"))
    dec = ref.PacketDecoder(2)
    for f in range(4):
        pkt = np.concatenate([[fx["toc"]], fx["payload"][f]]).astype(np.uint8)
        n, pcm = dec.decode(pkt)
        assert n == 960
        assert np.array_equal(pcm, fx["pcm"][f * 960:(f + 1) * 960])
    dec.close()


def test_hash_helper_matches_c_side(ref):
    h, pcm, _ = ref.decode_synth_stream(0xFC, 160, 2, 3)
    assert ref.fnv_hash(pcm, 2) == h
    h, pcm, _ = ref.decode_synth_stream(0xF8, 100, 1, 3)
    assert ref.fnv_hash(pcm, 1) == h
