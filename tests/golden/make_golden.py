#!/usr/bin/env python3
"""(Re)generate the golden fixtures from the reference oracle (oracle/_ref, built from /root/reference).

  survey_hashes.json   verified, never rewritten: the hashes SURVEY.md 8c recorded for the unmodified reference
  celt_fb_stereo.npz   4 frames of stream 0 of BASELINE config 2 (TOC 0xFC, 160 B packets): packets + PCM
  celt_short.npz       CELT-layer frames of 120/240/480 samples, mono and stereo: frames + PCM + final rng
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref  # noqa: E402


def xorshift_bytes(seed, n):
    s, out, M = seed, np.empty(n, np.uint8), (1 << 64) - 1
    for i in range(n):
        s ^= (s << 13) & M
        s ^= s >> 7
        s ^= (s << 17) & M
        out[i] = (s >> 11) & 0xFF
    return out, s


def main():
    g = json.load(open(os.path.join(HERE, "survey_hashes.json")))
    for e in g["packet_level"]:
        if e["frames"] > 1000:
            continue
        h, _, err = ref.decode_synth_stream(e["toc"], e["len"], e["channels"], e["frames"], want_pcm=False)
        assert err == 0 and "%016x" % h == e["hash"], e
    for e in g["celt_layer"]:
        h, _, rng = ref.decode_synth_celt(e["frame_size"], e["channels"], e["len"], e["frames"], want_pcm=False)
        assert "%016x" % h == e["hash"] and "%08x" % rng == e["rng"], e
    print("survey hashes reproduced")
    payload, _ = xorshift_bytes(ref.SEED, 4 * 159)
    _, pcm, err = ref.decode_synth_stream(0xFC, 160, 2, 4)
    assert err == 0
    np.savez_compressed(os.path.join(HERE, "celt_fb_stereo.npz"), payload=payload.reshape(4, 159), toc=np.uint8(0xFC), pcm=pcm)
    short = {}
    for fs, ch, ln in [(120, 1, 20), (120, 2, 30), (240, 2, 40), (480, 1, 60), (480, 2, 80)]:
        frames, _ = xorshift_bytes(ref.SEED, 6 * ln)
        _, pcm, rng = ref.decode_synth_celt(fs, ch, ln, 6)
        short["f_%d_%d" % (fs, ch)] = frames.reshape(6, ln)
        short["p_%d_%d" % (fs, ch)] = pcm
        short["r_%d_%d" % (fs, ch)] = np.uint32(rng)
    np.savez_compressed(os.path.join(HERE, "celt_short.npz"), **short)
    print("fixtures written")


if __name__ == "__main__":
    main()
