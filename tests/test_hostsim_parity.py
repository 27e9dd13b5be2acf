"""The kernels' stage bodies (csrc/*.cuh), compiled for the host by tests/hostsim, against the oracle.
This is a logic check of the very source the GPU runs (lanes = 1); the GPU parity tests proper are in
test_gpu_celt.py.  CPU only."""
import numpy as np
import pytest


@pytest.mark.parametrize("fs,ch,C,ln,n", [(960, 2, 2, 160, 120), (960, 1, 1, 100, 60), (480, 2, 2, 80, 60), (240, 2, 2, 40, 60),
                                           (120, 1, 1, 20, 60), (120, 2, 2, 2, 40), (960, 2, 2, 1275, 12), (960, 2, 1, 90, 40),
                                           (960, 1, 2, 90, 40), (960, 2, 2, 0, 80)])
def test_celt_frame_bit_exact(hostsim, ref, fs, ch, C, ln, n):
    rng = np.random.default_rng(fs * 7 + ch * 3 + C + ln)
    layer = ref.CeltLayer(ch, C)
    d = hostsim.hs_celt_new(ch)
    got = np.zeros(960 * 2, np.int16)
    for f in range(n):
        L = ln if ln > 0 else int(rng.integers(2, 500))
        frame = rng.integers(0, 256, L, dtype=np.uint8)
        r, want = layer.decode(frame, fs)
        s = hostsim.hs_celt_decode(d, frame.ctypes.data, L, C, fs, 0, got.ctypes.data, None)
        assert r == s
        assert np.array_equal(got[: fs * ch].reshape(fs, ch), want), "frame %d" % f
    hostsim.hs_celt_free(d)


def test_parse_final_range_matches(hostsim, ref):
    import ctypes

    rng = np.random.default_rng(11)
    ref.CeltLayer(2, 2)
    hdr = np.zeros(512, np.uint8)
    ops = np.zeros(1360 * 16, np.uint8)
    X = np.zeros(1920, np.int16)
    pcm = np.zeros(1920, np.int16)
    R = ref.lib()
    for f in range(200):
        L = int(rng.integers(2, 700))
        frame = rng.integers(0, 256, L, dtype=np.uint8)
        R.ref_celt_decode(frame.ctypes.data, L, pcm.ctypes.data, 960)
        hostsim.hs_celt_parse(frame.ctypes.data, L, 2, 3, 0, 0, hdr.ctypes.data, ops.ctypes.data, X.ctypes.data)
        assert int(hdr[:4].view(np.uint32)[0]) == R.ref_ec_rng()


# ---- Opus frame layer: SILK-only, hybrid, CELT-only and the transitions between them ----
def _opus_run(hostsim, ref, ch, plan, seed):
    """plan: list of (toc, payload_len, n_frames). Decodes the same packets with the oracle and the host simulation."""
    rng = np.random.default_rng(seed)
    dec = ref.PacketDecoder(ch)
    d = hostsim.hs_opus_new(ch)
    got = np.zeros(960 * 2, np.int16)
    k = 0
    try:
        for toc, ln, n in plan:
            for _ in range(n):
                L = ln if ln > 0 else int(rng.integers(3, 400))
                payload = rng.integers(0, 256, L, dtype=np.uint8)
                pkt = np.concatenate([np.array([toc], np.uint8), payload])
                r, want = dec.decode(pkt)
                s = hostsim.hs_opus_decode(d, toc, payload.ctypes.data, L, got.ctypes.data)
                assert r == s, "frame %d: status %d vs %d" % (k, r, s)
                if r > 0:
                    assert np.array_equal(got[: 960 * ch].reshape(960, ch), want), "frame %d toc %#x" % (k, toc)
                k += 1
    finally:
        dec.close()
        hostsim.hs_opus_free(d)


@pytest.mark.parametrize("toc,ch,ln", [(0x08, 1, 40), (0x28, 1, 50), (0x48, 1, 60), (0x4C, 2, 80), (0x0C, 2, 60), (0x2C, 2, 70),
                                       (0x68, 1, 100), (0x6C, 2, 120), (0x78, 1, 110), (0x7C, 2, 160), (0x48, 1, 0), (0x7C, 2, 0)])
def test_opus_single_mode_bit_exact(hostsim, ref, toc, ch, ln):
    _opus_run(hostsim, ref, ch, [(toc, ln, 60)], toc * 131 + ln)


@pytest.mark.parametrize("ch", [1, 2])
def test_opus_mode_switching_bit_exact(hostsim, ref, ch):
    s = 4 if ch == 2 else 0
    plan = [(0xF8 | s, 100, 6), (0x48 | s, 60, 6), (0x78 | s, 120, 6), (0x48 | s, 60, 5), (0x08 | s, 40, 5), (0x28 | s, 50, 5),
            (0x68 | s, 100, 5), (0xF8 | s, 100, 4), (0x78 | s, 120, 4), (0xD8 | s, 90, 4), (0x48 | s, 0, 8), (0x78 | s, 0, 8),
            (0x08 | s, 0, 6), (0xF8 | s, 0, 6)]
    _opus_run(hostsim, ref, ch, plan, 77 + ch)


def test_narrowed_parser_arrays_hold_their_values(hostsim):
    """The parser's per-band work arrays are int8 / int16 / uint16 (celt_parse.cuh `narrow`); the host build aborts the
    process if a value does not fit.  Parse-only stress over the extremes: maximum and minimum frame sizes, every frame
    size, mono and stereo, hybrid start band, random payloads (so boosts, trims and skip decisions of every kind occur)."""
    rng = np.random.default_rng(2024)
    hdr = np.zeros(512, np.uint8)
    ops = np.zeros(1360 * 16, np.uint8)
    X = np.zeros(1920, np.int16)
    n = 0
    for LM in range(4):
        for C in (1, 2):
            for start in (0, 17):
                for L in (1275, 1000, 640, 320, 160, 60, 20, 9, 3, 2):
                    for _ in range(40 if L >= 640 else 15):
                        frame = rng.integers(0, 256, L, dtype=np.uint8)
                        if rng.integers(0, 4) == 0:  # bias towards all-ones / all-zeros runs: long boosts, extreme trims
                            frame[: int(rng.integers(1, L + 1))] = int(rng.choice([0, 255]))
                        hostsim.hs_celt_parse(frame.ctypes.data, L, C, LM, start, 0, hdr.ctypes.data, ops.ctypes.data, X.ctypes.data)
                        n += 1
    assert n >= 3000
