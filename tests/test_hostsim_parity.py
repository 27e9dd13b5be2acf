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
