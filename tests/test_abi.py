"""The C-ABI library: loads on a CPU-only box, exports every symbol include/*.h declares,
refuses to decode without a device, and its host-side packet helpers match the reference."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in sorted(os.listdir(os.path.join(ROOT, "include"))):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
        src = re.sub(r"typedef[^;]*\(\s*\*[^;]*;", " ", src)
        names += re.findall(r"\b([a-z_][a-z0-9_]*)\s*\([^;{]*\)\s*;", src)
    return sorted(set(n for n in names if n.startswith(("dm_", "decode_many", "opusb_", "opfb_"))))


def test_exports_every_declared_symbol(pkg):
    L = pkg.lib()
    syms = declared_symbols()
    assert len(syms) >= 42, syms
    for s in syms:
        assert hasattr(L, s), "missing export: " + s


def test_no_cpu_fallback(pkg):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        pkg.BatchDecoder(4, 2)
    with pytest.raises(RuntimeError):
        pkg.OpusDecoder(48000, 2)


def test_frame_struct_layout(pkg):
    assert pkg.FRAME_DTYPE.itemsize == 8


def test_packet_helpers_match_reference(pkg, ref):
    R = ref.lib()
    rng = np.random.default_rng(3)
    P, I = ctypes.c_void_p, ctypes.c_int
    # mangled C++ symbols of the reference helpers (src/opus_decoder.cpp:460-557)
    bw = getattr(R, "_Z25opus_packet_get_bandwidthPh")
    spf = getattr(R, "_Z33opus_packet_get_samples_per_framePhi")
    nch = getattr(R, "_Z27opus_packet_get_nb_channelsPh")
    nfr = getattr(R, "_Z25opus_packet_get_nb_framesPhi")
    parse = getattr(R, "_Z17opus_packet_parsePhiS_PS_PsPi")
    for f in (bw, nch):
        f.argtypes = [P]
    spf.argtypes = [P, I]
    nfr.argtypes = [P, I]
    parse.argtypes = [P, I, P, P, P, P]
    for _ in range(600):
        n = int(rng.integers(1, 80))
        pkt = rng.integers(0, 256, n, dtype=np.uint8)
        if rng.random() < 0.5:
            pkt[0] = (pkt[0] & 0xFC) | 3
            if n > 1:
                pkt[1] = (pkt[1] & 0xC0) | int(rng.integers(0, 6))
        assert pkg.opus_packet_get_bandwidth(pkt) == bw(pkt.ctypes.data)
        assert pkg.opus_packet_get_samples_per_frame(pkt) == spf(pkt.ctypes.data, 48000)
        assert pkg.opus_packet_get_nb_channels(pkt) == nch(pkt.ctypes.data)
        assert pkg.opus_packet_get_nb_frames(pkt) == nfr(pkt.ctypes.data, n)
        toc = ctypes.c_uint8(0)
        sizes = (ctypes.c_int16 * 48)()
        frames = (P * 48)()
        off = I(0)
        want = parse(pkt.ctypes.data, n, ctypes.cast(ctypes.byref(toc), P), ctypes.cast(frames, P), ctypes.cast(sizes, P),
                     ctypes.cast(ctypes.byref(off), P))
        got, gtoc, gfr = pkg.opus_packet_parse(pkt)
        assert got == want
        if want > 0:
            assert gtoc == toc.value
            assert [s for _, s in gfr] == [sizes[i] for i in range(want)]
            assert [o for o, _ in gfr] == [frames[i] - pkt.ctypes.data for i in range(want)]


def test_synth_corpus_matches_oracle_generator(pkg, ref):
    data, frames = pkg.synth_corpus(3, 5, 159, 0xFC)
    R = ref.lib()
    for s in range(3):
        st = ctypes.c_uint64(ref.SEED ^ s)
        for f in range(5):
            buf = np.zeros(160, np.uint8)
            R.ref_gen_packet(ctypes.cast(ctypes.byref(st), ctypes.c_void_p), 0xFC, 160, buf.ctypes.data)
            fr = frames[s, f]
            assert fr["len"] == 159 and fr["toc"] == 0xFC
            assert np.array_equal(data[fr["offset"]:fr["offset"] + 159], buf[1:])
