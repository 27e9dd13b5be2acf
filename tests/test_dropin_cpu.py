"""The same-signature drop-in: libopusb200_dropin.so exports the reference's own C++ entry points
(include/opus_decoder.h, include/opusfile.h).  CPU-only checks: the library loads, every declared function is exported
under the Itanium-mangled name the reference's objects import, the reference's UNMODIFIED opusfile.o / ogg.o link against
it in place of opus_decoder.cpp + celt.cpp + silk.cpp, and without a GPU the entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import subprocess
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin"))
import build_harness as bh  # noqa: E402

# mangled names: what `nm oracle/_ref/opus_decoder.o celt.o silk.o opusfile.o` shows for the reference's definitions
# (src/opus_decoder.h:165-218, src/celt.h:662-719, src/silk.h:1043-1044, src/opusfile.h:141-156)
PACKET_API = [
    "_Z21opus_decoder_get_sizei", "_Z17opus_decoder_initP11OpusDecoderii", "_Z11opus_decodeP11OpusDecoderPhiPsi",
    "_Z16opus_decoder_ctlP11OpusDecoderiz", "_Z20opus_decoder_destroyP11OpusDecoder", "_Z17opus_packet_parsePhiS_PS_PsPi",
    "_Z25opus_packet_get_bandwidthPh", "_Z33opus_packet_get_samples_per_framePhi", "_Z27opus_packet_get_nb_channelsPh",
    "_Z25opus_packet_get_nb_framesPhi", "_Z26opus_packet_get_nb_samplesPhii", "_Z27opus_decoder_get_nb_samplesPK11OpusDecoderPhi",
    "_Z33opus_multistream_decoder_get_sizeii", "_Z31opus_multistream_decoder_createiiiiPKhPi",
    "_Z23opus_multistream_decodeP13OpusMSDecoderPhiPsi", "_Z28opus_multistream_decoder_ctlP13OpusMSDecoderiz",
    "_Z32opus_multistream_decoder_destroyP13OpusMSDecoder",
]
CODEC_API = [
    "_Z27CELTDecoder_AllocateBuffersv", "_Z23CELTDecoder_FreeBuffersv", "_Z27SILKDecoder_AllocateBuffersv", "_Z23SILKDecoder_FreeBuffersv",
    "_Z17celt_decoder_initi", "_Z16celt_decoder_ctliz", "_Z11ec_dec_initPhj", "_Z19celt_decode_with_ecPsi",
]
PLAYER_API = ["_Z17opus_init_decoderv", "_Z14op_read_stereoPsi", "_Z15opus_head_parseP8OpusHeadPhm"]


def _dyn_symbols(path, defined=True):
    out = subprocess.run(["nm", "-D", "--defined-only" if defined else "--undefined-only", path], check=True, capture_output=True, text=True).stdout
    return {ln.split()[-1] for ln in out.splitlines() if ln.strip()}


def test_dropin_exports_the_reference_symbols(pkg):
    pkg.lib()
    so = os.path.join(bh.LIBDIR, "libopusb200_dropin.so")
    ctypes.CDLL(so)  # loads (pulls in libopusb200.so through its rpath)
    have = _dyn_symbols(so)
    missing = [s for s in PACKET_API + CODEC_API + PLAYER_API if s not in have]
    assert not missing, missing


def test_mangled_names_are_the_references_own():
    """The names above are not typed in from memory: wherever the reference objects are present they are checked
    against what the reference itself defines."""
    objs = [os.path.join(bh.ROOT, "oracle", "_ref", f) for f in ("opus_decoder.o", "celt.o", "silk.o", "opusfile.o")]
    if not all(os.path.exists(o) for o in objs):
        pytest.skip("oracle/_ref objects not built")
    defined = set()
    for o in objs:
        out = subprocess.run(["nm", "--defined-only", o], check=True, capture_output=True, text=True).stdout
        defined |= {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    missing = [s for s in PACKET_API + CODEC_API + PLAYER_API if s not in defined]
    assert not missing, missing


def test_reference_player_objects_link_against_the_dropin(pkg):
    """opusfile.o + ogg.o of the unmodified reference need opus_multistream_*, opus_packet_get_*, *_AllocateBuffers
    from the decoder: the drop-in supplies all of them (an executable does not link with an unresolved symbol)."""
    pkg.lib()
    exe = bh.harness_reflink()
    if exe is None:
        pytest.skip("oracle/_ref objects not built")
    need = _dyn_symbols(exe, defined=False)
    assert "_Z23opus_multistream_decodeP13OpusMSDecoderPhiPsi" in need  # resolved at load time from the drop-in
    assert "_Z31opus_multistream_decoder_createiiiiPKhPi" in need


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present: covered by test_gpu_dropin.py")
def test_no_cpu_fallback_behind_the_dropin(pkg, tmp_path):
    pkg.lib()
    exe = bh.harness_dropin()
    out = tmp_path / "o.bin"
    r = subprocess.run([exe, "packet", "0xFC", "160", "2", "2", "1", str(out)], capture_output=True, text=True)
    assert r.returncode == 4 and "no usable CUDA device" in r.stderr
    r = subprocess.run([exe, "celt", "480", "2", "80", "2", "1", str(out)], capture_output=True, text=True)
    assert r.returncode == 4
