"""GPU parity of the same-signature drop-in, each case in a separate process so that its symbols cannot
clash with the oracle's:

* harness_reflink: the reference's UNMODIFIED opusfile.o + ogg.o (its own Ogg demux and player logic) linked against
  libopusb200_dropin.so in place of opus_decoder.cpp / celt.cpp / silk.cpp -- op_read_stereo return values and PCM must
  equal the full reference's on the player corpus of test_gpu_player.py;
* harness_dropin: opus_init_decoder / op_read_stereo, opus_decoder_get_size / init / opus_decode / ctl and the CELT-layer
  entry points (celt_decoder_init, ec_dec_init, celt_decode_with_ec: short frames) all from the drop-in."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oggutil as og
from test_gpu_player import mixed_packets
from test_player_demux import ref_returns

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin"))
import build_harness as bh  # noqa: E402

pytestmark = pytest.mark.gpu


def run_player(exe, blob, tmp_path, buf_size=2048):
    src, out = tmp_path / "in.ogg", tmp_path / "out.bin"
    src.write_bytes(bytes(blob))
    r = subprocess.run([exe, "player", str(src), str(out), str(buf_size)], capture_output=True, text=True, timeout=300)
    assert r.returncode in (0, 4), r.stderr
    raw = out.read_bytes()
    opened, nr = np.frombuffer(raw, np.int32, 2)
    ns = int(np.frombuffer(raw, np.int64, 1, 8)[0])
    rets = np.frombuffer(raw, np.int32, int(nr), 16).tolist()
    pcm = np.frombuffer(raw, np.int16, ns, 16 + 4 * int(nr))
    return bool(opened), rets, pcm


def corpus(case):
    rng = np.random.default_rng(100 + len(case))
    r = lambda n: rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    if case == "celt":
        return og.simple_file(og.rand_packets(rng, 0xFC, 40), per_page=5), 2048
    if case == "pre_skip_trim_small_buf":
        return og.simple_file(og.rand_packets(rng, 0xFC, 20), per_page=4, pre_skip=3840, end_trim=777), 500
    if case == "mono_silk":
        return og.simple_file(og.rand_packets(rng, 0x48, 40, 30, 90), channels=1, per_page=7, pre_skip=312), 2048
    if case == "mixed_modes":
        return og.simple_file(mixed_packets(rng, True, 60), per_page=6, pre_skip=400, end_trim=200), 2048
    if case == "chained":
        a = og.simple_file(mixed_packets(rng, True, 15), pre_skip=200, per_page=4, serial=1)
        b = og.simple_file(og.rand_packets(rng, 0xFC, 10), pre_skip=960, per_page=3, serial=2, end_trim=300)
        return a + b, 2048
    assert case == "multi_frame_and_failing"
    pk = og.rand_packets(rng, 0xFC, 3)
    pk.append(bytes([0xFC | 1]) + r(2 * 90))
    pk.append(bytes([0xFC | 3, 0x80 | 3, 60, 1]) + r(60) + r(1) + r(88))  # frame 2 undecodable: OP_EBADPACKET
    pk += og.rand_packets(rng, 0xFC, 2)
    return og.simple_file(pk, per_page=3, durations=[960] * 3 + [1920, 2880] + [960] * 2, pre_skip=500), 2048


CASES = ["celt", "pre_skip_trim_small_buf", "mono_silk", "mixed_modes", "chained", "multi_frame_and_failing"]


@pytest.mark.parametrize("case", CASES)
def test_reference_player_code_over_the_dropin_decoder(pkg, ref, tmp_path, case):
    pkg.lib()
    exe = bh.harness_reflink()
    if exe is None:
        pytest.skip("oracle/_ref objects not built")
    blob, buf = corpus(case)
    want_r, want_pcm = ref_returns(ref, blob, buf)
    opened, got_r, got_pcm = run_player(exe, blob, tmp_path, buf)
    assert opened and got_r == want_r
    assert np.array_equal(got_pcm, want_pcm)


@pytest.mark.parametrize("case", CASES)
def test_dropin_player_api(pkg, ref, tmp_path, case):
    pkg.lib()
    blob, buf = corpus(case)
    want_r, want_pcm = ref_returns(ref, blob, buf)
    opened, got_r, got_pcm = run_player(bh.harness_dropin(), blob, tmp_path, buf)
    assert opened and got_r == want_r
    assert np.array_equal(got_pcm, want_pcm)


def test_dropin_player_refuses_garbage(pkg, tmp_path):
    pkg.lib()
    opened, rets, pcm = run_player(bh.harness_dropin(), b"this is not an ogg file" * 50, tmp_path)
    assert not opened and rets == [] and pcm.size == 0


@pytest.mark.parametrize("toc,length,ch", [(0xFC, 160, 2), (0x48, 60, 1), (0x7C, 160, 2), (0x4C, 80, 2), (0xF8, 100, 1)])
def test_dropin_packet_api(pkg, ref, tmp_path, toc, length, ch):
    pkg.lib()
    n = 12
    out = tmp_path / "p.bin"
    r = subprocess.run([bh.harness_dropin(), "packet", hex(toc), str(length), str(ch), str(n), str(ref.SEED), str(out)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    raw = out.read_bytes()
    hdr = np.frombuffer(raw, np.int32, 10)
    ns = int(np.frombuffer(raw, np.int64, 1, 40)[0])
    rets = np.frombuffer(raw, np.int32, int(hdr[0]), 48).tolist()
    pcm = np.frombuffer(raw, np.int16, ns, 48 + 4 * int(hdr[0])).reshape(-1, ch)
    assert rets == [960] * (n + 2)
    _, want, err = ref.decode_synth_stream(toc, length, ch, n)
    assert err == 0
    assert np.array_equal(pcm[: n * 960], want)
    assert np.array_equal(pcm[n * 960:], want[: 2 * 960])  # opus_decoder_init on the same blob = a fresh stream
    bw = {0: 1101, 1: 1102, 2: 1103, 3: 1104}
    want_bw = (1105 if toc & 0x10 else 1104) if (toc & 0xE0) == 0x60 else (max(1101, 1102 + ((toc >> 5) & 3)) if toc & 0x80 else bw[(toc >> 5) & 3])
    if toc & 0x80 and want_bw == 1102:
        want_bw = 1101
    # ctl: bandwidth, last packet duration, sample rate, final range (always 0 in the reference), unknown request, nb_samples
    assert hdr[1:].tolist() == [0, want_bw, 0, 960, 0, 48000, 0, -5, 960]


@pytest.mark.parametrize("frame_size,ch,length", [(120, 2, 30), (240, 2, 40), (480, 2, 80), (960, 2, 160), (120, 1, 20), (480, 1, 60)])
def test_dropin_celt_layer_short_frames(pkg, ref, tmp_path, frame_size, ch, length):
    """celt_decode_with_ec through the drop-in against the reference's CELT layer, all four frame sizes."""
    pkg.lib()
    n = 40
    out = tmp_path / "c.bin"
    r = subprocess.run([bh.harness_dropin(), "celt", str(frame_size), str(ch), str(length), str(n), str(ref.SEED), str(out)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    raw = out.read_bytes()
    hdr = np.frombuffer(raw, np.int32, 2)
    ns = int(np.frombuffer(raw, np.int64, 1, 8)[0])
    rets = np.frombuffer(raw, np.int32, int(hdr[0]), 16).tolist()
    pcm = np.frombuffer(raw, np.int16, ns, 16 + 4 * int(hdr[0])).reshape(-1, ch)
    _, want, rng = ref.decode_synth_celt(frame_size, ch, length, n)
    assert rets == [frame_size] * n
    assert np.array_equal(pcm, want)
    assert int(np.uint32(hdr[1])) == rng
