"""ctypes binding of oracle/_ref/libopusref.so -- the UNMODIFIED reference decoder built for x86.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference arm.  The product (esp32-opus-player_b200) never imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_ref", "libopusref.so")
SEED = 0x9E3779B97F4A7C15
_P, _I = ctypes.c_void_p, ctypes.c_int
_lib = None


def build():
    """(Re)build the oracle when the reference tree is present; otherwise keep the shipped .so."""
    subprocess.run(["make", "-C", HERE, "-s"], check=True)
    return SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        L = ctypes.CDLL(SO)
        L.ref_opus_new.restype = _P
        L.ref_opus_new.argtypes = [_I]
        L.ref_opus_init.argtypes = [_P, _I]
        L.ref_opus_free.argtypes = [_P]
        L.ref_opus_decode.argtypes = [_P, _P, _I, _P, _I]
        L.ref_celt_decode.argtypes = [_P, _I, _P, _I]
        L.ref_ec_rng.restype = ctypes.c_uint32
        L.ref_celt_rng.restype = ctypes.c_uint32
        L.ref_celt_peek_energies.argtypes = [_P]
        L.ref_celt_peek_mem.argtypes = [_I, _P]
        L.ref_celt_peek_misc.argtypes = [_P]
        L.ref_decode_synth_stream.restype = ctypes.c_uint64
        L.ref_decode_synth_stream.argtypes = [_I, _I, _I, _I, ctypes.c_uint64, _P, _P]
        L.ref_decode_synth_celt.restype = ctypes.c_uint64
        L.ref_decode_synth_celt.argtypes = [_I, _I, _I, _I, ctypes.c_uint64, _P, _P]
        L.ref_bench_fork.restype = ctypes.c_double
        L.ref_bench_fork.argtypes = [_I, _I, _I, _I, _I, ctypes.c_uint64, _I, _P]
        L.ref_bench_fork_ids.restype = ctypes.c_double
        L.ref_bench_fork_ids.argtypes = [_I, _I, _I, _I, _P, _I, _I, ctypes.c_uint64, _I, _P]
        L.ref_ogg_open.argtypes = [_P, ctypes.c_size_t]
        L.ref_ogg_read_stereo.argtypes = [_P, _I]
        L.ref_gen_packet.argtypes = [_P, _I, _I, _P]
        _lib = L
    return _lib


def fnv_hash(pcm, channels):
    """Survey hash: FNV-1a over uint16 samples, interleaved L,R (mono duplicated)."""
    a = np.ascontiguousarray(pcm, np.int16).reshape(-1, channels)
    if channels == 1:
        a = np.repeat(a, 2, axis=1)
    h = 1469598103934665603
    for v in a.reshape(-1).view(np.uint16).tolist():
        h = ((h ^ v) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def decode_synth_stream(toc, length, channels, n_frames, seed=SEED, want_pcm=True):
    """Packet-level reference decode of one synthetic stream -> (hash, pcm[n_frames*960][channels], err)."""
    pcm = np.zeros((n_frames * 960, channels), np.int16) if want_pcm else None
    err = _I(0)
    h = lib().ref_decode_synth_stream(toc, length, channels, n_frames, seed, pcm.ctypes.data if want_pcm else None,
                                      ctypes.cast(ctypes.byref(err), _P))
    return h, pcm, err.value


def decode_synth_celt(frame_size, channels, length, n_frames, seed=SEED, want_pcm=True):
    """CELT-layer reference decode (all bytes random) -> (hash, pcm, final_rng)."""
    pcm = np.zeros((n_frames * frame_size, channels), np.int16) if want_pcm else None
    rng = ctypes.c_uint32(0)
    h = lib().ref_decode_synth_celt(frame_size, channels, length, n_frames, seed, pcm.ctypes.data if want_pcm else None,
                                    ctypes.cast(ctypes.byref(rng), _P))
    return h, pcm, rng.value


class CeltLayer:
    """celt_decoder_init + celt_decode_with_ec on caller-supplied frames (global singleton state!)."""

    def __init__(self, channels, stream_channels=None, start_band=0):
        L = lib()
        L.ref_celt_init(channels)
        L.ref_celt_set_stream_channels(stream_channels or channels)
        if start_band:
            L.ref_celt_set_start_band(start_band)
        self.channels = channels

    def decode(self, frame, frame_size):
        frame = np.ascontiguousarray(frame, np.uint8)
        pcm = np.zeros((frame_size, self.channels), np.int16)
        n = lib().ref_celt_decode(frame.ctypes.data, frame.size, pcm.ctypes.data, frame_size)
        return n, pcm


class PacketDecoder:
    """opus_decoder_init + opus_decode on caller-supplied packets (global singleton state!)."""

    def __init__(self, channels):
        self.channels = channels
        self.st = lib().ref_opus_new(channels)

    def decode(self, packet):
        packet = np.ascontiguousarray(packet, np.uint8)
        pcm = np.zeros((5760, self.channels), np.int16)
        n = lib().ref_opus_decode(self.st, packet.ctypes.data, packet.size, pcm.ctypes.data, 5760)
        return n, pcm[: max(n, 0)]

    def close(self):
        if self.st:
            lib().ref_opus_free(self.st)
            self.st = None


def bench_fork(toc, length, channels, n_streams, n_frames, seed_base=SEED, jobs=None):
    """CPU baseline: fork one worker per job, contiguous share of streams each. Returns (wall_s, hashes)."""
    jobs = jobs or os.cpu_count() or 1
    hashes = np.zeros(n_streams, np.uint64)
    wall = lib().ref_bench_fork(toc, length, channels, n_streams, n_frames, seed_base, jobs, hashes.ctypes.data)
    return wall, hashes


def bench_fork_ids(toc, length, channels, ids, n_frames, frame_size=0, seed_base=SEED, jobs=None):
    """bench_fork for an explicit list of stream ids; frame_size != 0 decodes at the CELT layer (raw frames of `length`
    random bytes).  Returns (wall_s, hashes)."""
    jobs = jobs or os.cpu_count() or 1
    ids = np.ascontiguousarray(ids, np.int64)
    hashes = np.zeros(ids.size, np.uint64)
    wall = lib().ref_bench_fork_ids(toc, length, channels, frame_size, ids.ctypes.data, ids.size, n_frames, seed_base, jobs, hashes.ctypes.data)
    return wall, hashes
