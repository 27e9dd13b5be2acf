"""ctypes binding of the C ABI in include/opus_b200.h.

Names follow the reference's packet API (src/opus_decoder.h:165-176): opus_decode, opus_packet_get_*;
`BatchDecoder.decode_many` is the batched call the reference lacks.  Errors are the reference's
negative ints; this layer raises only for misuse (wrong array shapes) and for a missing library.
"""
import ctypes
import os

import numpy as np

from .build import LIB, build

DM_MODE_OPUS = 0
DM_MODE_CELT_RAW = 1

FRAME_DTYPE = np.dtype([("offset", "<u4"), ("len", "<u2"), ("toc", "u1"), ("flags", "u1")])

_P = ctypes.c_void_p
_I = ctypes.c_int


class _Batch(ctypes.Structure):
    _fields_ = [("n_frames", _I), ("mode", _I), ("frame_size", _I), ("coded_channels", _I), ("start_band", _I),
                ("bytes", _P), ("n_bytes", ctypes.c_size_t), ("frames", _P)]


class _Out(ctypes.Structure):
    _fields_ = [("pcm", _P), ("status", _P)]


_lib = None


def lib():
    """Load libopusb200.so (building it if needed). Fails loudly if it cannot be built or loaded."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        build()
    L = ctypes.CDLL(os.environ.get("OPUSB_LIB") or LIB)  # OPUSB_LIB: a build variant of the same library (tuning sweeps)
    L.dm_create.argtypes = [ctypes.POINTER(_P), _I, _I, _I, _I]
    L.dm_destroy.argtypes = [_P]
    L.dm_reset.argtypes = [_P, _P, _I]
    L.decode_many.argtypes = [_P, ctypes.POINTER(_Batch), ctypes.POINTER(_Out)]
    L.decode_many_async.argtypes = [_P, ctypes.POINTER(_Batch), ctypes.POINTER(_Out), ctypes.POINTER(ctypes.c_longlong)]
    L.dm_wait.argtypes = [_P, ctypes.c_longlong]
    L.dm_debug_leaf_overflow.argtypes = [_P]
    L.dm_debug_deemph_redone.argtypes = [_P]
    L.dm_debug_deemph_redone.restype = ctypes.c_longlong
    L.dm_decode_device.argtypes = [_P, ctypes.POINTER(_Batch), ctypes.POINTER(_Out), _P]
    L.dm_probe_d2h.argtypes = [_P, _P, ctypes.c_size_t, _I, ctypes.POINTER(ctypes.c_float)]
    L.dm_host_alloc.restype = _P
    L.dm_host_alloc.argtypes = [ctypes.c_size_t]
    L.dm_host_free.argtypes = [_P]
    L.dm_launch_count.restype = ctypes.c_longlong
    L.dm_launch_count.argtypes = [_P]
    L.dm_device_bytes.restype = ctypes.c_size_t
    L.dm_device_bytes.argtypes = [_P]
    L.dm_decode_device_async.argtypes = [_P, ctypes.POINTER(_Batch), ctypes.POINTER(_Out), _P, ctypes.POINTER(_P)]
    L.dm_stream_wait.argtypes = [_P, _P]
    L.dm_debug_trace.argtypes = [_P, _I, _P]
    L.dm_stage_times.argtypes = [_P, _P, _I]
    L.dm_stage_name.restype = ctypes.c_char_p
    L.dm_stage_name.argtypes = [_I]
    L.dm_get_stream_state.argtypes = [_P, _I, _P, ctypes.c_size_t]
    L.dm_stream_state_size.restype = ctypes.c_size_t
    L.dm_synth_bytes.restype = ctypes.c_size_t
    L.dm_synth_bytes.argtypes = [_I, _I, _I]
    L.dm_synth_fill.argtypes = [_P, _P, _I, _I, _I, _I, ctypes.c_uint64, _I]
    L.dm_synth_fill_raw.argtypes = [_P, _P, _I, _I, _I, ctypes.c_uint64, _I]
    L.dm_fnv_update.argtypes = [_P, _P, _I, ctypes.c_size_t, _I, _I, _I]
    for fn in ("opusb_packet_get_mode", "opusb_packet_get_bandwidth", "opusb_packet_get_nb_channels"):
        getattr(L, fn).argtypes = [_P]
    L.opusb_packet_get_samples_per_frame.argtypes = [_P, ctypes.c_int32]
    L.opusb_packet_get_nb_frames.argtypes = [_P, ctypes.c_int32]
    L.opusb_packet_get_nb_samples.argtypes = [_P, ctypes.c_int32, ctypes.c_int32]
    L.opusb_packet_parse.argtypes = [_P, ctypes.c_int32, _P, _P, _P, _P]
    L.opusb_decoder_create.restype = _P
    L.opusb_decoder_create.argtypes = [ctypes.c_int32, _I, ctypes.POINTER(_I)]
    L.opusb_decoder_destroy.argtypes = [_P]
    L.opusb_decoder_init.argtypes = [_P, ctypes.c_int32, _I]
    L.opusb_decode.argtypes = [_P, _P, ctypes.c_int32, _P, _I]
    L.opusb_decoder_ctl.argtypes = [_P, _I, ctypes.POINTER(ctypes.c_int32)]
    L.opusb_multistream_decoder_create.restype = _P
    L.opusb_multistream_decoder_create.argtypes = [ctypes.c_int32, _I, _I, _I, _P, ctypes.POINTER(_I)]
    L.opusb_multistream_decode.argtypes = [_P, _P, ctypes.c_int32, _P, _I]
    L.opusb_multistream_decoder_ctl.argtypes = [_P, _I, ctypes.POINTER(ctypes.c_int32)]
    L.opusb_multistream_decoder_destroy.argtypes = [_P]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data if isinstance(a, np.ndarray) else int(a)


class PinnedArray:
    """numpy view over cudaHostAlloc memory (dm_host_alloc)."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(int(x) for x in np.atleast_1d(shape))
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self.ptr = lib().dm_host_alloc(max(self.nbytes, 1))
        if not self.ptr:
            raise MemoryError("dm_host_alloc(%d) failed" % self.nbytes)
        buf = (ctypes.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().dm_host_free(self.ptr)
            self.ptr = None


def make_frames(n_streams, n_frames):
    return np.zeros((n_streams, n_frames), dtype=FRAME_DTYPE)


def synth_corpus(n_streams, n_frames, payload_len, toc, seed_base=0x9E3779B97F4A7C15, first_stream=0, bytes_out=None, frames_out=None):
    """SURVEY 8c/8d synthetic streams: stream i = xorshift64(seed_base ^ i) payload bytes behind `toc`."""
    L = lib()
    nb = L.dm_synth_bytes(n_streams, n_frames, payload_len)
    b = bytes_out if bytes_out is not None else np.empty(nb, np.uint8)
    f = frames_out if frames_out is not None else make_frames(n_streams, n_frames)
    if toc is None or toc < 0:  # raw CELT frames: no TOC, every byte is generator output
        L.dm_synth_fill_raw(_ptr(b), _ptr(f), n_streams, n_frames, payload_len, ctypes.c_uint64(seed_base), first_stream)
    else:
        L.dm_synth_fill(_ptr(b), _ptr(f), n_streams, n_frames, payload_len, toc, ctypes.c_uint64(seed_base), first_stream)
    return b, f


FNV_START = 1469598103934665603


def fnv_update(hashes, pcm, channels, dup_mono=True, n_threads=0):
    """Carry the SURVEY 8c stream hashes (uint64[n_streams], start FNV_START) over one more call's PCM
    (int16[n_streams][...], C-contiguous) on host threads."""
    assert hashes.dtype == np.uint64 and pcm.dtype == np.int16 and pcm.flags["C_CONTIGUOUS"] and pcm.shape[0] == hashes.size
    lib().dm_fnv_update(hashes.ctypes.data, pcm.ctypes.data, hashes.size, pcm.size // hashes.size, channels, 1 if dup_mono else 0, n_threads)
    return hashes


class BatchDecoder:
    """n_streams independent decoders on one GPU (dm_ctx)."""

    def __init__(self, n_streams, max_frames, channels=2, device=0):
        self.n_streams, self.max_frames, self.channels, self.device = n_streams, max_frames, channels, device
        self._ctx = _P()
        rc = lib().dm_create(ctypes.byref(self._ctx), n_streams, max_frames, channels, device)
        if rc != 0:
            raise RuntimeError("dm_create failed: %d (no CUDA device => no decode path)" % rc)

    def close(self):
        if self._ctx:
            lib().dm_destroy(self._ctx)
            self._ctx = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, stream_ids=None):
        if stream_ids is None:
            return lib().dm_reset(self._ctx, None, 0)
        ids = np.ascontiguousarray(stream_ids, np.int32)
        return lib().dm_reset(self._ctx, ids.ctypes.data, len(ids))

    def _batch(self, bytes_ptr, n_bytes, frames_ptr, n_frames, mode, frame_size, coded_channels, start_band):
        b = _Batch()
        b.n_frames, b.mode, b.frame_size, b.coded_channels, b.start_band = n_frames, mode, frame_size, coded_channels, start_band
        b.bytes, b.n_bytes, b.frames = bytes_ptr, n_bytes, frames_ptr
        return b

    def decode_many(self, data, frames, mode=DM_MODE_OPUS, frame_size=960, coded_channels=2, start_band=0, pcm=None, status=None):
        """Host buffers in, host buffers out. data: uint8[n]; frames: FRAME_DTYPE[n_streams][n_frames].
        Returns (rc, pcm int16[n_streams][n_frames][frame_size][channels], status int32[n_streams][n_frames])."""
        frames = np.ascontiguousarray(frames)
        assert frames.dtype == FRAME_DTYPE and frames.shape[0] == self.n_streams
        n_frames = frames.shape[1]
        fs = frame_size if mode == DM_MODE_CELT_RAW else 960
        if pcm is None:
            pcm = np.zeros((self.n_streams, n_frames, fs, self.channels), np.int16)
        if status is None:
            status = np.zeros((self.n_streams, n_frames), np.int32)
        data = np.ascontiguousarray(data, np.uint8)
        b = self._batch(_ptr(data), data.size, _ptr(frames), n_frames, mode, frame_size, coded_channels, start_band)
        o = _Out(_ptr(pcm), _ptr(status))
        rc = lib().decode_many(self._ctx, ctypes.byref(b), ctypes.byref(o))
        return rc, pcm, status

    def decode_many_async(self, data, frames, pcm, status, mode=DM_MODE_OPUS, frame_size=960, coded_channels=2, start_band=0):
        """Enqueue one call and return (rc, ticket) at once; `wait(ticket)` blocks until pcm/status (pinned host arrays
        owned by the caller, one set per call in flight) are filled.  data/frames must stay alive until then."""
        assert frames.dtype == FRAME_DTYPE and frames.shape[0] == self.n_streams and frames.flags["C_CONTIGUOUS"]
        b = self._batch(_ptr(data), data.size, _ptr(frames), frames.shape[1], mode, frame_size, coded_channels, start_band)
        o = _Out(_ptr(pcm), _ptr(status))
        t = ctypes.c_longlong(0)
        rc = lib().decode_many_async(self._ctx, ctypes.byref(b), ctypes.byref(o), ctypes.byref(t))
        return rc, t.value

    def wait(self, ticket):
        return lib().dm_wait(self._ctx, ticket)

    def decode_device(self, d_bytes, n_bytes, d_frames, n_frames, d_pcm, d_status, cuda_stream=0, mode=DM_MODE_OPUS, frame_size=960,
                      coded_channels=2, start_band=0):
        """Device pointers (ints), enqueued on cuda_stream; no synchronisation."""
        b = self._batch(d_bytes, n_bytes, d_frames, n_frames, mode, frame_size, coded_channels, start_band)
        o = _Out(d_pcm, d_status)
        return lib().dm_decode_device(self._ctx, ctypes.byref(b), ctypes.byref(o), cuda_stream)

    def probe_d2h(self, host_array, nbytes, reps=5):
        """ms per bare D2H copy of nbytes from the context's PCM buffer into host_array (pinned)."""
        ms = ctypes.c_float(0)
        rc = lib().dm_probe_d2h(self._ctx, _ptr(host_array), nbytes, reps, ctypes.byref(ms))
        if rc:
            raise RuntimeError("dm_probe_d2h rc=%d" % rc)
        return ms.value / reps

    def decode_device_async(self, d_bytes, n_bytes, d_frames, n_frames, d_pcm, d_status, ready_event=None, mode=DM_MODE_OPUS, frame_size=960,
                            coded_channels=2, start_band=0):
        """Device pointers (ints); starts behind `ready_event` (a cudaEvent_t handle or None), returns (rc, done_event).
        Consecutive calls overlap: the tail of one call runs beside the head of the next."""
        b = self._batch(d_bytes, n_bytes, d_frames, n_frames, mode, frame_size, coded_channels, start_band)
        o = _Out(d_pcm, d_status)
        done = _P()
        rc = lib().dm_decode_device_async(self._ctx, ctypes.byref(b), ctypes.byref(o), ready_event, ctypes.byref(done))
        return rc, done.value

    def stream_wait(self, cuda_stream=0):
        """Make `cuda_stream` wait for every call enqueued on this context so far."""
        return lib().dm_stream_wait(self._ctx, cuda_stream)

    def trace(self, n):
        """Stage boundaries (ms) of the last n pipelined calls, float32[n][9]; needs OPUSB_TRACE=1 at creation."""
        out = np.zeros((n, 9), np.float32)
        rc = lib().dm_debug_trace(self._ctx, n, out.ctypes.data)
        if rc:
            raise RuntimeError("dm_debug_trace rc=%d" % rc)
        return out

    def stage_times(self):
        ms = (ctypes.c_float * 16)()
        n = lib().dm_stage_times(self._ctx, ctypes.cast(ms, _P), 16)
        return {lib().dm_stage_name(i).decode(): float(ms[i]) for i in range(n)}

    @property
    def launch_count(self):
        return int(lib().dm_launch_count(self._ctx))

    def deemph_redone(self):
        """Segments of de-emphasis chains done a second time since the decoder was created (dm_debug_deemph_redone)."""
        return int(lib().dm_debug_deemph_redone(self._ctx))

    def leaf_overflow(self):
        """1 if the last call took the band kernel's own leaf expansion (sorted leaf list full), else 0."""
        return int(lib().dm_debug_leaf_overflow(self._ctx))

    @property
    def device_bytes(self):
        return int(lib().dm_device_bytes(self._ctx))

    def stream_state(self, stream):
        n = lib().dm_stream_state_size()
        buf = np.zeros(n, np.uint8)
        rc = lib().dm_get_stream_state(self._ctx, stream, buf.ctypes.data, n)
        if rc:
            raise RuntimeError("dm_get_stream_state: %d" % rc)
        return buf


# ---- packet helpers with the reference's names (src/opus_decoder.h:169-176) ----
def _pkt(data):
    return np.ascontiguousarray(np.frombuffer(bytes(data), np.uint8) if not isinstance(data, np.ndarray) else data, np.uint8)


def opus_packet_get_mode(data):
    return lib().opusb_packet_get_mode(_pkt(data).ctypes.data)


def opus_packet_get_bandwidth(data):
    return lib().opusb_packet_get_bandwidth(_pkt(data).ctypes.data)


def opus_packet_get_nb_channels(data):
    return lib().opusb_packet_get_nb_channels(_pkt(data).ctypes.data)


def opus_packet_get_samples_per_frame(data, Fs=48000):
    return lib().opusb_packet_get_samples_per_frame(_pkt(data).ctypes.data, Fs)


def opus_packet_get_nb_frames(data):
    p = _pkt(data)
    return lib().opusb_packet_get_nb_frames(p.ctypes.data, p.size)


def opus_packet_get_nb_samples(data, Fs=48000):
    p = _pkt(data)
    return lib().opusb_packet_get_nb_samples(p.ctypes.data, p.size, Fs)


def opus_packet_parse(data):
    """Returns (count or negative error, toc, [(offset, size)...])."""
    p = _pkt(data)
    toc = ctypes.c_uint8(0)
    sizes = (ctypes.c_int16 * 48)()
    frames = (_P * 48)()
    off = _I(0)
    n = lib().opusb_packet_parse(p.ctypes.data, p.size, ctypes.cast(ctypes.byref(toc), _P), ctypes.cast(frames, _P), ctypes.cast(sizes, _P),
                                 ctypes.cast(ctypes.byref(off), _P))
    if n < 0:
        return n, 0, []
    base = p.ctypes.data
    return n, toc.value, [((frames[i] or base) - base, sizes[i]) for i in range(n)]


class OpusDecoder:
    """Single-stream drop-in for the reference's opus_decoder_init/opus_decode (src/opus_decoder.cpp:82, 351)."""

    def __init__(self, Fs=48000, channels=2):
        err = _I(0)
        self._st = lib().opusb_decoder_create(Fs, channels, ctypes.byref(err))
        self.channels = channels
        if not self._st:
            raise RuntimeError("opus_decoder_create failed: %d" % err.value)

    def init(self, Fs=48000, channels=None):
        return lib().opusb_decoder_init(self._st, Fs, channels or self.channels)

    def decode(self, data, frame_size=5760):
        """Returns (samples per channel or negative error, pcm int16[samples][channels])."""
        p = _pkt(data)
        pcm = np.zeros((frame_size, self.channels), np.int16)
        n = lib().opusb_decode(self._st, p.ctypes.data, p.size, pcm.ctypes.data, frame_size)
        return n, pcm[: max(n, 0)]

    def ctl(self, request, value=None):
        """opus_decoder_ctl (src/opus_decoder.cpp:361): returns (rc, value) -- value is read for SET requests."""
        v = ctypes.c_int32(0 if value is None else int(value))
        rc = lib().opusb_decoder_ctl(self._st, request, ctypes.byref(v))
        return rc, v.value

    def close(self):
        if self._st:
            lib().opusb_decoder_destroy(self._st)
            self._st = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class OpusMSDecoder:
    """opus_multistream_decoder_create / _decode / _ctl (src/opus_decoder.h:205-218), one Opus stream."""

    def __init__(self, Fs, channels, streams, coupled_streams, mapping):
        err = _I(0)
        m = np.ascontiguousarray(mapping, np.uint8)
        self._st = lib().opusb_multistream_decoder_create(Fs, channels, streams, coupled_streams, m.ctypes.data, ctypes.byref(err))
        self.channels, self.error = channels, err.value
        if not self._st:
            raise RuntimeError("opus_multistream_decoder_create failed: %d" % err.value)

    def decode(self, data, frame_size=5760):
        p = _pkt(data)
        pcm = np.zeros((frame_size, self.channels), np.int16)
        n = lib().opusb_multistream_decode(self._st, p.ctypes.data, p.size, pcm.ctypes.data, frame_size)
        return n, pcm[: max(n, 0)]

    def ctl(self, request, value=None):
        v = ctypes.c_int32(0 if value is None else int(value))
        rc = lib().opusb_multistream_decoder_ctl(self._st, request, ctypes.byref(v))
        return rc, v.value

    def close(self):
        if self._st:
            lib().opusb_multistream_decoder_destroy(self._st)
            self._st = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- player layer (include/opusfile_b200.h; ref src/opusfile.h:141-156) ----
OPFB_PENDING = -2100
OP_HOLE, OP_EREAD, OP_EBADPACKET, OP_ENOTFORMAT, OP_EBADHEADER, OP_EBADTIMESTAMP = -3, -128, -136, -132, -133, -139
READ_FUNC = ctypes.CFUNCTYPE(_I, _P, _P, _I)
_player_bound = False


def _plib():
    global _player_bound
    L = lib()
    if not _player_bound:
        L.opusb_head_parse.argtypes = [_P, _P, ctypes.c_size_t]
        L.opusb_init_decoder.restype = _P
        L.opusb_init_decoder.argtypes = [READ_FUNC, _P, _I, ctypes.POINTER(_I)]
        L.opusb_read_stereo.argtypes = [_P, _P, _I]
        L.opusb_free_decoder.argtypes = [_P]
        L.opfb_set_create.argtypes = [ctypes.POINTER(_P), _I, _I, _I, _I]
        L.opfb_set_destroy.argtypes = [_P]
        L.opfb_set_open_memory.argtypes = [_P, _I, _P, ctypes.c_size_t, _I]
        L.opfb_set_prepare.argtypes = [_P]
        L.opfb_set_open_result.argtypes = [_P, _I]
        L.opfb_set_decode_round.argtypes = [_P]
        L.opfb_set_decode.argtypes = [_P]
        L.opfb_set_read_stereo.argtypes = [_P, _I, _P, _I]
        L.opfb_set_pcm_total.restype = ctypes.c_int64
        L.opfb_set_pcm_total.argtypes = [_P, _I]
        L.opfb_set_stats.argtypes = [_P, _P, _I]
        L.opfb_set_drain.restype = ctypes.c_int64
        L.opfb_set_drain.argtypes = [_P, _I]
        L.opfb_debug_events.argtypes = [_P, ctypes.c_size_t, _I, _P, _I, ctypes.POINTER(_I)]
        L.opfb_synth_ogg.restype = ctypes.c_size_t
        L.opfb_synth_ogg.argtypes = [_P, ctypes.c_size_t, _I, _I, _I, _I, _I, _I, ctypes.c_uint64, ctypes.c_uint32]
        _player_bound = True
    return L


class OpusHead(ctypes.Structure):
    _fields_ = [("version", _I), ("channel_count", _I), ("pre_skip", ctypes.c_uint), ("input_sample_rate", ctypes.c_uint32),
                ("output_gain", _I), ("mapping_family", _I), ("stream_count", _I), ("coupled_count", _I), ("mapping", ctypes.c_uint8 * 2)]


def opus_head_parse(data):
    """ref opus_head_parse (src/opusfile.cpp:1333) -> (rc, OpusHead)."""
    p = _pkt(data)
    h = OpusHead()
    rc = _plib().opusb_head_parse(ctypes.cast(ctypes.byref(h), _P), p.ctypes.data, p.size)
    return rc, h


def demux_events(data, clean_eof=False, cap=1 << 16):
    """Host demux stage alone (no GPU): (open_result, int32[n][6] rows of kind, code, bytes, duration, od_pos, od_size)."""
    p = _pkt(data)
    out = np.zeros((cap, 6), np.int32)
    res = _I(0)
    n = _plib().opfb_debug_events(p.ctypes.data, p.size, int(clean_eof), out.ctypes.data, cap, ctypes.byref(res))
    return res.value, out[: min(n, cap)]


def synth_ogg(channels, toc, packet_len, n_packets, per_page=4, pre_skip=0, seed=0x9E3779B97F4A7C15, serial=1):
    """Synthetic Ogg Opus file as uint8 array (opfb_synth_ogg)."""
    L = _plib()
    n = L.opfb_synth_ogg(None, 0, channels, pre_skip, toc, packet_len, n_packets, per_page, ctypes.c_uint64(seed), serial)
    out = np.empty(n, np.uint8)
    L.opfb_synth_ogg(out.ctypes.data, n, channels, pre_skip, toc, packet_len, n_packets, per_page, ctypes.c_uint64(seed), serial)
    return out


class OggOpusFile:
    """Single-stream player with the reference's call sequence: opus_init_decoder() + op_read_stereo()
    (src/opusfile.cpp:784, 1293). `read` is SD_read: read(nbytes) -> bytes (b"" at the end)."""

    def __init__(self, source, device=0, eof_is_error=True):
        if isinstance(source, (bytes, bytearray, np.ndarray)):
            blob = bytes(source)
            pos = [0]

            def read(n):
                b = blob[pos[0]: pos[0] + n]
                pos[0] += len(b)
                return b
        else:
            read = source

        def cb(user, buf, n):
            if n == 0:
                return 0
            b = read(n)
            if not b:
                return -1 if eof_is_error else 0
            ctypes.memmove(buf, b, len(b))
            return len(b)

        self._cb = READ_FUNC(cb)
        err = _I(0)
        self._of = _plib().opusb_init_decoder(self._cb, None, device, ctypes.byref(err))
        self.error = err.value
        if not self._of:
            raise RuntimeError("opus_init_decoder failed: %d" % err.value)

    def read_stereo(self, buf_size=2048):
        """-> (ret, pcm int16[ret][2]) exactly like op_read_stereo(pcm, buf_size)."""
        pcm = np.zeros(buf_size, np.int16)
        n = _plib().opusb_read_stereo(self._of, pcm.ctypes.data, buf_size)
        return n, pcm[: 2 * max(n, 0)].reshape(-1, 2)

    def close(self):
        if self._of:
            _plib().opusb_free_decoder(self._of)
            self._of = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class OggOpusFileSet:
    """Many Ogg Opus files at once: host-thread demux + batched GPU decode + op_read_stereo per file."""

    def __init__(self, blobs, device=0, frames_per_round=50, n_threads=0, clean_eof=False):
        self._set = _P()
        self._blobs = [np.frombuffer(bytes(b), np.uint8) if not isinstance(b, np.ndarray) else np.ascontiguousarray(b, np.uint8) for b in blobs]
        L = _plib()
        rc = L.opfb_set_create(ctypes.byref(self._set), len(self._blobs), device, frames_per_round, n_threads)
        if rc:
            raise RuntimeError("opfb_set_create: %d" % rc)
        for i, b in enumerate(self._blobs):
            rc = L.opfb_set_open_memory(self._set, i, b.ctypes.data, b.size, int(clean_eof))
            if rc:
                raise RuntimeError("opfb_set_open_memory: %d" % rc)
        rc = L.opfb_set_prepare(self._set)
        if rc:
            raise RuntimeError("opfb_set_prepare: %d (no CUDA device => no decode path)" % rc)

    def open_result(self, i):
        return _plib().opfb_set_open_result(self._set, i)

    def decode_round(self):
        return _plib().opfb_set_decode_round(self._set)

    def read_stereo(self, i, buf_size=2048):
        pcm = np.zeros(buf_size, np.int16)
        n = _plib().opfb_set_read_stereo(self._set, i, pcm.ctypes.data, buf_size)
        return n, pcm[: 2 * max(n, 0)].reshape(-1, 2)

    def drain(self, buf_size=2048):
        """Benchmark consumer (opfb_set_drain): op_read_stereo over every file on the host threads; frames read."""
        return int(_plib().opfb_set_drain(self._set, buf_size))

    def stats(self):
        v = (ctypes.c_double * 12)()
        _plib().opfb_set_stats(self._set, ctypes.cast(v, _P), 12)
        return dict(zip(("demux_s", "decode_s", "packets", "frames", "ogg_bytes", "launches", "keep_s", "plan_s", "reset_s", "stage_s",
                         "gpu_wait_s", "finish_s"), [float(x) for x in v]))

    def close(self):
        if self._set:
            _plib().opfb_set_destroy(self._set)
            self._set = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
