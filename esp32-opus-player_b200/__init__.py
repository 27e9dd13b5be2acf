"""B200-native batched Opus decoder: ctypes binding of libopusb200.so (include/opus_b200.h).

The package holds only what the hot path needs: csrc/ (CUDA kernels + C ABI) and this host-side
mirror of the reference's packet API (opus_decoder.h) plus the batched decode_many(), and of its player API
(opusfile.h: opus_init_decoder / op_read_stereo) plus the many-files player.
There is no CPU decode path: importing works anywhere, decoding needs a CUDA device.
"""
from .api import (  # noqa: F401
    DM_MODE_CELT_RAW,
    DM_MODE_OPUS,
    FRAME_DTYPE,
    BatchDecoder,
    OggOpusFile,
    OggOpusFileSet,
    OpusDecoder,
    OpusMSDecoder,
    demux_events,
    lib,
    make_frames,
    opus_head_parse,
    opus_packet_get_bandwidth,
    opus_packet_get_mode,
    opus_packet_get_nb_channels,
    opus_packet_get_nb_frames,
    opus_packet_get_nb_samples,
    opus_packet_get_samples_per_frame,
    opus_packet_parse,
    synth_corpus,
    synth_ogg,
)
from .build import build  # noqa: F401
from . import api, shard  # noqa: F401,E402
