/* opusfile.h -- source-level drop-in for the reference's player API (libopusb200_dropin.so).
 *
 * Same file name and the same C++ signatures as schreibfaul1/ESP32-Opus-Player src/opusfile.h:19, 141, 144, 156:
 * the firmware's loop (src/main.cpp:270-282)
 *
 *     opus_init_decoder();
 *     while ((n = op_read_stereo(buf, 2048)) > 0) play(buf, n);
 *
 * with its own SD_read() compiles and links unchanged.  Ogg demux runs on the host (csrc/ogg_demux.hpp,
 * csrc/opus_player.cpp), every packet is decoded on the GPU; per call the return value and the samples are the
 * reference's (pre-skip, end trimming, chained links, OP_HOLE / OP_EBADPACKET / OP_EREAD), tests/test_gpu_dropin.py.
 * The extern "C" spellings are opusb_init_decoder / opusb_read_stereo of opusfile_b200.h.
 * The CUDA device is taken from the environment variable OPUSB_DEVICE (default 0).
 */
#ifndef OPUSB200_DROPIN_OPUSFILE_H
#define OPUSB200_DROPIN_OPUSFILE_H

#include <stddef.h>
#include <stdint.h>

#include "opus_decoder.h"

/* ref src/opusfile.h:19 -- the caller's data source: bytes read (> 0), 0 only for nbytes == 0, negative at the end of
 * the data (src/main.cpp:264-269). */
extern __attribute__((weak)) int SD_read(unsigned char* buff, int nbytes);

#define OPUS_CHANNEL_COUNT_MAX (2)

/* ref src/opusfile.h:42-52 */
typedef struct OpusHead {
    int version;
    int channel_count;
    unsigned pre_skip;
    uint32_t input_sample_rate;
    int output_gain;
    int mapping_family;
    int stream_count;
    int coupled_count;
    uint8_t mapping[OPUS_CHANNEL_COUNT_MAX];
} OpusHead_t;

/* ref src/opusfile.h:105-128 */
#define OP_FALSE (-1)
#define OP_EOF (-2)
#define OP_HOLE (-3)
#define OP_EREAD (-128)
#define OP_EFAULT (-129)
#define OP_EIMPL (-130)
#define OP_EINVAL (-131)
#define OP_ENOTFORMAT (-132)
#define OP_EBADHEADER (-133)
#define OP_EVERSION (-134)
#define OP_ENOTAUDIO (-135)
#define OP_EBADPACKET (-136)
#define OP_EBADLINK (-137)
#define OP_ENOSEEK (-138)
#define OP_EBADTIMESTAMP (-139)

/* The reference exposes its whole bookkeeping struct (src/opusfile.h:62-103); callers only ever test the pointer
 * against NULL (src/main.cpp:308), so it is opaque here. */
typedef struct OggOpusFile OggOpusFile_t;

/* ref src/opusfile.h:141, body src/opusfile.cpp:1333 */
int opus_head_parse(OpusHead_t* _head, uint8_t* _data, size_t _len);
/* ref src/opusfile.h:144, body src/opusfile.cpp:784: reads the headers through SD_read, validates the first audio page,
 * creates the decoder.  NULL on failure.  One file per process at a time, as in the reference (its state is a global);
 * a second call closes the first file. */
OggOpusFile_t* opus_init_decoder();
/* ref src/opusfile.h:156, body src/opusfile.cpp:1293: up to _buf_size/2 stereo frames of the current packet (mono is
 * duplicated); 0 at the end of the stream; OP_* < 0 on error. */
int op_read_stereo(int16_t* _pcm, int _buf_size);
/* Not in the reference (its firmware never closes the file): releases the GPU context of the open file. */
void opus_free_decoder();

#endif /* OPUSB200_DROPIN_OPUSFILE_H */
