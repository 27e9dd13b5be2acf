/* opus_decoder.h -- source-level drop-in for the reference's packet API (libopusb200_dropin.so).
 *
 * Same file name, same C++ free-function signatures (Itanium-mangled, no extern "C", no namespace) as
 * schreibfaul1/ESP32-Opus-Player src/opus_decoder.h:165-218, so that
 *   - a host build of the reference player compiles its own opusfile.cpp / ogg.cpp and LINKS them against
 *     libopusb200_dropin.so in place of opus_decoder.cpp + celt.cpp + silk.cpp (tests/test_gpu_dropin.py does exactly
 *     that with the unmodified reference objects), and
 *   - new code can include this header instead of the reference's and call opus_decode() unchanged.
 * Decoding runs on the GPU through libopusb200.so (include/opus_b200.h); there is no CPU decode path.
 * The extern "C" spellings of the same entry points (for cgo / JNI / ctypes) are the opusb_* functions of opus_b200.h.
 *
 * Differences a caller can observe, all on inputs where the reference has undefined behaviour:
 * a packet whose TOC frame size is not 20 ms returns OPUS_UNIMPLEMENTED (the reference asserts / overruns),
 * OPUS_GET_PITCH_REQUEST returns OPUS_UNIMPLEMENTED (it reads loss-concealment bookkeeping that is not kept).
 */
#ifndef OPUSB200_DROPIN_OPUS_DECODER_H
#define OPUSB200_DROPIN_OPUS_DECODER_H

#include <stddef.h>
#include <stdint.h>

/* error codes, ref src/opus_decoder.h:70-77 */
#define OPUS_OK 0
#define OPUS_BAD_ARG -1
#define OPUS_BUFFER_TOO_SMALL -2
#define OPUS_INTERNAL_ERROR -3
#define OPUS_INVALID_PACKET -4
#define OPUS_UNIMPLEMENTED -5
#define OPUS_INVALID_STATE -6
#define OPUS_ALLOC_FAIL -7
#define ERR_OPUS_CELT_BAD_ARG -18 /* ref errcode enum, src/opus_decoder.h:58 */

/* opus_decoder_ctl / opus_multistream_decoder_ctl requests the decoder answers, ref src/opus_decoder.h:84-123 */
#define OPUS_GET_BANDWIDTH_REQUEST 4009
#define OPUS_RESET_STATE 4028
#define OPUS_GET_SAMPLE_RATE_REQUEST 4029
#define OPUS_GET_FINAL_RANGE_REQUEST 4031
#define OPUS_GET_PITCH_REQUEST 4033
#define OPUS_SET_GAIN_REQUEST 4034
#define OPUS_GET_GAIN_REQUEST 4045
#define OPUS_GET_LAST_PACKET_DURATION_REQUEST 4039
#define OPUS_SET_PHASE_INVERSION_DISABLED_REQUEST 4046
#define OPUS_GET_PHASE_INVERSION_DISABLED_REQUEST 4047
#define OPUS_BANDWIDTH_NARROWBAND 1101
#define OPUS_BANDWIDTH_MEDIUMBAND 1102
#define OPUS_BANDWIDTH_WIDEBAND 1103
#define OPUS_BANDWIDTH_SUPERWIDEBAND 1104
#define OPUS_BANDWIDTH_FULLBAND 1105
#define MODE_SILK_ONLY 1000
#define MODE_HYBRID 1001
#define MODE_CELT_ONLY 1002

/* ref src/opus_decoder.h:146: the blob is opaque; callers size it with opus_decoder_get_size() */
typedef struct OpusDecoder OpusDecoder;

/* ref src/opus_decoder.h:155-160 (public fields; the decoder itself lives behind them) */
typedef struct OpusMSDecoder {
    int nb_channels;
    int nb_streams;
    int nb_coupled_streams;
    unsigned char mapping[256];
} OpusMSDecoder_t;

/* ---- packet API, ref src/opus_decoder.h:165-176, bodies src/opus_decoder.cpp:66, 82, 351, 361, 457-559 ---- */
int opus_decoder_get_size(int channels);
/* Initialises a caller-allocated blob of opus_decoder_get_size() bytes; calling it again on the same blob is the
 * reference's "new stream" reset.  Fs: 48000/24000/16000/12000/8000 accepted, output is always 48 kHz as in the
 * reference (src/opusfile.cpp:693 only ever passes 48000). */
int opus_decoder_init(OpusDecoder* st, int32_t Fs, int channels);
int opus_decode(OpusDecoder* st, uint8_t* data, int32_t len, int16_t* pcm, int frame_size);
int opus_decoder_ctl(OpusDecoder* st, int request, ...);
/* ref src/opus_decoder.cpp:457 is free(st); here it also releases the GPU context bound to the blob.  A blob that is
 * simply free()d instead leaks that context until the process ends. */
void opus_decoder_destroy(OpusDecoder* st);
int opus_packet_parse(uint8_t* data, int32_t len, unsigned char* out_toc, uint8_t* frames[48], int16_t size[48], int* payload_offset);
int opus_packet_get_bandwidth(uint8_t* data);
int opus_packet_get_samples_per_frame(uint8_t* data, int32_t Fs);
int opus_packet_get_nb_channels(uint8_t* data);
int opus_packet_get_nb_frames(uint8_t packet[], int32_t len);
int opus_packet_get_nb_samples(uint8_t packet[], int32_t len, int32_t Fs);
int opus_decoder_get_nb_samples(const OpusDecoder* dec, uint8_t packet[], int32_t len);

/* ---- multistream wrapper, ref src/opus_decoder.h:205-218, bodies src/opus_decoder.cpp:778-1046.  One Opus stream
 * (mono or coupled) mapped onto 1..255 output channels -- what the player creates (mapping family 0, or 1 with one
 * stream); streams != 1 fails with OPUS_UNIMPLEMENTED (the reference's shared global codec state cannot do it either,
 * SURVEY App. D 9). ---- */
int32_t opus_multistream_decoder_get_size(int streams, int coupled_streams);
OpusMSDecoder_t* opus_multistream_decoder_create(int32_t Fs, int channels, int streams, int coupled_streams, const uint8_t* mapping, int* error);
int opus_multistream_decode(OpusMSDecoder_t* st, uint8_t* data, int32_t len, int16_t* pcm, int frame_size);
int opus_multistream_decoder_ctl(OpusMSDecoder_t* st, int request, ...);
void opus_multistream_decoder_destroy(OpusMSDecoder_t* st);

/* ---- codec-layer entry points the reference player and its oracle call (ref src/celt.h:662-664, 704, 711-712, 719;
 * src/silk.h:1043-1045).  The reference keeps ONE global CELT decoder (src/celt.cpp:36-39); so does this layer: a
 * one-stream GPU context that celt_decoder_init() (re)creates.  celt_decode_with_ec() is the short-frame entry point
 * (frame_size 120/240/480/960). ---- */
#define CELT_SET_CHANNELS_REQUEST 10008
#define CELT_SET_START_BAND_REQUEST 10010
#define CELT_SET_END_BAND_REQUEST 10012 /* stored, never read -- as in the reference */
#define CELT_SET_SIGNALLING_REQUEST 10016
bool CELTDecoder_AllocateBuffers(void);
void CELTDecoder_FreeBuffers();
bool SILKDecoder_AllocateBuffers(void);
void SILKDecoder_FreeBuffers();
int32_t celt_decoder_init(int32_t channels);
int32_t celt_decoder_ctl(int32_t request, ...);
void ec_dec_init(uint8_t* _buf, uint32_t _storage);
int32_t celt_decode_with_ec(int16_t* pcm, int32_t frame_size);

#endif /* OPUSB200_DROPIN_OPUS_DECODER_H */
