/* opusfile_b200.h -- C ABI of the Ogg Opus player layer of libopusb200.so.
 *
 * Drop-in boundary for the player API of schreibfaul1/ESP32-Opus-Player (src/opusfile.h):
 *
 *   reference                                   here
 *   OggOpusFile_t* opus_init_decoder()  :144    opusb_init_decoder(read, user, device, &err)
 *   int op_read_stereo(pcm, buf_size)   :156    opusb_read_stereo(of, pcm, buf_size)
 *   int SD_read(buf, nbytes) (weak)     :19     the `read` callback (same return convention: bytes read, 0 only
 *                                               for nbytes == 0, < 0 at end of data -- src/main.cpp:264-269)
 *   int opus_head_parse(head, data, len) :141   opusb_head_parse(head, data, len)
 *
 * plus what the reference cannot do: opfb_set_* plays thousands of Ogg Opus files at once -- the Ogg page demux
 * (CRC, packet assembly, pre-skip / end-trim time stamps; src/ogg.cpp, src/opusfile.cpp:424-633, 835-1133) runs on
 * host threads, packets are packed into pinned staging buffers and decoded by decode_many() on one GPU, and every
 * file is then read with op_read_stereo() semantics (same samples, same return value per call).
 *
 * Return values are the reference's: frames written (0 at end of stream) or OP_* (< 0), src/opusfile.h:105-128.
 * Only what the reference itself supports is reproduced: unseekable streaming, channel mapping family 0/1 with one
 * (coupled) stream, 20 ms frames.  A packet the reference cannot decode without undefined behaviour
 * (frame size other than 20 ms, a SILK/hybrid packet whose channel count differs from the header's) is reported
 * as OP_EBADPACKET instead.  There is no CPU decode path.
 */
#ifndef OPUSFILE_B200_H
#define OPUSFILE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OPB_FALSE (-1)
#define OPB_EOF (-2)
#define OPB_HOLE (-3)
#define OPB_EREAD (-128)
#define OPB_EFAULT (-129)
#define OPB_EIMPL (-130)
#define OPB_EINVAL (-131)
#define OPB_ENOTFORMAT (-132)
#define OPB_EBADHEADER (-133)
#define OPB_EVERSION (-134)
#define OPB_EBADPACKET (-136)
#define OPB_EBADLINK (-137)
#define OPB_EBADTIMESTAMP (-139)
#define OPFB_PENDING (-2100) /* opfb_set_read_stereo: the next samples of this file are not decoded yet */

/* src/opusfile.h:42-52 */
typedef struct OpusHeadB200 {
    int version;
    int channel_count;
    unsigned pre_skip;
    uint32_t input_sample_rate;
    int output_gain;
    int mapping_family;
    int stream_count;
    int coupled_count;
    uint8_t mapping[2];
} OpusHeadB200;

/* ref opus_head_parse (src/opusfile.cpp:1333) */
int opusb_head_parse(OpusHeadB200* head, const uint8_t* data, size_t len);

/* ref SD_read (src/opusfile.h:19): returns bytes read (> 0), 0 only when nbytes == 0, negative at end of data.
 * A callback that returns 0 at the end instead gets a clean end of stream (op_read_stereo returns 0). */
typedef int (*opusb_read_func)(void* user, unsigned char* buf, int nbytes);

/* ---- single stream, same call sequence as the reference player (src/main.cpp:270-282) ---- */
typedef struct OggOpusFileB200 OggOpusFileB200;
/* ref opus_init_decoder (src/opusfile.cpp:784): reads the headers through `read`, validates the first audio page,
 * creates the decoder.  NULL on failure (*error = OP_* code). */
OggOpusFileB200* opusb_init_decoder(opusb_read_func read, void* user, int device, int* error);
/* ref op_read_stereo (src/opusfile.cpp:1293): up to buf_size/2 stereo frames of the current packet (mono is
 * duplicated to both channels); 0 at the end of the stream; OP_* < 0 on error (the next call carries on). */
int opusb_read_stereo(OggOpusFileB200* of, int16_t* pcm, int buf_size);
void opusb_free_decoder(OggOpusFileB200* of);
const OpusHeadB200* opusb_head(const OggOpusFileB200* of);

/* ---- many files at once ---- */
typedef struct opfb_set opfb_set;
/* frames_per_round: 20 ms frames per file decoded by one decode_many() call; n_threads: host demux threads
 * (0 = one per core). */
int opfb_set_create(opfb_set** out, int n_files, int device, int frames_per_round, int n_threads);
void opfb_set_destroy(opfb_set* set);
/* Attach file `index` as a memory image (borrowed until opfb_set_prepare returns).  End of data reads like the
 * reference's SD_read (negative), so a complete file ends with OP_EREAD exactly as the reference does, unless
 * clean_eof != 0. */
int opfb_set_open_memory(opfb_set* set, int index, const uint8_t* data, size_t len, int clean_eof);
/* Demux every attached file on the host threads: headers, packets, time stamps.  Returns 0; per-file open results
 * (0 or OP_*, what opus_init_decoder would have reported) via opfb_set_open_result. */
int opfb_set_prepare(opfb_set* set);
int opfb_set_open_result(const opfb_set* set, int index);
/* Decode the next frames_per_round frames of every file on the GPU (one decode_many() per channel-count group).
 * Returns 1 when something was decoded, 0 when every file is finished, DM_ERR_* on failure.  Decoded samples stay
 * queued per file until opfb_set_read_stereo consumes them. */
int opfb_set_decode_round(opfb_set* set);
/* All rounds at once (keeps the PCM of every file in host memory: small sets only). */
int opfb_set_decode(opfb_set* set);
/* op_read_stereo for file `index`; OPFB_PENDING when opfb_set_decode_round must run first. */
int opfb_set_read_stereo(opfb_set* set, int index, int16_t* pcm, int buf_size);
/* Benchmark consumer: reads every file with op_read_stereo calls of buf_size samples until it reports OPFB_PENDING or
 * its end, on the set's host threads.  Returns the stereo frames read (the samples are discarded). */
int64_t opfb_set_drain(opfb_set* set, int buf_size);
/* Total stereo frames op_read_stereo will deliver for file `index` (after pre-skip and end trimming). */
int64_t opfb_set_pcm_total(const opfb_set* set, int index);
/* Statistics of the last prepare/decode: [0] demux seconds, [1] decode seconds (H2D + kernels + D2H), [2] packets,
 * [3] frames, [4] bytes of Ogg data, [5] kernel launches, [6..11] host seconds of the round scheduler by phase: keeping
 * unread samples, planning, decoder resets, staging + enqueue, waiting for the GPU, handing results to the files. */
int opfb_set_stats(const opfb_set* set, double* out, int cap);

/* Synthetic Ogg Opus file (benchmarks, tests): OpusHead + OpusTags + n_packets code-0 packets of packet_len bytes (TOC +
 * xorshift64 payload seeded `seed`, the generator of dm_synth_fill), per_page packets per page, 960 samples per packet,
 * EOS on the last page.  Returns the size; writes only when `out` has room for it. */
size_t opfb_synth_ogg(uint8_t* out, size_t cap, int channels, int pre_skip, int toc, int packet_len, int n_packets, int per_page,
                      uint64_t seed, uint32_t serial);

/* Test access to the host demux stage alone (runs without a GPU): the player events of one memory image as rows of
 * (kind 0 packet | 1 return code | 2 OPUS_RESET_STATE | 3 new decoder, code, bytes, duration, od_pos, od_size). */
int opfb_debug_events(const uint8_t* data, size_t len, int clean_eof, int32_t* out, int cap, int* open_result);

#ifdef __cplusplus
}
#endif
#endif /* OPUSFILE_B200_H */
