/* opus_b200.h -- C ABI of the B200 batched Opus decoder (libopusb200.so).
 *
 * Drop-in boundary for the hot path "Opus packet -> int16 PCM" of schreibfaul1/ESP32-Opus-Player
 * (opus_decode_frame and everything below it, src/opus_decoder.cpp:154), plus the batched
 * decode_many() call the reference lacks.  Every entry point states the reference interface it
 * replaces (file:line under /root/reference).  Plain pointers and sizes only; no C++/torch types.
 *
 * Errors are the reference's negative ints (src/opus_decoder.h:40-77): OPUS_BAD_ARG -1,
 * OPUS_BUFFER_TOO_SMALL -2, OPUS_INTERNAL_ERROR -3, OPUS_INVALID_PACKET -4, OPUS_UNIMPLEMENTED -5,
 * ERR_OPUS_CELT_BAD_ARG -18 ..., plus DM_ERR_* (<= -2000) for CUDA/runtime failures of this library.
 * There is NO CPU fallback: without a usable CUDA device every decode entry point fails with
 * DM_ERR_NO_DEVICE.
 */
#ifndef OPUS_B200_H
#define OPUS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DM_ERR_NO_DEVICE (-2000)
#define DM_ERR_CUDA (-2001)
#define DM_ERR_NOMEM (-2002)
#define DM_ERR_CAPACITY (-2003)
#define DM_FRAME_SKIPPED (-2004) /* per-frame status: filler frame, or a frame behind a failed frame of its packet */

/* ------------------------------------------------------------------------------------------------
 * Batched decoder.  One context owns the decoder state of n_streams independent streams on one GPU
 * (the reference keeps exactly one such state in process-wide singletons, src/celt.cpp:36-39,
 * src/silk.cpp:18-29, src/opus_decoder.cpp:40-59).
 * ---------------------------------------------------------------------------------------------- */
typedef struct dm_ctx dm_ctx;

/* Replaces opus_decoder_get_size + opus_decoder_init (src/opus_decoder.cpp:66, 82) for n_streams
 * decoders at once: Fs is fixed at 48000 (the only rate the reference ever uses, src/opusfile.cpp:693),
 * `channels` (1|2) is the output channel count of every stream.  max_frames = frames per stream that
 * one decode call may carry.  device = CUDA ordinal. */
int dm_create(dm_ctx** out, int n_streams, int max_frames, int channels, int device);
void dm_destroy(dm_ctx* ctx);

/* opus_decoder_init / OPUS_RESET_STATE for a subset of streams (src/opus_decoder.cpp:82, 361):
 * stream_ids == NULL resets all. */
int dm_reset(dm_ctx* ctx, const int32_t* stream_ids, int n);
/* opus_decoder_ctl(OPUS_RESET_STATE) (src/opus_decoder.cpp:380-388) for a subset of streams (NULL = all): CELT soft
 * reset (src/celt.cpp:2479-2498 keeps the signal history, oldBandE and de-emphasis memory), silk_InitDecoder. */
int dm_reset_state(dm_ctx* ctx, const int32_t* stream_ids, int n);

/* How frames are interpreted. */
#define DM_MODE_OPUS 0     /* frames are Opus frames; toc[] gives mode/bandwidth/channels (player-level
                              semantics of opus_decode_frame: 20 ms only) */
#define DM_MODE_CELT_RAW 1 /* frames are raw CELT frames for celt_decode_with_ec (src/celt.cpp:2162):
                              frame_size 120/240/480/960, coded_channels, start_band given explicitly */

/* dm_frame.flags */
#define DMF_CONT 1 /* 2nd..nth frame of a multi-frame packet: not decoded if an earlier frame of the packet failed
                      (opus_decode_native returns at the first failing frame, src/opus_decoder.cpp:334-338) */
#define DMF_NULL 2 /* filler behind the last real frame of a stream in this call: no output, no state change;
                      must trail the stream's row and repeat the last real frame's toc */

typedef struct dm_frame {  /* one coded frame of one stream */
    uint32_t offset;       /* byte offset of the frame payload inside `bytes` (TOC already stripped) */
    uint16_t len;          /* payload bytes (<= 1275) */
    uint8_t toc;           /* the packet's TOC byte (DM_MODE_OPUS) */
    uint8_t flags;         /* DMF_* */
} dm_frame;

typedef struct dm_batch {
    int n_frames;             /* frames per stream in this call (<= max_frames) */
    int mode;                 /* DM_MODE_* */
    int frame_size;           /* DM_MODE_CELT_RAW: samples per frame; DM_MODE_OPUS: ignored (960) */
    int coded_channels;       /* DM_MODE_CELT_RAW: channels coded in the frames (1|2) */
    int start_band;           /* DM_MODE_CELT_RAW: 0, or 17 for the hybrid high band */
    const uint8_t* bytes;     /* all payload bytes */
    size_t n_bytes;
    const dm_frame* frames;   /* [n_streams][n_frames], stream-major */
} dm_batch;

typedef struct dm_out {
    int16_t* pcm;             /* [n_streams][n_frames][frame_size][channels] interleaved */
    int32_t* status;          /* [n_streams][n_frames]: samples per channel, or a negative error */
} dm_out;

/* Batched replacement of opus_decode() (src/opus_decoder.cpp:351) / celt_decode_with_ec()
 * (src/celt.cpp:2162) with HOST buffers: copies the batch to the GPU, runs the kernels, copies PCM
 * and status back, returns when they are in host memory.  Pinned buffers (dm_host_alloc) make the
 * copies asynchronous and pipelined. */
int decode_many(dm_ctx* ctx, const dm_batch* in, dm_out* out);

/* The same call split in two so that consecutive calls pipeline: decode_many_async() enqueues the copies and kernels
 * and returns at once with a ticket; dm_wait() blocks until that call's PCM and status are in host memory.  Up to 3
 * calls may be in flight; the input buffers of a call (pinned) must stay valid until dm_wait() returned for it, and
 * each call in flight needs its own output buffers.  With calls in flight the kernels of call n+1 overlap the
 * PCM read-back of call n, which is what bounds a single call (PCIe D2H of 3840 B per decoded frame). */
int decode_many_async(dm_ctx* ctx, const dm_batch* in, dm_out* out, long long* ticket);
int dm_wait(dm_ctx* ctx, long long ticket);

/* Same with every pointer in DEVICE memory, as ONE stream-ordered operation on `cuda_stream` (a cudaStream_t, NULL =
 * the legacy default stream), no host synchronisation: the kernel-only path.  The kernels run on the library's own
 * streams behind everything `cuda_stream` holds when the call is made; `cuda_stream` continues once they are done. */
int dm_decode_device(dm_ctx* ctx, const dm_batch* in_dev, dm_out* out_dev, void* cuda_stream);
/* The same with explicit dependencies instead of a stream, so that consecutive calls overlap: the call starts once
 * `ready_event` (a cudaEvent_t; NULL: the inputs are ready now) has fired and *done_event (a cudaEvent_t owned by the
 * library, valid until four more calls have been made) fires when PCM and status are complete.  Calls on a context are
 * ordered among themselves; calls on different contexts run beside each other (a mixed workload has one context per
 * channel count and call mode).  With OPUSB_TAIL=2 in the environment the serial tail of call k (post-filter,
 * de-emphasis) also runs beside the head of call k + 1 -- measured, and off by default because it wins nothing.  The
 * inputs of a call may be reused once its done event has fired, and a consumer must have finished with an output buffer
 * before a later call is given the same one.
 * dm_stream_wait makes a stream wait for everything enqueued so far. */
int dm_decode_device_async(dm_ctx* ctx, const dm_batch* in_dev, dm_out* out_dev, void* ready_event, void** done_event);
int dm_stream_wait(dm_ctx* ctx, void* cuda_stream);

/* Pinned host memory helpers (cudaHostAlloc / cudaFreeHost). */
void* dm_host_alloc(size_t bytes);
void dm_host_free(void* p);

/* Measurement aid: `reps` bare device-to-host copies of `bytes` from the context's PCM buffer into `host` (the copy
 * decode_many_async ends with, and nothing else), timed with CUDA events: the read-back ceiling of this box. */
int dm_probe_d2h(dm_ctx* ctx, void* host, size_t bytes, int reps, float* ms_total);
/* Number of kernel launches issued by this library since dm_create (bench evidence). */
long long dm_launch_count(const dm_ctx* ctx);
/* Bytes of device memory held by the context. */
size_t dm_device_bytes(const dm_ctx* ctx);
/* Device time in ms of each pipeline stage of the most recent pipelined call (dm_decode_device, dm_decode_device_async,
 * decode_many_async), measured with CUDA events on the streams the kernels run on; valid once that call has completed.
 * Returns the number of entries written (names via dm_stage_name): the eight stages, then "pipeline" = first kernel to
 * last.  With calls in flight a stage's span includes whatever ran beside it: for a breakdown use dm_decode_device. */
int dm_stage_times(dm_ctx* ctx, float* ms, int cap);
const char* dm_stage_name(int i);
/* Debug/test access: 1 if the most recent call had more coded PVQ leaves than the sorted leaf list holds and the band
 * kernel expanded its own leaves instead (slower, same result), 0 if not, < 0 on error.  Synchronises the device. */
int dm_debug_leaf_overflow(dm_ctx* ctx);
/* Debug/test access: how many segments of de-emphasis chains were done a second time since dm_create (the chains of a
 * 20 ms call are cut into runs of frames that warm up on the frame in front of them and are checked against the serial
 * order afterwards, csrc/kernels.cuh k_deemph / k_deemph_fix): 0 for streams of full CELT frames.  Synchronises the device. */
long long dm_debug_deemph_redone(dm_ctx* ctx);
/* Debug access (contexts created with OPUSB_TRACE=1 in the environment): stage boundaries of the last n <= 16 pipelined
 * calls in ms since the oldest one's start, out[n][9]: head start, end of entropy, sequence, leaves, bands, imdct, then
 * on the tail stream end of postfilter, deemph, final SILK sum.  Shows which calls overlapped.  Synchronises the device. */
int dm_debug_trace(dm_ctx* ctx, int n, float* out);
/* Debug access: header (CeltFrameHdr, 364 bytes) and op buffer (1360 x 16 bytes) of frame t = stream * n_frames + frame
 * of the most recent call, as the parser kernel left them.  Synchronises the device. */
int dm_debug_frame_script(dm_ctx* ctx, long long t, void* hdr_out, size_t hdr_cap, void* ops_out, size_t ops_cap);
/* Debug/test access: copies one stream's CELT state blob to host (checkpoint). */
int dm_get_stream_state(dm_ctx* ctx, int stream, void* dst, size_t cap);
size_t dm_stream_state_size(void);

/* ------------------------------------------------------------------------------------------------
 * Synthetic corpus: stream i, frame f payload = xorshift64 bytes seeded
 * seed_base ^ i, behind a fixed TOC.  Fills `bytes`/`frames` on the host.
 * ---------------------------------------------------------------------------------------------- */
size_t dm_synth_bytes(int n_streams, int n_frames, int payload_len);
void dm_synth_fill(uint8_t* bytes, dm_frame* frames, int n_streams, int n_frames, int payload_len, int toc,
                   uint64_t seed_base, int first_stream);

/* The same for raw CELT frames (DM_MODE_CELT_RAW): all frame_len bytes of a frame are generator output. */
void dm_synth_fill_raw(uint8_t* bytes, dm_frame* frames, int n_streams, int n_frames, int frame_len, uint64_t seed_base,
                       int first_stream);
/* Result check: the SURVEY 8c stream hash (FNV-1a over the uint16 view of the interleaved samples, a mono stream's
 * samples counted twice when dup_mono != 0, as the packet-level oracle does), carried on across calls.
 * pcm = [n_streams][samples_per_stream]; hashes[n_streams] is updated in place (start: 1469598103934665603).
 * n_threads host threads (0 = one per core). */
void dm_fnv_update(uint64_t* hashes, const int16_t* pcm, int n_streams, size_t samples_per_stream, int channels, int dup_mono,
                   int n_threads);

/* ------------------------------------------------------------------------------------------------
 * Packet helpers with the reference's semantics (src/opus_decoder.cpp:135, 460-680).
 * ---------------------------------------------------------------------------------------------- */
int opusb_packet_get_mode(const uint8_t* data);                 /* 1000 SILK, 1001 hybrid, 1002 CELT */
int opusb_packet_get_bandwidth(const uint8_t* data);            /* 1101..1105 */
int opusb_packet_get_samples_per_frame(const uint8_t* data, int32_t Fs);
int opusb_packet_get_nb_channels(const uint8_t* data);
int opusb_packet_get_nb_frames(const uint8_t* packet, int32_t len);
int opusb_packet_get_nb_samples(const uint8_t* packet, int32_t len, int32_t Fs);
int opusb_packet_parse(const uint8_t* data, int32_t len, uint8_t* out_toc, const uint8_t* frames[48], int16_t size[48],
                       int* payload_offset);

/* ------------------------------------------------------------------------------------------------
 * Single-stream drop-in (same argument meaning as the reference's packet API, src/opus_decoder.h:165-168).
 * The OpusDecoder blob is opaque; it holds a one-stream dm_ctx.
 * ---------------------------------------------------------------------------------------------- */
typedef struct OpusB200Decoder OpusB200Decoder;
OpusB200Decoder* opusb_decoder_create(int32_t Fs, int channels, int* error);
void opusb_decoder_destroy(OpusB200Decoder* st);
int opusb_decoder_init(OpusB200Decoder* st, int32_t Fs, int channels);
int opusb_decode(OpusB200Decoder* st, const uint8_t* data, int32_t len, int16_t* pcm, int frame_size);
/* opus_decoder_ctl (src/opus_decoder.cpp:361-454) without varargs: `value` is written by GET requests and read by SET
 * requests; request codes are the reference's (src/opus_decoder.h:84-123).  GET_BANDWIDTH, GET_FINAL_RANGE (always 0,
 * as in the reference), RESET_STATE, GET_SAMPLE_RATE, GET/SET_GAIN (stored, never applied), GET_LAST_PACKET_DURATION;
 * the phase-inversion requests return OPUS_UNIMPLEMENTED exactly as the reference's do; anything else too. */
int opusb_decoder_ctl(OpusB200Decoder* st, int request, int32_t* value);

/* Multistream wrapper (src/opus_decoder.h:205-218) for what the player creates: one Opus stream, mono or coupled,
 * mapped onto `channels` output channels (mapping[c] = 0/1 or 255 for silence).  streams != 1: OPUS_UNIMPLEMENTED. */
typedef struct OpusB200MSDecoder OpusB200MSDecoder;
OpusB200MSDecoder* opusb_multistream_decoder_create(int32_t Fs, int channels, int streams, int coupled_streams, const uint8_t* mapping,
                                                    int* error);
int opusb_multistream_decode(OpusB200MSDecoder* st, const uint8_t* data, int32_t len, int16_t* pcm, int frame_size);
int opusb_multistream_decoder_ctl(OpusB200MSDecoder* st, int request, int32_t* value);
void opusb_multistream_decoder_destroy(OpusB200MSDecoder* st);

#ifdef __cplusplus
}
#endif
#endif /* OPUS_B200_H */
