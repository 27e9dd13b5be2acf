// dropin.cpp -- libopusb200_dropin.so: the reference's own entry points, same C++ signatures (include/opus_decoder.h,
// include/opusfile.h), on top of the C ABI of libopusb200.so.  Host glue only: every sample is decoded on the GPU.
//
// What it replaces in a host build of schreibfaul1/ESP32-Opus-Player: src/opus_decoder.cpp + src/celt.cpp + src/silk.cpp
// (link the reference's unmodified opusfile.cpp / ogg.cpp against this library), or the whole of src/ below main.cpp
// (call opus_init_decoder / op_read_stereo from here).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "../../include/opus_b200.h"
#include "../../include/opusfile.h"
#include "../../include/opusfile_b200.h"

namespace {

int env_device() {
    const char* e = getenv("OPUSB_DEVICE");
    return e ? atoi(e) : 0;
}

// ---- OpusDecoder blobs ------------------------------------------------------------------------------------------
// The reference's callers malloc(opus_decoder_get_size(ch)) and hand the raw blob to opus_decoder_init
// (src/opus_decoder.cpp:66, 82); nothing tells the library when the blob is freed.  The GPU decoder of a blob is
// therefore kept in a registry keyed by the blob's address: a second init on the same address resets that decoder
// (or replaces it when the channel count changed), which is what the reference's init does to its globals.
struct Blob {  // what opus_decoder_get_size() accounts for; only `tag` is ever read back from caller memory
    uint64_t tag;
    int channels;
    int32_t Fs;
};
const uint64_t kTag = 0x6f70757362323030ull;  // "opusb200"

struct Entry { OpusB200Decoder* impl; int channels; int32_t Fs; };
std::mutex g_mu;
std::unordered_map<const void*, Entry> g_decoders;

Entry* find_decoder(const void* st) {
    auto it = g_decoders.find(st);
    return it == g_decoders.end() ? nullptr : &it->second;
}

bool valid_fs(int32_t Fs) { return Fs == 48000 || Fs == 24000 || Fs == 16000 || Fs == 12000 || Fs == 8000; }

// ---- multistream blobs: the public fields of the reference's struct first, the decoder behind them ----
struct MSBlob {
    OpusMSDecoder_t pub;
    OpusB200MSDecoder* impl;
};

// ---- the one open file of the player API (the reference keeps it in globals, src/opusfile.cpp:20-22) ----
OggOpusFileB200* g_file = nullptr;

int sd_read_trampoline(void*, unsigned char* buf, int nbytes) {
    if (!SD_read) return -1;  // no data source linked in: end of data
    return SD_read(buf, nbytes);
}

// ---- the one CELT-layer decoder (ref s_celtDec, src/celt.cpp:36) ----
struct CeltLayer {
    dm_ctx* ctx = nullptr;
    int channels = 0, stream_channels = 0, start = 0, end = 21, signalling = 1;
    const uint8_t* buf = nullptr;
    uint32_t storage = 0;
    uint8_t* h_bytes = nullptr;
    int16_t* h_pcm = nullptr;
    int32_t* h_status = nullptr;
    dm_frame* h_frame = nullptr;
} g_celt;

void celt_layer_free() {
    dm_host_free(g_celt.h_bytes); dm_host_free(g_celt.h_pcm); dm_host_free(g_celt.h_status); dm_host_free(g_celt.h_frame);
    dm_destroy(g_celt.ctx);
    g_celt = CeltLayer();
}

}  // namespace

// =====================================================================================================================
// packet API (ref src/opus_decoder.h:165-176)
// =====================================================================================================================
int opus_decoder_get_size(int channels) {
    if (channels < 1 || channels > 2) return 0;  // ref src/opus_decoder.cpp:68
    return (int)sizeof(Blob);
}

int opus_decoder_init(OpusDecoder* st, int32_t Fs, int channels) {
    if (!st || !valid_fs(Fs) || (channels != 1 && channels != 2)) return OPUS_BAD_ARG;  // ref :85-87
    std::lock_guard<std::mutex> lock(g_mu);
    Entry* e = find_decoder(st);
    int err = 0;
    if (e && e->channels == channels) {
        err = opusb_decoder_init(e->impl, Fs, channels);
        e->Fs = Fs;
    } else {
        if (e) { opusb_decoder_destroy(e->impl); g_decoders.erase(st); }
        OpusB200Decoder* impl = opusb_decoder_create(Fs, channels, &err);
        if (impl) g_decoders[st] = Entry{impl, channels, Fs};
    }
    if (err) return err;
    Blob* b = reinterpret_cast<Blob*>(st);
    b->tag = kTag; b->channels = channels; b->Fs = Fs;
    return OPUS_OK;
}

int opus_decode(OpusDecoder* st, uint8_t* data, int32_t len, int16_t* pcm, int frame_size) {
    if (frame_size <= 0) return OPUS_BAD_ARG;  // ref src/opus_decoder.cpp:354-356
    OpusB200Decoder* impl;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        Entry* e = find_decoder(st);
        if (!e) return OPUS_INVALID_STATE;  // never initialised
        impl = e->impl;
    }
    return opusb_decode(impl, data, len, pcm, frame_size);
}

int opus_decoder_ctl(OpusDecoder* st, int request, ...) {
    OpusB200Decoder* impl;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        Entry* e = find_decoder(st);
        if (!e) return OPUS_INVALID_STATE;
        impl = e->impl;
    }
    va_list ap;
    va_start(ap, request);
    int ret;
    switch (request) {  // argument shapes of ref src/opus_decoder.cpp:367-445
        case OPUS_RESET_STATE:
            ret = opusb_decoder_ctl(impl, request, nullptr);
            break;
        case OPUS_SET_GAIN_REQUEST:
        case OPUS_SET_PHASE_INVERSION_DISABLED_REQUEST: {
            int32_t v = va_arg(ap, int32_t);
            ret = opusb_decoder_ctl(impl, request, &v);
        } break;
        case OPUS_GET_FINAL_RANGE_REQUEST: {
            uint32_t* p = va_arg(ap, uint32_t*);
            int32_t v = 0;
            ret = p ? opusb_decoder_ctl(impl, request, &v) : OPUS_BAD_ARG;
            if (!ret) *p = (uint32_t)v;
        } break;
        case OPUS_GET_BANDWIDTH_REQUEST:
        case OPUS_GET_SAMPLE_RATE_REQUEST:
        case OPUS_GET_PITCH_REQUEST:
        case OPUS_GET_GAIN_REQUEST:
        case OPUS_GET_LAST_PACKET_DURATION_REQUEST:
        case OPUS_GET_PHASE_INVERSION_DISABLED_REQUEST: {
            int32_t* p = va_arg(ap, int32_t*);
            ret = p ? opusb_decoder_ctl(impl, request, p) : OPUS_BAD_ARG;
        } break;
        default:
            ret = OPUS_UNIMPLEMENTED;
            break;
    }
    va_end(ap);
    return ret;
}

void opus_decoder_destroy(OpusDecoder* st) {
    if (!st) return;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        if (Entry* e = find_decoder(st)) {
            opusb_decoder_destroy(e->impl);
            g_decoders.erase(st);
        }
    }
    free(st);  // ref src/opus_decoder.cpp:457
}

int opus_packet_parse(uint8_t* data, int32_t len, unsigned char* out_toc, uint8_t* frames[48], int16_t size[48], int* payload_offset) {
    return opusb_packet_parse(data, len, out_toc, const_cast<const uint8_t**>(frames), size, payload_offset);
}
int opus_packet_get_bandwidth(uint8_t* data) { return opusb_packet_get_bandwidth(data); }
int opus_packet_get_samples_per_frame(uint8_t* data, int32_t Fs) { return opusb_packet_get_samples_per_frame(data, Fs); }
int opus_packet_get_nb_channels(uint8_t* data) { return opusb_packet_get_nb_channels(data); }
int opus_packet_get_nb_frames(uint8_t packet[], int32_t len) { return opusb_packet_get_nb_frames(packet, len); }
int opus_packet_get_nb_samples(uint8_t packet[], int32_t len, int32_t Fs) { return opusb_packet_get_nb_samples(packet, len, Fs); }
int opus_decoder_get_nb_samples(const OpusDecoder* dec, uint8_t packet[], int32_t len) {  // ref src/opus_decoder.cpp:507
    std::lock_guard<std::mutex> lock(g_mu);
    Entry* e = find_decoder(dec);
    return opusb_packet_get_nb_samples(packet, len, e ? e->Fs : 48000);
}

// =====================================================================================================================
// multistream wrapper (ref src/opus_decoder.h:205-218)
// =====================================================================================================================
int32_t opus_multistream_decoder_get_size(int streams, int coupled_streams) {
    if (streams < 1 || coupled_streams > streams || coupled_streams < 0) return 0;  // ref src/opus_decoder.cpp:733
    return (int32_t)sizeof(MSBlob);
}

OpusMSDecoder_t* opus_multistream_decoder_create(int32_t Fs, int channels, int streams, int coupled_streams, const uint8_t* mapping,
                                                 int* error) {
    int err = 0;
    OpusB200MSDecoder* impl = opusb_multistream_decoder_create(Fs, channels, streams, coupled_streams, mapping, &err);
    MSBlob* b = nullptr;
    if (impl) {
        b = (MSBlob*)calloc(1, sizeof(MSBlob));
        if (!b) {
            opusb_multistream_decoder_destroy(impl);
            err = OPUS_ALLOC_FAIL;
        } else {
            b->pub.nb_channels = channels; b->pub.nb_streams = streams; b->pub.nb_coupled_streams = coupled_streams;
            memcpy(b->pub.mapping, mapping, (size_t)channels);
            b->impl = impl;
        }
    }
    if (error) *error = err;
    return b ? &b->pub : nullptr;
}

int opus_multistream_decode(OpusMSDecoder_t* st, uint8_t* data, int32_t len, int16_t* pcm, int frame_size) {
    if (!st) return OPUS_BAD_ARG;
    return opusb_multistream_decode(reinterpret_cast<MSBlob*>(st)->impl, data, len, pcm, frame_size);
}

int opus_multistream_decoder_ctl(OpusMSDecoder_t* st, int request, ...) {
    if (!st) return OPUS_BAD_ARG;
    OpusB200MSDecoder* impl = reinterpret_cast<MSBlob*>(st)->impl;
    va_list ap;
    va_start(ap, request);
    int ret;
    switch (request) {  // ref opus_multistream_decoder_ctl_va_list, src/opus_decoder.cpp:944-1031
        case OPUS_RESET_STATE:
            ret = opusb_multistream_decoder_ctl(impl, request, nullptr);
            break;
        case OPUS_SET_GAIN_REQUEST:
        case OPUS_SET_PHASE_INVERSION_DISABLED_REQUEST: {
            int32_t v = va_arg(ap, int32_t);
            ret = opusb_multistream_decoder_ctl(impl, request, &v);
        } break;
        case OPUS_GET_FINAL_RANGE_REQUEST: {
            uint32_t* p = va_arg(ap, uint32_t*);
            int32_t v = 0;
            ret = p ? opusb_multistream_decoder_ctl(impl, request, &v) : OPUS_BAD_ARG;
            if (!ret) *p = (uint32_t)v;
        } break;
        case OPUS_GET_BANDWIDTH_REQUEST:
        case OPUS_GET_SAMPLE_RATE_REQUEST:
        case OPUS_GET_GAIN_REQUEST:
        case OPUS_GET_LAST_PACKET_DURATION_REQUEST:
        case OPUS_GET_PHASE_INVERSION_DISABLED_REQUEST: {
            int32_t* p = va_arg(ap, int32_t*);
            ret = p ? opusb_multistream_decoder_ctl(impl, request, p) : OPUS_BAD_ARG;
        } break;
        default:
            ret = OPUS_UNIMPLEMENTED;
            break;
    }
    va_end(ap);
    return ret;
}

void opus_multistream_decoder_destroy(OpusMSDecoder_t* st) {
    if (!st) return;
    MSBlob* b = reinterpret_cast<MSBlob*>(st);
    opusb_multistream_decoder_destroy(b->impl);
    free(b);
}

// =====================================================================================================================
// codec layer (ref src/celt.h:662-664, 704, 711-712, 719; src/silk.h:1043-1044)
// =====================================================================================================================
// The reference allocates its global decoder state here; the GPU contexts are created by the *_init calls instead.
bool CELTDecoder_AllocateBuffers(void) { return true; }
void CELTDecoder_FreeBuffers() { celt_layer_free(); }
bool SILKDecoder_AllocateBuffers(void) { return true; }
void SILKDecoder_FreeBuffers() {}

int32_t celt_decoder_init(int32_t channels) {  // ref src/celt.cpp:1933
    if (channels < 0 || channels > 2) return ERR_OPUS_CELT_BAD_ARG;
    if (channels == 0) return ERR_OPUS_CELT_BAD_ARG;
    if (g_celt.ctx && g_celt.channels != channels) celt_layer_free();
    if (!g_celt.ctx) {
        int r = dm_create(&g_celt.ctx, 1, 1, channels, env_device());
        if (r) return r;
        g_celt.h_bytes = (uint8_t*)dm_host_alloc(1275 + 64);
        g_celt.h_pcm = (int16_t*)dm_host_alloc(sizeof(int16_t) * 960 * 2);
        g_celt.h_status = (int32_t*)dm_host_alloc(sizeof(int32_t));
        g_celt.h_frame = (dm_frame*)dm_host_alloc(sizeof(dm_frame));
        if (!g_celt.h_bytes || !g_celt.h_pcm || !g_celt.h_status || !g_celt.h_frame) { celt_layer_free(); return OPUS_ALLOC_FAIL; }
    } else {
        int r = dm_reset(g_celt.ctx, nullptr, 0);
        if (r) return r;
    }
    g_celt.channels = g_celt.stream_channels = channels;
    g_celt.start = 0; g_celt.end = 21; g_celt.signalling = 1;
    return OPUS_OK;
}

int32_t celt_decoder_ctl(int32_t request, ...) {  // ref src/celt.cpp:2448
    if (!g_celt.ctx) return OPUS_INVALID_STATE;
    va_list ap;
    va_start(ap, request);
    int32_t ret = OPUS_OK;
    switch (request) {
        case CELT_SET_START_BAND_REQUEST: {
            int32_t v = va_arg(ap, int32_t);
            if (v < 0 || v >= 21) ret = ERR_OPUS_CELT_BAD_ARG; else g_celt.start = v;
        } break;
        case CELT_SET_END_BAND_REQUEST: {
            int32_t v = va_arg(ap, int32_t);
            if (v < 1 || v > 21) ret = ERR_OPUS_CELT_BAD_ARG; else g_celt.end = v;  // stored, never read
        } break;
        case CELT_SET_CHANNELS_REQUEST: {
            int32_t v = va_arg(ap, int32_t);
            if (v < 1 || v > 2) ret = ERR_OPUS_CELT_BAD_ARG; else g_celt.stream_channels = v;
        } break;
        case CELT_SET_SIGNALLING_REQUEST:
            g_celt.signalling = va_arg(ap, int32_t);
            break;
        case OPUS_RESET_STATE:
            ret = dm_reset_state(g_celt.ctx, nullptr, 0);
            break;
        case OPUS_GET_FINAL_RANGE_REQUEST: {  // the decoder's rng after the last frame (ref s_celtDec->rng, :2437)
            uint32_t* p = va_arg(ap, uint32_t*);
            if (!p) { ret = ERR_OPUS_CELT_BAD_ARG; break; }
            void* blob = malloc(dm_stream_state_size());
            if (!blob) { ret = OPUS_ALLOC_FAIL; break; }
            ret = dm_get_stream_state(g_celt.ctx, 0, blob, dm_stream_state_size());
            if (!ret) memcpy(p, blob, sizeof(uint32_t));  // CeltStreamState begins with rng
            free(blob);
        } break;
        default:
            ret = OPUS_UNIMPLEMENTED;
            break;
    }
    va_end(ap);
    return ret;
}

void ec_dec_init(uint8_t* _buf, uint32_t _storage) {  // ref src/celt.cpp:2666; the bytes are read by the next celt_decode_with_ec
    g_celt.buf = _buf;
    g_celt.storage = _storage;
}

int32_t celt_decode_with_ec(int16_t* pcm, int32_t frame_size) {  // ref src/celt.cpp:2162
    if (!g_celt.ctx) return OPUS_INVALID_STATE;
    if (frame_size != 120 && frame_size != 240 && frame_size != 480 && frame_size != 960) return ERR_OPUS_CELT_BAD_ARG;  // ref :2208-2212
    if (g_celt.storage > 1275 || !pcm || !g_celt.buf) return ERR_OPUS_CELT_BAD_ARG;                                     // ref :2216
    if (g_celt.storage <= 1) return ERR_OPUS_CELT_BAD_ARG;                                                              // ref :2225
    memcpy(g_celt.h_bytes, g_celt.buf, g_celt.storage);
    g_celt.h_frame->offset = 0; g_celt.h_frame->len = (uint16_t)g_celt.storage; g_celt.h_frame->toc = 0; g_celt.h_frame->flags = 0;
    dm_batch in;
    memset(&in, 0, sizeof(in));
    in.n_frames = 1; in.mode = DM_MODE_CELT_RAW; in.frame_size = frame_size; in.coded_channels = g_celt.stream_channels;
    in.start_band = g_celt.start; in.bytes = g_celt.h_bytes; in.n_bytes = g_celt.storage; in.frames = g_celt.h_frame;
    dm_out out = {g_celt.h_pcm, g_celt.h_status};
    int r = decode_many(g_celt.ctx, &in, &out);
    if (r) return r;
    if (g_celt.h_status[0] < 0) return g_celt.h_status[0];
    memcpy(pcm, g_celt.h_pcm, sizeof(int16_t) * (size_t)frame_size * (size_t)g_celt.channels);
    return g_celt.h_status[0];  // samples per channel (downsample == 1, ref :2442)
}

// =====================================================================================================================
// player API (ref src/opusfile.h:141, 144, 156)
// =====================================================================================================================
int opus_head_parse(OpusHead_t* _head, uint8_t* _data, size_t _len) {
    static_assert(sizeof(OpusHead_t) == sizeof(OpusHeadB200), "OpusHead layout");
    return opusb_head_parse(reinterpret_cast<OpusHeadB200*>(_head), _data, _len);
}

OggOpusFile_t* opus_init_decoder() {
    if (g_file) { opusb_free_decoder(g_file); g_file = nullptr; }
    int err = 0;
    g_file = opusb_init_decoder(sd_read_trampoline, nullptr, env_device(), &err);
    return reinterpret_cast<OggOpusFile_t*>(g_file);
}

int op_read_stereo(int16_t* _pcm, int _buf_size) {
    if (!g_file) return OP_EINVAL;
    return opusb_read_stereo(g_file, _pcm, _buf_size);
}

void opus_free_decoder() {
    if (g_file) { opusb_free_decoder(g_file); g_file = nullptr; }
}
