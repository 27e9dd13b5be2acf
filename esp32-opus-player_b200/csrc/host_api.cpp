// host_api.cpp -- host-only part of libopusb200.so: Opus packet framing helpers, the synthetic corpus
// generator and the single-stream drop-in that mirrors the reference's packet API
// (src/opus_decoder.h:165-168) on top of decode_many().  No decoding happens on the CPU here.
#include <stdlib.h>
#include <string.h>

#include <thread>
#include <vector>

#include "../../include/opus_b200.h"

namespace {

const int kOpusBadArg = -1, kOpusBufferTooSmall = -2, kOpusInvalidPacket = -4, kOpusUnimplemented = -5;

// ref parse_size (src/opus_decoder.cpp:524)
int frame_size_field(const uint8_t* data, int32_t len, int16_t* size) {
    if (len < 1) { *size = -1; return -1; }
    if (data[0] < 252) { *size = data[0]; return 1; }
    if (len < 2) { *size = -1; return -1; }
    *size = (int16_t)(4 * data[1] + data[0]);
    return 2;
}

struct XorShift64 {  // SURVEY 8c generator
    uint64_t s;
    uint32_t next() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 11); }
};

}  // namespace

struct OpusB200Decoder {
    dm_ctx* ctx;
    int channels;
    int32_t Fs;
    // what opus_decoder_ctl reports (ref struct OpusDecoder, src/opus_decoder.cpp:40-59)
    int bandwidth, last_packet_duration, decode_gain;
    uint8_t* h_bytes;   // pinned
    int16_t* h_pcm;     // pinned
    int32_t* h_status;  // pinned
    dm_frame* h_frames; // pinned
};

extern "C" {

// ref opus_packet_get_mode (src/opus_decoder.cpp:135)
int opusb_packet_get_mode(const uint8_t* data) {
    if (data[0] & 0x80) return 1002;
    if ((data[0] & 0x60) == 0x60) return 1001;
    return 1000;
}
// ref opus_packet_get_bandwidth (src/opus_decoder.cpp:460)
int opusb_packet_get_bandwidth(const uint8_t* data) {
    int bw;
    if (data[0] & 0x80) {
        bw = 1102 + ((data[0] >> 5) & 0x3);
        if (bw == 1102) bw = 1101;
    } else if ((data[0] & 0x60) == 0x60) {
        bw = (data[0] & 0x10) ? 1105 : 1104;
    } else {
        bw = 1101 + ((data[0] >> 5) & 0x3);
    }
    return bw;
}
// ref opus_packet_get_samples_per_frame (src/opus_decoder.cpp:541)
int opusb_packet_get_samples_per_frame(const uint8_t* data, int32_t Fs) {
    int audiosize;
    if (data[0] & 0x80) {
        audiosize = ((data[0] >> 3) & 0x3);
        audiosize = (Fs << audiosize) / 400;
    } else if ((data[0] & 0x60) == 0x60) {
        audiosize = (data[0] & 0x08) ? Fs / 50 : Fs / 100;
    } else {
        audiosize = ((data[0] >> 3) & 0x3);
        audiosize = audiosize == 3 ? Fs * 60 / 1000 : (Fs << audiosize) / 100;
    }
    return audiosize;
}
int opusb_packet_get_nb_channels(const uint8_t* data) { return (data[0] & 0x4) ? 2 : 1; }
// ref opus_packet_get_nb_frames (src/opus_decoder.cpp:477)
int opusb_packet_get_nb_frames(const uint8_t* packet, int32_t len) {
    if (len < 1) return kOpusBadArg;
    int count = packet[0] & 0x3;
    if (count == 0) return 1;
    if (count != 3) return 2;
    if (len < 2) return kOpusInvalidPacket;
    return packet[1] & 0x3F;
}
// ref opus_packet_get_nb_samples (src/opus_decoder.cpp:492)
int opusb_packet_get_nb_samples(const uint8_t* packet, int32_t len, int32_t Fs) {
    int count = opusb_packet_get_nb_frames(packet, len);
    if (count < 0) return count;
    int samples = count * opusb_packet_get_samples_per_frame(packet, Fs);
    if (samples * 25 > Fs * 3) return kOpusInvalidPacket;
    return samples;
}

// ref opus_packet_parse_impl with self_delimited = 0 (src/opus_decoder.cpp:559): TOC, frame count code 0-3,
// padding, CBR/VBR sizes.  Returns the frame count or a negative error.
int opusb_packet_parse(const uint8_t* data, int32_t len, uint8_t* out_toc, const uint8_t* frames[48], int16_t size[48],
                       int* payload_offset) {
    if (!size || len < 0) return kOpusBadArg;
    if (len == 0) return kOpusInvalidPacket;
    const uint8_t* data0 = data;
    const int framesize = opusb_packet_get_samples_per_frame(data, 48000);
    const uint8_t toc = *data++;
    len--;
    int32_t last_size = len;
    int count, cbr = 0, bytes;
    switch (toc & 0x3) {
        case 0: count = 1; break;
        case 1:
            count = 2; cbr = 1;
            if (len & 0x1) return kOpusInvalidPacket;
            last_size = len / 2;
            size[0] = (int16_t)last_size;
            break;
        case 2:
            count = 2;
            bytes = frame_size_field(data, len, size);
            len -= bytes;
            if (size[0] < 0 || size[0] > len) return kOpusInvalidPacket;
            data += bytes;
            last_size = len - size[0];
            break;
        default: {
            if (len < 1) return kOpusInvalidPacket;
            const uint8_t ch = *data++;
            count = ch & 0x3F;
            if (count <= 0 || framesize * (int32_t)count > 5760) return kOpusInvalidPacket;
            len--;
            if (ch & 0x40) {  // padding
                int p;
                do {
                    if (len <= 0) return kOpusInvalidPacket;
                    p = *data++;
                    len--;
                    len -= p == 255 ? 254 : p;
                } while (p == 255);
            }
            if (len < 0) return kOpusInvalidPacket;
            cbr = !(ch & 0x80);
            if (!cbr) {
                last_size = len;
                for (int i = 0; i < count - 1; i++) {
                    bytes = frame_size_field(data, len, size + i);
                    len -= bytes;
                    if (size[i] < 0 || size[i] > len) return kOpusInvalidPacket;
                    data += bytes;
                    last_size -= bytes + size[i];
                }
                if (last_size < 0) return kOpusInvalidPacket;
            } else {
                last_size = len / count;
                if (last_size * count != len) return kOpusInvalidPacket;
                for (int i = 0; i < count - 1; i++) size[i] = (int16_t)last_size;
            }
        } break;
    }
    if (last_size > 1275) return kOpusInvalidPacket;
    size[count - 1] = (int16_t)last_size;
    if (payload_offset) *payload_offset = (int)(data - data0);
    for (int i = 0; i < count; i++) {
        if (frames) frames[i] = data;
        data += size[i];
    }
    if (out_toc) *out_toc = toc;
    return count;
}

// ---- synthetic corpus ----
size_t dm_synth_bytes(int n_streams, int n_frames, int payload_len) {
    return (size_t)n_streams * (size_t)n_frames * (size_t)payload_len;
}
void dm_synth_fill(uint8_t* bytes, dm_frame* frames, int n_streams, int n_frames, int payload_len, int toc, uint64_t seed_base,
                   int first_stream) {
    for (int s = 0; s < n_streams; s++) {
        XorShift64 g{seed_base ^ (uint64_t)(first_stream + s)};
        for (int f = 0; f < n_frames; f++) {
            const size_t t = (size_t)s * n_frames + f;
            uint8_t* p = bytes + t * (size_t)payload_len;
            for (int j = 0; j < payload_len; j++) p[j] = (uint8_t)g.next();
            frames[t].offset = (uint32_t)(t * (size_t)payload_len);
            frames[t].len = (uint16_t)payload_len;
            frames[t].toc = (uint8_t)toc;
            frames[t].flags = 0;
        }
    }
}

// Raw CELT frames (DM_MODE_CELT_RAW): every byte of a frame is payload (SURVEY 8c "CELT layer" generator).
void dm_synth_fill_raw(uint8_t* bytes, dm_frame* frames, int n_streams, int n_frames, int frame_len, uint64_t seed_base, int first_stream) {
    dm_synth_fill(bytes, frames, n_streams, n_frames, frame_len, 0, seed_base, first_stream);
}

// Result check of a benchmark or a server: the SURVEY 8c stream hash (FNV-1a over the uint16 view of the interleaved
// samples; a mono stream counts every sample twice when dup_mono is set, as the packet-level oracle does) carried on
// over consecutive calls.  pcm is [n_streams][samples_per_stream] with samples_per_stream = frames * frame_size *
// channels; hashes[n_streams] is updated in place (start value 1469598103934665603).  Streams are shared out over
// n_threads host threads (0 = one per core).
void dm_fnv_update(uint64_t* hashes, const int16_t* pcm, int n_streams, size_t samples_per_stream, int channels, int dup_mono, int n_threads) {
    if (!hashes || !pcm || n_streams <= 0) return;
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > n_streams) nt = n_streams;
    const bool twice = channels == 1 && dup_mono;
    auto work = [=](int lo, int hi) {
        for (int s = lo; s < hi; s++) {
            uint64_t h = hashes[s];
            const uint16_t* p = reinterpret_cast<const uint16_t*>(pcm + (size_t)s * samples_per_stream);
            if (twice) for (size_t i = 0; i < samples_per_stream; i++) { h = (h ^ p[i]) * 1099511628211ull; h = (h ^ p[i]) * 1099511628211ull; }
            else for (size_t i = 0; i < samples_per_stream; i++) h = (h ^ p[i]) * 1099511628211ull;
            hashes[s] = h;
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(work, (int)((long long)n_streams * t / nt), (int)((long long)n_streams * (t + 1) / nt));
    work(0, (int)((long long)n_streams / nt));
    for (auto& t : th) t.join();
}

// ---- single-stream drop-in (ref opus_decoder_init / opus_decode, src/opus_decoder.cpp:82, 351) ----
OpusB200Decoder* opusb_decoder_create(int32_t Fs, int channels, int* error) {
    int err = 0;
    OpusB200Decoder* st = nullptr;
    if ((Fs != 48000 && Fs != 24000 && Fs != 16000 && Fs != 12000 && Fs != 8000) || (channels != 1 && channels != 2)) {
        err = kOpusBadArg;
    } else {
        st = (OpusB200Decoder*)calloc(1, sizeof(*st));
        err = st ? dm_create(&st->ctx, 1, 48, channels, 0) : -7;
        if (!err) {
            st->channels = channels;
            st->Fs = Fs;
            st->h_bytes = (uint8_t*)dm_host_alloc(48 * 1275 + 64);
            st->h_pcm = (int16_t*)dm_host_alloc(sizeof(int16_t) * 48 * 960 * 2);
            st->h_status = (int32_t*)dm_host_alloc(sizeof(int32_t) * 48);
            st->h_frames = (dm_frame*)dm_host_alloc(sizeof(dm_frame) * 48);
            if (!st->h_bytes || !st->h_pcm || !st->h_status || !st->h_frames) err = -7;
        }
        if (err) { opusb_decoder_destroy(st); st = nullptr; }
    }
    if (error) *error = err;
    return st;
}
void opusb_decoder_destroy(OpusB200Decoder* st) {
    if (!st) return;
    dm_host_free(st->h_bytes); dm_host_free(st->h_pcm); dm_host_free(st->h_status); dm_host_free(st->h_frames);
    dm_destroy(st->ctx);
    free(st);
}
int opusb_decoder_init(OpusB200Decoder* st, int32_t Fs, int channels) {
    if (!st || (Fs != 48000 && Fs != 24000 && Fs != 16000 && Fs != 12000 && Fs != 8000) || channels != st->channels) return kOpusBadArg;
    st->Fs = Fs; st->bandwidth = 0; st->last_packet_duration = 0; st->decode_gain = 0;
    return dm_reset(st->ctx, nullptr, 0);
}
// ref opus_decode -> opus_decode_native (src/opus_decoder.cpp:280): parse the packet, decode each frame as
// 960 samples, return samples per channel.
int opusb_decode(OpusB200Decoder* st, const uint8_t* data, int32_t len, int16_t* pcm, int frame_size) {
    if (!st || frame_size <= 0) return kOpusBadArg;
    if (len <= 0 || !data) return kOpusBadArg;  // no PLC/FEC in the reference either
    uint8_t toc;
    int16_t size[48];
    const uint8_t* frames[48];
    int off;
    const int count = opusb_packet_parse(data, len, &toc, frames, size, &off);
    if (count < 0) return count;
    const int spf = opusb_packet_get_samples_per_frame(data, 48000);
    if (count * spf > frame_size) return kOpusBufferTooSmall;  // ref src/opus_decoder.cpp:320-321
    // Every frame is decoded as 960 samples whatever its TOC says.  For a TOC frame size other than 20 ms
    // the reference trips assert(ret == packet_frame_size) (src/opus_decoder.cpp:341) or, without asserts, overruns
    // the caller's buffer: such packets are refused here, and nothing is ever written past frame_size samples.
    if (spf != 960) return kOpusUnimplemented;
    size_t pos = 0;
    for (int i = 0; i < count; i++) {
        memcpy(st->h_bytes + pos, frames[i], (size_t)size[i]);
        st->h_frames[i].offset = (uint32_t)pos;
        st->h_frames[i].len = (uint16_t)size[i];
        st->h_frames[i].toc = toc;
        st->h_frames[i].flags = i ? DMF_CONT : 0;  // a failing frame ends the packet (ref src/opus_decoder.cpp:334-338)
        pos += (size_t)size[i];
    }
    dm_batch in;
    memset(&in, 0, sizeof(in));
    in.n_frames = count; in.mode = DM_MODE_OPUS; in.bytes = st->h_bytes; in.n_bytes = pos; in.frames = st->h_frames;
    dm_out out = {st->h_pcm, st->h_status};
    st->bandwidth = opusb_packet_get_bandwidth(data);  // set once the packet parsed, before its frames run (ref :325-329)
    int r = decode_many(st->ctx, &in, &out);
    if (r) return r;
    int n = 0;
    for (int i = 0; i < count; i++) {
        if (st->h_status[i] < 0) return st->h_status[i];
        if (n + st->h_status[i] > frame_size) return kOpusBufferTooSmall;  // cannot happen behind the checks above
        memcpy(pcm + (size_t)n * st->channels, st->h_pcm + (size_t)i * 960 * st->channels, sizeof(int16_t) * 960 * st->channels);
        n += st->h_status[i];
    }
    st->last_packet_duration = n;
    return n;
}

// ref opus_decoder_ctl (src/opus_decoder.cpp:361-454).  One int32 argument: written by GET requests, read by SET requests.
int opusb_decoder_ctl(OpusB200Decoder* st, int request, int32_t* value) {
    if (!st) return kOpusBadArg;
    switch (request) {
        case 4009:  // OPUS_GET_BANDWIDTH_REQUEST
            if (!value) return kOpusBadArg;
            *value = st->bandwidth;
            return 0;
        case 4031:  // OPUS_GET_FINAL_RANGE_REQUEST: the reference never stores it
            if (!value) return kOpusBadArg;
            *value = 0;
            return 0;
        case 4028:  // OPUS_RESET_STATE
            st->bandwidth = 0; st->last_packet_duration = 0;
            return dm_reset_state(st->ctx, nullptr, 0);
        case 4029:  // OPUS_GET_SAMPLE_RATE_REQUEST
            if (!value) return kOpusBadArg;
            *value = st->Fs;
            return 0;
        case 4045:  // OPUS_GET_GAIN_REQUEST (stored, never applied: SURVEY F4)
            if (!value) return kOpusBadArg;
            *value = st->decode_gain;
            return 0;
        case 4034:  // OPUS_SET_GAIN_REQUEST
            if (!value || *value < -32768 || *value > 32767) return kOpusBadArg;
            st->decode_gain = *value;
            return 0;
        case 4039:  // OPUS_GET_LAST_PACKET_DURATION_REQUEST
            if (!value) return kOpusBadArg;
            *value = st->last_packet_duration;
            return 0;
        case 4046:  // OPUS_SET_PHASE_INVERSION_DISABLED_REQUEST: the reference forwards the VALUE as the CELT request code
        case 4047:  // (src/opus_decoder.cpp:428, 435), which CELT answers with OPUS_UNIMPLEMENTED; phase inversion stays as
            if (request == 4046 && value && (*value < 0 || *value > 1)) return kOpusBadArg;  // celt_decoder_init set it
            if (request == 4047 && !value) return kOpusBadArg;
            return -5;
        default:    // incl. OPUS_GET_PITCH_REQUEST (the decoder state it reads is loss-concealment bookkeeping, not kept here)
            return -5;  // OPUS_UNIMPLEMENTED
    }
}

// ---- multistream wrapper (ref opus_multistream_decoder_create / _decode / _ctl / _destroy, src/opus_decoder.cpp:781-1046)
// for the layouts the player uses: ONE Opus stream (mono or coupled) mapped onto 1..255 output channels.
struct OpusB200MSDecoder {
    OpusB200Decoder* dec;
    int nb_channels, coupled;
    uint8_t mapping[256];
    int16_t* tmp;
};
OpusB200MSDecoder* opusb_multistream_decoder_create(int32_t Fs, int channels, int streams, int coupled_streams, const uint8_t* mapping,
                                                    int* error) {
    int err = 0;
    OpusB200MSDecoder* st = nullptr;
    if (channels > 255 || channels < 1 || coupled_streams > streams || streams < 1 || coupled_streams < 0 || streams > 255 - coupled_streams ||
        !mapping) {
        err = kOpusBadArg;
    } else if (streams != 1) {
        err = -5;  // several multiplexed streams: not needed by the player (mapping family 0, or 1 with one stream)
    } else {
        for (int c = 0; c < channels && !err; c++)
            if (mapping[c] != 255 && mapping[c] >= streams + coupled_streams) err = kOpusBadArg;  // ref validate_layout
    }
    if (!err) {
        st = (OpusB200MSDecoder*)calloc(1, sizeof(*st));
        if (!st) err = -7;
    }
    if (!err) {
        st->nb_channels = channels; st->coupled = coupled_streams;
        memcpy(st->mapping, mapping, (size_t)channels);
        st->tmp = (int16_t*)malloc(sizeof(int16_t) * 5760 * 2);
        st->dec = opusb_decoder_create(Fs, coupled_streams ? 2 : 1, &err);
        if (!st->tmp && !err) err = -7;
        if (err) { opusb_multistream_decoder_destroy(st); st = nullptr; }
    }
    if (error) *error = err;
    return st;
}
void opusb_multistream_decoder_destroy(OpusB200MSDecoder* st) {
    if (!st) return;
    opusb_decoder_destroy(st->dec);
    free(st->tmp);
    free(st);
}
int opusb_multistream_decode(OpusB200MSDecoder* st, const uint8_t* data, int32_t len, int16_t* pcm, int frame_size) {
    if (!st || frame_size <= 0 || len < 0) return kOpusBadArg;
    if (len < 1) return kOpusBadArg;  // len == 0 is packet-loss concealment, which the reference cannot do either
    if (frame_size > 5760) frame_size = 5760;
    const int n = opusb_decode(st->dec, data, len, st->tmp, frame_size);
    if (n <= 0) return n;
    const int dc = st->coupled ? 2 : 1;
    for (int c = 0; c < st->nb_channels; c++) {
        const int src = st->mapping[c];
        for (int i = 0; i < n; i++) pcm[(size_t)i * st->nb_channels + c] = src == 255 ? (int16_t)0 : st->tmp[(size_t)i * dc + src];
    }
    return n;
}
int opusb_multistream_decoder_ctl(OpusB200MSDecoder* st, int request, int32_t* value) {
    if (!st) return kOpusBadArg;
    return opusb_decoder_ctl(st->dec, request, value);  // one stream: every request goes to it (ref :944-1031)
}

}  // extern "C"
