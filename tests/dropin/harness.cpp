// tests/dropin/harness.cpp -- TEST INFRASTRUCTURE: a stand-in for the reference firmware's main.cpp on a host.
//
// Written against include/opusfile.h + include/opus_decoder.h only (the drop-in headers), it is built two ways by
// tests/test_dropin_cpu.py / tests/test_gpu_dropin.py:
//   harness_dropin   harness.cpp + libopusb200_dropin.so                      (player + packet + CELT-layer API from the drop-in)
//   harness_reflink  harness.cpp + the reference's UNMODIFIED opusfile.o, ogg.o (oracle/_ref) + libopusb200_dropin.so
//                    -- the reference's own player code calling opus_multistream_* / opus_packet_* of the drop-in in
//                    place of its opus_decoder.cpp, celt.cpp and silk.cpp.
// Modes (results go to a binary file the Python test reads back):
//   player <in.ogg> <out> <buf_size>                   opus_init_decoder + op_read_stereo loop (ref src/main.cpp:270-282)
//   packet <toc> <len> <channels> <frames> <seed> <out> opus_decoder_get_size/init + opus_decode on SURVEY 8c packets
//   celt <frame_size> <channels> <len> <frames> <seed> <out>  celt_decoder_init/ctl + ec_dec_init + celt_decode_with_ec
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "opusfile.h"

static std::vector<unsigned char> g_sd;
static size_t g_pos = 0;
int SD_read(unsigned char* buff, int nbytes) {  // the firmware's SD card (ref src/main.cpp:264-269)
    if (nbytes == 0) return 0;
    if (g_pos >= g_sd.size()) return -1;
    size_t n = g_sd.size() - g_pos;
    if (n > (size_t)nbytes) n = (size_t)nbytes;
    memcpy(buff, g_sd.data() + g_pos, n);
    g_pos += n;
    return (int)n;
}

struct XorShift {  // SURVEY 8c generator
    uint64_t s;
    uint32_t next() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 11); }
};

static void put(FILE* f, const void* p, size_t n) { if (n && fwrite(p, 1, n, f) != n) { perror("write"); exit(3); } }

static int run_player(const char* in, const char* out, int buf_size) {
    FILE* fi = fopen(in, "rb");
    if (!fi) { perror(in); return 2; }
    fseek(fi, 0, SEEK_END);
    g_sd.resize((size_t)ftell(fi));
    fseek(fi, 0, SEEK_SET);
    if (fread(g_sd.data(), 1, g_sd.size(), fi) != g_sd.size()) return 2;
    fclose(fi);
    OggOpusFile_t* of = opus_init_decoder();
    std::vector<int32_t> rets;
    std::vector<int16_t> pcm, buf((size_t)buf_size);
    if (of) {
        for (int k = 0; k < 1000000; k++) {
            const int r = op_read_stereo(buf.data(), buf_size);
            rets.push_back(r);
            if (r > 0) pcm.insert(pcm.end(), buf.begin(), buf.begin() + 2 * r);
            if (r == 0 || r == OP_EREAD) break;
        }
    }
    FILE* fo = fopen(out, "wb");
    if (!fo) { perror(out); return 2; }
    const int32_t opened = of ? 1 : 0, nr = (int32_t)rets.size();
    const int64_t ns = (int64_t)pcm.size();
    put(fo, &opened, 4); put(fo, &nr, 4); put(fo, &ns, 8);
    put(fo, rets.data(), rets.size() * 4);
    put(fo, pcm.data(), pcm.size() * 2);
    fclose(fo);
    return of ? 0 : 4;
}

static int run_packet(int toc, int len, int channels, int frames, uint64_t seed, const char* out) {
    OpusDecoder* st = (OpusDecoder*)malloc((size_t)opus_decoder_get_size(channels));
    int r = opus_decoder_init(st, 48000, channels);
    if (r) { fprintf(stderr, "opus_decoder_init: %d\n", r); return 4; }
    XorShift g{seed};
    std::vector<uint8_t> pkt((size_t)len);
    std::vector<int16_t> pcm(5760 * 2), all;
    std::vector<int32_t> rets;
    for (int f = 0; f < frames; f++) {
        pkt[0] = (uint8_t)toc;
        for (int j = 1; j < len; j++) pkt[j] = (uint8_t)g.next();
        const int n = opus_decode(st, pkt.data(), len, pcm.data(), 5760);
        rets.push_back(n);
        if (n > 0) all.insert(all.end(), pcm.begin(), pcm.begin() + (size_t)n * channels);
    }
    int32_t bw = 0, dur = 0, rate = 0;
    uint32_t rng = 77;
    const int c0 = opus_decoder_ctl(st, OPUS_GET_BANDWIDTH_REQUEST, &bw), c1 = opus_decoder_ctl(st, OPUS_GET_LAST_PACKET_DURATION_REQUEST, &dur);
    const int c2 = opus_decoder_ctl(st, OPUS_GET_SAMPLE_RATE_REQUEST, &rate), c3 = opus_decoder_ctl(st, OPUS_GET_FINAL_RANGE_REQUEST, &rng);
    const int c4 = opus_decoder_ctl(st, 12345);
    const int nbs = opus_decoder_get_nb_samples(st, pkt.data(), len);
    // a second stream on the same blob: init again = the reference's reset between streams (SURVEY 8c "multi-stream use")
    r = opus_decoder_init(st, 48000, channels);
    XorShift g2{seed};
    for (int f = 0; f < 2 && !r; f++) {
        pkt[0] = (uint8_t)toc;
        for (int j = 1; j < len; j++) pkt[j] = (uint8_t)g2.next();
        const int n = opus_decode(st, pkt.data(), len, pcm.data(), 5760);
        rets.push_back(n);
        if (n > 0) all.insert(all.end(), pcm.begin(), pcm.begin() + (size_t)n * channels);
    }
    opus_decoder_destroy(st);
    FILE* fo = fopen(out, "wb");
    if (!fo) { perror(out); return 2; }
    const int32_t hdr[10] = {(int32_t)rets.size(), c0, bw, c1, dur, c2, rate, c3 ? c3 : (int32_t)rng, c4, nbs};
    const int64_t ns = (int64_t)all.size();
    put(fo, hdr, sizeof(hdr)); put(fo, &ns, 8);
    put(fo, rets.data(), rets.size() * 4);
    put(fo, all.data(), all.size() * 2);
    fclose(fo);
    return 0;
}

static int run_celt(int frame_size, int channels, int len, int frames, uint64_t seed, const char* out) {
    if (!CELTDecoder_AllocateBuffers()) return 4;
    int r = celt_decoder_init(channels);
    if (r) { fprintf(stderr, "celt_decoder_init: %d\n", r); return 4; }
    celt_decoder_ctl(CELT_SET_SIGNALLING_REQUEST, 0);
    celt_decoder_ctl(CELT_SET_CHANNELS_REQUEST, channels);
    XorShift g{seed};
    std::vector<uint8_t> pkt((size_t)len);
    std::vector<int16_t> pcm(960 * 2), all;
    std::vector<int32_t> rets;
    for (int f = 0; f < frames; f++) {
        for (int j = 0; j < len; j++) pkt[j] = (uint8_t)g.next();
        ec_dec_init(pkt.data(), (uint32_t)len);
        const int n = celt_decode_with_ec(pcm.data(), frame_size);
        rets.push_back(n);
        if (n > 0) all.insert(all.end(), pcm.begin(), pcm.begin() + (size_t)n * channels);
    }
    uint32_t rng = 0;
    celt_decoder_ctl(OPUS_GET_FINAL_RANGE_REQUEST, &rng);
    CELTDecoder_FreeBuffers();
    FILE* fo = fopen(out, "wb");
    if (!fo) { perror(out); return 2; }
    const int32_t hdr[2] = {(int32_t)rets.size(), (int32_t)rng};
    const int64_t ns = (int64_t)all.size();
    put(fo, hdr, sizeof(hdr)); put(fo, &ns, 8);
    put(fo, rets.data(), rets.size() * 4);
    put(fo, all.data(), all.size() * 2);
    fclose(fo);
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 5 && !strcmp(argv[1], "player")) return run_player(argv[2], argv[3], atoi(argv[4]));
    if (argc >= 8 && !strcmp(argv[1], "packet"))
        return run_packet((int)strtol(argv[2], 0, 0), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), strtoull(argv[6], 0, 0), argv[7]);
    if (argc >= 8 && !strcmp(argv[1], "celt"))
        return run_celt(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), strtoull(argv[6], 0, 0), argv[7]);
    fprintf(stderr, "usage: see the head of harness.cpp\n");
    return 1;
}
