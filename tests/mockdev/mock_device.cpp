// tests/mockdev/mock_device.cpp -- TEST INFRASTRUCTURE: a CPU stand-in for the dm_* C ABI (include/opus_b200.h) behind
// csrc/opus_player.cpp, so that the host side of the player layer -- RoundScheduler (plan / stage / finish, rounds in
// flight, decoder resets at link boundaries, private copies of unread samples) and Track::read -- can be tested and
// timed without a GPU.  It decodes nothing: the "PCM" of a frame is a checksum of exactly what the scheduler handed the
// device for it -- the frame's bytes, its TOC, and how many frames the stream has decoded since its last reset -- so a
// test can tell from the samples op_read_stereo returns whether every frame reached the device once, in order, with the
// right bytes and behind the right resets.  Built with opus_player.cpp + host_api.cpp into libplayer_mock.so
// (tests/conftest.py); never part of the product.
//
// Rules (mirrored by the model in tests/test_player_scheduler_mock.py):
//   status   960, or -4 for a frame whose payload starts with 0xEE; a DMF_CONT frame behind a failed frame of its packet and
//            DMF_NULL fillers get -2004 and change nothing
//   pcm      sample i of channel c = (int16)(h + 131 i + 7 c), h = FNV-1a over (seq, toc, payload bytes)
//   seq      frames decoded on the stream since dm_create / dm_reset (-> 0) or dm_reset_state (-> 1 << 20)
//   async    decode_many_async does the work at once; dm_wait has nothing to wait for
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/opus_b200.h"

struct dm_ctx {
    int n_streams, max_frames, channels;
    std::vector<uint32_t> seq;
    std::vector<uint8_t> failed;  // the multi-frame packet being decoded has a failed frame
    long long calls;
};

extern "C" {

int dm_create(dm_ctx** out, int n_streams, int max_frames, int channels, int device) {
    (void)device;
    if (!out || n_streams < 1 || max_frames < 1 || channels < 1 || channels > 2) return -1;
    dm_ctx* c = new dm_ctx();
    c->n_streams = n_streams; c->max_frames = max_frames; c->channels = channels; c->calls = 0;
    c->seq.assign((size_t)n_streams, 0);
    c->failed.assign((size_t)n_streams, 0);
    *out = c;
    return 0;
}
void dm_destroy(dm_ctx* c) { delete c; }
static int reset_to(dm_ctx* c, const int32_t* ids, int n, uint32_t v) {
    if (!c) return -1;
    if (!ids) { for (auto& s : c->seq) s = v; for (auto& f : c->failed) f = 0; return 0; }
    for (int i = 0; i < n; i++) {
        if (ids[i] < 0 || ids[i] >= c->n_streams) return -1;
        c->seq[(size_t)ids[i]] = v; c->failed[(size_t)ids[i]] = 0;
    }
    return 0;
}
int dm_reset(dm_ctx* c, const int32_t* ids, int n) { return reset_to(c, ids, n, 0); }
int dm_reset_state(dm_ctx* c, const int32_t* ids, int n) { return reset_to(c, ids, n, 1u << 20); }

int decode_many(dm_ctx* c, const dm_batch* in, dm_out* out) {
    if (!c || !in || !out || in->n_frames < 1 || in->n_frames > c->max_frames || in->mode != DM_MODE_OPUS) return -1;
    const int F = in->n_frames, CC = c->channels;
    for (int s = 0; s < c->n_streams; s++)
        for (int f = 0; f < F; f++) {
            const dm_frame& fr = in->frames[(size_t)s * F + f];
            int32_t& st = out->status[(size_t)s * F + f];
            int16_t* pcm = out->pcm + ((size_t)s * F + f) * 960 * CC;
            if (fr.flags & DMF_NULL) { st = -2004; continue; }
            if (!(fr.flags & DMF_CONT)) c->failed[(size_t)s] = 0;
            if (c->failed[(size_t)s]) { st = -2004; continue; }
            if ((size_t)fr.offset + fr.len > in->n_bytes) return -2;
            const uint8_t* p = in->bytes + fr.offset;
            static const bool fast = getenv("MOCK_DEVICE_FAST") != nullptr;  // benchmarks of the host side: every frame "decodes", no samples written
            if (fast) { c->seq[(size_t)s]++; st = 960; continue; }
            if (fr.len > 0 && p[0] == 0xEE) { st = -4; c->failed[(size_t)s] = 1; continue; }
            uint32_t h = 2166136261u;
            auto mix = [&](uint32_t v) { h = (h ^ v) * 16777619u; };
            const uint32_t q = c->seq[(size_t)s]++;
            mix(q & 255); mix((q >> 8) & 255); mix((q >> 16) & 255); mix(fr.toc);
            for (int i = 0; i < fr.len; i++) mix(p[i]);
            for (int i = 0; i < 960; i++)
                for (int ch = 0; ch < CC; ch++) pcm[(size_t)i * CC + ch] = (int16_t)(h + 131u * (uint32_t)i + 7u * (uint32_t)ch);
            st = 960;
        }
    c->calls++;
    return 0;
}
int decode_many_async(dm_ctx* c, const dm_batch* in, dm_out* out, long long* ticket) {
    const int r = decode_many(c, in, out);
    if (!r && ticket) *ticket = c->calls;
    return r;
}
int dm_wait(dm_ctx* c, long long ticket) { return (c && ticket >= 0 && ticket <= c->calls) ? 0 : -1; }
void* dm_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void dm_host_free(void* p) { free(p); }
long long dm_launch_count(const dm_ctx* c) { return c ? c->calls : 0; }

}  // extern "C"
