/* oracle/ref_api.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the shipped library).
 *
 * extern "C" veneer over the *unmodified* reference decoder (schreibfaul1/ESP32-Opus-Player,
 * /root/reference/src/{celt,silk,opus_decoder,ogg,opusfile}.cpp) so that Python tests (ctypes) and
 * bench.py's cpu_baseline / --impl reference arm can call it.  The reference sources are compiled
 * where they lie by oracle/Makefile; only this file and oracle/shim/Arduino.h are ours.
 *
 * The reference keeps every piece of codec state in process-wide singletons (src/celt.cpp:36-39,
 * src/silk.cpp:18-29), so this library decodes ONE stream at a time per process.  Streams
 * are decoded back to back with opus_decoder_init() in between; parallelism across cores is by fork()
 * (ref_bench_fork below), exactly as BASELINE.md section 3 prescribes.
 */
#include <Arduino.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include "celt.h"
#include "silk.h"
#include "opus_decoder.h"
#include "opusfile.h"

extern celt_raw_ptr<CELTDecoder_t> s_celtDec; /* src/celt.cpp:36 */
uint32_t celt_pvq_u_row(uint32_t row, uint32_t data); /* src/celt.cpp:653 (not in celt.h) */

static bool g_setup = false;

/* ---- in-memory "SD card" for the Ogg player path (src/opusfile.h:19, src/main.cpp:264-269) ---- */
static const unsigned char* g_sd_data = nullptr;
static size_t g_sd_size = 0, g_sd_pos = 0;
int SD_read(unsigned char* buff, int nbytes) {
    if (nbytes == 0) return 0;
    if (g_sd_pos >= g_sd_size) return -1;
    size_t n = g_sd_size - g_sd_pos;
    if (n > (size_t)nbytes) n = (size_t)nbytes;
    memcpy(buff, g_sd_data + g_sd_pos, n);
    g_sd_pos += n;
    return (int)n;
}

/* Synthetic corpus generator shared by every arm: xorshift64, 21 high bits dropped. */
struct XorShift {
    uint64_t s;
    inline uint32_t next() {
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        return (uint32_t)(s >> 11);
    }
};

static inline uint64_t fnv_step(uint64_t h, int16_t v) { return (h ^ (uint16_t)v) * 1099511628211ULL; }

extern "C" {

int ref_setup(void) {
    if (g_setup) return 0;
    if (!SILKDecoder_AllocateBuffers()) return -1; /* src/silk.cpp:429 */
    if (!CELTDecoder_AllocateBuffers()) return -2; /* src/celt.cpp:664 */
    g_setup = true;
    return 0;
}

/* ---- packet level: opus_decoder_init + opus_decode (src/opus_decoder.cpp:82, 351) ---- */
void* ref_opus_new(int channels) {
    if (ref_setup()) return nullptr;
    int sz = opus_decoder_get_size(channels);
    if (sz <= 0) return nullptr;
    OpusDecoder* st = (OpusDecoder*)malloc(sz);
    if (st && opus_decoder_init(st, 48000, channels) != 0) {
        free(st);
        st = nullptr;
    }
    return st;
}
int ref_opus_init(void* st, int channels) { return opus_decoder_init((OpusDecoder*)st, 48000, channels); }
void ref_opus_free(void* st) { free(st); }
int ref_opus_decode(void* st, const uint8_t* pkt, int len, int16_t* pcm, int frame_size) {
    return opus_decode((OpusDecoder*)st, (uint8_t*)pkt, len, pcm, frame_size);
}

/* ---- CELT layer, any frame size (src/celt.cpp:1933, 2448, 2666, 2162) ---- */
int ref_celt_init(int channels) {
    if (ref_setup()) return -1;
    int r = celt_decoder_init(channels);
    if (r) return r;
    return celt_decoder_ctl(10016 /*CELT_SET_SIGNALLING*/, 0);
}
int ref_celt_set_stream_channels(int c) { return celt_decoder_ctl(10008, c); }
int ref_celt_set_start_band(int b) { return celt_decoder_ctl(10010, b); }
int ref_celt_reset_state(void) { return celt_decoder_ctl(4028 /*OPUS_RESET_STATE*/); }
int ref_celt_decode(const uint8_t* buf, int len, int16_t* pcm, int frame_size) {
    ec_dec_init((uint8_t*)buf, (uint32_t)len);
    return celt_decode_with_ec(pcm, frame_size);
}

/* ---- observable decoder state for stage-level parity checks ---- */
uint32_t ref_ec_rng(void) { return s_ec.rng; }
int ref_ec_error(void) { return s_ec.error; }
int ref_ec_tell(void) { return ec_tell(); }
uint32_t ref_celt_rng(void) { return s_celtDec->rng; }
/* energies: oldBandE, oldLogE, oldLogE2, backgroundLogE, each int16[2*21] (src/celt.cpp:2202-2206) */
void ref_celt_peek_energies(int16_t* out /*4*42*/) {
    const int CC = s_celtDec->channels;
    int16_t* lpc = (int16_t*)(s_celtDec->_decode_mem + (2048 + 120) * CC);
    memcpy(out, lpc + CC * 24, 4 * 42 * sizeof(int16_t));
}
/* post-filtered synthesis history: decode_mem[c][0 .. 2048+120) (src/celt.cpp:2221) */
void ref_celt_peek_mem(int c, int32_t* out /*2168*/) {
    memcpy(out, s_celtDec->_decode_mem + c * (2048 + 120), (2048 + 120) * sizeof(int32_t));
}
void ref_celt_peek_misc(int32_t* out /*10*/) {
    out[0] = s_celtDec->postfilter_period;
    out[1] = s_celtDec->postfilter_period_old;
    out[2] = s_celtDec->postfilter_gain;
    out[3] = s_celtDec->postfilter_gain_old;
    out[4] = s_celtDec->postfilter_tapset;
    out[5] = s_celtDec->postfilter_tapset_old;
    out[6] = s_celtDec->preemph_memD[0];
    out[7] = s_celtDec->preemph_memD[1];
    out[8] = (int32_t)s_celtDec->rng;
    out[9] = s_celtDec->error;
}

/* ---- primitives exported one by one for unit-level parity (all are global in the reference) ---- */
int32_t ref_cwrsi(int n, int k, uint32_t i, int32_t* y) { return cwrsi(n, k, i, y); }
uint32_t ref_pvq_v(int n, int k) {
    uint32_t a = celt_pvq_u_row(min(n, k), max(n, k));
    uint32_t b = celt_pvq_u_row(min(n, k + 1), max(n, k + 1));
    return a + b;
}
void ref_exp_rotation(int16_t* X, int len, int dir, int stride, int K, int spread) { exp_rotation(X, len, dir, stride, K, spread); }
void ref_normalise_residual(int32_t* iy, int16_t* X, int N, int32_t Ryy, int16_t gain) { normalise_residual(iy, X, N, Ryy, gain); }
void ref_renormalise_vector(int16_t* X, int N, int16_t gain) { renormalise_vector(X, N, gain); }
void ref_haar1(int16_t* X, int N0, int stride) { haar1(X, N0, stride); }
void ref_deinterleave_hadamard(int16_t* X, int N0, int stride, int h) { deinterleave_hadamard(X, N0, stride, h); }
void ref_interleave_hadamard(int16_t* X, int N0, int stride, int h) { interleave_hadamard(X, N0, stride, h); }
void ref_stereo_merge(int16_t* X, int16_t* Y, int16_t mid, int N) { stereo_merge(X, Y, mid, N); }
int16_t ref_bitexact_cos(int16_t x) { return bitexact_cos(x); }
int32_t ref_bitexact_log2tan(int32_t isin, int32_t icos) { return bitexact_log2tan(isin, icos); }
int16_t ref_celt_rsqrt_norm(int32_t x) { return celt_rsqrt_norm(x); }
int32_t ref_celt_sqrt(int32_t x) { return celt_sqrt(x); }
int16_t ref_celt_cos_norm(int32_t x) { return celt_cos_norm(x); }
int32_t ref_celt_rcp(int32_t x) { return celt_rcp(x); }
int32_t ref_celt_exp2(int16_t x) { return celt_exp2(x); }
uint32_t ref_isqrt32(uint32_t v) { return isqrt32(v); }
void ref_denormalise_bands(const int16_t* X, int32_t* freq, const int16_t* bandLogE, int start, int end, int M, int silence) {
    denormalise_bands(X, freq, bandLogE, start, end, M, 1, silence);
}
void ref_mdct_backward(int32_t* in, int32_t* out, int overlap, int shift, int stride) { clt_mdct_backward(in, out, overlap, shift, stride); }
void ref_comb_filter(int32_t* y, int32_t* x, int T0, int T1, int N, int16_t g0, int16_t g1, int tap0, int tap1) {
    comb_filter(y, x, T0, T1, N, g0, g1, tap0, tap1);
}
int ref_compute_allocation(int start, int end, const int32_t* offsets, const int32_t* cap, int alloc_trim, int32_t* intensity,
                           int32_t* dual_stereo, int32_t total, int32_t* balance, int32_t* pulses, int32_t* ebits,
                           int32_t* fine_priority, int C, int LM, const uint8_t* buf, int len) {
    ec_dec_init((uint8_t*)buf, (uint32_t)len);
    return clt_compute_allocation(start, end, offsets, cap, alloc_trim, intensity, dual_stereo, total, balance, pulses, ebits,
                                  fine_priority, C, LM, 0, 0, 0);
}
void ref_init_caps(int32_t* cap, int LM, int C) { init_caps(cap, LM, C); }

/* ---- Ogg player path (src/opusfile.cpp:784, 1293) from a memory image ---- */
int ref_ogg_open(const uint8_t* data, size_t size) {
    g_sd_data = data;
    g_sd_size = size;
    g_sd_pos = 0;
    OggOpusFile_t* f = opus_init_decoder();
    g_setup = true; /* opus_init_decoder allocates the codec singletons itself */
    return f ? 0 : -1;
}
int ref_ogg_read_stereo(int16_t* pcm, int buf_size) { return op_read_stereo(pcm, buf_size); }

/* ---- synthetic streams: packet = TOC byte + (len-1) xorshift bytes ---- */
void ref_gen_packet(uint64_t* state, int toc /* <0: no TOC byte, all random */, int len, uint8_t* out) {
    XorShift g{*state};
    int j = 0;
    if (toc >= 0) out[j++] = (uint8_t)toc;
    for (; j < len; j++) out[j] = (uint8_t)g.next();
    *state = g.s;
}

/* Decode one synthetic stream at packet level; returns FNV hash over interleaved L,R (mono duplicated).
 * pcm_out (optional) receives n_frames*960*channels samples. */
uint64_t ref_decode_synth_stream(int toc, int len, int channels, int n_frames, uint64_t seed, int16_t* pcm_out, int* err) {
    static OpusDecoder* st = nullptr;
    static int st_ch = 0;
    if (err) *err = 0;
    if (ref_setup()) { if (err) *err = -1; return 0; }
    if (!st || st_ch != channels) {
        free(st);
        st = (OpusDecoder*)malloc(opus_decoder_get_size(channels));
        st_ch = channels;
    }
    opus_decoder_init(st, 48000, channels);
    XorShift g{seed};
    uint8_t pkt[1500];
    int16_t pcm[5760 * 2];
    uint64_t h = 1469598103934665603ULL;
    for (int f = 0; f < n_frames; f++) {
        pkt[0] = (uint8_t)toc;
        for (int j = 1; j < len; j++) pkt[j] = (uint8_t)g.next();
        int n = opus_decode(st, pkt, len, pcm, 5760);
        if (n < 0) { if (err) *err = n; return h; }
        for (int i = 0; i < n; i++) {
            if (channels == 2) { h = fnv_step(h, pcm[2 * i]); h = fnv_step(h, pcm[2 * i + 1]); }
            else { h = fnv_step(h, pcm[i]); h = fnv_step(h, pcm[i]); }
        }
        if (pcm_out) { memcpy(pcm_out, pcm, sizeof(int16_t) * n * channels); pcm_out += n * channels; }
    }
    return h;
}

/* CELT-layer synthetic stream (all bytes random, no TOC); hash over ret*channels samples. */
uint64_t ref_decode_synth_celt(int frame_size, int channels, int len, int n_frames, uint64_t seed, int16_t* pcm_out, uint32_t* final_rng) {
    ref_celt_init(channels);
    ref_celt_set_stream_channels(channels);
    XorShift g{seed};
    uint8_t pkt[1500];
    int16_t pcm[960 * 2];
    uint64_t h = 1469598103934665603ULL;
    for (int f = 0; f < n_frames; f++) {
        for (int j = 0; j < len; j++) pkt[j] = (uint8_t)g.next();
        int n = ref_celt_decode(pkt, len, pcm, frame_size);
        if (n < 0) break;
        for (int i = 0; i < n * channels; i++) h = fnv_step(h, pcm[i]);
        if (pcm_out) { memcpy(pcm_out, pcm, sizeof(int16_t) * n * channels); pcm_out += n * channels; }
    }
    if (final_rng) *final_rng = s_ec.rng;
    return h;
}

/* ---- CPU baseline: one forked worker per job, contiguous share of streams each.
 * Stream i uses seed (seed_base ^ i).  Returns wall seconds of the slowest worker; hashes[i] per stream. */
static double now_s() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
double ref_bench_fork(int toc, int len, int channels, int n_streams, int n_frames, uint64_t seed_base, int jobs, uint64_t* hashes) {
    if (jobs < 1) jobs = 1;
    if (jobs > n_streams) jobs = n_streams;
    size_t shm_sz = sizeof(uint64_t) * n_streams + sizeof(double) * jobs;
    void* shm = mmap(nullptr, shm_sz, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (shm == MAP_FAILED) return -1.0;
    uint64_t* sh_hash = (uint64_t*)shm;
    double* sh_t = (double*)(sh_hash + n_streams);
    ref_setup();
    double t0 = now_s();
    pid_t* pids = (pid_t*)calloc(jobs, sizeof(pid_t));
    for (int j = 0; j < jobs; j++) {
        pid_t p = fork();
        if (p == 0) {
            int lo = (int)((int64_t)n_streams * j / jobs), hi = (int)((int64_t)n_streams * (j + 1) / jobs);
            double a = now_s();
            for (int s = lo; s < hi; s++) sh_hash[s] = ref_decode_synth_stream(toc, len, channels, n_frames, seed_base ^ (uint64_t)s, nullptr, nullptr);
            sh_t[j] = now_s() - a;
            _exit(0);
        }
        pids[j] = p;
    }
    for (int j = 0; j < jobs; j++) { int stt; if (pids[j] > 0) waitpid(pids[j], &stt, 0); }
    double wall = now_s() - t0;
    if (hashes) memcpy(hashes, sh_hash, sizeof(uint64_t) * n_streams);
    free(pids);
    munmap(shm, shm_sz);
    return wall;
}

/* Same for an explicit list of stream ids (stream ids[k] uses seed seed_base ^ ids[k]), at packet level (frame_size == 0:
 * packets of `len` bytes behind `toc`, hash with mono duplicated as ref_decode_synth_stream does) or at the CELT layer
 * (frame_size 120/240/480/960: `len` random bytes per frame, no TOC, as ref_decode_synth_celt does). */
double ref_bench_fork_ids(int toc, int len, int channels, int frame_size, const int64_t* ids, int n_streams, int n_frames, uint64_t seed_base,
                          int jobs, uint64_t* hashes) {
    if (n_streams <= 0) return 0.0;
    if (jobs < 1) jobs = 1;
    if (jobs > n_streams) jobs = n_streams;
    size_t shm_sz = sizeof(uint64_t) * n_streams;
    void* shm = mmap(nullptr, shm_sz, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (shm == MAP_FAILED) return -1.0;
    uint64_t* sh_hash = (uint64_t*)shm;
    ref_setup();
    double t0 = now_s();
    pid_t* pids = (pid_t*)calloc(jobs, sizeof(pid_t));
    for (int j = 0; j < jobs; j++) {
        pid_t p = fork();
        if (p == 0) {
            int lo = (int)((int64_t)n_streams * j / jobs), hi = (int)((int64_t)n_streams * (j + 1) / jobs);
            for (int s = lo; s < hi; s++) {
                const uint64_t seed = seed_base ^ (uint64_t)ids[s];
                sh_hash[s] = frame_size ? ref_decode_synth_celt(frame_size, channels, len, n_frames, seed, nullptr, nullptr)
                                        : ref_decode_synth_stream(toc, len, channels, n_frames, seed, nullptr, nullptr);
            }
            _exit(0);
        }
        pids[j] = p;
    }
    for (int j = 0; j < jobs; j++) { int stt; if (pids[j] > 0) waitpid(pids[j], &stt, 0); }
    double wall = now_s() - t0;
    if (hashes) memcpy(hashes, sh_hash, sizeof(uint64_t) * n_streams);
    free(pids);
    munmap(shm, shm_sz);
    return wall;
}

/* ---- demux only: the reference's page framer + CRC, packetizer and time-stamp logic over one memory
 * image, nothing decoded -- opus_init_decoder(), then op_fetch_and_process_page() (src/opusfile.cpp:835) page after page with
 * the buffered packets marked as consumed.  Returns the audio packets seen (< 0: the file does not open). */
} /* extern "C" */
int op_fetch_and_process_page(ogg_page* _og, int64_t _page_offset, int _spanp, int _ignore_holes); /* src/opusfile.cpp:835 */
extern OggOpusFile_t* m_OggOpusFile; /* src/opusfile.cpp:20 */
extern "C" {

long ref_ogg_demux_only(const uint8_t* data, size_t size, long long* packet_bytes) {
    if (ref_ogg_open(data, size) != 0) return -1;
    long n = 0;
    long long nb = 0;
    for (;;) {
        for (int i = m_OggOpusFile->op_pos; i < m_OggOpusFile->op_count; i++) { n++; nb += m_OggOpusFile->op[i].bytes; }
        m_OggOpusFile->op_pos = m_OggOpusFile->op_count;
        if (op_fetch_and_process_page(NULL, -1, 1, 0) < 0) break; /* OP_EOF, or OP_EREAD at the end of the image */
    }
    if (packet_bytes) *packet_bytes = nb;
    return n;
}

/* The same image `n_files` times, shared out over `jobs` forked workers (the reference's demux state is a process-wide
 * singleton too): wall seconds, packets of one file in *packets. */
double ref_demux_fork(const uint8_t* data, size_t size, int n_files, int jobs, long* packets) {
    if (jobs < 1) jobs = 1;
    if (jobs > n_files) jobs = n_files;
    long* sh = (long*)mmap(nullptr, sizeof(long), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (sh == MAP_FAILED) return -1.0;
    ref_setup();
    double t0 = now_s();
    pid_t* pids = (pid_t*)calloc(jobs, sizeof(pid_t));
    for (int j = 0; j < jobs; j++) {
        pid_t p = fork();
        if (p == 0) {
            int lo = (int)((int64_t)n_files * j / jobs), hi = (int)((int64_t)n_files * (j + 1) / jobs);
            long n = 0;
            for (int f = lo; f < hi; f++) n = ref_ogg_demux_only(data, size, nullptr);
            if (j == 0) *sh = n;
            _exit(0);
        }
        pids[j] = p;
    }
    for (int j = 0; j < jobs; j++) { int stt; if (pids[j] > 0) waitpid(pids[j], &stt, 0); }
    double wall = now_s() - t0;
    if (packets) *packets = *sh;
    free(pids);
    munmap(sh, sizeof(long));
    return wall;
}

} /* extern "C" */
