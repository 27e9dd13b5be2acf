// opus_b200.cu -- host side of libopusb200.so: context, launch sequence, pinned-buffer pipeline and
// the C ABI declared in include/opus_b200.h.  Build: nvcc -gencode arch=compute_100a,code=sm_100a.
// There is no CPU decode path in this library: without a CUDA device every decode call fails.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "../../include/opus_b200.h"
#include "kernels.cuh"

using namespace opusb200;

static_assert(sizeof(FrameDesc) == sizeof(dm_frame), "dm_frame layout");
static_assert(sizeof(BandOp) == 16, "BandOp layout");

enum { kStages = 8, kMaxChunks = 4, kRing = 4 };
static const char* kStageNames[kStages] = {"entropy", "sequence", "leaves", "bands", "imdct", "postfilter", "deemph", "silk"};

struct dm_ctx {
    int n_streams, max_frames, channels, device, n_sm;
    int leaf_slots;  // capacity of the sorted leaf list per frame (OPUSB_LEAF_SLOTS lowers it: tests of the overflow path)
    int silk_lanes;  // chains per warp of the SILK lane-per-chain kernels; 0 = silk_chain_lanes() (OPUSB_SILK_LANES overrides, for tuning)
    // per-stream state
    CeltStreamState* d_state;
    SilkStreamState* d_sstate;
    // per-frame intermediates
    CeltFrameHdr* d_hdr;
    BandOp* d_ops;
    SilkFrame* d_silk;
    SilkCtl* d_sctl;         // SILK stage hand-offs (silk_kernels.cuh)
    SilkPost* d_spost;
    int16_t* d_sxq;
    int16_t* d_slr;
    int16_t* d_sup;
    int* d_any_silk;         // [kMaxChunks] a frame of the stream range of the call has a SILK layer
    int* d_fix_count;        // per stream range of a call (kMaxChunks): filters queued for k_silk_fix
    FixEntry* d_sfix;        // LPC filters that failed their first stability check (two per SILK frame-channel at most)
    int16_t* d_X;
    LeafSort* d_leafsort;    // [kMaxChunks]: counting sort of the coded leaves of a launch by cost class
    uint32_t* d_leaf_order;  // the sorted leaf list, kLeafSlotsPerFrame entries per frame of the context
    FrameEnergy* d_fe;
    int32_t* d_raw;
    int32_t* d_syn;
    // staging for the host-buffer path
    uint8_t* d_bytes[2];     // packet bytes, double buffered so that call n+1 can upload while call n still parses
    size_t d_bytes_cap[2];
    FrameDesc* d_frames;
    int16_t* d_pcm;
    int32_t* d_status;
    int16_t* d_pcm2;         // second output buffer, allocated on the first decode_many_async
    int32_t* d_status2;
    cudaStream_t streams[kMaxChunks];
    cudaStream_t side[kMaxChunks];        // per stream range: the SILK kernels of a call run beside its CELT kernels
    cudaEvent_t fork_ev[kMaxChunks], join_ev[kMaxChunks];
    cudaEvent_t bytes_ready;
    // calls in flight (decode_many_async): per chunk stream, "packets parsed" and "results in host memory"
    cudaEvent_t parsed[kRing][kMaxChunks], done[kRing][kMaxChunks];
    cudaEvent_t uploaded[kRing], kernels_done[kRing];
    long long next_ticket;
    cudaEvent_t ev[kStages + 1];
    long long launches;
    size_t dev_bytes;
    int last_cuda_error;
};

#define CU_TRY(ctx, expr)                                                                   \
    do {                                                                                    \
        cudaError_t e__ = (expr);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            if (ctx) (ctx)->last_cuda_error = (int)e__;                                     \
            fprintf(stderr, "opusb200: CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); \
            return e__ == cudaErrorMemoryAllocation ? DM_ERR_NOMEM : DM_ERR_CUDA;           \
        }                                                                                   \
    } while (0)

template <class T>
static int dev_alloc(dm_ctx* c, T** p, size_t n) {
    CU_TRY(c, cudaMalloc((void**)p, n * sizeof(T)));
    c->dev_bytes += n * sizeof(T);
    return 0;
}

static void init_state_host(CeltStreamState& s) {
    memset(&s, 0, sizeof(s));
    for (int i = 0; i < kNbEBands; i++)
        for (int c = 0; c < 2; c++) s.band[i].logE[c] = s.band[i].logE2[c] = -(28 << 10);  // ref src/celt.cpp:2496
}

extern "C" {

}  // extern "C"

// Everything dm_create allocates; on failure the caller tears the half-built context down with dm_destroy.
static int create_impl(dm_ctx* c, int n_streams, int max_frames, int channels, int device) {
    CU_TRY(c, cudaSetDevice(device));
    CU_TRY(c, cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, device));
    c->leaf_slots = kLeafSlotsPerFrame;  // (the context was just zeroed)
    if (const char* e = getenv("OPUSB_LEAF_SLOTS")) { const int v = atoi(e); if (v >= 0 && v <= kLeafSlotsPerFrame) c->leaf_slots = v; }
    if (const char* e = getenv("OPUSB_SILK_LANES")) { const int v = atoi(e); if (v == 8 || v == 16 || v == 32) c->silk_lanes = v; }
    // The entropy kernel keeps its per-thread allocation arrays and the explicit partition stack in local memory.  The
    // stack limit is a DEVICE-wide setting shared with every other CUDA user of the process (torch): it is only ever
    // raised, and only as far as this library's largest frame needs (no recursion anywhere, so ptxas knows the size).
    {
        cudaFuncAttributes fa;
        CU_TRY(c, cudaFuncGetAttributes(&fa, k_entropy));
        size_t cur = 0;
        CU_TRY(c, cudaDeviceGetLimit(&cur, cudaLimitStackSize));
        const size_t need = (fa.localSizeBytes + 255) / 256 * 256;
        if (need > cur) CU_TRY(c, cudaDeviceSetLimit(cudaLimitStackSize, need));
    }
    const size_t nfr = (size_t)n_streams * max_frames;
    int r;
    if ((r = dev_alloc(c, &c->d_state, (size_t)n_streams))) return r;
    if ((r = dev_alloc(c, &c->d_sstate, (size_t)n_streams))) return r;
    if ((r = dev_alloc(c, &c->d_silk, nfr))) return r;
    if ((r = dev_alloc(c, &c->d_sctl, nfr * 2))) return r;
    if ((r = dev_alloc(c, &c->d_spost, nfr))) return r;
    if ((r = dev_alloc(c, &c->d_sxq, nfr * 2 * kXqRow))) return r;
    if ((r = dev_alloc(c, &c->d_slr, nfr * 2 * kXqRow))) return r;
    if ((r = dev_alloc(c, &c->d_sup, nfr * 2 * kUpRow))) return r;
    if ((r = dev_alloc(c, &c->d_any_silk, kMaxChunks))) return r;
    if ((r = dev_alloc(c, &c->d_fix_count, kMaxChunks))) return r;
    if ((r = dev_alloc(c, &c->d_sfix, nfr * 4))) return r;
    CU_TRY(c, cudaFuncSetAttribute(k_silk_core<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, silk_core_smem(32)));
    if ((r = dev_alloc(c, &c->d_hdr, nfr))) return r;
    if ((r = dev_alloc(c, &c->d_ops, nfr * kMaxOps))) return r;
    if ((r = dev_alloc(c, &c->d_X, nfr * 2 * kMaxFrame))) return r;
    if ((r = dev_alloc(c, &c->d_leafsort, (size_t)kMaxChunks))) return r;
    if ((r = dev_alloc(c, &c->d_leaf_order, nfr * kLeafSlotsPerFrame))) return r;
    if ((r = dev_alloc(c, &c->d_fe, nfr))) return r;
    if ((r = dev_alloc(c, &c->d_raw, nfr * channels * kMaxFrame))) return r;
    if ((r = dev_alloc(c, &c->d_syn, nfr * channels * kMaxFrame))) return r;
    if ((r = dev_alloc(c, &c->d_frames, nfr))) return r;
    if ((r = dev_alloc(c, &c->d_pcm, nfr * channels * kMaxFrame))) return r;
    if ((r = dev_alloc(c, &c->d_status, nfr))) return r;
    for (int i = 0; i < kMaxChunks; i++) CU_TRY(c, cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking));
    for (int i = 0; i < kMaxChunks; i++) {
        // the side stream carries few, long-running warps: give its blocks the first free slots
        int prio_lo = 0, prio_hi = 0;
        CU_TRY(c, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CU_TRY(c, cudaStreamCreateWithPriority(&c->side[i], cudaStreamNonBlocking, prio_hi));
        CU_TRY(c, cudaEventCreateWithFlags(&c->fork_ev[i], cudaEventDisableTiming));
        CU_TRY(c, cudaEventCreateWithFlags(&c->join_ev[i], cudaEventDisableTiming));
    }
    CU_TRY(c, cudaEventCreateWithFlags(&c->bytes_ready, cudaEventDisableTiming));
    for (int r = 0; r < kRing; r++)
        for (int k = 0; k < kMaxChunks; k++) {
            CU_TRY(c, cudaEventCreateWithFlags(&c->parsed[r][k], cudaEventDisableTiming));
            CU_TRY(c, cudaEventCreateWithFlags(&c->done[r][k], cudaEventDisableTiming));
        }
    for (int r = 0; r < kRing; r++) {
        CU_TRY(c, cudaEventCreateWithFlags(&c->uploaded[r], cudaEventDisableTiming));
        CU_TRY(c, cudaEventCreateWithFlags(&c->kernels_done[r], cudaEventDisableTiming));
    }
    for (int i = 0; i <= kStages; i++) CU_TRY(c, cudaEventCreate(&c->ev[i]));
    CU_TRY(c, cudaMemset(c->d_spost, 0, nfr * sizeof(SilkPost)));  // has_out of frames no SILK kernel has visited yet
    CU_TRY(c, cudaMemset(c->d_any_silk, 0, kMaxChunks * sizeof(int)));
    return dm_reset(c, nullptr, 0);
}

extern "C" {

int dm_create(dm_ctx** out, int n_streams, int max_frames, int channels, int device) {
    if (!out || n_streams < 1 || max_frames < 1 || (channels != 1 && channels != 2)) return OPUSB_BAD_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        fprintf(stderr, "opusb200: no usable CUDA device (this library has no CPU path)\n");
        return DM_ERR_NO_DEVICE;
    }
    dm_ctx* c = new (std::nothrow) dm_ctx();
    if (!c) return DM_ERR_NOMEM;
    memset(c, 0, sizeof(*c));
    c->n_streams = n_streams; c->max_frames = max_frames; c->channels = channels; c->device = device;
    const int r = create_impl(c, n_streams, max_frames, channels, device);
    if (r) {
        dm_destroy(c);  // frees whatever was allocated (every pointer of the zeroed context is null or valid)
        return r;
    }
    *out = c;
    return 0;
}

void dm_destroy(dm_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    cudaFree(c->d_state); cudaFree(c->d_sstate); cudaFree(c->d_silk); cudaFree(c->d_sctl); cudaFree(c->d_spost); cudaFree(c->d_sxq); cudaFree(c->d_slr);
    cudaFree(c->d_sup); cudaFree(c->d_any_silk); cudaFree(c->d_sfix); cudaFree(c->d_fix_count); cudaFree(c->d_hdr); cudaFree(c->d_ops); cudaFree(c->d_X); cudaFree(c->d_leafsort); cudaFree(c->d_leaf_order); cudaFree(c->d_fe);
    cudaFree(c->d_raw); cudaFree(c->d_syn); cudaFree(c->d_bytes[0]); cudaFree(c->d_bytes[1]); cudaFree(c->d_frames); cudaFree(c->d_pcm); cudaFree(c->d_pcm2); cudaFree(c->d_status2);
    cudaFree(c->d_status);
    for (int i = 0; i < kMaxChunks; i++) if (c->streams[i]) cudaStreamDestroy(c->streams[i]);
    for (int i = 0; i < kMaxChunks; i++) {
        if (c->side[i]) cudaStreamDestroy(c->side[i]);
        if (c->fork_ev[i]) cudaEventDestroy(c->fork_ev[i]);
        if (c->join_ev[i]) cudaEventDestroy(c->join_ev[i]);
    }
    if (c->bytes_ready) cudaEventDestroy(c->bytes_ready);
    for (int r = 0; r < kRing; r++)
        for (int k = 0; k < kMaxChunks; k++) {
            if (c->parsed[r][k]) cudaEventDestroy(c->parsed[r][k]);
            if (c->done[r][k]) cudaEventDestroy(c->done[r][k]);
        }
    for (int r = 0; r < kRing; r++) {
        if (c->uploaded[r]) cudaEventDestroy(c->uploaded[r]);
        if (c->kernels_done[r]) cudaEventDestroy(c->kernels_done[r]);
    }
    for (int i = 0; i <= kStages; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    delete c;
}

int dm_reset(dm_ctx* c, const int32_t* ids, int n) {
    if (!c) return OPUSB_BAD_ARG;
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaDeviceSynchronize());  // calls in flight (decode_many_async) finish first
    CeltStreamState* init = (CeltStreamState*)malloc(sizeof(CeltStreamState));
    if (!init) return DM_ERR_NOMEM;
    init_state_host(*init);
    // SILK layer: silk_InitDecoder (ref src/silk.cpp:1792) + zeroed resamplers
    SilkStreamState* si = (SilkStreamState*)calloc(1, sizeof(SilkStreamState));
    if (!si) { free(init); return DM_ERR_NOMEM; }
    for (int n = 0; n < 2; n++) { si->ch[n].prev_gain_Q16 = 65536; si->ch[n].first_frame_after_reset = 1; }  // ref silk_init_decoder :2192
    // one pristine blob of each kind goes to the device, a kernel copies them over the chosen streams
    const size_t wc = sizeof(CeltStreamState) / 4, ws = sizeof(SilkStreamState) / 4;
    static_assert(sizeof(CeltStreamState) % 4 == 0 && sizeof(SilkStreamState) % 4 == 0, "state blobs are copied word-wise");
    if (ids) for (int i = 0; i < n; i++) if (ids[i] < 0 || ids[i] >= c->n_streams) { free(init); free(si); return OPUSB_BAD_ARG; }
    const int count = ids ? n : c->n_streams;
    uint32_t* d_tmpl = nullptr;
    int32_t* d_ids = nullptr;
    cudaError_t e = count > 0 ? cudaMalloc((void**)&d_tmpl, (wc + ws) * 4) : cudaSuccess;
    if (e == cudaSuccess && count > 0) e = cudaMemcpy(d_tmpl, init, wc * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && count > 0) e = cudaMemcpy(d_tmpl + wc, si, ws * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && ids && count > 0) {
        e = cudaMalloc((void**)&d_ids, sizeof(int32_t) * (size_t)count);
        if (e == cudaSuccess) e = cudaMemcpy(d_ids, ids, sizeof(int32_t) * (size_t)count, cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess && count > 0) {
        k_init_streams<<<count, 256>>>(d_ids, (uint32_t*)c->d_state, (uint32_t*)c->d_sstate, d_tmpl, d_tmpl + wc, (int)wc, (int)ws);
        c->launches++;
        e = cudaDeviceSynchronize();
    }
    cudaFree(d_tmpl); cudaFree(d_ids);
    free(init); free(si);
    CU_TRY(c, e);
    return 0;
}

int dm_reset_state(dm_ctx* c, const int32_t* ids, int n) {
    if (!c) return OPUSB_BAD_ARG;
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaDeviceSynchronize());  // calls in flight (decode_many_async) finish first
    std::vector<int32_t> all;
    if (!ids) {
        all.resize((size_t)c->n_streams);
        for (int i = 0; i < c->n_streams; i++) all[(size_t)i] = i;
        ids = all.data();
        n = c->n_streams;
    }
    if (n <= 0) return 0;
    for (int i = 0; i < n; i++)
        if (ids[i] < 0 || ids[i] >= c->n_streams) return OPUSB_BAD_ARG;
    int32_t* d_ids = nullptr;
    CU_TRY(c, cudaMalloc((void**)&d_ids, sizeof(int32_t) * (size_t)n));
    cudaError_t e = cudaMemcpy(d_ids, ids, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        k_reset_state<<<(n + 127) / 128, 128>>>(d_ids, n, c->d_state, c->d_sstate);
        c->launches++;
        e = cudaDeviceSynchronize();
    }
    cudaFree(d_ids);
    CU_TRY(c, e);
    return 0;
}

static int check_batch(const dm_ctx* c, const dm_batch* in) {
    if (!c || !in || !in->frames || (!in->bytes && in->n_bytes)) return OPUSB_BAD_ARG;
    if (in->n_frames < 1 || in->n_frames > c->max_frames) return DM_ERR_CAPACITY;
    if (in->mode == DM_MODE_CELT_RAW) {
        if (in->frame_size != 120 && in->frame_size != 240 && in->frame_size != 480 && in->frame_size != 960) return OPUSB_BAD_ARG;
        if (in->coded_channels != 1 && in->coded_channels != 2) return OPUSB_BAD_ARG;
        if (in->start_band < 0 || in->start_band >= kNbEBands) return OPUSB_BAD_ARG;
    } else if (in->mode != DM_MODE_OPUS) {
        return OPUSB_BAD_ARG;
    }
    return 0;
}

static int reserve_bytes(dm_ctx* c, int bi, size_t n) {
    if (n <= c->d_bytes_cap[bi]) return 0;
    if (c->d_bytes[bi]) { cudaFree(c->d_bytes[bi]); c->dev_bytes -= c->d_bytes_cap[bi]; }
    c->d_bytes[bi] = nullptr; c->d_bytes_cap[bi] = 0;
    const size_t cap = n + n / 4 + 4096;
    CU_TRY(c, cudaMalloc((void**)&c->d_bytes[bi], cap));
    c->d_bytes_cap[bi] = cap; c->dev_bytes += cap;
    return 0;
}

// Enqueue the kernels of one call for streams [first, first+n_sel) on `s`.  Pointers are device pointers.
static int launch_pipeline(dm_ctx* c, const dm_batch* in, const uint8_t* d_bytes, const FrameDesc* d_frames, int16_t* d_pcm,
                           int32_t* d_status, int first, int n_sel, cudaStream_t s, bool timed, cudaEvent_t parsed = nullptr, int chunk = 0) {
    DevParams P;
    P.n_streams = c->n_streams; P.n_frames = in->n_frames; P.mode = in->mode;
    P.frame_size = in->mode == DM_MODE_CELT_RAW ? in->frame_size : 960;
    P.coded_channels = in->coded_channels; P.start_band = in->start_band; P.CC = c->channels;
    P.first_stream = first; P.n_sel = n_sel;
    const int nfr = n_sel * in->n_frames;
    const int wb = 32 * KW_WARPS;
    if (timed) cudaEventRecord(c->ev[0], s);
    int* any_silk = c->d_any_silk + chunk;  // per stream range: ranges of one call run concurrently on their own streams
    cudaMemsetAsync(any_silk, 0, sizeof(int), s);
    k_entropy<<<(nfr + K1_THREADS - 1) / K1_THREADS, K1_THREADS, 0, s>>>(d_bytes, d_frames, P, c->d_hdr, c->d_ops, c->d_silk, c->d_state, any_silk);
    if (timed) cudaEventRecord(c->ev[1], s);
    if (parsed) cudaEventRecord(parsed, s);
    k_sequence<<<(n_sel + KW_WARPS - 1) / KW_WARPS, wb, 0, s>>>(c->d_hdr, d_frames, c->d_fe, c->d_state, P);
    if (timed) cudaEventRecord(c->ev[2], s);
    // The SILK layer of a call needs nothing from its CELT layer until the final sum, and its two long kernels (LTP/LPC
    // synthesis, up-sampler) are single-wave recurrences that leave most issue slots idle: they run on a side stream
    // beside the issue-bound CELT kernels and meet them again in front of k_silk_fir.  Every one of these returns at
    // once unless k_entropy saw a frame with a SILK layer.
    const bool silk = in->mode == DM_MODE_OPUS;
    SilkBufs B = {c->d_silk, c->d_sctl, c->d_spost, c->d_sxq, c->d_slr, c->d_sup, c->d_sstate, any_silk};
    const int wide = c->n_sm * 8;  // grid-stride kernels: a fixed grid, so an idle call costs a few empty launches
    if (silk) {
        cudaStream_t ss = c->side[chunk];
        cudaEventRecord(c->fork_ev[chunk], s);
        cudaStreamWaitEvent(ss, c->fork_ev[chunk], 0);
        const int chains = n_sel * P.CC;
        // the stream range's own slice of the fix queue (at most two filters per frame and channel) and its own counter
        FixEntry* fq = c->d_sfix + (size_t)first * in->n_frames * 4;
        cudaMemsetAsync(c->d_fix_count + chunk, 0, sizeof(int), ss);
        k_silk_params<<<c->n_sm * 12, 128, 0, ss>>>(c->d_hdr, c->d_fe, B, P, fq, c->d_fix_count + chunk);
        k_silk_fix<<<c->n_sm * 8, 128, 0, ss>>>(any_silk, fq, c->d_fix_count + chunk);
        const int lanes = c->silk_lanes > 0 ? c->silk_lanes : silk_chain_lanes(chains, c->n_sm);
        if (lanes == 8) k_silk_core<8><<<(chains + 7) / 8, 32, silk_core_smem(8), ss>>>(c->d_hdr, c->d_fe, B, P);
        else if (lanes == 16) k_silk_core<16><<<(chains + 15) / 16, 32, silk_core_smem(16), ss>>>(c->d_hdr, c->d_fe, B, P);
        else k_silk_core<32><<<(chains + 31) / 32, 32, silk_core_smem(32), ss>>>(c->d_hdr, c->d_fe, B, P);
        k_silk_unmix<<<wide, 256, 0, ss>>>(B, P);
        if (lanes == 8) k_silk_up2<8><<<(chains + 7) / 8, 32, 0, ss>>>(B, P);
        else if (lanes == 16) k_silk_up2<16><<<(chains + 15) / 16, 32, 0, ss>>>(B, P);
        else k_silk_up2<32><<<(chains + 31) / 32, 32, 0, ss>>>(B, P);
        cudaEventRecord(c->join_ev[chunk], ss);
    }
    // coded leaves: sort by cost class, expand one per thread
    LeafSort* ls = c->d_leafsort + chunk;
    uint32_t* order = c->d_leaf_order + (size_t)first * in->n_frames * kLeafSlotsPerFrame;
    cudaMemsetAsync(ls, 0, sizeof(LeafSort), s);
    k_leaf_count<<<(nfr + KL_THREADS - 1) / KL_THREADS, KL_THREADS, 0, s>>>(c->d_hdr, c->d_ops, ls, P);
    k_leaf_place<<<(nfr + KL_THREADS - 1) / KL_THREADS, KL_THREADS, 0, s>>>(c->d_hdr, c->d_ops, ls, order, nfr * c->leaf_slots, P);
    k_leaf_expand<<<c->n_sm * KL_GRID, KL_THREADS, 0, s>>>(c->d_hdr, c->d_ops, ls, order, c->d_X, P);
    if (timed) cudaEventRecord(c->ev[3], s);
    k_bands<<<(nfr + KB_WARPS - 1) / KB_WARPS, 32 * KB_WARPS, 0, s>>>(c->d_hdr, c->d_ops, c->d_fe, c->d_X, ls, P);
    if (timed) cudaEventRecord(c->ev[4], s);
    k_imdct<<<(nfr * P.CC + KI_WARPS - 1) / KI_WARPS, 32 * KI_WARPS, 0, s>>>(c->d_hdr, c->d_fe, c->d_X, c->d_raw, P);
    if (timed) cudaEventRecord(c->ev[5], s);
    k_postfilter<<<(n_sel * P.CC + PF_WARPS - 1) / PF_WARPS, 32 * PF_WARPS, 0, s>>>(c->d_hdr, c->d_fe, c->d_raw, c->d_syn, c->d_state, P);
    if (timed) cudaEventRecord(c->ev[6], s);
    k_deemph<<<(n_sel * P.CC + 31) / 32, 32, 0, s>>>(c->d_fe, c->d_hdr, c->d_syn, d_pcm, d_status, c->d_state, P);
    if (timed) cudaEventRecord(c->ev[7], s);
    if (silk) {  // "silk" stage time: what is left of the SILK layer once the CELT kernels are done, plus the final sum
        cudaStreamWaitEvent(s, c->join_ev[chunk], 0);
        k_silk_fir<<<wide, 32 * FIR_WARPS, 0, s>>>(B, c->d_fe, d_pcm, P);
        c->launches += 6;
    }
    if (timed) cudaEventRecord(c->ev[8], s);
    c->launches += 9;  // entropy, sequence, 3 leaf kernels, bands, imdct, postfilter, deemph
    CU_TRY(c, cudaGetLastError());
    return 0;
}

int dm_decode_device(dm_ctx* c, const dm_batch* in, dm_out* out, void* cuda_stream) {
    int r = check_batch(c, in);
    if (r) return r;
    if (!out || !out->pcm || !out->status) return OPUSB_BAD_ARG;
    CU_TRY(c, cudaSetDevice(c->device));
    return launch_pipeline(c, in, in->bytes, (const FrameDesc*)in->frames, out->pcm, out->status, 0, c->n_streams,
                           (cudaStream_t)cuda_stream, true);
}

// Blocking call: per stream range (chunk) frames H2D -> kernels -> PCM/status D2H, each chunk on its own CUDA stream so
// that the read-back of one range overlaps the kernels of the next.
int decode_many(dm_ctx* c, const dm_batch* in, dm_out* out) {
    int r = check_batch(c, in);
    if (r) return r;
    if (!out || !out->pcm || !out->status) return OPUSB_BAD_ARG;
    CU_TRY(c, cudaSetDevice(c->device));
    if (c->next_ticket) CU_TRY(c, cudaDeviceSynchronize());  // asynchronous calls in flight finish first
    r = reserve_bytes(c, 0, in->n_bytes);
    if (r) return r;
    const int nf = in->n_frames, CC = c->channels;
    const int fs = in->mode == DM_MODE_CELT_RAW ? in->frame_size : 960;
    // packets are small: one copy on stream 0, the other chunk streams wait for it
    CU_TRY(c, cudaMemcpyAsync(c->d_bytes[0], in->bytes, in->n_bytes, cudaMemcpyHostToDevice, c->streams[0]));
    CU_TRY(c, cudaEventRecord(c->bytes_ready, c->streams[0]));
    const int n_chunks = c->n_streams >= 64 * kMaxChunks ? kMaxChunks : 1;
    for (int k = 0; k < n_chunks; k++) {
        const int first = (int)((long long)c->n_streams * k / n_chunks), last = (int)((long long)c->n_streams * (k + 1) / n_chunks);
        const int n_sel = last - first;
        cudaStream_t s = c->streams[k];
        const size_t f0 = (size_t)first * nf, fn = (size_t)n_sel * nf;
        if (k) CU_TRY(c, cudaStreamWaitEvent(s, c->bytes_ready, 0));
        CU_TRY(c, cudaMemcpyAsync(c->d_frames + f0, in->frames + f0, fn * sizeof(dm_frame), cudaMemcpyHostToDevice, s));
        r = launch_pipeline(c, in, c->d_bytes[0], c->d_frames, c->d_pcm, c->d_status, first, n_sel, s, false, nullptr, k);
        if (r) return r;
        CU_TRY(c, cudaMemcpyAsync(out->pcm + f0 * fs * CC, c->d_pcm + f0 * fs * CC, fn * fs * CC * sizeof(int16_t), cudaMemcpyDeviceToHost, s));
        CU_TRY(c, cudaMemcpyAsync(out->status + f0, c->d_status + f0, fn * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    }
    for (int k = 0; k < n_chunks; k++) CU_TRY(c, cudaStreamSynchronize(c->streams[k]));
    return 0;
}

// Pipelined call: upload, kernels and read-back each have their own CUDA stream and every buffer that crosses a
// stage is double buffered, so with two calls in flight the GPU decodes call n+1 while call n's PCM crosses PCIe and
// call n+2's packets arrive.  The kernels run over ALL streams at once: the stages that walk a stream's frames in
// order (sequence, post-filter, de-emphasis, SILK) take as long for 1024 streams as for 4096.
int decode_many_async(dm_ctx* c, const dm_batch* in, dm_out* out, long long* ticket_out) {
    int r = check_batch(c, in);
    if (r) return r;
    if (!out || !out->pcm || !out->status) return OPUSB_BAD_ARG;
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t nfr_cap = (size_t)c->n_streams * c->max_frames;
    if (!c->d_pcm2) {
        if ((r = dev_alloc(c, &c->d_pcm2, nfr_cap * c->channels * kMaxFrame))) return r;
        if ((r = dev_alloc(c, &c->d_status2, nfr_cap))) return r;
    }
    const long long ticket = c->next_ticket;
    const int slot = (int)(ticket % kRing), bi = (int)(ticket & 1);
    if (ticket >= kRing) CU_TRY(c, cudaEventSynchronize(c->done[slot][0]));  // the slot's previous call is over
    if (in->n_bytes > c->d_bytes_cap[bi]) CU_TRY(c, cudaDeviceSynchronize());
    r = reserve_bytes(c, bi, in->n_bytes);
    if (r) return r;
    const int nf = in->n_frames, CC = c->channels;
    const int fs = in->mode == DM_MODE_CELT_RAW ? in->frame_size : 960;
    const size_t fn = (size_t)c->n_streams * nf;
    cudaStream_t up = c->streams[2], run = c->streams[0], down = c->streams[1];
    int16_t* d_pcm = bi ? c->d_pcm2 : c->d_pcm;
    int32_t* d_status = bi ? c->d_status2 : c->d_status;
    FrameDesc* d_frames = c->d_frames;
    // upload: this byte buffer was last parsed two calls ago; the frame descriptors by the previous call
    if (ticket >= 2) CU_TRY(c, cudaStreamWaitEvent(up, c->parsed[(ticket - 2) % kRing][0], 0));
    CU_TRY(c, cudaMemcpyAsync(c->d_bytes[bi], in->bytes, in->n_bytes, cudaMemcpyHostToDevice, up));
    if (ticket >= 1) CU_TRY(c, cudaStreamWaitEvent(up, c->kernels_done[(ticket - 1) % kRing], 0));
    CU_TRY(c, cudaMemcpyAsync(d_frames, in->frames, fn * sizeof(dm_frame), cudaMemcpyHostToDevice, up));
    CU_TRY(c, cudaEventRecord(c->uploaded[slot], up));
    // kernels: the output buffer of this parity was last read back two calls ago
    CU_TRY(c, cudaStreamWaitEvent(run, c->uploaded[slot], 0));
    if (ticket >= 2) CU_TRY(c, cudaStreamWaitEvent(run, c->done[(ticket - 2) % kRing][0], 0));
    r = launch_pipeline(c, in, c->d_bytes[bi], d_frames, d_pcm, d_status, 0, c->n_streams, run, true, c->parsed[slot][0]);
    if (r) return r;
    CU_TRY(c, cudaEventRecord(c->kernels_done[slot], run));
    // read-back
    CU_TRY(c, cudaStreamWaitEvent(down, c->kernels_done[slot], 0));
    CU_TRY(c, cudaMemcpyAsync(out->pcm, d_pcm, fn * fs * CC * sizeof(int16_t), cudaMemcpyDeviceToHost, down));
    CU_TRY(c, cudaMemcpyAsync(out->status, d_status, fn * sizeof(int32_t), cudaMemcpyDeviceToHost, down));
    CU_TRY(c, cudaEventRecord(c->done[slot][0], down));
    c->next_ticket = ticket + 1;
    if (ticket_out) *ticket_out = ticket;
    return 0;
}

int dm_wait(dm_ctx* c, long long ticket) {
    if (!c) return OPUSB_BAD_ARG;
    if (ticket < 0 || ticket >= c->next_ticket) return OPUSB_BAD_ARG;
    if (ticket + kRing <= c->next_ticket - 1) return 0;  // long finished: its ring slot was recycled after a synchronise
    CU_TRY(c, cudaEventSynchronize(c->done[(int)(ticket % kRing)][0]));
    return 0;
}

// The bare read-back of decode_many_async, nothing else: `reps` cudaMemcpyAsync D2H of `bytes` from the context's PCM
// buffer to `host` on the read-back stream, timed with CUDA events.  What the PCIe link and the host memory system allow.
int dm_probe_d2h(dm_ctx* c, void* host, size_t bytes, int reps, float* ms_total) {
    if (!c || !host || reps < 1 || !ms_total) return OPUSB_BAD_ARG;
    const size_t cap = (size_t)c->n_streams * c->max_frames * c->channels * kMaxFrame * sizeof(int16_t);
    if (bytes > cap) return DM_ERR_CAPACITY;
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaDeviceSynchronize());
    cudaStream_t down = c->streams[1];
    CU_TRY(c, cudaMemcpyAsync(host, c->d_pcm, bytes, cudaMemcpyDeviceToHost, down));  // warm-up
    CU_TRY(c, cudaEventRecord(c->ev[0], down));
    for (int i = 0; i < reps; i++) CU_TRY(c, cudaMemcpyAsync(host, c->d_pcm, bytes, cudaMemcpyDeviceToHost, down));
    CU_TRY(c, cudaEventRecord(c->ev[1], down));
    CU_TRY(c, cudaStreamSynchronize(down));
    CU_TRY(c, cudaEventElapsedTime(ms_total, c->ev[0], c->ev[1]));
    return 0;
}

void* dm_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void dm_host_free(void* p) { if (p) cudaFreeHost(p); }

long long dm_launch_count(const dm_ctx* c) { return c ? c->launches : 0; }
size_t dm_device_bytes(const dm_ctx* c) { return c ? c->dev_bytes : 0; }
const char* dm_stage_name(int i) { return (i >= 0 && i < kStages) ? kStageNames[i] : ""; }
int dm_stage_times(dm_ctx* c, float* ms, int cap) {
    if (!c || !ms) return 0;
    int n = cap < kStages ? cap : kStages;
    for (int i = 0; i < n; i++) {
        ms[i] = 0.f;
        if (cudaEventElapsedTime(&ms[i], c->ev[i], c->ev[i + 1]) != cudaSuccess) ms[i] = -1.f;
    }
    return n;
}
int dm_debug_leaf_overflow(dm_ctx* c) {
    if (!c) return OPUSB_BAD_ARG;
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaDeviceSynchronize());
    LeafSort ls[kMaxChunks];
    CU_TRY(c, cudaMemcpy(ls, c->d_leafsort, sizeof(ls), cudaMemcpyDeviceToHost));
    int any = 0;
    for (int k = 0; k < kMaxChunks; k++) any |= ls[k].overflow;
    return any;
}
size_t dm_stream_state_size(void) { return sizeof(CeltStreamState); }
int dm_get_stream_state(dm_ctx* c, int stream, void* dst, size_t cap) {
    if (!c || !dst || stream < 0 || stream >= c->n_streams || cap < sizeof(CeltStreamState)) return OPUSB_BAD_ARG;
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaDeviceSynchronize());
    CU_TRY(c, cudaMemcpy(dst, c->d_state + stream, sizeof(CeltStreamState), cudaMemcpyDeviceToHost));
    return 0;
}

}  // extern "C"
