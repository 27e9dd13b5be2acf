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

enum { kStages = 8, kMaxChunks = 4, kRing = 4, kTraceCalls = 16, kMaxPieces = 8 };
static const char* kStageNames[kStages + 1] = {"entropy", "sequence", "leaves", "bands", "imdct", "postfilter", "deemph", "silk", "pipeline"};

// The per-frame intermediates that cross from the head of a call (parse, spectral stage) into its tail (post-filter,
// de-emphasis) and its SILK stage.  There are two sets, used by consecutive calls in turn, so that the head of call k + 1
// can run while the tail of call k still reads what the head of call k left.
struct FrameSet {
    CeltFrameHdr* hdr;
    FrameEnergy* fe;
    int32_t* raw;
    int32_t* syn;            // post-filter -> de-emphasis
    SilkFrame* silk;
    SilkCtl* sctl;           // SILK stage hand-offs (silk_kernels.cuh)
    SilkPost* spost;
    int16_t* sxq;
    int16_t* slr;
    int16_t* sup;
    FixEntry* sfix;          // LPC filters that failed their first stability check (two per SILK frame-channel at most)
    int* any_silk;           // [kMaxChunks] a frame of that stream range of the call has a SILK layer
    int* fix_count;          // [kMaxChunks] filters queued for k_silk_fix
};

struct dm_ctx {
    int n_streams, max_frames, channels, device, n_sm;
    int leaf_slots;  // capacity of the sorted leaf list per frame (OPUSB_LEAF_SLOTS lowers it: tests of the overflow path)
    int tail_mode;   // pipelined calls: 0 = one stream (default: overlapping the tail of call k with the head of call k + 1 was measured
                     // and wins nothing, profiles/r02_call_overlap_trace.txt), 2 = post-filter + de-emphasis on a tail stream beside the next
                     // call's head (two alternating frame sets), 1 = de-emphasis only (OPUSB_TAIL, to repeat the measurement)
    int silk_lanes;  // chains per warp of the SILK lane-per-chain kernels; 0 = silk_chain_lanes() (OPUSB_SILK_LANES overrides, for tuning)
    // per-stream state
    CeltStreamState* d_state;
    SilkStreamState* d_sstate;
    FrameSet set[2];         // set[1] is allocated by the first pipelined call
    // per-frame intermediates of the head alone ...
    BandOp* d_ops;
    int16_t* d_X;
    uint32_t* d_parse_order; // work order of the CELT-only parser kernel (k_classify), one entry per frame of the context
    int* d_parse_count;      // [kMaxChunks][2]: frames per class of a launch
    int sort_frames;         // 1: the parser takes its frames sorted by kind (OPUSB_SORT_FRAMES=1; off by default)
    LeafSort* d_leafsort;    // [kMaxChunks]: counting sort of the coded leaves of a launch by cost class
    uint32_t* d_leaf_order;  // the sorted leaf list, kLeafSlotsPerFrame entries per frame of the context
    // staging for the host-buffer path: double buffered so that call n+1 can upload while call n still parses
    uint8_t* d_bytes[2];
    size_t d_bytes_cap[2];
    FrameDesc* d_frames[2];
    int16_t* d_pcm;
    int32_t* d_status;
    int16_t* d_pcm2;         // second output buffer, allocated on the first decode_many_async
    int32_t* d_status2;
    cudaStream_t streams[kMaxChunks];     // blocking decode_many: one per stream range; pipelined calls: [0] head, [1] read-back, [2] upload
    cudaStream_t tail;                    // pipelined calls: post-filter, de-emphasis, final SILK sum
    cudaStream_t side[kMaxChunks];        // per stream range: the SILK kernels of a call run beside its CELT kernels
    cudaStream_t de_s[kMaxChunks];        // per stream range: the de-emphasis of one run of frames beside the post-filter of the next
    cudaEvent_t pf_ev[kMaxChunks][kMaxPieces], de_ev[kMaxChunks];
    int tail_pieces;                      // runs of frames the tail of a call is cut into (OPUSB_TAIL_PIECES; 1 = no overlap)
    DeemphSeg* d_deseg;                   // [stream][channel][DE_MAX_SEGS]: what the segments of a de-emphasis chain hand to k_deemph_fix
    int32_t* d_demem;                     // [stream][channel][frame]: memory of a segment's run at the end of every frame (k_deemph_fix stops there)
    int* d_deemph_redone;                 // segments k_deemph_fix had to do again, since dm_create (dm_debug_deemph_redone)
    int deemph_segs, deemph_warm;         // OPUSB_DEEMPH_SEGS (1 = the serial kernel alone), OPUSB_DEEMPH_WARM (warm-up tiles of 32 samples)
    cudaEvent_t fork_ev[kMaxChunks], join_ev[kMaxChunks];
    cudaEvent_t bytes_ready;
    // pipelined calls (dm_decode_device, dm_decode_device_async, decode_many_async), a ring over the call number
    long long calls;
    cudaEvent_t ready_ev[kRing], seq_ev[kRing], head_ev[kRing], x_ev[kRing], done_ev[kRing];
    // decode_many_async: "packets parsed", "results in host memory", "packets uploaded"
    cudaEvent_t parsed[kRing], host_done[kRing], uploaded[kRing];
    long long next_ticket;
    cudaEvent_t ev[kStages + 1];
    cudaEvent_t probe_ev[2];
    cudaEvent_t* trace;      // OPUSB_TRACE=1: kTraceCalls x (kStages + 1) timing events, one row per pipelined call (dm_debug_trace)
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

static int alloc_set(dm_ctx* c, FrameSet& F) {
    if (F.hdr) return 0;
    const size_t nfr = (size_t)c->n_streams * c->max_frames;
    int r;
    if ((r = dev_alloc(c, &F.hdr, nfr))) return r;
    if ((r = dev_alloc(c, &F.fe, nfr))) return r;
    if ((r = dev_alloc(c, &F.raw, nfr * c->channels * kMaxFrame))) return r;
    if ((r = dev_alloc(c, &F.syn, nfr * c->channels * kMaxFrame))) return r;
    if ((r = dev_alloc(c, &F.silk, nfr))) return r;
    if ((r = dev_alloc(c, &F.sctl, nfr * 2))) return r;
    if ((r = dev_alloc(c, &F.spost, nfr))) return r;
    if ((r = dev_alloc(c, &F.sxq, nfr * 2 * kXqRow))) return r;
    if ((r = dev_alloc(c, &F.slr, nfr * 2 * kXqRow))) return r;
    if ((r = dev_alloc(c, &F.sup, nfr * 2 * kUpRow))) return r;
    if ((r = dev_alloc(c, &F.sfix, nfr * 4))) return r;
    if ((r = dev_alloc(c, &F.any_silk, kMaxChunks))) return r;
    if ((r = dev_alloc(c, &F.fix_count, kMaxChunks))) return r;
    CU_TRY(c, cudaMemset(F.spost, 0, nfr * sizeof(SilkPost)));  // has_out of frames no SILK kernel has visited yet
    CU_TRY(c, cudaMemset(F.any_silk, 0, kMaxChunks * sizeof(int)));
    return 0;
}
static void free_set(FrameSet& F) {
    cudaFree(F.hdr); cudaFree(F.fe); cudaFree(F.raw); cudaFree(F.syn); cudaFree(F.silk); cudaFree(F.sctl); cudaFree(F.spost); cudaFree(F.sxq); cudaFree(F.slr);
    cudaFree(F.sup); cudaFree(F.sfix); cudaFree(F.any_silk); cudaFree(F.fix_count);
    memset(&F, 0, sizeof(F));
}

// Everything dm_create allocates; on failure the caller tears the half-built context down with dm_destroy.
static int create_impl(dm_ctx* c, int n_streams, int max_frames, int channels, int device) {
    CU_TRY(c, cudaSetDevice(device));
    CU_TRY(c, cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, device));
    c->leaf_slots = kLeafSlotsPerFrame;  // (the context was just zeroed)
    if (const char* e = getenv("OPUSB_LEAF_SLOTS")) { const int v = atoi(e); if (v >= 0 && v <= kLeafSlotsPerFrame) c->leaf_slots = v; }
    c->tail_mode = 0;
    c->tail_pieces = 1;  // measured: 2 / 4 / 8 runs make the tail 0.1 - 0.7 ms slower (the two kernels slow each other down side by side)
    if (const char* e = getenv("OPUSB_TAIL_PIECES")) { const int v = atoi(e); if (v >= 1 && v <= kMaxPieces) c->tail_pieces = v; }
    if (const char* e = getenv("OPUSB_TAIL")) { const int v = atoi(e); if (v >= 0 && v <= 2) c->tail_mode = v; }
    c->deemph_segs = 2;  // measured: 0.81 ms (1) -> 0.47 (2) -> 1.3 (3) / 0.9 (4) / 0.8 (8); at 2 the kernel moves 5.1 TB/s
    c->deemph_warm = DE_WARM / DE_TW;
    if (const char* e = getenv("OPUSB_DEEMPH_SEGS")) { const int v = atoi(e); if (v >= 1 && v <= DE_MAX_SEGS) c->deemph_segs = v; }
    if (const char* e = getenv("OPUSB_DEEMPH_WARM")) { const int v = atoi(e); if (v >= 1 && v <= DE_WARM / DE_TW) c->deemph_warm = v; }  // (tests: a short warm-up forces the redo path)
    if (const char* e = getenv("OPUSB_SILK_LANES")) { const int v = atoi(e); if (v == 8 || v == 16 || v == 32) c->silk_lanes = v; }
    // The entropy kernel keeps its per-thread allocation arrays and the explicit partition stack in local memory.  The
    // stack limit is a DEVICE-wide setting shared with every other CUDA user of the process (torch): it is only ever
    // raised, and only as far as this library's largest frame needs (no recursion anywhere, so ptxas knows the size).
    {
        cudaFuncAttributes fa;
        CU_TRY(c, cudaFuncGetAttributes(&fa, k_entropy<true>));
        size_t cur = 0;
        CU_TRY(c, cudaDeviceGetLimit(&cur, cudaLimitStackSize));
        const size_t need = (fa.localSizeBytes + 255) / 256 * 256;
        if (need > cur) CU_TRY(c, cudaDeviceSetLimit(cudaLimitStackSize, need));
    }
    const size_t nfr = (size_t)n_streams * max_frames;
    int r;
    if ((r = dev_alloc(c, &c->d_state, (size_t)n_streams))) return r;
    if ((r = dev_alloc(c, &c->d_sstate, (size_t)n_streams))) return r;
    if ((r = alloc_set(c, c->set[0]))) return r;
    CU_TRY(c, cudaFuncSetAttribute(k_silk_core<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, silk_core_smem(32)));
    CU_TRY(c, cudaFuncSetAttribute(k_imdct_long, cudaFuncAttributeMaxDynamicSharedMemorySize, kImdctLongSmem));
    if ((r = dev_alloc(c, &c->d_ops, nfr * kMaxOps))) return r;
    if ((r = dev_alloc(c, &c->d_X, nfr * 2 * kMaxFrame))) return r;
    if ((r = dev_alloc(c, &c->d_leafsort, (size_t)kMaxChunks))) return r;
    if ((r = dev_alloc(c, &c->d_deseg, (size_t)n_streams * 2 * DE_MAX_SEGS))) return r;
    if ((r = dev_alloc(c, &c->d_demem, nfr * 2))) return r;
    if ((r = dev_alloc(c, &c->d_deemph_redone, (size_t)1))) return r;
    CU_TRY(c, cudaMemset(c->d_deemph_redone, 0, sizeof(int)));
    if ((r = dev_alloc(c, &c->d_parse_order, nfr))) return r;
    if ((r = dev_alloc(c, &c->d_parse_count, (size_t)kMaxChunks * 2))) return r;
    c->sort_frames = 0;  // measured: 3.85 ms without, 3.91 ms with (the classification pass costs what the fuller warps win)
    if (const char* e = getenv("OPUSB_SORT_FRAMES")) c->sort_frames = atoi(e) != 0;
    if ((r = dev_alloc(c, &c->d_leaf_order, nfr * kLeafSlotsPerFrame))) return r;
    if ((r = dev_alloc(c, &c->d_frames[0], nfr))) return r;
    if ((r = dev_alloc(c, &c->d_pcm, nfr * channels * kMaxFrame))) return r;
    if ((r = dev_alloc(c, &c->d_status, nfr))) return r;
    // the streams that carry few, long-running warps (the serial tail, the SILK recurrences) get the first free slots
    int prio_lo = 0, prio_hi = 0;
    CU_TRY(c, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    for (int i = 0; i < kMaxChunks; i++) CU_TRY(c, cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking));
    CU_TRY(c, cudaStreamCreateWithPriority(&c->tail, cudaStreamNonBlocking, prio_hi));
    for (int i = 0; i < kMaxChunks; i++) {
        CU_TRY(c, cudaStreamCreateWithPriority(&c->side[i], cudaStreamNonBlocking, prio_hi));
        CU_TRY(c, cudaStreamCreateWithPriority(&c->de_s[i], cudaStreamNonBlocking, prio_hi));
        CU_TRY(c, cudaEventCreateWithFlags(&c->de_ev[i], cudaEventDisableTiming));
        for (int g = 0; g < kMaxPieces; g++) CU_TRY(c, cudaEventCreateWithFlags(&c->pf_ev[i][g], cudaEventDisableTiming));
        CU_TRY(c, cudaEventCreateWithFlags(&c->fork_ev[i], cudaEventDisableTiming));
        CU_TRY(c, cudaEventCreateWithFlags(&c->join_ev[i], cudaEventDisableTiming));
    }
    CU_TRY(c, cudaEventCreateWithFlags(&c->bytes_ready, cudaEventDisableTiming));
    for (int r = 0; r < kRing; r++) {
        cudaEvent_t* evs[] = {&c->ready_ev[r], &c->seq_ev[r], &c->head_ev[r], &c->x_ev[r], &c->done_ev[r], &c->parsed[r], &c->host_done[r], &c->uploaded[r]};
        for (cudaEvent_t* e : evs) CU_TRY(c, cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    }
    for (int i = 0; i <= kStages; i++) CU_TRY(c, cudaEventCreate(&c->ev[i]));
    for (int i = 0; i < 2; i++) CU_TRY(c, cudaEventCreate(&c->probe_ev[i]));
    if (const char* e = getenv("OPUSB_TRACE")) {
        if (atoi(e) > 0) {
            c->trace = (cudaEvent_t*)calloc((size_t)kTraceCalls * (kStages + 1), sizeof(cudaEvent_t));
            if (!c->trace) return DM_ERR_NOMEM;
            for (int i = 0; i < kTraceCalls * (kStages + 1); i++) CU_TRY(c, cudaEventCreate(&c->trace[i]));
        }
    }
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
    cudaFree(c->d_state); cudaFree(c->d_sstate);
    free_set(c->set[0]); free_set(c->set[1]);
    cudaFree(c->d_ops); cudaFree(c->d_X); cudaFree(c->d_deseg); cudaFree(c->d_demem); cudaFree(c->d_deemph_redone); cudaFree(c->d_leafsort); cudaFree(c->d_leaf_order); cudaFree(c->d_parse_order); cudaFree(c->d_parse_count);
    cudaFree(c->d_bytes[0]); cudaFree(c->d_bytes[1]); cudaFree(c->d_frames[0]); cudaFree(c->d_frames[1]);
    cudaFree(c->d_pcm); cudaFree(c->d_pcm2); cudaFree(c->d_status); cudaFree(c->d_status2);
    for (int i = 0; i < kMaxChunks; i++) if (c->streams[i]) cudaStreamDestroy(c->streams[i]);
    if (c->tail) cudaStreamDestroy(c->tail);
    for (int i = 0; i < kMaxChunks; i++) {
        if (c->side[i]) cudaStreamDestroy(c->side[i]);
        if (c->de_s[i]) cudaStreamDestroy(c->de_s[i]);
        if (c->de_ev[i]) cudaEventDestroy(c->de_ev[i]);
        for (int g = 0; g < kMaxPieces; g++) if (c->pf_ev[i][g]) cudaEventDestroy(c->pf_ev[i][g]);
        if (c->fork_ev[i]) cudaEventDestroy(c->fork_ev[i]);
        if (c->join_ev[i]) cudaEventDestroy(c->join_ev[i]);
    }
    if (c->bytes_ready) cudaEventDestroy(c->bytes_ready);
    for (int r = 0; r < kRing; r++) {
        cudaEvent_t evs[] = {c->ready_ev[r], c->seq_ev[r], c->head_ev[r], c->x_ev[r], c->done_ev[r], c->parsed[r], c->host_done[r], c->uploaded[r]};
        for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
    }
    for (int i = 0; i <= kStages; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < 2; i++) if (c->probe_ev[i]) cudaEventDestroy(c->probe_ev[i]);
    if (c->trace) {
        for (int i = 0; i < kTraceCalls * (kStages + 1); i++) if (c->trace[i]) cudaEventDestroy(c->trace[i]);
        free(c->trace);
    }
    delete c;
}

int dm_reset(dm_ctx* c, const int32_t* ids, int n) {
    if (!c) return OPUSB_BAD_ARG;
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaDeviceSynchronize());  // calls in flight finish first
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

// Where the kernels of one call go.  Blocking calls put everything on one stream (H == T); pipelined calls give the head
// (parse + spectral stage) and the tail (post-filter, de-emphasis, final SILK sum) a stream each, so that the tail of
// call k -- a handful of long-running warps that leave the SMs almost empty -- runs beside the head of call k + 1.
struct CallPlan {
    cudaStream_t H, T, X;     // head, tail, SILK side stream
    FrameSet* F;
    int chunk;
    cudaEvent_t seq_ev;       // recorded on H behind k_sequence: the SILK stage may start
    cudaEvent_t head_ev;      // recorded on H behind k_imdct: the tail may start (unused when H == T)
    cudaEvent_t x_ev;         // recorded on X behind the SILK stage: the final sum may run
    cudaEvent_t parsed;       // optional, recorded on H once the packet bytes and frame descriptors have been read
    bool timed;
};

// Enqueue the kernels of one call for streams [first, first+n_sel).  Pointers are device pointers.
static int launch_pipeline(dm_ctx* c, const dm_batch* in, const uint8_t* d_bytes, const FrameDesc* d_frames, int16_t* d_pcm,
                           int32_t* d_status, int first, int n_sel, const CallPlan& pl) {
    DevParams P;
    P.n_streams = c->n_streams; P.n_frames = in->n_frames; P.mode = in->mode;
    P.frame_size = in->mode == DM_MODE_CELT_RAW ? in->frame_size : 960;
    P.coded_channels = in->coded_channels; P.start_band = in->start_band; P.CC = c->channels;
    P.first_stream = first; P.n_sel = n_sel;
    P.f0 = 0; P.nf = in->n_frames;
    const FrameSet& F = *pl.F;
    const cudaStream_t H = pl.H, T = pl.T, X = pl.X;
    const bool timed = pl.timed;
    const int chunk = pl.chunk;
    cudaEvent_t* tr = (c->trace && timed) ? c->trace + (size_t)(c->calls % kTraceCalls) * (kStages + 1) : nullptr;
    auto mark = [&](int i, cudaStream_t st) {
        cudaEventRecord(c->ev[i], st);
        if (tr) cudaEventRecord(tr[i], st);
    };
    const int nfr = n_sel * in->n_frames;
    const int wb = 32 * KW_WARPS;
    if (timed) mark(0, H);
    int* any_silk = F.any_silk + chunk;  // per stream range: ranges of one call run concurrently on their own streams
    cudaMemsetAsync(any_silk, 0, sizeof(int), H);
    const uint32_t* order = nullptr;
    if (c->sort_frames && nfr >= 4096) {  // (a small launch has too few transient frames to fill warps with)
        uint32_t* ord = c->d_parse_order + (size_t)first * in->n_frames;
        cudaMemsetAsync(c->d_parse_count + 2 * chunk, 0, 2 * sizeof(int), H);
        k_classify<<<(nfr + 255) / 256, 256, 0, H>>>(d_bytes, d_frames, P, ord, c->d_parse_count + 2 * chunk);
        c->launches++;
        order = ord;
    }
    k_entropy<false><<<(nfr + K1_THREADS - 1) / K1_THREADS, K1_THREADS, 0, H>>>(d_bytes, d_frames, P, F.hdr, c->d_ops, F.silk, c->d_state, any_silk, order);
    if (in->mode == DM_MODE_OPUS) {  // the frames with a SILK layer
        k_entropy<true><<<(nfr + K1_THREADS - 1) / K1_THREADS, K1_THREADS, 0, H>>>(d_bytes, d_frames, P, F.hdr, c->d_ops, F.silk, c->d_state, any_silk, nullptr);
        c->launches++;
    }
    if (timed) mark(1, H);
    k_sequence<<<(n_sel + KW_WARPS - 1) / KW_WARPS, wb, 0, H>>>(F.hdr, d_frames, F.fe, c->d_state, P);
    if (timed) mark(2, H);
    if (pl.parsed) cudaEventRecord(pl.parsed, H);
    // The SILK layer of a call needs nothing from its CELT layer until the final sum, and its two long kernels (LTP/LPC
    // synthesis, up-sampler) are single-wave recurrences that leave most issue slots idle: they run on a side stream
    // beside the issue-bound CELT kernels and meet them again in front of k_silk_fir.  Every one of these returns at
    // once unless k_entropy saw a frame with a SILK layer.
    const bool silk = in->mode == DM_MODE_OPUS;
    SilkBufs B = {F.silk, F.sctl, F.spost, F.sxq, F.slr, F.sup, c->d_sstate, any_silk};
    const int wide = c->n_sm * 8;  // grid-stride kernels: a fixed grid, so an idle call costs a few empty launches
    if (silk) {
        cudaEventRecord(pl.seq_ev, H);
        cudaStreamWaitEvent(X, pl.seq_ev, 0);
        const int chains = n_sel * P.CC;
        // the stream range's own slice of the fix queue (at most two filters per frame and channel) and its own counter
        FixEntry* fq = F.sfix + (size_t)first * in->n_frames * 4;
        cudaMemsetAsync(F.fix_count + chunk, 0, sizeof(int), X);
        k_silk_params<<<c->n_sm * 12, 128, 0, X>>>(F.hdr, F.fe, B, P, fq, F.fix_count + chunk);
        k_silk_fix<<<c->n_sm * 8, 128, 0, X>>>(any_silk, fq, F.fix_count + chunk);
        const int lanes = c->silk_lanes > 0 ? c->silk_lanes : silk_chain_lanes(chains, c->n_sm);
        if (lanes == 8) k_silk_core<8><<<(chains + 7) / 8, 32, silk_core_smem(8), X>>>(F.hdr, F.fe, B, P);
        else if (lanes == 16) k_silk_core<16><<<(chains + 15) / 16, 32, silk_core_smem(16), X>>>(F.hdr, F.fe, B, P);
        else k_silk_core<32><<<(chains + 31) / 32, 32, silk_core_smem(32), X>>>(F.hdr, F.fe, B, P);
        k_silk_unmix<<<wide, 256, 0, X>>>(B, P);
        if (lanes == 8) k_silk_up2<8><<<(chains + 7) / 8, 32, 0, X>>>(B, P);
        else if (lanes == 16) k_silk_up2<16><<<(chains + 15) / 16, 32, 0, X>>>(B, P);
        else k_silk_up2<32><<<(chains + 31) / 32, 32, 0, X>>>(B, P);
        cudaEventRecord(pl.x_ev, X);
    }
    // coded leaves: sort by cost class, expand one per thread
    LeafSort* ls = c->d_leafsort + chunk;
    uint32_t* lorder = c->d_leaf_order + (size_t)first * in->n_frames * kLeafSlotsPerFrame;
    cudaMemsetAsync(ls, 0, sizeof(LeafSort), H);
    k_leaf_count<<<(nfr + KL_THREADS - 1) / KL_THREADS, KL_THREADS, 0, H>>>(F.hdr, c->d_ops, ls, P);
    k_leaf_place<<<(nfr + KL_THREADS - 1) / KL_THREADS, KL_THREADS, 0, H>>>(F.hdr, c->d_ops, ls, lorder, nfr * c->leaf_slots, P);
    k_leaf_expand<<<c->n_sm * KL_GRID, KL_THREADS, 0, H>>>(F.hdr, c->d_ops, ls, lorder, c->d_X, P);
    if (timed) mark(3, H);
    k_bands<<<(nfr + KB_WARPS - 1) / KB_WARPS, 32 * KB_WARPS, 0, H>>>(F.hdr, c->d_ops, F.fe, c->d_X, ls, P);
    if (timed) mark(4, H);
    if (P.frame_size == 960) {  // long blocks: persistent blocks, a warp per frame (both channels), grid-stride over the frames
        const int want = (nfr + KI_WARPS - 1) / KI_WARPS, cap = c->n_sm * KI_GRID;
        k_imdct_long<<<want < cap ? want : cap, 32 * KI_WARPS, kImdctLongSmem, H>>>(F.hdr, F.fe, c->d_X, F.raw, P);
        c->launches++;
    }
    k_imdct<<<(nfr * P.CC + KI_WARPS - 1) / KI_WARPS, 32 * KI_WARPS, 0, H>>>(F.hdr, F.fe, c->d_X, F.raw, P);  // everything else
    if (timed) mark(5, H);
    const cudaStream_t Tp = (T != H && c->tail_mode == 1) ? H : T;  // the post-filter's stream
    if (T != H && Tp == T) {
        cudaEventRecord(pl.head_ev, H);
        cudaStreamWaitEvent(T, pl.head_ev, 0);
    }
    // The two serial kernels of the tail cost per frame in order, not per stream: the post-filter (16 warps per SM, bound by
    // its shared-memory rings) and the de-emphasis (one warp per 32 chains: 256 warps on the whole GPU).  The frames of the
    // call are cut into runs, and the de-emphasis of one run goes to a stream of its own beside the post-filter of the next.
    {
        int pieces = c->tail_pieces;
        if (pieces > in->n_frames / 4) pieces = in->n_frames / 4;
        if (pieces < 1) pieces = 1;
        const int per = (in->n_frames + pieces - 1) / pieces;
        const cudaStream_t D = pieces > 1 ? c->de_s[chunk] : T;
        DevParams Q = P;
        for (int g = 0; g < pieces; g++) {
            Q.f0 = g * per;
            Q.nf = in->n_frames - Q.f0 < per ? in->n_frames - Q.f0 : per;
            if (Q.nf <= 0) break;
            k_postfilter<<<(n_sel * P.CC + PF_WARPS - 1) / PF_WARPS, 32 * PF_WARPS, 0, Tp>>>(F.hdr, F.fe, F.raw, F.syn, c->d_state, Q);
            if (D != Tp) {
                cudaEventRecord(c->pf_ev[chunk][g], Tp);
                cudaStreamWaitEvent(D, c->pf_ev[chunk][g], 0);
            }
            if (timed && Q.f0 + Q.nf >= in->n_frames) mark(6, Tp);
            // the chains of a 20 ms call in G segments side by side, put together again by k_deemph_fix (kernels.cuh)
            const int G = (pieces == 1 && P.frame_size == 960 && Q.nf >= 2 * c->deemph_segs) ? c->deemph_segs : 1;
            const int dgrid = ((n_sel * P.CC + 31) / 32) * G;
            if (P.frame_size == 960 && (DE_BURST > 1 || DE_GROUPS != 4))  // (the burst fetch needs frames longer than its look-ahead)
                k_deemph<DE_BURST, DE_GROUPS, DE_COOP != 0><<<dgrid, 32, 0, D>>>(F.fe, F.hdr, F.syn, d_pcm, d_status, c->d_state, Q, c->d_deseg, c->d_demem, G, c->deemph_warm);
            else
                k_deemph<1, 4, DE_COOP != 0><<<dgrid, 32, 0, D>>>(F.fe, F.hdr, F.syn, d_pcm, d_status, c->d_state, Q, c->d_deseg, c->d_demem, G, c->deemph_warm);
            if (G > 1) {
                k_deemph_fix<<<(n_sel * P.CC + 127) / 128, 128, 0, D>>>(F.fe, F.syn, d_pcm, c->d_state, Q, c->d_deseg, c->d_demem, G, c->d_deemph_redone);
                c->launches++;
            }
            c->launches += g ? 2 : 0;
        }
        if (D != T) {
            cudaEventRecord(c->de_ev[chunk], D);
            cudaStreamWaitEvent(T, c->de_ev[chunk], 0);
        }
    }
    if (timed) mark(7, T);
    if (silk) {  // "silk" stage time: what is left of the SILK layer once the CELT kernels are done, plus the final sum
        cudaStreamWaitEvent(T, pl.x_ev, 0);
        k_silk_fir<<<wide, 32 * FIR_WARPS, 0, T>>>(B, F.fe, d_pcm, P);
        c->launches += 6;
    }
    if (timed) mark(8, T);
    c->launches += 9;  // entropy, sequence, 3 leaf kernels, bands, imdct, postfilter, deemph
    CU_TRY(c, cudaGetLastError());
    return 0;
}

// One pipelined call over all streams: the head on streams[0] behind `ready`, the tail on c->tail, the SILK stage on
// side[0]; *done fires when PCM and status are complete.  Consecutive calls alternate between the two frame sets.
static int enqueue_pipelined(dm_ctx* c, const dm_batch* in, const uint8_t* d_bytes, const FrameDesc* d_frames, int16_t* d_pcm, int32_t* d_status,
                             cudaEvent_t ready, cudaEvent_t parsed, cudaEvent_t* done) {
    const long long q = c->calls;
    const int r = (int)(q % kRing);
    FrameSet& F = c->set[c->tail_mode ? (q & 1) : 0];
    int rc = alloc_set(c, F);
    if (rc) return rc;
    const cudaStream_t H = c->streams[0];
    if (ready) CU_TRY(c, cudaStreamWaitEvent(H, ready, 0));
    // this frame set was last read by the tail and the SILK stage of the call before the previous one
    if (c->tail_mode && q >= 2) CU_TRY(c, cudaStreamWaitEvent(H, c->done_ev[(q - 2) % kRing], 0));
    CallPlan pl = {H, c->tail_mode ? c->tail : H, c->side[0], &F, 0, c->seq_ev[r], c->head_ev[r], c->x_ev[r], parsed, true};
    rc = launch_pipeline(c, in, d_bytes, d_frames, d_pcm, d_status, 0, c->n_streams, pl);
    if (rc) return rc;
    CU_TRY(c, cudaEventRecord(c->done_ev[r], pl.T));
    c->calls = q + 1;
    if (done) *done = c->done_ev[r];
    return 0;
}

int dm_decode_device_async(dm_ctx* c, const dm_batch* in, dm_out* out, void* ready_event, void** done_event) {
    int r = check_batch(c, in);
    if (r) return r;
    if (!out || !out->pcm || !out->status) return OPUSB_BAD_ARG;
    CU_TRY(c, cudaSetDevice(c->device));
    cudaEvent_t done = nullptr;
    r = enqueue_pipelined(c, in, in->bytes, (const FrameDesc*)in->frames, out->pcm, out->status, (cudaEvent_t)ready_event, nullptr, &done);
    if (done_event) *done_event = (void*)done;
    return r;
}

int dm_stream_wait(dm_ctx* c, void* cuda_stream) {
    if (!c) return OPUSB_BAD_ARG;
    if (c->calls == 0) return 0;
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaStreamWaitEvent((cudaStream_t)cuda_stream, c->done_ev[(c->calls - 1) % kRing], 0));
    return 0;
}

int dm_decode_device(dm_ctx* c, const dm_batch* in, dm_out* out, void* cuda_stream) {
    int r = check_batch(c, in);
    if (r) return r;
    if (!out || !out->pcm || !out->status) return OPUSB_BAD_ARG;
    CU_TRY(c, cudaSetDevice(c->device));
    // one stream-ordered operation on the caller's stream: the internal streams start behind what that stream holds so
    // far, and it waits for the call's last kernel
    const cudaStream_t s = (cudaStream_t)cuda_stream;
    const int rr = (int)(c->calls % kRing);
    CU_TRY(c, cudaEventRecord(c->ready_ev[rr], s));
    cudaEvent_t done = nullptr;
    r = enqueue_pipelined(c, in, in->bytes, (const FrameDesc*)in->frames, out->pcm, out->status, c->ready_ev[rr], nullptr, &done);
    if (r) return r;
    CU_TRY(c, cudaStreamWaitEvent(s, done, 0));
    return 0;
}

// Blocking call: per stream range (chunk) frames H2D -> kernels -> PCM/status D2H, each chunk on its own CUDA stream so
// that the read-back of one range overlaps the kernels of the next.
int decode_many(dm_ctx* c, const dm_batch* in, dm_out* out) {
    int r = check_batch(c, in);
    if (r) return r;
    if (!out || !out->pcm || !out->status) return OPUSB_BAD_ARG;
    CU_TRY(c, cudaSetDevice(c->device));
    if (c->next_ticket || c->calls) CU_TRY(c, cudaDeviceSynchronize());  // pipelined calls in flight finish first
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
        CU_TRY(c, cudaMemcpyAsync(c->d_frames[0] + f0, in->frames + f0, fn * sizeof(dm_frame), cudaMemcpyHostToDevice, s));
        CallPlan pl = {s, s, c->side[k], &c->set[0], k, c->fork_ev[k], nullptr, c->join_ev[k], nullptr, false};
        r = launch_pipeline(c, in, c->d_bytes[0], c->d_frames[0], c->d_pcm, c->d_status, first, n_sel, pl);
        if (r) return r;
        CU_TRY(c, cudaMemcpyAsync(out->pcm + f0 * fs * CC, c->d_pcm + f0 * fs * CC, fn * fs * CC * sizeof(int16_t), cudaMemcpyDeviceToHost, s));
        CU_TRY(c, cudaMemcpyAsync(out->status + f0, c->d_status + f0, fn * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    }
    for (int k = 0; k < n_chunks; k++) CU_TRY(c, cudaStreamSynchronize(c->streams[k]));
    return 0;
}

// Pipelined call with host buffers: upload, head, tail and read-back each have their own CUDA stream and every buffer
// that crosses a stage is double buffered, so with calls in flight the GPU runs the head of call n+1 beside the tail of
// call n while call n-1's PCM crosses PCIe and call n+2's packets arrive.  The kernels run over ALL streams at once: the
// stages that walk a stream's frames in order (sequence, post-filter, de-emphasis, SILK) take as long for 1024 streams
// as for 4096.
int decode_many_async(dm_ctx* c, const dm_batch* in, dm_out* out, long long* ticket_out) {
    int r = check_batch(c, in);
    if (r) return r;
    if (!out || !out->pcm || !out->status) return OPUSB_BAD_ARG;
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t nfr_cap = (size_t)c->n_streams * c->max_frames;
    if (!c->d_pcm2) {
        if ((r = dev_alloc(c, &c->d_pcm2, nfr_cap * c->channels * kMaxFrame))) return r;
        if ((r = dev_alloc(c, &c->d_status2, nfr_cap))) return r;
        if ((r = dev_alloc(c, &c->d_frames[1], nfr_cap))) return r;
    }
    const long long ticket = c->next_ticket;
    const int slot = (int)(ticket % kRing), bi = (int)(ticket & 1);
    if (ticket >= kRing) CU_TRY(c, cudaEventSynchronize(c->host_done[slot]));  // the slot's previous call is over
    if (in->n_bytes > c->d_bytes_cap[bi]) CU_TRY(c, cudaDeviceSynchronize());
    r = reserve_bytes(c, bi, in->n_bytes);
    if (r) return r;
    const int nf = in->n_frames, CC = c->channels;
    const int fs = in->mode == DM_MODE_CELT_RAW ? in->frame_size : 960;
    const size_t fn = (size_t)c->n_streams * nf;
    cudaStream_t up = c->streams[2], down = c->streams[1];
    int16_t* d_pcm = bi ? c->d_pcm2 : c->d_pcm;
    int32_t* d_status = bi ? c->d_status2 : c->d_status;
    // upload: these input buffers were last read (entropy, sequence) two calls ago
    if (ticket >= 2) CU_TRY(c, cudaStreamWaitEvent(up, c->parsed[(ticket - 2) % kRing], 0));
    CU_TRY(c, cudaMemcpyAsync(c->d_bytes[bi], in->bytes, in->n_bytes, cudaMemcpyHostToDevice, up));
    CU_TRY(c, cudaMemcpyAsync(c->d_frames[bi], in->frames, fn * sizeof(dm_frame), cudaMemcpyHostToDevice, up));
    CU_TRY(c, cudaEventRecord(c->uploaded[slot], up));
    // kernels: the output buffer of this parity was last read back two calls ago
    if (ticket >= 2) CU_TRY(c, cudaStreamWaitEvent(c->streams[0], c->host_done[(ticket - 2) % kRing], 0));
    cudaEvent_t done = nullptr;
    r = enqueue_pipelined(c, in, c->d_bytes[bi], c->d_frames[bi], d_pcm, d_status, c->uploaded[slot], c->parsed[slot], &done);
    if (r) return r;
    // read-back
    CU_TRY(c, cudaStreamWaitEvent(down, done, 0));
    CU_TRY(c, cudaMemcpyAsync(out->pcm, d_pcm, fn * fs * CC * sizeof(int16_t), cudaMemcpyDeviceToHost, down));
    CU_TRY(c, cudaMemcpyAsync(out->status, d_status, fn * sizeof(int32_t), cudaMemcpyDeviceToHost, down));
    CU_TRY(c, cudaEventRecord(c->host_done[slot], down));
    c->next_ticket = ticket + 1;
    if (ticket_out) *ticket_out = ticket;
    return 0;
}

int dm_wait(dm_ctx* c, long long ticket) {
    if (!c) return OPUSB_BAD_ARG;
    if (ticket < 0 || ticket >= c->next_ticket) return OPUSB_BAD_ARG;
    if (ticket + kRing <= c->next_ticket - 1) return 0;  // long finished: its ring slot was recycled after a synchronise
    CU_TRY(c, cudaEventSynchronize(c->host_done[(int)(ticket % kRing)]));
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
    CU_TRY(c, cudaEventRecord(c->probe_ev[0], down));
    for (int i = 0; i < reps; i++) CU_TRY(c, cudaMemcpyAsync(host, c->d_pcm, bytes, cudaMemcpyDeviceToHost, down));
    CU_TRY(c, cudaEventRecord(c->probe_ev[1], down));
    CU_TRY(c, cudaStreamSynchronize(down));
    CU_TRY(c, cudaEventElapsedTime(ms_total, c->probe_ev[0], c->probe_ev[1]));
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
const char* dm_stage_name(int i) { return (i >= 0 && i <= kStages) ? kStageNames[i] : ""; }
int dm_stage_times(dm_ctx* c, float* ms, int cap) {
    if (!c || !ms) return 0;
    int n = cap < kStages + 1 ? cap : kStages + 1;
    for (int i = 0; i < n; i++) {
        ms[i] = 0.f;
        const cudaEvent_t a = i == kStages ? c->ev[0] : c->ev[i], b = i == kStages ? c->ev[kStages] : c->ev[i + 1];
        if (cudaEventElapsedTime(&ms[i], a, b) != cudaSuccess) ms[i] = -1.f;
    }
    return n;
}
// OPUSB_TRACE=1: the stage boundaries of the last n pipelined calls (n <= 16) in ms since the first boundary of the oldest
// of them, out[call][kStages + 1]: head start, entropy, sequence, leaves, bands, imdct (head end), postfilter, deemph, end.
int dm_debug_trace(dm_ctx* c, int n, float* out) {
    if (!c || !c->trace || !out || n < 1 || n > kTraceCalls || n > c->calls) return OPUSB_BAD_ARG;
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaDeviceSynchronize());
    const long long q0 = c->calls - n;
    const cudaEvent_t base = c->trace[(size_t)(q0 % kTraceCalls) * (kStages + 1)];
    for (int k = 0; k < n; k++)
        for (int i = 0; i <= kStages; i++)
            if (cudaEventElapsedTime(&out[k * (kStages + 1) + i], base, c->trace[(size_t)((q0 + k) % kTraceCalls) * (kStages + 1) + i]) != cudaSuccess)
                out[k * (kStages + 1) + i] = -1.f;
    return 0;
}

// Debug access: the script of frame t of the most recent call (header and op buffer) as the parser left it.
int dm_debug_frame_script(dm_ctx* c, long long t, void* hdr_out, size_t hdr_cap, void* ops_out, size_t ops_cap) {
    if (!c || t < 0 || t >= (long long)c->n_streams * c->max_frames) return OPUSB_BAD_ARG;
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaDeviceSynchronize());
    if (hdr_out && hdr_cap >= sizeof(CeltFrameHdr)) CU_TRY(c, cudaMemcpy(hdr_out, c->set[0].hdr + t, sizeof(CeltFrameHdr), cudaMemcpyDeviceToHost));
    if (ops_out && ops_cap >= sizeof(BandOp) * kMaxOps) CU_TRY(c, cudaMemcpy(ops_out, c->d_ops + t * kMaxOps, sizeof(BandOp) * kMaxOps, cudaMemcpyDeviceToHost));
    return 0;
}

// Segments of de-emphasis chains that k_deemph_fix did again since dm_create (a segment behind a frame without full CELT
// output, or one whose warm-up had not met the serial run): -1 on error.
long long dm_debug_deemph_redone(dm_ctx* c) {
    if (!c) return -1;
    if (cudaSetDevice(c->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return -1;
    int v = 0;
    if (cudaMemcpy(&v, c->d_deemph_redone, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return v;
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
