// kernels.cuh -- the CUDA kernels of the decode path (sm_100a), one per subsystem.
//
//   k_entropy   thread per (stream, frame)   range decoder + allocation + PVQ leaves -> script     [a1-a5]
//               (two instantiations: frames without / with a SILK layer)
//   k_sequence  warp   per stream            energy recurrence, noise seed, post-filter history    [a2,a12]
//   k_bands     warp   per (stream, frame)   folding, Haar/Hadamard, stereo merge, anti-collapse   [a4-a7]
//   k_imdct_long warp  per (stream, frame)   long blocks: denormalise + inverse MDCT in registers  [a8,a9]
//   k_imdct     warp   per (stream, frame, ch) every other frame: the same in shared memory         [a8,a9]
//   k_postfilter warp  per (stream, ch)      TDAC overlap-add + comb filter, frames in order       [a9,a10]
//   k_deemph    lane   per (stream, ch)      1-pole IIR + int16 PCM                                 [a11]
//   k_silk_*    see silk_kernels.cuh         SILK parameters, LTP/LPC synthesis, unmix, resampler   [a14-a20]
//
// Frames are indexed t = stream * n_frames + frame.  All intermediates live in HBM between kernels
// (<= 36 KB per frame: 14 GB of DRAM traffic per 204 800-frame step, 17 % of the measured HBM peak over the step and
// 37 % inside k_deemph -- the path is issue / latency bound, SURVEY 8d, but not by orders of magnitude), per-stream
// state is a POD blob in HBM (CeltStreamState).
#pragma once
#include <cuda_pipeline.h>

#include "celt_imdct_regs.cuh"
#include "opus_frame.cuh"

namespace opusb200 {

struct FrameDesc {  // mirrors dm_frame in include/opus_b200.h
    uint32_t offset;
    uint16_t len;
    uint8_t toc;
    uint8_t flags;
};

struct DevParams {
    int n_streams, n_frames;
    int mode;            // 0 Opus frames (TOC), 1 raw CELT
    int frame_size;      // samples per frame per channel in the output layout
    int coded_channels;  // raw CELT only
    int start_band;      // raw CELT only
    int CC;              // decoder (output) channels
    int first_stream, n_sel;  // stream range handled by this launch
    int f0, nf;               // frame range [f0, f0 + nf) of every stream handled by this launch (a "frame group": the host
                              // runs the groups of a call as a software pipeline, see launch_pipeline in opus_b200.cu)
    // work item `local` of a frame-parallel launch -> stream, frame, frame index t
    __device__ __forceinline__ int n_local() const { return n_sel * nf; }
    __device__ __forceinline__ int stream_of(int local) const { return first_stream + local / nf; }
    __device__ __forceinline__ int frame_of(int local) const { return f0 + local % nf; }
    __device__ __forceinline__ size_t index_of(int local) const { return (size_t)stream_of(local) * n_frames + frame_of(local); }
};

#ifndef K1_THREADS
#define K1_THREADS 128
#endif
#ifndef K1_BLOCKS
#define K1_BLOCKS 6
#endif
enum { KW_WARPS = 4 };
// warps per block of the warp-per-frame kernels: a block keeps its shared memory until its slowest warp is done, so the
// band kernel (frames of very different length) runs better in pairs than in fours: 4.83 -> 4.54 ms
#ifndef KB_WARPS
#define KB_WARPS 2
#endif
#ifndef KI_WARPS
#define KI_WARPS 4
#endif
#ifndef KI_GRID
#define KI_GRID 8  // blocks per SM of the persistent grid of k_imdct_long
#endif
#ifndef KIL_BLOCKS
#define KIL_BLOCKS 4  // resident blocks per SM the register allocation of k_imdct_long is held to (128 registers)
#endif
enum { FD_CONT = 1, FD_NULL = 2 };  // dm_frame.flags (include/opus_b200.h)
#ifndef OPUSB_STAGE_PACKETS
#define OPUSB_STAGE_PACKETS 0  // k_entropy: 1 = packets through shared memory (one TMA bulk copy per packet) instead of byte loads from
                               // global memory.  Bit-exact and measured 3-9 % slower (profiles/r02_packet_staging.txt): the parser needs
                               // the shared-memory space as L1 for its work arrays more than it needs the packet bytes closer
#endif
enum { kStageBytes = 176 };  // per packet: covers a 160-byte packet at any alignment (11 x 16 bytes)

__device__ __forceinline__ bool frame_skipped(int status) {
    return status == OPUSB_CELT_BAD_ARG || status == OPUSB_UNIMPLEMENTED || status == OPUSB_BAD_ARG;
}

// ------------------------------------------------------------------------------------------------
// Frames of one kind side by side.  In a thread-per-frame parser the lanes of a warp only work together while their frames
// go the same way, and the one decision that changes the way most is the transient flag: short blocks (B = 8) mean other
// transforms per band and a collapse-mask walk for every coded leaf, which 7 lanes in 8 then sit out.  k_classify reads
// just the first symbols of every CELT frame (celt_parse_head) and writes the work order of the parser kernel: the
// transient frames of the launch first (the longest first), then the others.  Two counters, one atomic per warp and class;
// which slot a frame gets inside its class is left to chance -- frames are independent, the results do not depend on it.
__global__ void __launch_bounds__(256)
k_classify(const uint8_t* __restrict__ bytes, const FrameDesc* __restrict__ fr, DevParams P, uint32_t* __restrict__ order, int* __restrict__ counters) {
    const int local = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = P.n_local();
    const bool in_range = local < n;
    bool transient = false;
    if (in_range) {
        const FrameDesc d = fr[P.index_of(local)];
        const bool celt_kind = P.mode != 0 || (!(d.flags & FD_NULL) && (d.toc & 0x80));
        if (celt_kind && !(P.mode == 0 && (d.flags & FD_NULL)) && d.len > 1 && d.len <= kMaxPacket) {
            RangeDec rd;
            rd_init(rd, bytes + d.offset, d.len);
            const int LM = P.mode == 0 ? 3 : (P.frame_size == 120 ? 0 : P.frame_size == 240 ? 1 : P.frame_size == 480 ? 2 : 3);
            transient = celt_parse_head(rd, P.mode == 0 ? 0 : P.start_band, LM).transient != 0;
        }
    }
    const unsigned active = __ballot_sync(0xffffffffu, in_range);
    const unsigned tr = __ballot_sync(0xffffffffu, in_range && transient);
    if (!in_range) return;
    const unsigned lane = threadIdx.x & 31, below = (1u << lane) - 1u;
    const unsigned mine = transient ? tr : (active & ~tr);
    const int leader = __ffs(mine) - 1;
    int base = 0;
    if ((int)lane == leader) base = atomicAdd(&counters[transient ? 1 : 0], __popc(mine));
    base = __shfl_sync(mine, base, leader);
    const int rank = base + __popc(mine & below);
    // transient frames fill the order from the front, the others from the back
    order[transient ? rank : n - 1 - rank] = (uint32_t)local;
}

// Two instantiations share the frames of a call between them by TOC: kSilk = false takes the frames without a SILK layer
// (CELT-only TOCs, raw CELT frames, fillers), kSilk = true the SILK-only and hybrid frames.  The CELT-only kernel -- the
// whole of the headline workload -- then carries neither the SILK parser's code nor its registers; a warp of the other
// kernel that finds no frame of its kind returns at once.
template <bool kSilk>
__global__ void __launch_bounds__(K1_THREADS, K1_BLOCKS)  // 80 registers: 24 warps per SM measured faster than 16 (128 regs) or 32 (64 regs, spills)
k_entropy(const uint8_t* __restrict__ bytes, const FrameDesc* __restrict__ fr, DevParams P, CeltFrameHdr* __restrict__ hdr,
          BandOp* __restrict__ ops, SilkFrame* __restrict__ silk, const CeltStreamState* __restrict__ st, int* __restrict__ any_silk,
          const uint32_t* __restrict__ order) {
    const int slot = blockIdx.x * K1_THREADS + threadIdx.x;
    bool in_range = slot < P.n_local();
    const int local = (order && in_range) ? (int)order[slot] : slot;  // the work order of k_classify, or frame order
    const size_t t = P.index_of(in_range ? local : 0);
    const FrameDesc d = fr[t];
    // which of the two kernels owns the frame
    const bool silk_kind = P.mode == 0 && !(d.flags & FD_NULL) && !(d.toc & 0x80);
    in_range = in_range && (silk_kind == kSilk);
    const unsigned warp = __ballot_sync(0xffffffffu, in_range);  // the lanes of this warp that own a frame
    if (!in_range) return;
    // ---- packet staging: the first kStageBytes of every packet of the warp go to shared memory, one bulk copy
    // (cp.async.bulk, the TMA unit) per packet, all completing on one mbarrier per warp.  A bulk copy moves whole 16-byte
    // units between 16-byte aligned addresses, so a lane copies the aligned span around its packet and remembers where
    // payload byte 0 landed.  (The span never leaves the allocation the packet lies in: allocations are 256-byte units.)
    u32 sbase = 0, staged = 0;
#if OPUSB_STAGE_PACKETS
    {
        __shared__ __align__(16) uint8_t stage_sm[K1_THREADS][kStageBytes];
        __shared__ __align__(8) unsigned long long stage_bar[K1_THREADS / 32];
        const u32 bar = (u32)__cvta_generic_to_shared(&stage_bar[threadIdx.x >> 5]);
        const int leader = __ffs(warp) - 1;
        if ((int)(threadIdx.x & 31) == leader) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(__popc(warp)) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp(warp);
        const bool has_bytes = !(d.flags & FD_NULL) && d.len > 0;
        const uintptr_t src = reinterpret_cast<uintptr_t>(bytes + d.offset);
        const u32 pre = (u32)(src & 15);
        u32 span = has_bytes ? ((pre + d.len + 15u) & ~15u) : 0u;
        if (span > (u32)kStageBytes) span = kStageBytes;
        const u32 dst = (u32)__cvta_generic_to_shared(&stage_sm[threadIdx.x][0]);
        unsigned long long token;
        if (span) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 %0, [%1], %2;" : "=l"(token) : "r"(bar), "r"(span) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                         "l"(src - pre), "r"(span), "r"(bar)
                         : "memory");
            sbase = dst + pre;
            staged = span - pre < d.len ? span - pre : d.len;
        } else {
            asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(token) : "r"(bar) : "memory");
        }
        (void)token;
        u32 ready = 0;
        while (!ready)  // phase 0 of the barrier: every lane has arrived and every byte has landed
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ready) : "r"(bar) : "memory");
    }
#endif
    CeltFrameHdr& h = hdr[t];
    ParseCtx p;
    p.end = kNbEBands; p.disable_inv = (P.CC == 1);
    p.band = 0; p.intensity = 0; p.spread = 0; p.tf_change = 0; p.remaining_bits = 0;
    p.ops = ops + t * kMaxOps; p.n_ops = 0; p.n_leaves = 0; p.ops_cap = kMaxOps; p.overflow = 0; p.lcg_steps = 0;
    p.C = 1; p.LM = 3; p.start = 0; p.prep_on = 0;
    int oflags = 0;
    bool celt = true, fade = false;
    if (d.flags & FD_NULL) {  // filler behind the last frame of a stream: no output, no state change
        h.status = OPUSB_FRAME_SKIPPED; h.flags = 0; h.n_ops = 0; h.n_leaves = 0; h.LM = 3; h.C = 1;
        h.start = 0; h.end = kNbEBands;
        oflags = OF_NO_CELT;
        celt = false;
    } else if (P.mode == 0) {
        const int f = P.frame_of(local);  // (the previous frame's TOC is call input, whichever group it belongs to)
        const int prev_mode = f > 0 ? opus_toc_mode(fr[t - 1].toc) : st[P.stream_of(local)].prev_mode;
        if constexpr (kSilk) {
            // the lanes with a SILK layer that this decoder can play go through its parser in lock step (silk_parse_frame)
            const unsigned silk_lanes = __ballot_sync(warp, opus_has_silk_layer(d.toc, P.CC));
            oflags = opus_parse_prologue(p, bytes + d.offset, d.len, d.toc, prev_mode, P.CC, h, silk[t], &celt, &fade, silk_lanes, sbase, staged);
        } else {
            oflags = opus_parse_prologue_celt(p, bytes + d.offset, d.len, d.toc, prev_mode, h, sbase, staged);
        }
    } else {
        rd_init(p.rd, bytes + d.offset, d.len, sbase, staged);
        p.C = P.coded_channels;
        p.LM = P.frame_size == 120 ? 0 : P.frame_size == 240 ? 1 : P.frame_size == 480 ? 2 : 3;
        p.start = P.start_band;
    }
    // the lanes that go through the CELT parser step through its band loops together; a frame the parser refuses
    // (ref src/celt.cpp:2216, 2225) leaves it before the first meeting point
    const bool refused = celt && (p.rd.storage > (u32)kMaxPacket || p.rd.storage <= 1);
    p.wmask = __ballot_sync(warp, celt && !refused);
    if (celt) {
        celt_parse_frame(p, h);  // the one call site: the parser is by far the largest body of code in this kernel
        if (fade) {
            if (h.status == OPUSB_CELT_BAD_ARG) oflags |= OF_NO_CELT;
            h.status = 960;
        }
    }
    h.oflags = (uint8_t)oflags;
    if (kSilk && (oflags & OF_HAS_SILK)) *any_silk = 1;  // the SILK kernels of this call have work
}

// ------------------------------------------------------------------------------------------------
// Coded PVQ leaves (alg_unquant, a5).  A leaf is serial work of very uneven length (cwrsi is O(N + K), the spreading
// rotation a recurrence over N; a frame has ~30 leaves of 2 .. 176 coefficients), so with a lane per leaf of ONE frame
// a warp waits for its longest leaf with 3.8 of 32 lanes active.  Here all leaf records of a step are sorted by cost
// class (counting sort: k_leaf_count, k_leaf_place; per-block counts in shared memory, one global atomic per class and
// block) and k_leaf_expand takes them in that order, one leaf per thread: the leaves of a warp cost about the same.
// The vectors go to the frames' X rows, where the band kernel picks them up.
// LeafSort.n[c]: leaves of class c; cursor[c]: next free slot of class c; total / overflow (more leaves than the list
// holds: the band kernel expands its own leaves, as the host simulation does).
struct LeafSort { int n[kLeafClasses], cursor[kLeafClasses], total, overflow; };
enum { kLeafSlotsPerFrame = 48 };  // capacity of the sorted list per frame of the context (a 64 kb/s stereo frame has ~30 leaves)
#ifndef KL_THREADS
#define KL_THREADS 128
#endif
#ifndef KL_GRID
#define KL_GRID 32
#endif

__global__ void __launch_bounds__(KL_THREADS)
k_leaf_count(const CeltFrameHdr* __restrict__ hdr, const BandOp* __restrict__ ops, LeafSort* __restrict__ ls, DevParams P) {
    __shared__ int cnt[kLeafClasses];
    if (threadIdx.x < kLeafClasses) cnt[threadIdx.x] = 0;
    __syncthreads();
    const int local = blockIdx.x * KL_THREADS + threadIdx.x;
    if (local < P.n_local()) {
        const size_t t = P.index_of(local);
        const int n = hdr[t].n_leaves, spread = hdr[t].spread;
        const BandOp* last = ops + t * kMaxOps + (kMaxOps - 1);
#pragma unroll 1
        for (int l = 0; l < n; l++) atomicAdd(&cnt[leaf_class(last[-l], spread)], 1);
    }
    __syncthreads();
    if (threadIdx.x < kLeafClasses && cnt[threadIdx.x]) atomicAdd(&ls->n[threadIdx.x], cnt[threadIdx.x]);
}

// order[] entry: frame index relative to the launch's first frame << 11 | leaf index (kMaxOps < 2048)
__global__ void __launch_bounds__(KL_THREADS)
k_leaf_place(const CeltFrameHdr* __restrict__ hdr, const BandOp* __restrict__ ops, LeafSort* __restrict__ ls, uint32_t* __restrict__ order,
             int cap, DevParams P) {
    __shared__ int cnt[kLeafClasses], base[kLeafClasses];
    if (threadIdx.x < kLeafClasses) cnt[threadIdx.x] = 0;
    __syncthreads();
    int total = 0;
#pragma unroll
    for (int c = 0; c < kLeafClasses; c++) total += ls->n[c];
    if (total > cap || P.n_local() > (1 << 21)) {  // the list is full, or a frame index does not fit an entry
        if (blockIdx.x == 0 && threadIdx.x == 0) { ls->overflow = 1; ls->total = 0; }
        return;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) ls->total = total;
    const int local = blockIdx.x * KL_THREADS + threadIdx.x;
    const bool in = local < P.n_local();
    const size_t t = P.index_of(in ? local : 0);
    const int n = in ? hdr[t].n_leaves : 0, spread = hdr[t].spread;
    const BandOp* last = ops + t * kMaxOps + (kMaxOps - 1);
#pragma unroll 1
    for (int l = 0; l < n; l++) atomicAdd(&cnt[leaf_class(last[-l], spread)], 1);
    __syncthreads();
    if (threadIdx.x < kLeafClasses) {
        // class c starts behind all more expensive classes; this block's share of it
        int start = 0;
        for (int c = kLeafClasses - 1; c > (int)threadIdx.x; c--) start += ls->n[c];
        base[threadIdx.x] = start + (cnt[threadIdx.x] ? atomicAdd(&ls->cursor[threadIdx.x], cnt[threadIdx.x]) : 0);
        cnt[threadIdx.x] = 0;
    }
    __syncthreads();
#pragma unroll 1
    for (int l = 0; l < n; l++) {
        const int c = leaf_class(last[-l], spread);
        order[base[c] + atomicAdd(&cnt[c], 1)] = (uint32_t)local << 11 | (uint32_t)l;
    }
}

__global__ void __launch_bounds__(KL_THREADS)
k_leaf_expand(const CeltFrameHdr* __restrict__ hdr, const BandOp* __restrict__ ops, const LeafSort* __restrict__ ls,
              const uint32_t* __restrict__ order, int16_t* __restrict__ X, DevParams P) {
    const int total = ls->total;
    __align__(16) int16_t y[kMaxLeaf];
#pragma unroll 1
    for (int k = blockIdx.x * KL_THREADS + threadIdx.x; k < total; k += gridDim.x * KL_THREADS) {
        const uint32_t ent = order[k];
        const size_t t = P.index_of((int)(ent >> 11));
        const BandOp o = ops[t * kMaxOps + (kMaxOps - 1) - (ent & 2047u)];
        leaf_expand(o, hdr[t].spread, y);
        int16_t* dst = X + t * (2 * kMaxFrame) + o.off;
        const int N = o.n;
        if (((o.off | N) & 7) == 0) {
#pragma unroll 1
            for (int i = 0; i < N; i += 8) *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<const uint4*>(y + i);
        } else if (((o.off | N) & 1) == 0) {
#pragma unroll 1
            for (int i = 0; i < N; i += 2) *reinterpret_cast<uint32_t*>(dst + i) = *reinterpret_cast<const uint32_t*>(y + i);
        } else {
#pragma unroll 1
            for (int i = 0; i < N; i++) dst[i] = y[i];
        }
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * KW_WARPS)
k_sequence(const CeltFrameHdr* __restrict__ hdr, const FrameDesc* __restrict__ fr, FrameEnergy* __restrict__ fe,
           CeltStreamState* __restrict__ st, DevParams P) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= P.n_sel) return;
    const int s = P.first_stream + w;
    CeltStreamState& S = st[s];
    CeltBandState b;
    if (lane < kNbEBands) b = S.band[lane];
    u32 rng = S.rng;
    int pfp = S.pf_period, pfpo = S.pf_period_old, pfg = S.pf_gain, pfgo = S.pf_gain_old, pft = S.pf_tapset, pfto = S.pf_tapset_old;
    bool pkt_failed = S.pkt_failed != 0;  // a multi-frame packet may straddle two launches (frame groups of a call)
    int last_toc = -1;
    for (int f = P.f0; f < P.f0 + P.nf; f++) {
        const size_t t = (size_t)s * P.n_frames + f;
        const CeltFrameHdr& h = hdr[t];
        FrameEnergy& o = fe[t];
        // a frame that fails ends its packet: the reference returns from opus_decode_native without touching the
        // remaining frames (ref src/opus_decoder.cpp:334-338)
        const int fl = fr[t].flags;
        if (!(fl & FD_CONT)) pkt_failed = false;
        const bool dropped = (fl & FD_NULL) || pkt_failed;
        if (!dropped) {
            last_toc = fr[t].toc;
            if (h.status < 0) pkt_failed = true;
        }
        const bool failed = dropped || frame_skipped(h.status);
        const bool skip = failed || (h.oflags & OF_NO_CELT);
        if (!dropped && (h.oflags & OF_RESET_CELT)) {  // OPUS_RESET_STATE (ref src/celt.cpp:2479-2498)
            rng = 0; pfp = pfpo = pfg = pfgo = pft = pfto = 0;
            if (lane < kNbEBands) for (int c = 0; c < 2; c++) b.logE[c] = b.logE2[c] = -(28 << 10);
        }
        if (lane == 0) {
            o.skip = (uint8_t)((skip ? FS_NO_CELT : 0) | ((h.oflags & OF_HAS_SILK) ? FS_HAS_SILK : 0) | (failed ? FS_FAILED : 0) |
                               (dropped ? FS_DROPPED : 0));
            o.celt_n = skip ? 0 : (uint16_t)(kShortMdct << h.LM);
            o.seed = rng;
            o.pf_period = (int16_t)pfp; o.pf_period_old = (int16_t)pfpo;
            o.pf_gain = (int16_t)pfg; o.pf_gain_old = (int16_t)pfgo;
            o.pf_tapset = (uint8_t)pft; o.pf_tapset_old = (uint8_t)pfto;
        }
        if (skip) continue;
        if (lane < kNbEBands) energy_step(b, lane, h, o);
        rng = h.final_rng;
        // post-filter history (ref src/celt.cpp:2392-2402)
        pfpo = pfp; pfgo = pfg; pfto = pft;
        pfp = h.pf_pitch; pfg = h.pf_gain; pft = h.pf_tapset;
        if (h.LM != 0) { pfpo = pfp; pfgo = pfg; pfto = pft; }
    }
    if (lane < kNbEBands) S.band[lane] = b;
    if (lane == 0) {
        if (P.mode == 0 && last_toc >= 0) S.prev_mode = opus_toc_mode(last_toc);
        S.pkt_failed = pkt_failed;
        S.rng = rng;
        S.pf_period = pfp; S.pf_period_old = pfpo; S.pf_gain = pfg; S.pf_gain_old = pfgo; S.pf_tapset = pft; S.pf_tapset_old = pfto;
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * KB_WARPS)
k_bands(const CeltFrameHdr* __restrict__ hdr, const BandOp* __restrict__ ops, const FrameEnergy* __restrict__ fe,
        int16_t* __restrict__ X, const LeafSort* __restrict__ ls, DevParams P) {
    __shared__ __align__(16) int16_t Wsm[KB_WARPS][W_SIZE];
    const int local = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (local >= P.n_local()) return;
    const size_t t = P.index_of(local);
    const CeltFrameHdr& h = hdr[t];
    const FrameEnergy& e = fe[t];
    if (e.skip & FS_NO_CELT) return;
    Lanes L;
    int16_t* W = Wsm[threadIdx.x >> 5];
    const int words = (h.C * (kShortMdct << h.LM)) >> 1;
    int32_t* Xw = reinterpret_cast<int32_t*>(X + t * (2 * kMaxFrame));
    int32_t* Ww = reinterpret_cast<int32_t*>(W);
    const BandOp* fops = ops + t * kMaxOps;
    if (ls->overflow) {
        leaf_stage(L, W, fops + kMaxOps - 1, h.n_leaves, h.spread);
    } else {
        // the coded leaves are already in the frame's X row (k_leaf_expand); everything else in it is overwritten by the ops
        for (int j = L.lane; j < words; j += 32) Ww[j] = Xw[j];
        L.sync();
    }
    u32 seed = band_stage_ops(L, W, fops, h.n_ops, e.seed);
    if (h.flags & FF_ANTICOLLAPSE) band_stage_anti_collapse(L, W, h, e, seed);
    L.sync();
    for (int j = L.lane; j < words; j += 32) Xw[j] = Ww[j];
}

// ------------------------------------------------------------------------------------------------
// The long block of a 20 ms frame that is not transient (seven frames in eight of the headline workload) goes through the
// register FFT of celt_imdct_regs.cuh: one warp per frame, the two channels side by side in pass A (30 lanes), one after
// the other in pass B.  Blocks are persistent (grid-stride over the frames) so that the rotation and twiddle tables are
// staged in shared memory once per block.  A kernel of its own: the straight-line butterfly code is long, and next to
// the generic FFT (k_imdct, which takes every other frame) it would not stay in the instruction cache.
enum {
    kImdctWarpSmem = 2 * kMaxFrame * 2 + 2 * kImdctBuf * 4 + 2 * 24 * 4,          // spectrum + two transpose buffers + gains
    kImdctLongSmem = KI_WARPS * kImdctWarpSmem + (960 + 3 * 480) * 2,            // + the tables
};
__device__ __forceinline__ bool imdct_is_long(const CeltFrameHdr& h) { return h.LM == 3 && !(h.flags & FF_TRANSIENT); }

__global__ void __launch_bounds__(32 * KI_WARPS, KIL_BLOCKS)
k_imdct_long(const CeltFrameHdr* __restrict__ hdr, const FrameEnergy* __restrict__ fe, const int16_t* __restrict__ X,
             int32_t* __restrict__ raw, DevParams P) {
    // dynamic shared memory (kImdctLongSmem bytes, above the 48 KB static limit): per warp the frame's normalised spectrum
    // (fetched one frame ahead), a transpose buffer per channel and the band gains; per block the tables
    extern __shared__ __align__(16) unsigned char imdct_smem[];
    Lanes L;
    const int warp = threadIdx.x >> 5;
    unsigned char* mine_base = imdct_smem + (size_t)warp * kImdctWarpSmem;
    int16_t* Xs = reinterpret_cast<int16_t*>(mine_base);
    Cpx* T0 = reinterpret_cast<Cpx*>(mine_base + 2 * kMaxFrame * sizeof(int16_t));
    Cpx* T1 = T0 + kImdctBuf / 2;
    int32_t* gs0 = reinterpret_cast<int32_t*>(T1 + kImdctBuf / 2);
    int32_t* gs1 = gs0 + 24;
    int16_t* trig_s = reinterpret_cast<int16_t*>(imdct_smem + (size_t)KI_WARPS * kImdctWarpSmem);
    int16_t* twr_s = trig_s + 960;
    int16_t* twi_s = twr_s + 480;
    int16_t* rev_s = twi_s + 480;
    for (int i = threadIdx.x; i < 960; i += blockDim.x) trig_s[i] = kMdctTrig[i];
    for (int i = threadIdx.x; i < 480; i += blockDim.x) { twr_s[i] = kFftTwRe[i]; twi_s[i] = kFftTwIm[i]; rev_s[i] = kBitrev480[i]; }
    __syncthreads();
    const int CC = P.CC, stride = gridDim.x * KI_WARPS, n_local = P.n_local();
    auto fetch = [&](int local, bool wanted) {  // the X row of frame `local` -> Xs, 16 bytes per lane and copy
        if (wanted && local < n_local) {
            const int16_t* src = X + P.index_of(local) * (2 * kMaxFrame);
            for (int i = 8 * L.lane; i < 2 * kMaxFrame; i += 256) __pipeline_memcpy_async(Xs + i, src + i, 16);
        }
        __pipeline_commit();
    };
    // what the warp needs to know about a frame (and lane b its band energies) is read one frame ahead, like the spectrum
    struct Info { int skip, LM, flags, C, start, end; int16_t E0, E1; };
    auto load_info = [&](int loc) {
        Info v = {FS_NO_CELT, 0, 0, 1, 0, 0, 0, 0};
        if (loc < n_local) {
            const size_t tt = P.index_of(loc);
            const CeltFrameHdr& h = hdr[tt];
            const FrameEnergy& e = fe[tt];
            v.skip = e.skip; v.LM = h.LM; v.flags = h.flags; v.C = h.C; v.start = h.start; v.end = h.end;
            if (L.lane < kNbEBands) { v.E0 = e.E[0][L.lane]; v.E1 = e.E[1][L.lane]; }
        }
        return v;
    };
    auto is_mine = [](const Info& v) { return !(v.skip & FS_NO_CELT) && v.LM == 3 && !(v.flags & FF_TRANSIENT); };
    int local = blockIdx.x * KI_WARPS + warp;
    Info cur = load_info(local), nxt = load_info(local + stride);
    fetch(local, is_mine(cur));
#pragma unroll 1
    for (; local < n_local; local += stride, cur = nxt, nxt = load_info(local + stride)) {
        const size_t t = P.index_of(local);
        const bool mine = is_mine(cur);
        const int C = cur.C, start = (cur.flags & FF_SILENCE) ? 0 : cur.start, end = (cur.flags & FF_SILENCE) ? 0 : cur.end;
        if (mine && L.lane <= kNbEBands) {  // band gains
            gs0[L.lane] = band_gain_packed(L.lane, cur.E0, start, end);
            gs1[L.lane] = band_gain_packed(L.lane, cur.E1, start, end);
        }
        __pipeline_wait_prior(0);
        L.sync();
        if (!mine) {
            fetch(local + stride, is_mine(nxt));
            continue;
        }
        // spectra of the output channels (ref celt_synthesis :2085-2109)
        const bool same = CC == 2 && C == 1;     // mono frame in a stereo decoder: one transform, stored twice
        const bool downmix = CC == 1 && C == 2;  // stereo frame in a mono decoder: the average of the two spectra
        const int nch = (CC == 2 && !same) ? 2 : 1;
        imdct480_prerotate(L.lane, Xs, gs0, gs1, trig_s, rev_s, T0, downmix);
        if (nch == 2) imdct480_prerotate(L.lane, Xs + kMaxFrame, gs1, gs1, trig_s, rev_s, T1, false);
        L.sync();
        fetch(local + stride, is_mine(nxt));  // Xs is free: the next frame's spectrum arrives while this one is transformed
        // pass A: lane = 15 * channel + block, each on its own row
        {
            const int ch = L.lane >= 15 ? 1 : 0, b = L.lane - 15 * ch;
            if (L.lane < 15 * nch) imdct480_pass_a((ch ? T1 : T0) + kImdctRow * b);
        }
        L.sync();
        int32_t* out0 = raw + (t * CC) * kMaxFrame;
#pragma unroll 1
        for (int c = 0; c < nch; c++) imdct480_pass_b(c ? T1 : T0, trig_s, twr_s, twi_s, L.lane, out0 + c * kMaxFrame);
        if (same) {
            L.sync();
            for (int i = L.lane; i < kMaxFrame; i += 32) out0[kMaxFrame + i] = out0[i];
        }
    }
    __pipeline_wait_prior(0);
}

// Every other frame (short blocks, the shorter frame sizes): the generic shared-memory FFT, one warp per (frame, channel).
__global__ void __launch_bounds__(32 * KI_WARPS)
k_imdct(const CeltFrameHdr* __restrict__ hdr, const FrameEnergy* __restrict__ fe, const int16_t* __restrict__ X,
        int32_t* __restrict__ raw, DevParams P) {
    __shared__ __align__(16) int32_t fsm[KI_WARPS][2][kMaxFrame];
    __shared__ int32_t gsm[KI_WARPS][kNbEBands + 1];
    const int local = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (local >= P.n_local() * P.CC) return;
    const int c = local % P.CC;
    const size_t t = P.index_of(local / P.CC);
    const CeltFrameHdr& h = hdr[t];
    const FrameEnergy& e = fe[t];
    if ((e.skip & FS_NO_CELT) || imdct_is_long(h)) return;
    Lanes L;
    int32_t* freq = fsm[threadIdx.x >> 5][0];
    int32_t* work = fsm[threadIdx.x >> 5][1];
    int32_t* gs = gsm[threadIdx.x >> 5];
    const int LM = h.LM, M = 1 << LM, N = kShortMdct << LM, C = h.C, CC = P.CC;
    const int transient = (h.flags & FF_TRANSIENT) != 0, silence = (h.flags & FF_SILENCE) != 0;
    const int16_t* Xf = X + t * (2 * kMaxFrame);
    if (CC == 2 && C == 1) {
        denormalise(L, Xf, freq, e.E[0], h.start, h.end, M, silence, gs);
    } else if (CC == 1 && C == 2) {
        denormalise(L, Xf, freq, e.E[0], h.start, h.end, M, silence, gs);
        denormalise(L, Xf + N, work, e.E[1], h.start, h.end, M, silence, gs);
        for (int i = L.lane; i < N; i += 32) freq[i] = (freq[i] >> 1) + (work[i] >> 1);
        L.sync();
    } else {
        denormalise(L, Xf + c * N, freq, e.E[c], h.start, h.end, M, silence, gs);
    }
    imdct_raw(L, freq, work, transient ? 3 : 3 - LM, transient ? M : 1);
    int32_t* out = raw + (t * CC + c) * kMaxFrame;
    for (int i = L.lane; i < N; i += 32) out[i] = work[i];
}

// ------------------------------------------------------------------------------------------------
// One warp walks the frames of one (stream, channel) in order.  The signal history the comb filter reaches back
// into (<= 1026 samples) and the frame being filtered live side by side in one linear shared-memory buffer --
// [history | frame] -- so that every access of the filter is a plain offset; after a frame the newest 1028 samples are
// moved down to the front (33 loads + 33 stores per lane, cheaper than masking every index of a ring: 18 % of the
// kernel's instructions went into that).  The next frame's raw IMDCT output is fetched with cp.async into a staging row
// while the current frame is being filtered (the warp has nothing else to overlap a global-memory round trip with).
#ifndef PF_WARPS
#define PF_WARPS 4
#endif
enum { kPfLin = kCombHist + kMaxFrame, kPfMove = 352 };  // kPfMove: samples moved per step of the history shift (11 per lane)

__global__ void __launch_bounds__(32 * PF_WARPS)
k_postfilter(const CeltFrameHdr* __restrict__ hdr, const FrameEnergy* __restrict__ fe, const int32_t* __restrict__ raw,
             int32_t* __restrict__ syn, CeltStreamState* __restrict__ st, DevParams P) {
    __shared__ __align__(16) int32_t lsm[PF_WARPS][kPfLin];
    __shared__ __align__(16) int32_t ssm[PF_WARPS][kMaxFrame];
    __shared__ int32_t tsm[PF_WARPS][kOverlap / 2];
    const int local = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (local >= P.n_sel * P.CC) return;
    const int c = local % P.CC, s = P.first_stream + local / P.CC;
    Lanes L;
    int32_t* lin = lsm[threadIdx.x >> 5];
    int32_t* stage = ssm[threadIdx.x >> 5];
    int32_t* tail = tsm[threadIdx.x >> 5];
    CeltStreamState& S = st[s];
    for (int i = L.lane; i < kCombHist; i += 32) lin[i] = S.hist[c][i];
    for (int i = L.lane; i < kOverlap / 2; i += 32) tail[i] = S.tail[c][i];
    const LinView x{lin + kCombHist};  // the frame being filtered; x[-1] is the newest sample of the history
    const size_t t0 = (size_t)s * P.n_frames;
    const int f_end = P.f0 + P.nf;

    // What the warp needs to know about a frame is read one frame ahead (every lane the same words): loaded in front of
    // the frame it would cost a DRAM round trip per frame with nothing to hide it behind.
    struct Info { int skip, celt_n, LM, flags, T0, T1, T2, g0, g1, g2, tap0, tap1, tap2; };
    auto load_info = [&](int f) {
        Info v = {FS_NO_CELT, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (f < f_end) {
            const FrameEnergy& e = fe[t0 + f];
            const CeltFrameHdr& h = hdr[t0 + f];
            v.skip = e.skip; v.celt_n = e.celt_n; v.LM = h.LM; v.flags = h.flags;
            v.T0 = e.pf_period_old; v.T1 = e.pf_period; v.T2 = h.pf_pitch;
            v.g0 = e.pf_gain_old; v.g1 = e.pf_gain; v.g2 = h.pf_gain;
            v.tap0 = e.pf_tapset_old; v.tap1 = e.pf_tapset; v.tap2 = h.pf_tapset;
        }
        return v;
    };
    // stage <- raw output of frame f (16 bytes per lane and copy), if that frame has a CELT layer
    auto fetch = [&](int f, const Info& v) {
        if (f < f_end && !(v.skip & FS_NO_CELT)) {
            const int32_t* src = raw + ((t0 + f) * P.CC + c) * kMaxFrame;
            for (int i = 4 * L.lane; i < v.celt_n; i += 128) __pipeline_memcpy_async(stage + i, src + i, 16);
        }
        __pipeline_commit();
    };
    Info cur = load_info(P.f0), nxt = load_info(P.f0 + 1);
    fetch(P.f0, cur);
    for (int f = P.f0; f < f_end; f++, cur = nxt, nxt = load_info(f + 1)) {
        const size_t t = t0 + f;
        __pipeline_wait_prior(0);
        L.sync();
        if (cur.skip & FS_NO_CELT) {
            fetch(f + 1, nxt);
            continue;
        }
        const int LM = cur.LM, N = kShortMdct << LM;
        const int nblk = (cur.flags & FF_TRANSIENT) ? (1 << LM) : 1;
        tdac_frame(L, stage, tail, x, N / nblk, nblk);  // ends with a warp barrier: the staging row is free again
        fetch(f + 1, nxt);
        comb_filter_inplace(L, x, cur.T0, cur.T1, kShortMdct, (s16)cur.g0, (s16)cur.g1, cur.tap0, cur.tap1);
        if (LM != 0) comb_filter_inplace(L, x + kShortMdct, cur.T1, cur.T2, N - kShortMdct, (s16)cur.g1, (s16)cur.g2, cur.tap1, cur.tap2);
        L.sync();
        int32_t* out = syn + (t * P.CC + c) * kMaxFrame;
        for (int i = L.lane; i < N; i += 32) out[i] = x[i];
        // the newest kCombHist samples become the history: shift down by N, front to back, each step's loads before its stores
#pragma unroll 1
        for (int base = 0; base < kCombHist; base += kPfMove) {
            int32_t v[kPfMove / 32];
#pragma unroll
            for (int k = 0; k < kPfMove / 32; k++) {
                const int i = base + 32 * k + L.lane;
                v[k] = i < kCombHist ? lin[i + N] : 0;
            }
            L.sync();
#pragma unroll
            for (int k = 0; k < kPfMove / 32; k++) {
                const int i = base + 32 * k + L.lane;
                if (i < kCombHist) lin[i] = v[k];
            }
            L.sync();
        }
    }
    __pipeline_wait_prior(0);
    L.sync();
    for (int i = L.lane; i < kCombHist; i += 32) S.hist[c][i] = lin[i];
    for (int i = L.lane; i < kOverlap / 2; i += 32) S.tail[c][i] = tail[i];
}

// ------------------------------------------------------------------------------------------------
// De-emphasis is a strictly serial 1-pole IIR with per-step truncation (ref src/celt.cpp:1977-1982):
// one lane owns one (stream, channel) chain, a warp owns 32 chains -- there is no more parallelism to be had
// (S x CC chains per GPU), so the kernel is built around the LATENCY of one warp: each lane streams its own row
// 32 samples at a time through a private shared-memory ring (8 x 16-byte cp.async per tile, 4 tiles ahead, no
// address arithmetic on the critical path), pulls a tile into registers with 8 LDS.128, runs the recurrence
// (add -> 64-bit multiply -> shift per sample), and the stereo pair of lanes swaps halves with 8 shuffles so that
// every lane ends up with 16 consecutive interleaved frames = 64 contiguous bytes, written as 4 x int16x8 stores.
//
// Segments.  The chain is serial, but it FORGETS: the memory is scaled by 0.85 per sample, so two runs over the same samples
// from different memories differ by less than one after ~110 samples and -- both being integers -- soon meet exactly and
// stay together (simulated on noise of every loudness: always within 170 samples).  So a chain's frames of a call are cut
// into G runs of frames ("segments") that different warps work on at the same time: the first starts from the carried
// memory, every other one from memory 0 `DE_WARM` samples inside the frame in front of it, and records the memory it had
// arrived at where its own frames begin (`spec`) and where they end.  k_deemph_fix then walks the G segments of a chain in order:
// a segment whose `spec` equals the true memory at that point (the end of the segment before it) IS the serial result, bit
// for bit; one that does not (never seen with the full warm-up; forced in the tests with a short one), or that had no full
// frame in front of it to warm up on, is done again from the true memory.  Measured (profiles/r02_sweep_deemph.txt): G = 2
// takes 0.47 ms where the serial kernel takes 0.81; three and more segments are SLOWER again (0.8 - 1.3 ms: beyond about four
// of these warps per SM every one of them slows down by more than their number grows), so G = 2 is the default.
//
// Fetching: a lane's row goes through a ring of tiles in shared memory with cp.async, in groups of kBurst tiles, kGroups
// groups in the ring (<1, 4>: one tile at a time, three ahead; <4, 2>: 512 contiguous bytes of a row at a time, which needs
// frames longer than the look-ahead and is used for 20 ms frames).  The group size hardly matters; who copies does (kCoop).
#ifndef DE_BURST
#define DE_BURST 4
#endif
#ifndef DE_GROUPS
#define DE_GROUPS 2
#endif
#ifndef DE_COOP
#define DE_COOP 1
#endif
enum { DE_TW = 32, DE_ROW = 36, DE_WARM = 256, DE_MAX_SEGS = 8 };
enum { DSEG_EXACT = 0, DSEG_SPEC = 1, DSEG_REDO = 2 };
struct DeemphSeg { int32_t spec, end, kind, pad; };  // [chain of the context][segment]

// frames [a, b) of segment g of G over a range of nf frames
__host__ __device__ __forceinline__ int deemph_seg_begin(int nf, int g, int G) { return (int)((long long)nf * g / G); }

template <int kBurst, int kGroups, bool kCoop>
__global__ void __launch_bounds__(32)
k_deemph(const FrameEnergy* __restrict__ fe, const CeltFrameHdr* __restrict__ hdr, const int32_t* __restrict__ syn,
         int16_t* __restrict__ pcm, int32_t* __restrict__ status, CeltStreamState* __restrict__ st, DevParams P,
         DeemphSeg* __restrict__ segs, int32_t* __restrict__ memf, int G, int warm_tiles) {
    constexpr int RS = kBurst * kGroups;  // ring slots
    __shared__ __align__(16) int32_t ring[RS][32 * DE_ROW];
    const int lane = threadIdx.x;
    const int CC = P.CC;
    const int n_chains = P.n_sel * CC;          // chains are (stream - first_stream) * CC + c
    const int warps_per_seg = (n_chains + 31) >> 5;  // a warp works on ONE segment: its lanes walk the same frames
    const int g = blockIdx.x / warps_per_seg;
    const int my = (blockIdx.x - g * warps_per_seg) * 32 + lane;
    const bool live = my < n_chains;
    const int my_s = P.first_stream + (live ? my / CC : 0), my_c = live ? my % CC : 0;
    s32 mem = (live && g == 0) ? st[my_s].preemph_mem[my_c] : 0;
    const int N = P.frame_size;
    const int fa = P.f0 + deemph_seg_begin(P.nf, g, G), f_end = P.f0 + deemph_seg_begin(P.nf, g + 1, G);  // this warp's frames
    const int WT = g ? warm_tiles : 0;          // warm-up tiles in front of them
    const int tpf = (N + DE_TW - 1) / DE_TW, T = tpf * (f_end - fa);
    const int32_t* row0 = syn + ((size_t)my_s * P.n_frames * CC + my_c) * kMaxFrame;  // frame f of this chain: + f * CC * kMaxFrame
    int16_t* out0 = pcm + (size_t)my_s * P.n_frames * N * CC;                            // frame f of this stream: + f * N * CC

    // What the chain needs to know about a frame -- samples with CELT output, flags, status -- is read one frame ahead:
    // a load in front of every frame would cost the warp a DRAM round trip 50 times per call with nothing to hide it behind.
    int cur_n = 0, cur_sk = 0, cur_st = 0, nxt_n = 0, nxt_sk = 0, nxt_st = 0;
    auto load_frame = [&](int ff, int& n, int& sk, int& stt) {
        n = 0; sk = 0; stt = 0;
        if (live && ff < f_end) {
            const size_t tt = (size_t)my_s * P.n_frames + ff;
            n = fe[tt].celt_n; sk = fe[tt].skip; stt = hdr[tt].status;
        }
    };
    load_frame(fa, cur_n, cur_sk, cur_st);
    load_frame(fa + 1, nxt_n, nxt_sk, nxt_st);
    int f = fa;  // the frame the recurrence is in
    // warm-up (segments behind the first): the last WT tiles of the frame in front of this segment, which must be a full one
    int kind = DSEG_EXACT;
    if (g) kind = (live && fe[(size_t)my_s * P.n_frames + fa - 1].celt_n == N) ? DSEG_SPEC : DSEG_REDO;
    int wq = 0;  // warm-up tiles prefetched so far
    // prefetch cursor (tile index -> frame, offset) kept incrementally; it runs at most one frame ahead of f
    int pf_q = 0, pf_f = fa, pf_j = 0, pf_n = 0;
    int ps = 0;  // ring slot the next tile is fetched into
    // kCoop: the warp copies a tile of its 32 rows TOGETHER -- copy k of a lane is 16 bytes of row 4k + lane / 8, so one
    // instruction covers four whole 128-byte lines instead of 16 bytes of 32 different ones (the LSU handles a line per
    // cycle: lane-private copies made it the limit of the kernel as soon as an SM held more than three of these warps).
    // Whole tiles are copied whatever a row's frame holds (stale rows are never looked at), unless no lane needs the tile.
    const int32_t* rowk[kCoop ? DE_TW / 4 : 1];
    if (kCoop) {
#pragma unroll
        for (int k = 0; k < DE_TW / 4; k++) {
            int ch = (blockIdx.x - g * warps_per_seg) * 32 + 4 * k + (lane >> 3);
            if (ch >= n_chains) ch = 0;  // (a ragged last warp: any row that exists)
            rowk[k] = syn + ((size_t)(P.first_stream + ch / CC) * P.n_frames * CC + ch % CC) * kMaxFrame + 4 * (lane & 7);
        }
    }
    auto issue_tile = [&]() {
        int32_t* dst = kCoop ? &ring[ps][(lane >> 3) * DE_ROW + 4 * (lane & 7)] : &ring[ps][lane * DE_ROW];
        if (++ps == RS) ps = 0;
        if (wq < WT) {  // (warp-uniform)
            const size_t off = (size_t)(fa - 1) * CC * kMaxFrame + (N - WT * DE_TW) + wq * DE_TW;
            if (kCoop) {
#pragma unroll
                for (int k = 0; k < DE_TW / 4; k++) __pipeline_memcpy_async(dst + 4 * k * DE_ROW, rowk[k] + off, 16);
            } else if (live) {
                const int32_t* src = row0 + off;
#pragma unroll
                for (int k = 0; k < DE_TW / 4; k++) __pipeline_memcpy_async(dst + 4 * k, src + 4 * k, 16);
            }
            wq++;
            return;
        }
        if (kCoop) {
            if (pf_j == 0) pf_n = (pf_q < T && live) ? (pf_f == f ? cur_n : nxt_n) : 0;
            if (__any_sync(0xffffffffu, pf_n > pf_j)) {
                const size_t off = (size_t)pf_f * CC * kMaxFrame + pf_j;
#pragma unroll
                for (int k = 0; k < DE_TW / 4; k++) __pipeline_memcpy_async(dst + 4 * k * DE_ROW, rowk[k] + off, 16);
            }
        } else if (pf_q < T && live) {
            if (pf_j == 0) pf_n = pf_f == f ? cur_n : nxt_n;  // nothing to fetch for frames without a CELT layer
            const int w = min(DE_TW, pf_n - pf_j);
            const int32_t* src = row0 + (size_t)pf_f * CC * kMaxFrame + pf_j;
#pragma unroll
            for (int k = 0; k < DE_TW / 4; k++)
                if (4 * k < w) __pipeline_memcpy_async(dst + 4 * k, src + 4 * k, 16);
        }
        pf_q++;
        pf_j += DE_TW;
        if (pf_j >= N) { pf_j = 0; pf_f++; }
    };
    auto issue_group = [&]() {
#pragma unroll
        for (int k = 0; k < kBurst; k++) issue_tile();
        __pipeline_commit();
    };
    for (int q = 0; q < kGroups - 1; q++) issue_group();
    int gt = 0;  // tile of the current group the recurrence is in
    int cs = 0;  // ring slot it reads
    // in front of every tile: at the head of a group, ask for the group kGroups - 1 ahead and wait for this one
    auto fetch_sync = [&]() -> const int4* {
        if (gt == 0) {
            if (kCoop) __syncwarp();  // every lane is done with the slots about to be refilled
            issue_group();
            __pipeline_wait_prior(kGroups - 1);
            if (kCoop) __syncwarp();  // ... and sees the copies of the others
        }
        if (++gt == kBurst) gt = 0;
        const int4* src = reinterpret_cast<const int4*>(&ring[cs][lane * DE_ROW]);
        if (++cs == RS) cs = 0;
        return src;
    };

    // the warm-up run: the recurrence alone, nothing stored
#pragma unroll 1
    for (int q = 0; q < WT; q++) {
        const int4* src = fetch_sync();
#pragma unroll
        for (int k = 0; k < DE_TW / 4; k++) {
            const int4 v = src[k];
            deemph_round(v.x, mem); deemph_round(v.y, mem); deemph_round(v.z, mem); deemph_round(v.w, mem);
        }
    }
    const s32 spec = mem;

    int j0 = 0, my_n = 0;
    bool silk_only = false;
    for (int q = 0; q < T; q++) {
        const int w = min(DE_TW, N - j0);
        if (j0 == 0) {
            const size_t my_t = (size_t)my_s * P.n_frames + f;
            my_n = cur_n;  // samples of this frame that carry CELT output (0, or 120 for a fade-out frame)
            // a frame with no CELT layer whose SILK layer is decoded gets its PCM from k_silk_fir alone: nothing to store here
            silk_only = live && my_n == 0 && (cur_sk & FS_HAS_SILK) && !(cur_sk & FS_DROPPED);
            if (live && my_c == 0) status[my_t] = (cur_sk & FS_DROPPED) ? OPUSB_FRAME_SKIPPED : cur_st;
        }
        // tile -> registers (a lane only ever reads what it copied itself: no warp synchronisation needed)
        const int4* src = fetch_sync();
        s32 xv[DE_TW];
#pragma unroll
        for (int k = 0; k < DE_TW / 4; k++) {
            const int4 v = src[k];
            xv[4 * k] = v.x; xv[4 * k + 1] = v.y; xv[4 * k + 2] = v.z; xv[4 * k + 3] = v.w;
        }
        // the recurrence; two samples per 32-bit word.  A tile is either all signal (the common case: no test per sample,
        // and the two saturations + pack of a pair are ONE cvt.pack.sat), all silence, or the ragged end of a short frame.
        const int live_n = my_n - j0;
        uint32_t pk[DE_TW / 2];
        if (live_n >= DE_TW) {
#pragma unroll
            for (int j = 0; j < DE_TW; j += 2) {
                const s32 a = deemph_round(xv[j], mem), b = deemph_round(xv[j + 1], mem);
                asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(pk[j >> 1]) : "r"(b), "r"(a));  // (sat16(b) << 16) | sat16(a)
            }
        } else if (live_n <= 0) {
#pragma unroll
            for (int j = 0; j < DE_TW / 2; j++) pk[j] = 0;
        } else {
#pragma unroll
            for (int j = 0; j < DE_TW; j += 2) {
                s32 a = 0, b = 0;
                if (j < live_n) a = deemph_sample(xv[j], mem);
                if (j + 1 < live_n) b = deemph_sample(xv[j + 1], mem);
                pk[j >> 1] = ((uint32_t)a & 0xffffu) | ((uint32_t)b << 16);
            }
        }
        if (CC == 2) {
            // lanes 2k (left) and 2k+1 (right): the left lane keeps frames 0..15 of the tile, the right lane 16..31
            uint32_t other[DE_TW / 4];
            const bool right = lane & 1;
#pragma unroll
            for (int k = 0; k < DE_TW / 4; k++) {
                const uint32_t give = right ? pk[k] : pk[k + DE_TW / 4];
                other[k] = __shfl_xor_sync(0xffffffffu, give, 1);
            }
            uint32_t ow[DE_TW / 2];  // 16 interleaved frames: (R << 16) | L
#pragma unroll
            for (int k = 0; k < DE_TW / 4; k++) {
                const uint32_t l = right ? other[k] : pk[k], r = right ? pk[k + DE_TW / 4] : other[k];
                ow[2 * k] = __byte_perm(l, r, 0x5410);
                ow[2 * k + 1] = __byte_perm(l, r, 0x7632);
            }
            if (live && !silk_only) {
                const int h0 = right ? DE_TW / 2 : 0;  // first frame of this lane's half inside the tile
                uint32_t* dst = reinterpret_cast<uint32_t*>(out0 + (size_t)f * N * 2) + j0 + h0;
                if (w == DE_TW) {
#pragma unroll
                    for (int k = 0; k < DE_TW / 8; k++)
                        *reinterpret_cast<uint4*>(dst + 4 * k) = make_uint4(ow[4 * k], ow[4 * k + 1], ow[4 * k + 2], ow[4 * k + 3]);
                } else {
#pragma unroll
                    for (int k = 0; k < DE_TW / 2; k++)
                        if (h0 + k < w) dst[k] = ow[k];
                }
            }
        } else if (live && !silk_only) {
            uint32_t* dst = reinterpret_cast<uint32_t*>(out0 + (size_t)f * N + j0);  // N and j0 are even: 4-byte aligned
            if (w == DE_TW && (((size_t)f * N + j0) & 7) == 0) {
#pragma unroll
                for (int k = 0; k < DE_TW / 8; k++)
                    *reinterpret_cast<uint4*>(dst + 4 * k) = make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
            } else {
#pragma unroll
                for (int k = 0; k < DE_TW / 2; k++)
                    if (2 * k < w) dst[k] = pk[k];
            }
        }
        j0 += DE_TW;
        if (j0 >= N) {
            // (segments) the memory this run had at the end of the frame: where k_deemph_fix may stop doing a segment again
            if (G > 1 && live) memf[((size_t)my_s * CC + my_c) * P.n_frames + f] = mem;
            j0 = 0; f++;
            cur_n = nxt_n; cur_sk = nxt_sk; cur_st = nxt_st;
            load_frame(f + 1, nxt_n, nxt_sk, nxt_st);
        }
    }
    if (live) {
        if (G == 1) {
            st[my_s].preemph_mem[my_c] = mem;
        } else {
            DeemphSeg& S = segs[((size_t)my_s * CC + my_c) * DE_MAX_SEGS + g];
            S.spec = spec; S.end = mem; S.kind = kind;
        }
    }
}

// The segments of a chain in order (see k_deemph): thread per chain.  A segment that has to be done again is done frame by
// frame from the true memory, and only until the memory at the end of a frame equals the one the segment's own run had
// there (`memf`): from that frame on its output stands.  That is one frame as a rule -- the case that does occur is a warm-up
// over digital silence (a frame with the silence flag, 2^-15 of random frames): on zeros a negative memory settles at -1, the
// warm-up from 0 stays at 0, and the two meet a few dozen samples into the next frame that has signal.  `redone` counts.
__global__ void __launch_bounds__(128)
k_deemph_fix(const FrameEnergy* __restrict__ fe, const int32_t* __restrict__ syn, int16_t* __restrict__ pcm, CeltStreamState* __restrict__ st,
             DevParams P, const DeemphSeg* __restrict__ segs, const int32_t* __restrict__ memf, int G, int* __restrict__ redone) {
    const int CC = P.CC, N = P.frame_size;
    const int my = blockIdx.x * blockDim.x + threadIdx.x;
    if (my >= P.n_sel * CC) return;
    const int my_s = P.first_stream + my / CC, my_c = my % CC;
    const DeemphSeg* S = segs + ((size_t)my_s * CC + my_c) * DE_MAX_SEGS;
    const int32_t* mf = memf + ((size_t)my_s * CC + my_c) * P.n_frames;
    s32 mem = S[0].end;  // the first segment started from the carried memory
#pragma unroll 1
    for (int g = 1; g < G; g++) {
        const DeemphSeg sg = S[g];
        if (sg.kind == DSEG_SPEC && sg.spec == mem) { mem = sg.end; continue; }
        atomicAdd(redone, 1);
        const int fa = P.f0 + deemph_seg_begin(P.nf, g, G), fb = P.f0 + deemph_seg_begin(P.nf, g + 1, G);
#pragma unroll 1
        for (int f = fa; f < fb; f++) {
            const size_t t = (size_t)my_s * P.n_frames + f;
            const int n = fe[t].celt_n;  // samples with CELT output (a multiple of 8): the others are zeros whatever the memory
            const int4* x = reinterpret_cast<const int4*>(syn + (t * CC + my_c) * kMaxFrame);
            int16_t* out = pcm + (t * N) * CC + my_c;
#pragma unroll 1
            for (int j = 0; j < n; j += 32) {  // 32 samples at a time: eight independent loads, then the chain
                int4 v[8];
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] = (j + 4 * k < n) ? x[(j >> 2) + k] : make_int4(0, 0, 0, 0);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    if (j + 4 * k < n) {
                        int16_t* o = out + (size_t)(j + 4 * k) * CC;
                        o[0] = (int16_t)deemph_sample(v[k].x, mem);
                        o[CC] = (int16_t)deemph_sample(v[k].y, mem);
                        o[2 * CC] = (int16_t)deemph_sample(v[k].z, mem);
                        o[3 * CC] = (int16_t)deemph_sample(v[k].w, mem);
                    }
                }
            }
            if (mem == mf[f]) { mem = sg.end; break; }  // met the segment's own run: the rest of it stands
        }
    }
    st[my_s].preemph_mem[my_c] = mem;
}

// ------------------------------------------------------------------------------------------------
// OPUS_RESET_STATE for a list of streams (ref opus_decoder_ctl, src/opus_decoder.cpp:380-388): CELT soft reset
// (src/celt.cpp:2479-2498: range seed, post-filter, oldLogE/oldLogE2 -- not the history, oldBandE or de-emphasis
// memory), silk_InitDecoder, prev_mode = 0.
__global__ void k_reset_state(const int32_t* __restrict__ ids, int n, CeltStreamState* __restrict__ st, SilkStreamState* __restrict__ sst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    CeltStreamState& S = st[ids[i]];
    S.rng = 0;
    S.pf_period = S.pf_period_old = S.pf_gain = S.pf_gain_old = S.pf_tapset = S.pf_tapset_old = 0;
    S.prev_mode = 0;
    for (int b = 0; b < kNbEBands; b++)
        for (int c = 0; c < 2; c++) S.band[b].logE[c] = S.band[b].logE2[c] = -(28 << 10);
    silk_stream_reinit(sst[ids[i]]);
}

// opus_decoder_init for a list of streams (ids == nullptr: stream = block index): a pristine CELT and SILK state blob
// copied word-wise over each of them, one block per stream.
__global__ void k_init_streams(const int32_t* __restrict__ ids, uint32_t* __restrict__ st, uint32_t* __restrict__ sst,
                               const uint32_t* __restrict__ tc, const uint32_t* __restrict__ ts, int wc, int ws) {
    const size_t s = ids ? (size_t)ids[blockIdx.x] : (size_t)blockIdx.x;
    for (int i = threadIdx.x; i < wc; i += blockDim.x) st[s * wc + i] = tc[i];
    for (int i = threadIdx.x; i < ws; i += blockDim.x) sst[s * ws + i] = ts[i];
}

}  // namespace opusb200

#include "silk_kernels.cuh"
