// silk_kernels.cuh -- the five SILK synthesis kernels (stage bodies in silk_stages.cuh).
//
//   k_silk_params  thread per (stream, frame, channel)   NLSF -> LPC, pitch, LTP filter                [a16]
//   k_silk_core    lane   per (stream, channel) chain    gains, excitation, LTP + LPC synthesis          [a16, a17]
//   k_silk_unmix   thread per internal-rate sample       mid/side -> left/right + 1-sample delay         [a19]
//   k_silk_up2     lane   per (stream, channel) chain    input delay + 2x all-pass up-sampler            [a20]
//   k_silk_fir     thread per 48 kHz output sample       12-phase FIR interpolation, + CELT PCM          [a20, a13]
//
// All of them return at once when the call holds no frame with a SILK layer (*any_silk == 0, set by k_entropy).
#pragma once
#include "silk_stages.cuh"

namespace opusb200 {

enum { kXqRow = kSilkMaxFrame, kUpRow = 2 * kSilkMaxFrame + 8 };

struct SilkBufs {
    SilkFrame* frames;        // [frame]                 entropy -> params, core
    SilkCtl* ctl;             // [frame][2]              params -> core
    SilkPost* post;           // [frame]                 core -> unmix, up2, fir
    int16_t* xq;              // [frame][2][kXqRow]      core -> unmix (mid/side or mono at the internal rate)
    int16_t* lr;              // [frame][2][kXqRow]      unmix -> up2 (what the resampler is fed)
    int16_t* up;              // [frame][2][kUpRow]      up2 -> fir (8 samples of history + 2x up-sampled frame)
    SilkStreamState* state;   // [stream]
    const int* any_silk;
};

__device__ __forceinline__ bool silk_frame_live(const CeltFrameHdr* hdr, const FrameEnergy* fe, size_t t) {
    return (hdr[t].oflags & OF_HAS_SILK) && !(fe[t].skip & FS_DROPPED);
}

// ------------------------------------------------------------------------------------------------
// Thread per (frame, channel).  NLSF -> LPC ends in a stability check; the filters that fail it need up to 16 rounds
// of bandwidth expansion + re-check (each an O(order^2) recursion in 64-bit arithmetic).  Done in place, every warp
// waited for its unluckiest lane with 5 of 32 lanes busy.  So k_silk_params queues the failing filters (global memory,
// one atomic per warp) and k_silk_fix runs the expansion loop with one queued filter per thread: the warps there are
// full and differ only in trip count (params 2.7 ms -> 1.9 + 0.2 ms per 819 200 frames; moving the first check into
// the second kernel as well was measured too: 1.1 + 1.6 ms, no better than in place).
struct FixEntry { s32 a32[kSilkMaxOrder]; int16_t* dst0; int16_t* dst1; int order; int pad; };
struct GlobalFixQueue {
    FixEntry* ents;
    int* count;
    __device__ __forceinline__ void push(int16_t* dst0, int16_t* dst1, const s32* a32, int order) const {
        // warp-aggregated slot allocation: the lanes that push right now share one atomic
        const unsigned m = __activemask();
        const int leader = __ffs(m) - 1, lane = threadIdx.x & 31;
        int base = 0;
        if (lane == leader) base = atomicAdd(count, __popc(m));
        base = __shfl_sync(m, base, leader);
        FixEntry& e = ents[base + __popc(m & ((1u << lane) - 1u))];
#pragma unroll
        for (int i = 0; i < kSilkMaxOrder; i++) e.a32[i] = a32[i];
        e.dst0 = dst0; e.dst1 = dst1; e.order = order;
    }
    __device__ __forceinline__ void operator()(int16_t* dst, int16_t* a, s32* a32, int order, bool stable) const {
        if (stable) {
#pragma unroll 1
            for (int i = 0; i < kSilkMaxOrder; i++) dst[i] = i < order ? a[i] : (int16_t)0;
        } else {
            push(dst, nullptr, a32, order);
        }
    }
    __device__ __forceinline__ void both(int16_t* dst0, int16_t* dst1, int16_t* a, s32* a32, int order, bool stable) const {
        if (stable) {
#pragma unroll 1
            for (int i = 0; i < kSilkMaxOrder; i++) dst0[i] = dst1[i] = i < order ? a[i] : (int16_t)0;
        } else {
            push(dst0, dst1, a32, order);
        }
    }
};

__global__ void __launch_bounds__(128)
k_silk_params(const CeltFrameHdr* __restrict__ hdr, const FrameEnergy* __restrict__ fe, SilkBufs B, DevParams P, FixEntry* fix_ents,
              int* fix_count) {
    if (*B.any_silk == 0) return;
    const GlobalFixQueue queue{fix_ents, fix_count};
    const int total = P.n_local() * 2;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int local = idx >> 1, n = idx & 1;
        if (n >= P.CC) continue;
        const int f = P.frame_of(local), s = P.stream_of(local);
        const size_t t = (size_t)s * P.n_frames + f;
        if (!silk_frame_live(hdr, fe, t)) continue;
        const SilkFrame& F = B.frames[t];
        if (n == 1 && F.mid_only) continue;
        const int fs = F.fs_kHz;
        // the NLSF vector this frame interpolates from: the previous frame this chain decoded, unless a reset lies between
        int16_t prev[kSilkMaxOrder];
        bool have_prev = false;
        if (!(hdr[t].oflags & OF_SILK_REINIT)) {
            int u = f - 1;  // look back inside this launch's frame group only: in front of it the carried state speaks
            while (u >= P.f0 && !silk_frame_live(hdr, fe, t - (size_t)(f - u))) u--;
            if (u >= P.f0) {
                const SilkFrame& Q = B.frames[t - (size_t)(f - u)];
                if (Q.fs_kHz == fs && !(n == 1 && Q.mid_only)) {
                    silk_nlsf_decode(prev, Q.idx[n].nlsf, silk_cb(fs));
                    have_prev = true;
                }
            } else {
                const SilkStreamState& S = B.state[s];
                const SilkChannel& C = S.ch[n];
                if (!C.first_frame_after_reset && C.fs_kHz == fs && !(n == 1 && S.prev_decode_only_middle)) {
                    for (int i = 0; i < kSilkMaxOrder; i++) prev[i] = C.prevNLSF_Q15[i];
                    have_prev = true;
                }
            }
        }
        silk_params_item(F.idx[n], fs, have_prev ? prev : nullptr, B.ctl[t * 2 + n], queue);
    }
}

__global__ void __launch_bounds__(128)
k_silk_fix(const int* __restrict__ any_silk, FixEntry* __restrict__ ents, const int* __restrict__ count) {
    if (*any_silk == 0) return;
    const int qn = *count;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < qn; e += gridDim.x * blockDim.x) {
        const FixEntry& E = ents[e];
        s32 a32[kSilkMaxOrder];
        int16_t a[kSilkMaxOrder];
#pragma unroll
        for (int i = 0; i < kSilkMaxOrder; i++) a32[i] = E.a32[i];
        const int order = E.order;
        silk_nlsf2a_fix(a, a32, order);
#pragma unroll 1
        for (int i = 0; i < kSilkMaxOrder; i++) {
            const int16_t v = i < order ? a[i] : (int16_t)0;
            E.dst0[i] = v;
            if (E.dst1) E.dst1[i] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// One warp = kLanes chains (the other lanes only help in the coop steps).  A chain is one long dependent instruction
// stream and all warps of a batch are resident at once, so the kernel takes as long as ONE warp takes for its frames,
// however many warps there are: what shortens it is a shorter stream per warp.  With 16 chains per warp the coop steps
// serve half as many voiced chains per frame (12.5 -> 12.2 ms for the SILK stages of 16384 chains); with 8 the extra
// warps compete for issue slots (16.3 ms).  So: 16 while the batch still fits the machine in one wave that way, else
// 32 (OPUSB_SILK_LANES overrides, for tuning).
// Dynamic shared memory: the LTP-state ring (int32) of every chain, lane-interleaved (kLanes * 1.5 KB), and the
// staging window of the coop re-whitening.  The chain's output history, which the re-whitening reads, is the xq rows
// this lane wrote earlier (global memory, L1/L2).
inline int silk_chain_lanes(int chains, int n_sm) { return chains <= n_sm * 8 * 16 ? 16 : 32; }

// SerialCoop's two loops shared out over the warp: one voiced chain after the other, 32 elements per trip.
// The chain's input window is staged once in shared memory (`win`), so an output costs 17 conflict-free reads instead
// of 17 selects between the two history spans; the results go to the chain's column of the lane-interleaved ring
// (a 32-way bank conflict per store, 10 stores per chain: cheap next to what the serial loop cost).
template <int kLanes>
struct WarpCoop {
    int32_t* ring_base;   // chain c of the warp owns ring_base[c + kLanes * i]
    int16_t* win;         // kSilkWin samples
    int lane;

    template <class T>
    __device__ __forceinline__ static T* bcast_ptr(T* p, int src) {
        return reinterpret_cast<T*>(__shfl_sync(0xffffffffu, (unsigned long long)p, src));
    }
    template <class LtpRing>
    __device__ __forceinline__ void rewhiten(bool on, const ChainHistory& hist, int n_cur, int n_out, int ltp_pos, s32 inv_gain_Q31,
                                             const int16_t* A_Q12, LtpRing) const {
        unsigned m = __ballot_sync(0xffffffffu, on);
        if (m == 0) return;
        __syncwarp();  // the chains' own earlier writes (xq rows) become visible to the lanes that now read them
#pragma unroll 1
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const int16_t* prev_end = bcast_ptr(hist.prev_end, src);
            const int16_t* cur = bcast_ptr(hist.cur, src);
            const int16_t* a = bcast_ptr(A_Q12, src);
            const int nc = __shfl_sync(0xffffffffu, n_cur, src), no = __shfl_sync(0xffffffffu, n_out, src);
            const int pos = __shfl_sync(0xffffffffu, ltp_pos, src);
            const s32 ig = __shfl_sync(0xffffffffu, inv_gain_Q31, src);
            // window: history positions [nc - no - 16, nc) relative to this frame's row
            const int p0 = nc - no - kSilkMaxOrder;
#pragma unroll 1
            for (int i = lane; i < no + kSilkMaxOrder; i += 32) {
                const int p = p0 + i;
                win[i] = p >= 0 ? cur[p] : (prev_end ? prev_end[p] : (int16_t)0);
            }
            const uint4 a0 = *reinterpret_cast<const uint4*>(a), a1 = *reinterpret_cast<const uint4*>(a + 8);
            const uint32_t aw[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            s32 a16[kSilkMaxOrder];
#pragma unroll
            for (int j = 0; j < kSilkMaxOrder; j++) a16[j] = (int16_t)(j & 1 ? aw[j >> 1] >> 16 : aw[j >> 1] & 0xffff);
            __syncwarp();
            int32_t* ring = ring_base + src;
#pragma unroll 1
            for (int o = lane; o < no; o += 32) {
                const int16_t* x = win + o;  // x[16] is the sample whitened, x[15 - j] the one j + 1 before it
                s32 acc[4] = {0, 0, 0, 0};   // silk_SMLABB_ovflw: int16 x int16 products, wrapping sum
#pragma unroll
                for (int j = 0; j < kSilkMaxOrder; j++) acc[j & 3] = (s32)((u32)acc[j & 3] + (u32)((s32)x[kSilkMaxOrder - 1 - j] * a16[j]));
                const s32 sum = (s32)(((u32)acc[0] + (u32)acc[1]) + ((u32)acc[2] + (u32)acc[3]));
                const s32 r = sub_w(shl32((s32)x[kSilkMaxOrder], 12), sum);
                const s32 white = sat16(rshift_round(r, 12));
                ring[kLanes * ChainRing<int32_t, kLanes>::wrap(pos - no + o)] = smulwb(ig, white);
            }
            __syncwarp();  // win is reused by the next chain; the ring column is read by its owner
        }
    }
    template <class LtpRing>
    __device__ __forceinline__ void rescale(bool on, int ltp_pos, int count, s32 gain_adj_Q16, LtpRing) const {
        unsigned m = __ballot_sync(0xffffffffu, on);
        if (m == 0) return;
        __syncwarp();
#pragma unroll 1
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const int pos = __shfl_sync(0xffffffffu, ltp_pos, src), cnt = __shfl_sync(0xffffffffu, count, src);
            const s32 g = __shfl_sync(0xffffffffu, gain_adj_Q16, src);
            int32_t* ring = ring_base + src;
#pragma unroll 1
            for (int i = lane; i < cnt; i += 32) {
                int32_t& e = ring[kLanes * ChainRing<int32_t, kLanes>::wrap(pos - i - 1)];
                e = smulww(g, e);
            }
        }
        __syncwarp();
    }
};
enum { kSilkWin = 320 };  // >= max lag (288) + 2 + order (16)
constexpr int silk_core_smem(int lanes) { return kSilkRing * lanes * (int)sizeof(int32_t) + kSilkWin * (int)sizeof(int16_t); }

template <int kLanes>
__global__ void __launch_bounds__(32)
k_silk_core(const CeltFrameHdr* __restrict__ hdr, const FrameEnergy* __restrict__ fe, SilkBufs B, DevParams P) {
    if (*B.any_silk == 0) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int32_t* ltp_base = reinterpret_cast<int32_t*>(smem_raw);
    const int lane = threadIdx.x;
    const int chain = blockIdx.x * kLanes + lane;
    const bool alive = lane < kLanes && chain < P.n_sel * P.CC;  // the other lanes only help in the coop steps
    const int n = alive ? chain % P.CC : 0, s = alive ? P.first_stream + chain / P.CC : P.first_stream;
    const ChainRing<int32_t, kLanes> ltp{ltp_base + (alive ? lane : 0)};
    const WarpCoop<kLanes> coop{ltp_base, reinterpret_cast<int16_t*>(ltp_base + kSilkRing * kLanes), lane};
    SilkStreamState& S = B.state[s];
    SilkChannel& C = S.ch[n];
    SilkChainRegs R;
    SilkChainCtl K;
    for (int j = 0; j < kSilkMaxOrder; j++) R.h[j] = C.sLPC_Q14_buf[kSilkMaxOrder - 1 - j];
    R.prev_gain_Q16 = C.prev_gain_Q16;
    R.ltp_pos = 0;
    K.fs_kHz = C.fs_kHz; K.LastGainIndex = C.LastGainIndex; K.first = C.first_frame_after_reset;
    const int16_t* hist_end = C.outBuf + kSilkMaxFrame;  // carried history: the newest 320 output samples, oldest first
    int prev_mid_only = S.prev_decode_only_middle;
    int last_out_f = -1, last_dec_f = -1;
    bool hist_zero = false;  // a re-init happened in this call and no frame has produced output since
    const size_t t0 = (size_t)s * P.n_frames;
#pragma unroll 1
    for (int f = P.f0; f < P.f0 + P.nf; f++) {
        const size_t t = t0 + f;
        SilkPost& post = B.post[t];
        const bool live = alive && silk_frame_live(hdr, fe, t);
        if (alive && !live && n == 0) post.has_out = 0;
        const SilkFrame& F = B.frames[t];
        bool decode = false, rs_reset = false, reinit = false;
        int fs = 16;
        s32 gains[4] = {0, 0, 0, 0};
        int16_t* row = B.xq + (t * 2 + n) * kXqRow;
        if (live) {
            fs = F.fs_kHz;
            reinit = (hdr[t].oflags & OF_SILK_REINIT) != 0;
            if (reinit) { prev_mid_only = 0; hist_zero = true; }
            bool cleared;
            decode = silk_chain_begin_frame(K, R, n, reinit, fs, F.mid_only, prev_mid_only, &rs_reset, &cleared);
            if (cleared) hist_end = nullptr;
            if (decode) {
                int8_t lgi = (int8_t)K.LastGainIndex;
                silk_gains_dequant(gains, F.idx[n].gains, &lgi);
                K.LastGainIndex = lgi;
            }
        }
        silk_core_frame(R, ChainHistory{hist_end, row}, ltp, F.pulses[n], &B.ctl[t * 2 + n], gains, decode ? F.idx[n].signalType : 0,
                        decode ? F.idx[n].quantOffsetType : 0, decode ? F.idx[n].seed : 0, fs, row, decode, coop);
        if (decode) {
            hist_end = row + 20 * fs;
            K.first = 0;
            last_dec_f = f;
        }
        if (!live) continue;
        post.rs_reset[n] = rs_reset;
        if (n == 0) {
            post.prev = hist_zero ? (int16_t)0 : last_out_f >= 0 ? (int16_t)(last_out_f - f) : (int16_t)1;
            post.has_out = 1;
            post.fs_kHz = (uint8_t)fs;
            post.dec_mask = (uint8_t)(1 | ((P.CC == 2 && !F.mid_only) ? 2 : 0));
            post.reinit = reinit;
        }
        prev_mid_only = F.mid_only;
        last_out_f = f;
        hist_zero = false;
    }
    if (!alive) return;
    for (int j = 0; j < kSilkMaxOrder; j++) C.sLPC_Q14_buf[kSilkMaxOrder - 1 - j] = R.h[j];
    C.prev_gain_Q16 = R.prev_gain_Q16;
    if (hist_end != C.outBuf + kSilkMaxFrame) {  // the history moved on: keep its newest frame (zeros after a clear)
        const int L = 20 * K.fs_kHz;
        for (int i = 0; i < L; i++) C.outBuf[kSilkMaxFrame - L + i] = hist_end ? hist_end[i - L] : (int16_t)0;
    }
    C.fs_kHz = K.fs_kHz; C.LastGainIndex = (int8_t)K.LastGainIndex; C.first_frame_after_reset = (int8_t)K.first;
    if (last_dec_f >= 0) {
        const SilkCtl& Lc = B.ctl[(t0 + last_dec_f) * 2 + n];
        for (int i = 0; i < kSilkMaxOrder; i++) C.prevNLSF_Q15[i] = Lc.nlsf_Q15[i];
    }
    if (n == 0) S.prev_decode_only_middle = prev_mid_only;
}

// ------------------------------------------------------------------------------------------------
// history of the mid / side signals in front of frame t (ref stereo_dec_state sMid, sSide, pred_prev_Q13)
struct SilkHist { s32 m2, m1, s1, p0, p1; };
__device__ __forceinline__ SilkHist silk_hist(const SilkBufs& B, size_t t, int s, int CC) {
    SilkHist h = {0, 0, 0, 0, 0};
    const SilkPost& post = B.post[t];
    if (post.prev < 0) {
        const size_t u = t + post.prev;
        const SilkPost& pp = B.post[u];
        const int Lp = 20 * pp.fs_kHz;
        const int16_t* m = B.xq + (u * 2) * kXqRow;
        h.m2 = m[Lp - 2]; h.m1 = m[Lp - 1];
        if (CC == 2) {
            if (pp.dec_mask & 2) h.s1 = B.xq[(u * 2 + 1) * kXqRow + Lp - 1];
            h.p0 = (int16_t)B.frames[u].ms_pred_Q13[0];
            h.p1 = (int16_t)B.frames[u].ms_pred_Q13[1];
        }
    } else if (post.prev == 1) {
        const SilkStreamState& S = B.state[s];
        h.m2 = S.sMid[0]; h.m1 = S.sMid[1]; h.s1 = S.sSide[1];
        h.p0 = S.pred_prev_Q13[0]; h.p1 = S.pred_prev_Q13[1];
    }
    return h;
}

__global__ void __launch_bounds__(256)
k_silk_unmix(SilkBufs B, DevParams P) {
    if (*B.any_silk == 0) return;
    // one warp per frame: the frames are independent and each is a few hundred samples, a block per frame would spend
    // its time waiting for the handful of dependent loads in front of every frame
    const int nfr = P.n_local();
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int local = blockIdx.x * wpb + (threadIdx.x >> 5); local < nfr; local += gridDim.x * wpb) {
        const int s = P.stream_of(local);
        const size_t t = P.index_of(local);
        const SilkPost& post = B.post[t];
        if (!post.has_out) continue;
        const int fs = post.fs_kHz, L = 20 * fs;
        const SilkHist h = silk_hist(B, t, s, P.CC);
        const int16_t* mid = B.xq + (t * 2) * kXqRow;
        if (P.CC == 1) {
            int16_t* o = B.lr + (t * 2) * kXqRow;
            for (int i = lane; i < L; i += 32) o[i] = i == 0 ? (int16_t)h.m1 : mid[i - 1];
        } else {
            const int16_t* side = B.xq + (t * 2 + 1) * kXqRow;
            const bool has_side = post.dec_mask & 2;
            SilkUnmix U;
            U.setup(B.frames[t].ms_pred_Q13, h.p0, h.p1, fs);
            int16_t* ol = B.lr + (t * 2) * kXqRow;
            int16_t* orr = B.lr + (t * 2 + 1) * kXqRow;
            for (int i = lane; i < L; i += 32) {
                const s32 x1a = i >= 2 ? mid[i - 2] : (i == 1 ? h.m1 : h.m2);
                const s32 x1b = i >= 1 ? mid[i - 1] : h.m1;
                const s32 x1c = mid[i];
                const s32 x2b = i >= 1 ? (has_side ? side[i - 1] : 0) : h.s1;
                U.sample(i, x1a, x1b, x1c, x2b, &ol[i], &orr[i]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
template <int kLanes>
__global__ void __launch_bounds__(32)
k_silk_up2(SilkBufs B, DevParams P) {
    if (*B.any_silk == 0) return;
    const int chain = blockIdx.x * kLanes + threadIdx.x;
    if ((int)threadIdx.x >= kLanes || chain >= P.n_sel * P.CC) return;
    const int n = chain % P.CC, s = P.first_stream + chain / P.CC;
    SilkStreamState& S = B.state[s];
    SilkResampler& Rs = S.ch[n].rs;
    SilkUp2Regs R;
    uint32_t fh[4];  // the newest 8 up-sampled samples, two per word
    for (int i = 0; i < 6; i++) R.S[i] = Rs.sIIR[i];
    int fs = Rs.Fs_in_kHz, D = silk_input_delay(fs);
    for (int j = 0; j < 8; j++) R.d[j] = 0;
    for (int i = 0; i < D; i++) R.d[D - 1 - i] = Rs.delayBuf[i];
    for (int i = 0; i < 4; i++) fh[i] = (uint32_t)(uint16_t)Rs.sFIR[2 * i] | (uint32_t)(uint16_t)Rs.sFIR[2 * i + 1] << 16;
    const size_t t0 = (size_t)s * P.n_frames;
    int last_out_f = -1;
    for (int f = P.f0; f < P.f0 + P.nf; f++) {
        const size_t t = t0 + f;
        const SilkPost& post = B.post[t];
        if (!post.has_out) continue;
        if (post.rs_reset[n]) {  // ref silk_resampler_init :3590
            for (int i = 0; i < 6; i++) R.S[i] = 0;
            for (int j = 0; j < 8; j++) R.d[j] = 0;
            for (int j = 0; j < 4; j++) fh[j] = 0;
        }
        fs = post.fs_kHz;
        D = silk_input_delay(fs);
        const int L = 20 * fs;
        const int16_t* in = B.lr + (t * 2 + n) * kXqRow;
        int16_t* u = B.up + (t * 2 + n) * kUpRow;
        *reinterpret_cast<uint4*>(u) = make_uint4(fh[0], fh[1], fh[2], fh[3]);
        // The lane is alone with its recurrence: a load in front of every four samples would cost it a memory round trip
        // each time (measured: ~800 of the ~1000 cycles per trip).  The input is read kUp2Ahead trips ahead into a
        // register queue; L / 4 is 40, 60 or 80 trips, a multiple of the queue length.
        enum { kUp2Ahead = 4 };
        const int trips = L >> 2;
        uint2 q[kUp2Ahead];
#pragma unroll
        for (int k = 0; k < kUp2Ahead; k++) q[k] = *reinterpret_cast<const uint2*>(in + 4 * k);
#pragma unroll 1
        for (int i0 = 0; i0 < trips; i0 += kUp2Ahead) {
#pragma unroll
            for (int kk = 0; kk < kUp2Ahead; kk++) {
                const uint2 raw = q[kk];
                const int i = i0 + kk;
                if (i + kUp2Ahead < trips) q[kk] = *reinterpret_cast<const uint2*>(in + 4 * (i + kUp2Ahead));
                const int16_t x[4] = {(int16_t)(raw.x & 0xffff), (int16_t)(raw.x >> 16), (int16_t)(raw.y & 0xffff), (int16_t)(raw.y >> 16)};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    int16_t a, b;
                    silk_up2_sample(R, x[k], D, &a, &b);
                    fh[k] = (uint32_t)(uint16_t)a | (uint32_t)(uint16_t)b << 16;
                }
                *reinterpret_cast<uint4*>(u + 8 + 8 * i) = make_uint4(fh[0], fh[1], fh[2], fh[3]);
            }
        }
        last_out_f = f;
    }
    for (int i = 0; i < 6; i++) Rs.sIIR[i] = R.S[i];
    for (int i = 0; i < 4; i++) { Rs.sFIR[2 * i] = (int16_t)(fh[i] & 0xffff); Rs.sFIR[2 * i + 1] = (int16_t)(fh[i] >> 16); }
    for (int i = 0; i < D; i++) Rs.delayBuf[i] = (int16_t)R.d[D - 1 - i];
    Rs.Fs_in_kHz = fs;
    if (n == 0 && last_out_f >= 0) {  // stereo history for the next call (ref silk_stereo_MS_to_LR :4040-4043, :4066-4067)
        const size_t tl = t0 + last_out_f;
        const SilkPost& pl = B.post[tl];
        const int L = 20 * pl.fs_kHz;
        const int16_t* m = B.xq + (tl * 2) * kXqRow;
        S.sMid[0] = m[L - 2]; S.sMid[1] = m[L - 1];
        if (P.CC == 2) {
            const int16_t* sd = B.xq + (tl * 2 + 1) * kXqRow;
            const bool has_side = pl.dec_mask & 2;
            S.sSide[0] = has_side ? sd[L - 2] : (int16_t)0;
            S.sSide[1] = has_side ? sd[L - 1] : (int16_t)0;
            S.pred_prev_Q13[0] = (int16_t)B.frames[tl].ms_pred_Q13[0];
            S.pred_prev_Q13[1] = (int16_t)B.frames[tl].ms_pred_Q13[1];
        }
    }
}

// ------------------------------------------------------------------------------------------------
#ifndef FIR_WARPS
#define FIR_WARPS 16
#endif
__global__ void __launch_bounds__(32 * FIR_WARPS)
k_silk_fir(SilkBufs B, const FrameEnergy* __restrict__ fe, int16_t* __restrict__ pcm, DevParams P) {
    if (*B.any_silk == 0) return;
    // one warp per frame; the up-sampled rows of the frame are staged in shared memory (every output reads 8
    // neighbouring taps of them), warps run independently of each other
    __shared__ __align__(16) int16_t rows[FIR_WARPS][2][kUpRow];
    const int nfr = P.n_local();
    const int CC = P.CC;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int16_t(*my)[kUpRow] = rows[warp];
    for (int local = blockIdx.x * FIR_WARPS + warp; local < nfr; local += gridDim.x * FIR_WARPS) {
        const size_t t = P.index_of(local);
        const SilkPost& post = B.post[t];
        if (!post.has_out) continue;
        const int fs = post.fs_kHz;
        const int words = (2 * 20 * fs + 8) / 8;  // 16-byte words per row
        __syncwarp();
        for (int n = 0; n < CC; n++)
            for (int w = lane; w < words; w += 32)
                reinterpret_cast<uint4*>(my[n])[w] = reinterpret_cast<const uint4*>(B.up + (t * 2 + n) * kUpRow)[w];
        __syncwarp();
        const s32 inv = silk_inv_ratio_Q16(fs);
        const bool add = fe[t].celt_n != 0;  // else k_deemph left this frame's PCM to us (no CELT layer)
        // two interleaved int16 samples (one stereo frame, or two mono samples) per lane and store
        uint32_t* o = reinterpret_cast<uint32_t*>(pcm + t * (size_t)(kMaxFrame * CC));
#pragma unroll 2
        for (int i = lane; i < kMaxFrame * CC / 2; i += 32) {
            const uint32_t old = add ? o[i] : 0u;
            s32 a, b;
            if (CC == 2) { a = silk_fir_sample(my[0], i, fs, inv); b = silk_fir_sample(my[1], i, fs, inv); }
            else { a = silk_fir_sample(my[0], 2 * i, fs, inv); b = silk_fir_sample(my[0], 2 * i + 1, fs, inv); }
            const s32 lo = sat16((s32)(int16_t)(old & 0xffff) + a), hi = sat16((s32)(int16_t)(old >> 16) + b);  // ref src/opus_decoder.cpp:271-273
            o[i] = (uint32_t)(lo & 0xffff) | (uint32_t)hi << 16;
        }
    }
}

}  // namespace opusb200
