// silk_stages.cuh -- SILK synthesis split into GPU-shaped stages.
//
// The reference decodes one frame of one stream from top to bottom (silk_decode_frame, src/silk.cpp:1974).  Here the
// same arithmetic is cut where its dependencies allow:
//
//   params   per (stream, frame, channel), frame-parallel   NLSF -> LPC (both halves), pitch lags, LTP filter
//                                                            (ref silk_decode_parameters :827 minus the gains)
//   core     per (stream, channel) chain, frames in order    gains, excitation, LTP + LPC synthesis
//                                                            (ref silk_gains_dequant :2148, silk_decode_core :1806)
//   unmix    per output sample of the internal rate          mid/side -> left/right, 1-sample delay
//                                                            (ref silk_stereo_MS_to_LR :4028, silk_Decode :1714-1731)
//   up2      per chain, frames in order                      input delay + 2x all-pass up-sampler
//                                                            (ref silk_resampler :3676, _up2_HQ :3515)
//   fir      per 48 kHz output sample                        12-phase 8-tap interpolation, added to the CELT PCM
//                                                            (ref silk_resampler_private_IIR_FIR :3475)
//
// Only `core` and `up2` are serial in time.  They run one chain per LANE (32 independent recurrences per warp, every
// lane fully busy); the long per-chain histories of `core` (output history for the LTP re-whitening, the LTP state
// itself) live in shared memory in lane-interleaved rings -- element i of lane l at word 32*i + l -- so that a lane
// only ever touches its own bank whatever its pitch lag is.
// Everything is written against small accessor types so that tests/hostsim runs the very same code with stride 1.
#pragma once
#include "silk.cuh"

namespace opusb200 {

OPUS_HD uint2 make_uint2_(uint32_t x, uint32_t y) { uint2 r; r.x = x; r.y = y; return r; }

enum { kSilkRing = 384 };  // >= max pitch lag (288) + 2 + one sub-frame (80), a multiple of 4

// element i of a chain-private ring, i within one ring length of [0, kSilkRing); stride = 32 (lane-interleaved shared
// memory) or 1 (host)
template <class T, int kStride>
struct ChainRing {
    T* base;
    OPUS_HD static int wrap(int i) { return i < 0 ? i + kSilkRing : (i >= kSilkRing ? i - kSilkRing : i); }
    OPUS_HD T& operator[](int i) const { return base[wrap(i) * kStride]; }
    // elements i .. i+3 for i a multiple of 4 (a block never wraps): p[0], p[kStride], p[2 * kStride], p[3 * kStride]
    OPUS_HD T* block4(int i) const { return base + wrap(i) * kStride; }
    enum { stride = kStride };
};

// The chain's own output history, newest last, as two spans of global memory: what the previous decoded frame left
// (its xq row, or the history carried from the previous call) and this frame's row so far.  The LTP re-whitening reads
// at most lag + order + 2 <= frame length samples back (ref ltp_mem_length = one frame), so two spans always suffice.
struct ChainHistory {
    const int16_t* prev_end;  // one past the newest sample before this frame; nullptr: the history is all zero
    const int16_t* cur;       // this frame's xq row
    OPUS_HD s32 back(int n_cur, int d) const {  // the sample d >= 1 positions before sample n_cur of this frame
        return d <= n_cur ? (s32)cur[n_cur - d] : (prev_end ? (s32)prev_end[n_cur - d] : 0);
    }
};

// per (frame, channel) output of the params stage
struct alignas(16) SilkCtl {
    int16_t A_Q12[2][kSilkMaxOrder];  // [0] first two sub-frames, [1] last two; zero beyond the LPC order
    int16_t B_Q14[20];
    int16_t pitchL[4];
    int16_t nlsf_Q15[kSilkMaxOrder];
    int32_t LTP_scale_Q14;
    uint8_t interp_flag;              // NLSF_interpolation_flag of silk_decode_core
    uint8_t pad[3];
};

// per frame record of the core stage for the frame-parallel stages behind it
struct SilkPost {
    int16_t prev;          // frame offset (< 0) of the previous frame with SILK output in this call, 0: none since a
                           // re-init (history is zero), 1: the stream state carried from the previous call
    uint8_t has_out;       // this frame carries a decoded SILK layer
    uint8_t fs_kHz;
    uint8_t dec_mask;      // bit n: channel n was decoded (bit 1 clear on mid-only frames)
    uint8_t reinit;        // silk_InitDecoder ran before this frame: stereo history and predictor are zero
    uint8_t rs_reset[2];   // resampler of channel n is set up afresh before this frame
};

// ---- params ----
// An LPC filter that fails its first stability check goes to `fix`: SerialFix runs the bandwidth-expansion loop on the
// spot (host simulation), the params kernel queues the filter for k_silk_fix.
struct SerialFix {
    OPUS_HD void operator()(int16_t* dst, int16_t* a, s32* a32_QA1, int order, bool stable) const {
        if (!stable) silk_nlsf2a_fix(a, a32_QA1, order);
#pragma unroll 1
        for (int i = 0; i < kSilkMaxOrder; i++) dst[i] = i < order ? a[i] : (int16_t)0;
    }
    OPUS_HD void both(int16_t* dst0, int16_t* dst1, int16_t* a, s32* a32_QA1, int order, bool stable) const {
        (*this)(dst1, a, a32_QA1, order, stable);
#pragma unroll 1
        for (int i = 0; i < kSilkMaxOrder; i++) dst0[i] = dst1[i];
    }
};

// prevNLSF == nullptr: first frame after a reset (no interpolation, ref :838)
template <class Fix>
OPUS_HD void silk_params_item(const SilkIdx& ix, int fs_kHz, const int16_t* prevNLSF, SilkCtl& o, const Fix& fix) {
    const SilkCb cb = silk_cb(fs_kHz);
    const int order = cb.order;
    int16_t cur[kSilkMaxOrder], a[kSilkMaxOrder];
    s32 a32[kSilkMaxOrder];
    silk_nlsf_decode(cur, ix.nlsf, cb);
    bool stable = silk_nlsf2a_first(a, a32, cur, order);
#pragma unroll 1
    for (int i = 0; i < kSilkMaxOrder; i++) o.nlsf_Q15[i] = i < order ? cur[i] : (int16_t)0;
    const int interpQ2 = prevNLSF ? ix.interpQ2 : 4;
    if (interpQ2 < 4) {
        fix(o.A_Q12[1], a, a32, order, stable);
        int16_t n0[kSilkMaxOrder];
#pragma unroll 1
        for (int i = 0; i < order; i++) n0[i] = (int16_t)(prevNLSF[i] + (mul_w(interpQ2, cur[i] - prevNLSF[i]) >> 2));
        stable = silk_nlsf2a_first(a, a32, n0, order);
        fix(o.A_Q12[0], a, a32, order, stable);
    } else {
        fix.both(o.A_Q12[0], o.A_Q12[1], a, a32, order, stable);  // no interpolation: both halves use the same filter
    }
    o.interp_flag = interpQ2 < 4;
    if (ix.signalType == 2) {  // ref silk_decode_pitch :2055 and the LTP codebooks :871-886
        const int8_t* cbk = fs_kHz == 8 ? kCbLagsStage2 : kCbLagsStage3;
        const int cbk_size = fs_kHz == 8 ? 11 : 34;
        const int min_lag = 2 * fs_kHz, max_lag = 18 * fs_kHz;
        const int lag = min_lag + ix.lagIndex;
        for (int k = 0; k < 4; k++) o.pitchL[k] = (int16_t)limit32(lag + cbk[k * cbk_size + ix.contour], min_lag, max_lag);
        const int8_t* vq = ix.perIndex == 0 ? kLtpVq0 : ix.perIndex == 1 ? kLtpVq1 : kLtpVq2;
        for (int k = 0; k < 4; k++)
            for (int i = 0; i < 5; i++) o.B_Q14[k * 5 + i] = (int16_t)shl32(vq[ix.ltp[k] * 5 + i], 7);
        o.LTP_scale_Q14 = ix.ltpScaleIndex == 0 ? 15565 : ix.ltpScaleIndex == 1 ? 12288 : 8192;
    } else {
        for (int k = 0; k < 4; k++) o.pitchL[k] = 0;
        for (int k = 0; k < 20; k++) o.B_Q14[k] = 0;
        o.LTP_scale_Q14 = 0;
    }
}

// ---- core ----
// silk_SMLAWB with the 16-bit factor pre-shifted into the upper half: one multiply-high-add
OPUS_HD s32 smlawt(s32 a, s32 b, s32 c_hi) {
#if defined(__CUDA_ARCH__)
    return a + __mulhi(b, c_hi);
#else
    return (s32)((s64)a + (((s64)b * (s64)c_hi) >> 32));
#endif
}

struct SilkChainRegs {       // what a chain carries from frame to frame in registers
    s32 h[kSilkMaxOrder];    // h[j] = sLPC_Q14 of sample (now - 1 - j)  (ref sLPC_Q14_buf, newest first)
    s32 prev_gain_Q16;
    int ltp_pos;             // ring index one past the newest LTP state sample, in [0, kSilkRing)
};

// The two loops of silk_decode_core that run over the LTP state instead of over the frame's samples.  Neither has a
// dependency from one element to the next, so a warp can share them out over its lanes (WarpCoop, silk_kernels.cuh)
// instead of leaving each chain's lane to walk up to 290 elements alone while the unvoiced lanes wait; SerialCoop is
// the chain-local form (host simulation, and the definition of what WarpCoop must compute).
//
// rewhiten: the last n_out = lag + 2 output samples, filtered with this half's LPC (ref silk_LPC_analysis_filter
// :2268) and scaled into the LTP state (ref :1877-1888).  The reference's input window outBuf[start_idx .. ltp_mem)
// (+ the first two sub-frames when k == 2) is simply the newest lag + order + 2 samples of the output history.
// rescale: the newest `count` state samples times gain_adj_Q16 (ref :1892-1896).
struct SerialCoop {
    template <class LtpRing>
    OPUS_HD void rewhiten(bool on, const ChainHistory& hist, int n_cur, int n_out, int ltp_pos, s32 inv_gain_Q31, const int16_t* A_Q12,
                          LtpRing ltp) const {
        if (!on) return;
        s32 a16[kSilkMaxOrder];
#pragma unroll
        for (int j = 0; j < kSilkMaxOrder; j++) a16[j] = A_Q12[j];
        s32 w[kSilkMaxOrder];  // w[j] = in[ix - 1 - j]; the sample whitened first lies n_out back
#pragma unroll
        for (int j = 0; j < kSilkMaxOrder; j++) w[j] = hist.back(n_cur, n_out + 1 + j);
#pragma unroll 1
        for (int o0 = 0; o0 < n_out; o0 += 4) {
            s32 c[4];
#pragma unroll
            for (int u = 0; u < 4; u++) c[u] = o0 + u < n_out ? hist.back(n_cur, n_out - o0 - u) : 0;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                s32 acc[4] = {0, 0, 0, 0};  // silk_SMLABB_ovflw: int16 x int16 products, wrapping sum
#pragma unroll
                for (int j = 0; j < kSilkMaxOrder; j++) {
                    const s32 v = (u - 1 - j >= 0) ? c[(u - 1 - j) & 3] : w[(j - u) & (kSilkMaxOrder - 1)];
                    acc[j & 3] = (s32)((u32)acc[j & 3] + (u32)(v * a16[j]));
                }
                const s32 sum = (s32)(((u32)acc[0] + (u32)acc[1]) + ((u32)acc[2] + (u32)acc[3]));
                const s32 r = sub_w(shl32(c[u], 12), sum);
                const s32 white = sat16(rshift_round(r, 12));
                if (o0 + u < n_out) ltp[ltp_pos - n_out + o0 + u] = smulwb(inv_gain_Q31, white);
            }
#pragma unroll
            for (int j = kSilkMaxOrder - 1; j >= 4; j--) w[j] = w[j - 4];
#pragma unroll
            for (int j = 0; j < 4; j++) w[j] = c[3 - j];
        }
    }
    template <class LtpRing>
    OPUS_HD void rescale(bool on, int ltp_pos, int count, s32 gain_adj_Q16, LtpRing ltp) const {
        if (!on) return;
#pragma unroll 2
        for (int i = 0; i < count; i++) {
            s32& e = ltp[ltp_pos - i - 1];
            e = smulww(gain_adj_Q16, e);
        }
    }
};

// One frame of one chain (ref silk_decode_core, src/silk.cpp:1806).  `hist` is the chain's output history (ref outBuf),
// `ltp` the LTP state (ref sLTP_Q15); xq receives frame_length samples at the internal rate.
// Every lane of the warp calls this for every frame, `active` or not (no frame to synthesise: nothing is read or
// written through the arguments), because the coop steps are warp-wide.
// A chain runs alone on its lane and a warp has few neighbours on its SM (the rings take 48 KB per warp), so the
// code is written for the latency of one instruction stream: the 16-tap sums are split over four accumulators (the
// reference's 32-bit wrapping adds are associative), the next block's pulses are loaded before the current block is
// computed, ring addresses are formed once per block of four samples.
OPUS_LOCAL_TABLE s32 kQuantOffsets[2][2] = {{100, 240}, {32, 100}};  // ref silk_Quantization_Offsets_Q10, src/silk.cpp
template <class LtpRing, class Coop>
OPUS_HD void silk_core_frame(SilkChainRegs& R, ChainHistory hist, LtpRing ltp, const int16_t* pulses, const SilkCtl* ctl, const s32* Gains_Q16,
                             int signalType, int quantOffsetType, int seed, int fs_kHz, int16_t* xq, bool active, const Coop& coop) {
    const int order = fs_kHz == 16 ? 16 : 10;
    const int subfr = 5 * fs_kHz;
    const s32 offset_Q14 = active ? kQuantOffsets[signalType >> 1][quantOffsetType] << 4 : 0;
    u32 rand_seed = (u32)(s32)seed;
    int lag = 0;
    const uint2* p4 = reinterpret_cast<const uint2*>(pulses);  // four pulses per load (rows are 16-byte aligned)
    uint2* x4 = reinterpret_cast<uint2*>(xq);
    uint2 pnext = make_uint2_(0, 0);
    if (active) pnext = p4[0];
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
        s32 A[kSilkMaxOrder];  // A_Q12 << 16 (multiply-high form)
        s32 B[5] = {0, 0, 0, 0, 0};
        s32 Gain_Q10 = 0, inv_gain_Q31 = 0, gain_adj_Q16 = 1 << 16;
        bool rewhiten = false, rescale = false;
        if (active) {
#pragma unroll
            for (int j = 0; j < kSilkMaxOrder; j++) A[j] = (s32)((u32)(uint16_t)ctl->A_Q12[k >> 1][j] << 16);
            const s32 gain_Q16 = Gains_Q16[k];
            Gain_Q10 = gain_Q16 >> 6;
            inv_gain_Q31 = inverse32_varq(gain_Q16, 47);
            if (gain_Q16 != R.prev_gain_Q16) {
                gain_adj_Q16 = div32_varq(R.prev_gain_Q16, gain_Q16, 16);
#pragma unroll
                for (int j = 0; j < kSilkMaxOrder; j++) R.h[j] = smulww(gain_adj_Q16, R.h[j]);
            }
            R.prev_gain_Q16 = gain_Q16;
            if (signalType == 2) {
                lag = ctl->pitchL[k];
#pragma unroll
                for (int t = 0; t < 5; t++) B[t] = (s32)((u32)(uint16_t)ctl->B_Q14[k * 5 + t] << 16);
                rewhiten = k == 0 || (k == 2 && ctl->interp_flag);
                rescale = !rewhiten && gain_adj_Q16 != 1 << 16;
                if (k == 0) inv_gain_Q31 = shl32(smulwb(inv_gain_Q31, ctl->LTP_scale_Q14), 2);
            }
        } else {
#pragma unroll
            for (int j = 0; j < kSilkMaxOrder; j++) A[j] = 0;
        }
        coop.rewhiten(rewhiten, hist, k * subfr, lag + 2, R.ltp_pos, inv_gain_Q31, rewhiten ? ctl->A_Q12[k >> 1] : nullptr, ltp);
        coop.rescale(rescale, R.ltp_pos, lag + 2, gain_adj_Q16, ltp);
        if (!active) continue;
#pragma unroll 1
        for (int i = 0; i < subfr; i += 4) {
            const uint2 pw = pnext;
            const int blk = (k * subfr + i) >> 2;
            if (blk + 1 < 20 * fs_kHz / 4) pnext = p4[blk + 1];
            const s32 pl4[4] = {(s32)(int16_t)(pw.x & 0xffff), (s32)(int16_t)(pw.x >> 16), (s32)(int16_t)(pw.y & 0xffff), (s32)(int16_t)(pw.y >> 16)};
            s32 nw[4];
            s32 yq[4];
            // LTP taps of the whole block first: they lie at least lag - 2 >= 14 samples back, never inside the block
            s32 lt[8];
            if (signalType == 2) {
#pragma unroll
                for (int t = 0; t < 8; t++) lt[t] = ltp[R.ltp_pos - lag - 2 + t];  // state samples q-4 .. q+3 of u = 0
            }
            int32_t* lw = ltp.block4(R.ltp_pos);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                // excitation (ref :1827-1842)
                const s32 pl = pl4[u];
                rand_seed = 907633515u + rand_seed * 196314165u;
                s32 e = shl32(pl, 14);
                if (e > 0) e -= 80 << 4;
                else if (e < 0) e += 80 << 4;
                e += offset_Q14;
                if ((s32)rand_seed < 0) e = -e;
                rand_seed += (u32)pl;
                // long-term prediction (ref :1907-1920): taps q, q-1, .., q-4 with q = pos - lag + 2 (+ u)
                s32 res = e;
                if (signalType == 2) {
                    const s32 pa = smlawt(smlawt(smlawt(2, lt[u + 4], B[0]), lt[u + 3], B[1]), lt[u + 2], B[2]);
                    const s32 pb = smlawt(smlawt(0, lt[u + 1], B[3]), lt[u], B[4]);
                    res = e + shl32(pa + pb, 1);
                    lw[u * LtpRing::stride] = shl32(res, 1);
                }
                // short-term prediction (ref :1925-1960): taps beyond the order are zero
                s32 acc[4] = {order >> 1, 0, 0, 0};
#pragma unroll
                for (int j = 0; j < kSilkMaxOrder; j++) {
                    const s32 v = (u - 1 - j >= 0) ? nw[(u - 1 - j) & 3] : R.h[(j - u) & (kSilkMaxOrder - 1)];
                    acc[j & 3] = smlawt(acc[j & 3], v, A[j]);
                }
                const s32 pred = add_w(add_w(acc[0], acc[1]), add_w(acc[2], acc[3]));  // wraps like the reference's running sum (:1925-1960)
                const s32 s = add_sat32(res, lshift_sat32(pred, 4));
                nw[u] = s;
                yq[u] = sat16(rshift_round(smulww(s, Gain_Q10), 8));
            }
            if (signalType == 2) R.ltp_pos = LtpRing::wrap(R.ltp_pos + 4);
            x4[blk] = make_uint2_((u32)(yq[0] & 0xffff) | (u32)yq[1] << 16, (u32)(yq[2] & 0xffff) | (u32)yq[3] << 16);
#pragma unroll
            for (int j = kSilkMaxOrder - 1; j >= 4; j--) R.h[j] = R.h[j - 4];
#pragma unroll
            for (int j = 0; j < 4; j++) R.h[j] = nw[3 - j];
        }
    }
}

// Per-chain bookkeeping that decides what happens before a frame is synthesised (ref silk_Decode :1512-1563, 1639-1646,
// silk_decoder_set_fs :978, silk_init_decoder :2192).  Returns whether this channel is decoded in this frame.
struct SilkChainCtl {
    int fs_kHz, LastGainIndex;
    int first;  // no frame decoded since the last clear (ref first_frame_after_reset)
};
OPUS_HD void silk_chain_clear(SilkChainRegs& R) {
#pragma unroll
    for (int j = 0; j < kSilkMaxOrder; j++) R.h[j] = 0;
}
// *cleared: the output history (ref outBuf) and the LPC state were zeroed before this frame
OPUS_HD bool silk_chain_begin_frame(SilkChainCtl& c, SilkChainRegs& R, int n, bool reinit, int fs_kHz, int mid_only, int prev_mid_only,
                                    bool* rs_reset, bool* cleared) {
    *cleared = false;
    if (reinit) {  // silk_InitDecoder: both channels as new; the resampler follows through the rate check below
        *cleared = true;
        R.prev_gain_Q16 = 65536;
        c.LastGainIndex = 0;
        c.fs_kHz = 0;
    }
    *rs_reset = c.fs_kHz != fs_kHz;
    if (c.fs_kHz != fs_kHz) {  // silk_decoder_set_fs
        *cleared = true;
        c.LastGainIndex = 10;
        c.fs_kHz = fs_kHz;
    }
    if (n == 1 && !mid_only && prev_mid_only == 1) {  // the side channel comes back after mid-only frames
        *cleared = true;
        c.LastGainIndex = 10;
    }
    if (*cleared) {
        silk_chain_clear(R);
        c.first = 1;
    }
    return n == 0 || !mid_only;
}

// ---- unmix: one sample of the signal the resampler sees (ref silk_stereo_MS_to_LR :4028; mono: the one-sample delay
// of silk_Decode :1727-1731).  m(i), s(i): mid / side sample i of this frame for i >= 0, history for i = -2, -1.
struct SilkUnmix {
    s32 pred0_prev, pred1_prev, pred0, pred1, delta0, delta1;
    int n_interp;
    OPUS_HD void setup(const int32_t* pred_Q13, s32 p0_prev, s32 p1_prev, int fs_kHz) {
        pred0_prev = p0_prev; pred1_prev = p1_prev; pred0 = pred_Q13[0]; pred1 = pred_Q13[1];
        const s32 denom_Q16 = ((s32)1 << 16) / (8 * fs_kHz);
        delta0 = rshift_round(smulbb(pred0 - pred0_prev, denom_Q16), 16);
        delta1 = rshift_round(smulbb(pred1 - pred1_prev, denom_Q16), 16);
        n_interp = 8 * fs_kHz;
    }
    // x1a, x1b, x1c = mid[n-2], mid[n-1], mid[n]; x2b = side[n-1]  ->  left/right sample n-1 (output index n)
    OPUS_HD void sample(int n, s32 x1a, s32 x1b, s32 x1c, s32 x2b, int16_t* l, int16_t* r) const {
        const s32 p0 = n < n_interp ? pred0_prev + delta0 * (n + 1) : pred0;
        const s32 p1 = n < n_interp ? pred1_prev + delta1 * (n + 1) : pred1;
        s32 sum = shl32(x1a + x1c + shl32(x1b, 1), 9);
        sum = smlawb(shl32(x2b, 8), sum, p0);
        sum = smlawb(sum, shl32(x1b, 11), p1);
        const s32 side = sat16(rshift_round(sum, 8));
        *l = sat16(x1b + side);
        *r = sat16(x1b - side);
    }
};

// ---- up2: input delay line + 2x all-pass up-sampler, one input sample -> two output samples ----
struct SilkUp2Regs {
    s32 S[6];
    s32 d[8];  // d[j] = input sample (now - 1 - j)
};
OPUS_HD void silk_up2_sample(SilkUp2Regs& R, s32 xin, int delay, int16_t* o0, int16_t* o1) {
    // the resampler consumes its input `delay` samples late (ref silk_resampler :3688-3699: 1 ms from the delay buffer
    // first, the newest `inputDelay` samples are kept for the next call)
    const s32 x = delay == 0 ? xin : delay == 4 ? R.d[3] : R.d[6];
#pragma unroll
    for (int j = 7; j > 0; j--) R.d[j] = R.d[j - 1];
    R.d[0] = xin;
    s32* S = R.S;
    const s32 in32 = shl32(x, 10);
    s32 Y = in32 - S[0];
    s32 X = smulwb(Y, 1746);
    s32 out32_1 = S[0] + X;
    S[0] = in32 + X;
    Y = out32_1 - S[1];
    X = smulwb(Y, 14986);
    s32 out32_2 = S[1] + X;
    S[1] = out32_1 + X;
    Y = out32_2 - S[2];
    X = smlawb(Y, Y, -26453);
    out32_1 = S[2] + X;
    S[2] = out32_2 + X;
    *o0 = sat16(rshift_round(out32_1, 10));
    Y = in32 - S[3];
    X = smulwb(Y, 6854);
    out32_1 = S[3] + X;
    S[3] = in32 + X;
    Y = out32_1 - S[4];
    X = smulwb(Y, 25769);
    out32_2 = S[4] + X;
    S[4] = out32_1 + X;
    Y = out32_2 - S[5];
    X = smlawb(Y, Y, -9994);
    out32_1 = S[5] + X;
    S[5] = out32_2 + X;
    *o1 = sat16(rshift_round(out32_1, 10));
}
OPUS_HD int silk_input_delay(int fs_kHz) { return fs_kHz == 8 ? 0 : fs_kHz == 12 ? 4 : 7; }  // ref delay_matrix_dec :326

// ---- fir: 48 kHz output sample o (0..959) of a 20 ms frame from the 2x up-sampled row `u` whose first 8 entries are
// the last 8 up-sampled samples of the previous frame (ref silk_resampler_private_IIR_FIR :3475 as called from
// silk_resampler: 1 ms, then batches of 10 ms, each batch restarting its phase accumulator).
OPUS_HD s32 silk_inv_ratio_Q16(int fs_kHz) {  // ref silk_resampler_init :3640-3646
    const s32 Fs_in = fs_kHz * 1000, Fs_out = 48000;
    s32 r = shl32(shl32(Fs_in, 14 + 1) / Fs_out, 2);
    while (smulww(r, Fs_out) < shl32(Fs_in, 1)) r++;
    return r;
}
OPUS_HD int16_t silk_fir_sample(const int16_t* u, int o, int fs_kHz, s32 inv_ratio_Q16) {
    // segment of the frame this output falls into: 48 samples from the first millisecond, then 480 per 10 ms batch
    int k, base;
    if (o < 48) { k = o; base = 0; }
    else if (o < 528) { k = o - 48; base = 2 * fs_kHz; }
    else { k = o - 528; base = 22 * fs_kHz; }
    const s32 index_Q16 = k * inv_ratio_Q16;
    const s32 ti = smulwb(index_Q16 & 0xFFFF, 12);
    const int16_t* bp = u + base + (index_Q16 >> 16);
    s32 res = smulbb(bp[0], kResamplerFir12[ti * 4 + 0]);
    res = smlabb(res, bp[1], kResamplerFir12[ti * 4 + 1]);
    res = smlabb(res, bp[2], kResamplerFir12[ti * 4 + 2]);
    res = smlabb(res, bp[3], kResamplerFir12[ti * 4 + 3]);
    res = smlabb(res, bp[4], kResamplerFir12[(11 - ti) * 4 + 3]);
    res = smlabb(res, bp[5], kResamplerFir12[(11 - ti) * 4 + 2]);
    res = smlabb(res, bp[6], kResamplerFir12[(11 - ti) * 4 + 1]);
    res = smlabb(res, bp[7], kResamplerFir12[(11 - ti) * 4 + 0]);
    return sat16(rshift_round(res, 15));
}

}  // namespace opusb200
