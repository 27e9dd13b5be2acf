// celt_synth.cuh -- energy recurrence, denormalisation, inverse MDCT, TDAC overlap-add, comb
// post-filter and de-emphasis stage bodies.
#pragma once
#include "celt_bands.cuh"

namespace opusb200 {

// ---------------------------------------------------------------------------------------------
// Per-stream CELT state carried between frames.  Plain POD in HBM.
// ---------------------------------------------------------------------------------------------
enum { kCombHist = 1024 + 4 };  // post-filter reaches back T+2 <= 1024 samples (ref src/celt.cpp:874-884)

struct CeltBandState {   // one per (stream, band): both channels
    int16_t oldE[2], logE[2], logE2[2];
};

struct CeltStreamState {
    uint32_t rng;                 // previous frame's final range = noise seed (ref :2437)
    int32_t pf_period, pf_period_old;
    int32_t pf_gain, pf_gain_old;
    int32_t pf_tapset, pf_tapset_old;
    int32_t prev_mode;            // Opus layer: mode of the previous frame (ref st->prev_mode, src/opus_decoder.cpp:275)
    int32_t pkt_failed;           // a frame of the multi-frame packet being decoded has failed: its remaining frames are dropped
    int32_t preemph_mem[2];       // de-emphasis IIR memory (ref :1972-1985)
    int32_t tail[2][kOverlap / 2];  // raw IMDCT tail awaiting the next frame's TDAC (ref :2349, 3281)
    int32_t hist[2][kCombHist];     // last post-filtered samples, oldest first (ref decode_mem)
    CeltBandState band[kNbEBands];
};

// ---- energy recurrence: one lane per band, frames in order (ref :2234-2237, 3613-3662, 2373-2436) ----
OPUS_HD void energy_step(CeltBandState& s, int i, const CeltFrameHdr& h, FrameEnergy& out) {
    const int C = h.C, LM = h.LM;
    const int intra = (h.flags & FF_INTRA) != 0;
    if (C == 1) s.oldE[0] = (int16_t)imax(s.oldE[0], s.oldE[1]);
    if (i >= h.start && i < h.end) {
        const s16 coef = intra ? (s16)0 : (LM == 0 ? (s16)29440 : LM == 1 ? (s16)26112 : LM == 2 ? (s16)21248 : (s16)16384);
        for (int c = 0; c < C; c++) {
            s32 old = imax(-9216, s.oldE[c]);
            s32 tmp = pshr32(mul16(coef, old), 8) + h.pq[c][i];
            tmp = imax(-(28 << 17), tmp);
            s16 e = (s16)pshr32(tmp, 7);
            s.oldE[c] = (s16)(e + h.fine[c][i]);
        }
    }
    for (int c = 0; c < 2; c++) {
        s16 p1 = s.logE[c], p2 = s.logE2[c];
        if (C == 1) { p1 = (s16)imax(p1, s.logE[1]); p2 = (s16)imax(p2, s.logE2[1]); }
        out.pmin[c][i] = (s16)imin(p1, p2);
    }
    // anti_collapse reads the energies before the silence override; silence never co-occurs with it.
    if (h.flags & FF_SILENCE) for (int c = 0; c < C; c++) s.oldE[c] = -(28 << 10);
    out.E[0][i] = s.oldE[0];
    out.E[1][i] = s.oldE[1];
    if (C == 1) s.oldE[1] = s.oldE[0];
    if (!(h.flags & FF_TRANSIENT)) {
        for (int c = 0; c < 2; c++) { s.logE2[c] = s.logE[c]; s.logE[c] = s.oldE[c]; }
    } else {
        for (int c = 0; c < 2; c++) s.logE[c] = (s16)imin(s.logE[c], s.oldE[c]);
    }
    if (i < h.start || i >= h.end) {
        for (int c = 0; c < 2; c++) { s.oldE[c] = 0; s.logE[c] = s.logE2[c] = -(28 << 10); }
    }
}

// ---- denormalise_bands (ref src/celt.cpp:948): freq[j] = X[j] * 2^(E+eMeans) ----
// band of every short-block bin (eBands5ms, ref src/celt.cpp:200); 21 = above the last band
OPUS_TABLE uint8_t kBinBand[120] = {
    0, 1, 2, 3, 4, 5, 6, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 12, 12, 13, 13, 13, 13, 14, 14, 14, 14, 15, 15, 15, 15, 15, 15, 16, 16, 16, 16, 16, 16, 17, 17, 17, 17, 17, 17, 17, 17, 18, 18, 18, 18, 18, 18, 18, 18, 18, 18, 18, 18, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 21, 21, 21, 21, 21, 21, 21, 21, 21, 21, 21, 21, 21, 21, 21, 21, 21, 21, 21, 21,
};
// Two phases: lane b works out the gain mantissa and shift of band b once (`gs`, kNbEBands + 1 words of warp-private
// scratch), then all lanes sweep the N bins in one flat loop -- the bands are 8 to 176 bins wide at LM = 3, a
// per-band loop leaves most lanes idle in two thirds of them.
OPUS_HD void denormalise(const Lanes& L, const int16_t* X, int32_t* freq, const int16_t* bandLogE, int start, int end, int M,
                         int silence, int32_t* gs) {
    const int N = M * kShortMdct;
    const int LM = ilog2p(M);
    if (silence) start = end = 0;
    for (int i = L.lane; i <= kNbEBands; i += kLanes) {
        int32_t packed = 0;  // g = 0: bins outside [start, end) come out as zero
        if (i >= start && i < end) {
            s16 lg = sat16((s32)bandLogE[i] + shl32((s32)kEMeans[i], 6));
            int shift = 16 - (lg >> 10);
            s16 g;
            if (shift > 31) { shift = 0; g = 0; }
            else g = (s16)exp2_frac((s16)(lg & 1023));
            if (shift <= -2) { g = 16384; shift = -2; }
            packed = (int32_t)(((u32)(uint16_t)g << 16) | (u32)(shift & 0xff));
        }
        gs[i] = packed;
    }
    L.sync();
    for (int j = L.lane; j < N; j += kLanes) {
        const int32_t pk = gs[kBinBand[j >> LM]];
        const s16 g = (s16)(pk >> 16);
        const int shift = (int)(int8_t)(pk & 0xff);
        const s32 v = mul16(X[j], g);
        freq[j] = shift < 0 ? shl32(v, -shift) : v >> shift;
    }
    L.sync();
}

// ---- fixed-point FFT butterflies (ref kf_bfly2/3/4/5, src/celt.cpp:2794-2995) ----
struct alignas(8) Cpx { s32 r, i; };  // one 64-bit shared-memory access per complex value (work buffers are 16-byte aligned)
OPUS_HD s32 smul(s32 a, s16 b) { return mul16x32_q15(b, a); }  // ref S_MUL(a,b), src/celt.h:192
OPUS_HD Cpx cmul(Cpx a, int tw) {                                // ref C_MUL with twiddle index tw
    const s16 br = kFftTwRe[tw], bi = kFftTwIm[tw];
    Cpx m;
    m.r = sub_w(smul(a.r, br), smul(a.i, bi));
    m.i = add_w(smul(a.r, bi), smul(a.i, br));
    return m;
}
OPUS_HD Cpx cadd(Cpx a, Cpx b) { Cpx c; c.r = add_w(a.r, b.r); c.i = add_w(a.i, b.i); return c; }
OPUS_HD Cpx csub(Cpx a, Cpx b) { Cpx c; c.r = sub_w(a.r, b.r); c.i = sub_w(a.i, b.i); return c; }

struct FftStage { uint8_t radix, m, N, fstride, mm; };

// radix-4, unit (i, j) = (t / m, t % m), t in [0, N*m)
OPUS_HD void bfly4(Cpx* f, int i, int j, const FftStage& s) {
    const int m = s.m;
    if (m == 1) {
        Cpx* F = f + 4 * i;
        Cpx a0 = F[0], a1 = F[1], a2 = F[2], a3 = F[3];
        Cpx s0 = csub(a0, a2);
        a0 = cadd(a0, a2);
        Cpx s1 = cadd(a1, a3);
        F[2] = csub(a0, s1);
        F[0] = cadd(a0, s1);
        s1 = csub(a1, a3);
        F[1].r = add_w(s0.r, s1.i); F[1].i = sub_w(s0.i, s1.r);
        F[3].r = sub_w(s0.r, s1.i); F[3].i = add_w(s0.i, s1.r);
    } else {
        Cpx* F = f + i * s.mm + j;
        Cpx c0 = cmul(F[m], j * s.fstride);
        Cpx c1 = cmul(F[2 * m], 2 * j * s.fstride);
        Cpx c2 = cmul(F[3 * m], 3 * j * s.fstride);
        Cpx a0 = F[0];
        Cpx c5 = csub(a0, c1);
        a0 = cadd(a0, c1);
        Cpx c3 = cadd(c0, c2);
        Cpx c4 = csub(c0, c2);
        F[2 * m] = csub(a0, c3);
        F[0] = cadd(a0, c3);
        F[m].r = add_w(c5.r, c4.i); F[m].i = sub_w(c5.i, c4.r);
        F[3 * m].r = sub_w(c5.r, c4.i); F[3 * m].i = add_w(c5.i, c4.r);
    }
}
// radix-2 (always m == 4), unit t in [0, N*4)
OPUS_HD void bfly2(Cpx* f, int t) {
    const int i = t >> 2, k = t & 3;
    Cpx* F = f + 8 * i + k;
    Cpx a = F[0], b = F[4], tt;
    const s16 tw = 23170;
    if (k == 0) tt = b;
    else if (k == 1) { tt.r = smul(add_w(b.r, b.i), tw); tt.i = smul(sub_w(b.i, b.r), tw); }
    else if (k == 2) { tt.r = b.i; tt.i = neg_w(b.r); }
    else { tt.r = smul(sub_w(b.i, b.r), tw); tt.i = smul(neg_w(add_w(b.i, b.r)), tw); }
    F[4] = csub(a, tt);
    F[0] = cadd(a, tt);
}
// radix-3, unit (i, j) = (t / m, t % m)
OPUS_HD void bfly3(Cpx* f, int i, int j, const FftStage& s) {
    const int m = s.m;
    Cpx* F = f + i * s.mm + j;
    Cpx c1 = cmul(F[m], j * s.fstride);
    Cpx c2 = cmul(F[2 * m], 2 * j * s.fstride);
    Cpx c3 = cadd(c1, c2);
    Cpx c0 = csub(c1, c2);
    Cpx a0 = F[0], fm;
    fm.r = sub_w(a0.r, c3.r >> 1);
    fm.i = sub_w(a0.i, c3.i >> 1);
    c0.r = smul(c0.r, -28378);
    c0.i = smul(c0.i, -28378);
    F[0] = cadd(a0, c3);
    F[2 * m].r = add_w(fm.r, c0.i); F[2 * m].i = sub_w(fm.i, c0.r);
    F[m].r = sub_w(fm.r, c0.i);     F[m].i = add_w(fm.i, c0.r);
}
// radix-5, unit (i, u) = (t / m, t % m)
OPUS_HD void bfly5(Cpx* f, int i, int u, const FftStage& s) {
    const int m = s.m;
    Cpx* F = f + i * s.mm + u;
    const s16 yar = 10126, yai = -31164, ybr = -26510, ybi = -19261;
    Cpx s0 = F[0];
    Cpx s1 = cmul(F[m], u * s.fstride);
    Cpx s2 = cmul(F[2 * m], 2 * u * s.fstride);
    Cpx s3 = cmul(F[3 * m], 3 * u * s.fstride);
    Cpx s4 = cmul(F[4 * m], 4 * u * s.fstride);
    Cpx s7 = cadd(s1, s4), s10 = csub(s1, s4), s8 = cadd(s2, s3), s9 = csub(s2, s3);
    F[0].r = add_w(s0.r, add_w(s7.r, s8.r));
    F[0].i = add_w(s0.i, add_w(s7.i, s8.i));
    Cpx s5, s6, s11, s12;
    s5.r = add_w(s0.r, add_w(smul(s7.r, yar), smul(s8.r, ybr)));
    s5.i = add_w(s0.i, add_w(smul(s7.i, yar), smul(s8.i, ybr)));
    s6.r = add_w(smul(s10.i, yai), smul(s9.i, ybi));
    s6.i = neg_w(add_w(smul(s10.r, yai), smul(s9.r, ybi)));
    F[m] = csub(s5, s6);
    F[4 * m] = cadd(s5, s6);
    s11.r = add_w(s0.r, add_w(smul(s7.r, ybr), smul(s8.r, yar)));
    s11.i = add_w(s0.i, add_w(smul(s7.i, ybr), smul(s8.i, yar)));
    s12.r = sub_w(smul(s9.i, yai), smul(s10.i, ybi));
    s12.i = sub_w(smul(s10.r, ybi), smul(s9.r, yai));
    F[2 * m] = cadd(s11, s12);
    F[3 * m] = csub(s11, s12);
}

// One FFT stage with its shape known at compile time: the unit -> (block, i, j) arithmetic folds into shifts and
// multiply-highs, and each butterfly is compiled for its own m / twiddle stride.
template <int kRadix, int kM, int kN, int kFstride, int kMM>
OPUS_HD void fft_stage(const Lanes& L, int32_t* work, int N2, int nblk) {
    constexpr int units = kRadix == 2 ? kN * 4 : kN * kM;
    const FftStage g = FftStage{(uint8_t)kRadix, (uint8_t)kM, (uint8_t)kN, (uint8_t)kFstride, (uint8_t)kMM};
    for (int t = L.lane; t < nblk * units; t += kLanes) {
        const int blk = nblk > 1 ? t / units : 0;
        Cpx* f = (Cpx*)(work + blk * N2);
        const int u = t - blk * units;
        const int i = u / kM, j = u - i * kM;
        if (kRadix == 4) bfly4(f, i, j, g);
        else if (kRadix == 2) bfly2(f, u);
        else if (kRadix == 3) bfly3(f, i, j, g);
        else bfly5(f, i, j, g);
    }
    L.sync();
}

// Inverse MDCT of `nblk` interleaved blocks (ref clt_mdct_backward :3204 without the TDAC mirror, which
// needs the previous block and is done by tdac_* below).  freq: N coefficients, block b at freq[b + k*nblk].
// work: N int32 (nblk FFTs of N4 complex points).  On return work[b*N2 + k] = raw post-rotated y_b[k].
// FFT plans for nfft = 480, 240, 120, 60 (shift 0..3), in execution order (ref opus_fft_impl :2997 with the
// factor lists of fft_state48000_960_{0..3}, :589-626); the twiddle stride already includes << shift.
OPUS_HD void imdct_raw(const Lanes& L, const int32_t* freq, int32_t* work, int shift, int nblk) {
    const int Nm = 1920 >> shift, N2 = Nm >> 1, N4 = Nm >> 2;
    const int16_t* trig = kMdctTrig + (shift == 0 ? 0 : shift == 1 ? 960 : shift == 2 ? 1440 : 1680);
    const int16_t* bitrev = shift == 0 ? kBitrev480 : shift == 1 ? kBitrev240 : shift == 2 ? kBitrev120 : kBitrev60;
    const SmallDiv byN4(N4);
    // pre-rotation into bit-reversed order (ref :3219-3239)
    for (int t = L.lane; t < nblk * N4; t += kLanes) {
        const int b = nblk > 1 ? byN4.div(t) : 0, i = t - b * N4;
        const int32_t* in = freq + b;
        s32 x1 = in[2 * i * nblk], x2 = in[nblk * (N2 - 1 - 2 * i)];
        s16 t0 = trig[i], t1 = trig[N4 + i];
        s32 yr = add_w(smul(x2, t0), smul(x1, t1));
        s32 yi = sub_w(smul(x1, t0), smul(x2, t1));
        Cpx y;
        y.r = yi; y.i = yr;  // the FFT input has real and imaginary parts swapped (ref :3236-3237)
        reinterpret_cast<Cpx*>(work + b * N2)[bitrev[i]] = y;
    }
    L.sync();
    if (shift == 0) {
        fft_stage<4, 1, 120, 120, 4>(L, work, N2, nblk);
        fft_stage<2, 4, 60, 0, 0>(L, work, N2, nblk);
        fft_stage<4, 8, 15, 15, 32>(L, work, N2, nblk);
        fft_stage<3, 32, 5, 5, 96>(L, work, N2, nblk);
        fft_stage<5, 96, 1, 1, 1>(L, work, N2, nblk);
    } else if (shift == 1) {
        fft_stage<4, 1, 60, 120, 4>(L, work, N2, nblk);
        fft_stage<4, 4, 15, 30, 16>(L, work, N2, nblk);
        fft_stage<3, 16, 5, 10, 48>(L, work, N2, nblk);
        fft_stage<5, 48, 1, 2, 1>(L, work, N2, nblk);
    } else if (shift == 2) {
        fft_stage<4, 1, 30, 120, 4>(L, work, N2, nblk);
        fft_stage<2, 4, 15, 0, 0>(L, work, N2, nblk);
        fft_stage<3, 8, 5, 20, 24>(L, work, N2, nblk);
        fft_stage<5, 24, 1, 4, 1>(L, work, N2, nblk);
    } else {
        fft_stage<4, 1, 15, 120, 4>(L, work, N2, nblk);
        fft_stage<3, 4, 5, 40, 12>(L, work, N2, nblk);
        fft_stage<5, 12, 1, 8, 1>(L, work, N2, nblk);
    }
    // post-rotation, in place from both ends (ref :3244-3277)
    const int half = (N4 + 1) >> 1;
    const SmallDiv byHalf(half);
    for (int t = L.lane; t < nblk * half; t += kLanes) {
        const int b = nblk > 1 ? byHalf.div(t) : 0, i = t - b * half;
        Cpx* yp0 = reinterpret_cast<Cpx*>(work + b * N2) + i;
        Cpx* yp1 = reinterpret_cast<Cpx*>(work + b * N2) + N4 - 1 - i;  // never the same element: N4 is even
        const Cpx a = *yp0, z = *yp1;  // .r = imaginary part, .i = real part (swapped, as above)
        s16 t0 = trig[i], t1 = trig[N4 + i];
        Cpx o0, o1;
        o0.r = add_w(smul(a.i, t0), smul(a.r, t1));
        o1.i = sub_w(smul(a.i, t1), smul(a.r, t0));
        t0 = trig[N4 - i - 1];
        t1 = trig[N2 - i - 1];
        o1.r = add_w(smul(z.i, t0), smul(z.r, t1));
        o0.i = sub_w(smul(z.i, t1), smul(z.r, t0));
        *yp0 = o0;
        *yp1 = o1;
    }
    L.sync();
}

// Signal buffers of the TDAC / post-filter stage are addressed through a view: a plain pointer, or a power-of-two ring
// (the post-filter kernel keeps each stream's last 1028 samples plus the current frame in a 2048-entry ring, so the
// history never has to be moved).
struct LinView {
    int32_t* p;
    OPUS_HD int32_t& operator[](int i) const { return p[i]; }
    OPUS_HD LinView operator+(int k) const { return LinView{p + k}; }
};
enum { kRingSize = 2048 };
struct RingView {
    int32_t* base;
    int pos;  // ring index of element 0
    OPUS_HD int32_t& operator[](int i) const { return base[(pos + i) & (kRingSize - 1)]; }
    OPUS_HD RingView operator+(int k) const { return RingView{base, pos + k}; }
};

// TDAC mirror + window for one frame (ref :3279-3295 applied block by block, then SATURATE :2121-2125).
// raw: this frame's nblk * N2 raw values; tail: previous block's last overlap/2 raw values (in/out).
// out[0..N) receives the frame's time-domain signal (before the post-filter).
template <class View>
OPUS_HD void tdac_frame(const Lanes& L, const int32_t* raw, int32_t* tail, View out, int N2, int nblk) {
    const int N = N2 * nblk;
    const SmallDiv byN2(N2);
    for (int n = L.lane; n < N; n += kLanes) {
        const int b = nblk > 1 ? byN2.div(n) : 0, k = n - b * N2;
        s32 v;
        if (k < kOverlap) {
            const int i = k < kOverlap / 2 ? k : kOverlap - 1 - k;
            s32 x2 = b == 0 ? tail[i] : raw[(b - 1) * N2 + N2 - kOverlap / 2 + i];  // previous block, un-windowed
            s32 x1 = raw[b * N2 + kOverlap / 2 - 1 - i];
            s16 w1 = kWindow120[i], w2 = kWindow120[kOverlap - 1 - i];
            if (k < kOverlap / 2) v = sub_w(mul16x32_q15(w2, x2), mul16x32_q15(w1, x1));
            else v = add_w(mul16x32_q15(w1, x2), mul16x32_q15(w2, x1));
        } else {
            v = raw[b * N2 + k - kOverlap / 2];
        }
        out[n] = sat_sym(v, 300000000);
    }
    L.sync();
    for (int i = L.lane; i < kOverlap / 2; i += kLanes) tail[i] = raw[(nblk - 1) * N2 + N2 - kOverlap / 2 + i];
    L.sync();
}

OPUS_LOCAL_TABLE s16 kCombGains[3][3] = {{10048, 7112, 4248}, {15200, 8784, 0}, {26208, 3280, 0}};  // ref gains[3][3], :854-857

// ---- comb post-filter (ref comb_filter :848, comb_filter_const_c :830), in place on a history buffer ----
// x points at the first sample to filter inside a buffer that holds >= T+2 valid earlier samples.
// In-place IIR: y[i] depends on y[i-T+2 .. i-T-2], T >= 15, so `chunk` = min(T0,T1)-2 >= 13 samples
// are independent and are computed by the lanes in parallel.
template <class View>
OPUS_HD void comb_filter_inplace(const Lanes& L, View x, int T0, int T1, int N, s16 g0, s16 g1, int tapset0, int tapset1) {
    if (g0 == 0 && g1 == 0) return;
    T0 = imax(T0, 15);
    T1 = imax(T1, 15);
    const s16 g00 = (s16)mul16_p15(g0, kCombGains[tapset0][0]), g01 = (s16)mul16_p15(g0, kCombGains[tapset0][1]),
              g02 = (s16)mul16_p15(g0, kCombGains[tapset0][2]);
    const s16 g10 = (s16)mul16_p15(g1, kCombGains[tapset1][0]), g11 = (s16)mul16_p15(g1, kCombGains[tapset1][1]),
              g12 = (s16)mul16_p15(g1, kCombGains[tapset1][2]);
    int overlap = kOverlap;
    if (g0 == g1 && T0 == T1 && tapset0 == tapset1) overlap = 0;
    if (overlap > N) overlap = N;
    const int chunk = imin(imin(T0, T1) - 2, kLanes);
    int i0 = 0;
    for (; i0 < overlap; i0 += chunk) {
        const int i = i0 + L.lane;
        if (L.lane < chunk && i < overlap) {
            s16 f = (s16)mul16_q15(kWindow120[i], kWindow120[i]);
            s16 nf = (s16)(32767 - f);
            s32 v = x[i]
                + mul16x32_q15(mul16_q15(nf, g00), x[i - T0])
                + mul16x32_q15(mul16_q15(nf, g01), x[i - T0 + 1] + x[i - T0 - 1])
                + mul16x32_q15(mul16_q15(nf, g02), x[i - T0 + 2] + x[i - T0 - 2])
                + mul16x32_q15(mul16_q15(f, g10), x[i - T1])
                + mul16x32_q15(mul16_q15(f, g11), x[i - T1 + 1] + x[i - T1 - 1])
                + mul16x32_q15(mul16_q15(f, g12), x[i - T1 + 2] + x[i - T1 - 2]);
            x[i] = sat_sym(v, 300000000);
        }
        L.sync();
    }
    if (g1 == 0) return;
    const int chunk1 = imin(T1 - 2, kLanes);
    for (i0 = overlap; i0 < N; i0 += chunk1) {
        const int i = i0 + L.lane;
        if (L.lane < chunk1 && i < N) {
            s32 v = x[i] + mul16x32_q15(g10, x[i - T1]) + mul16x32_q15(g11, x[i - T1 + 1] + x[i - T1 - 1])
                  + mul16x32_q15(g12, x[i - T1 + 2] + x[i - T1 - 2]);
            x[i] = sat_sym(v, 300000000);
        }
        L.sync();
    }
}

// ---- de-emphasis + int16 conversion, strictly serial per channel (ref deemphasis :1988, sig2word16 :413) ----
// The memory update MULT16_32_Q15(27853, tmp) = floor(27853 * tmp / 2^15) is written as tmp + floor(-4915 * tmp / 2^15)
// (27853 = 2^15 - 4915; exact for every int32 tmp): the second term is ONE multiply-high by -4915 << 17, and the next
// sample's tmp = x + tmp + hi is one three-input add behind it -- the dependent chain per sample is two instructions.
// The same without the final clamp (the caller saturates two samples at once): PSHR32(tmp, 12), ref sig2word16 :413.
OPUS_HD s32 deemph_round(s32 x, s32& mem) {
    s32 tmp = (s32)((u32)x + (u32)mem);
#if defined(__CUDA_ARCH__)
    mem = (s32)((u32)tmp + (u32)__mulhi(tmp, -(4915 << 17)));
#else
    mem = mul16x32_q15(27853, tmp);
#endif
    return pshr32(tmp, 12);
}
OPUS_HD s32 deemph_sample(s32 x, s32& mem) {
    s32 tmp = (s32)((u32)x + (u32)mem);
#if defined(__CUDA_ARCH__)
    mem = (s32)((u32)tmp + (u32)__mulhi(tmp, -(4915 << 17)));
#else
    mem = mul16x32_q15(27853, tmp);
#endif
    s32 v = pshr32(tmp, 12);
    return v > 32767 ? 32767 : (v < -32768 ? -32768 : v);
}

}  // namespace opusb200
