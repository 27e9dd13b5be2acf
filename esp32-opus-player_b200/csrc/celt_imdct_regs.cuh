// celt_imdct_regs.cuh -- the long-block inverse MDCT of a 20 ms frame (N = 1920: 480-point complex FFT) with the
// butterflies in registers (ref clt_mdct_backward src/celt.cpp:3204, opus_fft_impl :2997).
//
// 480 = 15 x 32.  The reference's stages are radix 4, 2, 4 (m = 1, 4, 8), which only ever combine points inside a run of
// 32 consecutive FFT positions, then radix 3 and 5 (m = 32, 96), which only ever combine the 15 positions p = 32 k + j
// of one residue j.  So:
//   sweep    all lanes: denormalise + pre-rotate (ref :948, :3219-3239) straight from the int16 spectrum into the
//            digit-reversed position of a buffer of 15 rows x 33 complex values (rows padded: conflict-free both ways)
//   pass A   lane b (15 per channel, both channels of a frame side by side = 30 lanes) owns row b = positions 32 b ..
//            32 b + 31: three stages on 32 complex registers; every index and every stage-3 twiddle is a compile-time constant
//   pass B   lane j (32 per channel) owns positions 32 k + j, k < 15: radix 3, radix 5, then the post-rotation
//            (ref :3244-3277), whose partner element 479 - p lives in lane 31 - j: one shuffle per element
// instead of five shared-memory read + write passes and a scatter into bit-reversed order.  Every S_MUL keeps its
// operand pairing and every butterfly its order of additions (the sums wrap, so that would not even matter).
#pragma once
#include "celt_synth.cuh"

namespace opusb200 {

#if defined(__CUDACC__)

enum { kImdctRow = 33, kImdctBuf = 992 };  // transpose row stride (complex values); words per channel buffer (>= 2 * 15 * 33)

OPUS_HD Cpx cmul_c(Cpx a, s16 br, s16 bi) {  // ref C_MUL (src/celt.h:193) with the twiddle given by value
    Cpx m;
    m.r = sub_w(smul(a.r, br), smul(a.i, bi));
    m.i = add_w(smul(a.r, bi), smul(a.i, br));
    return m;
}

// fft_twiddles48000_960[15 * j * k] (kFftTwRe / kFftTwIm), j < 8, k = 1..3: the radix-4 stage with m = 8 (ref kf_bfly4)
__device__ constexpr int16_t kTwS3Re[8][3] = {{32767, 32767, 32767}, {32138, 30274, 27246}, {30274, 23171, 12540}, {27246, 12540, -6393},
                                              {23171, 0, -23171},    {18205, -12540, -32138}, {12540, -23171, -30274}, {6393, -30274, -18205}};
__device__ constexpr int16_t kTwS3Im[8][3] = {{0, 0, 0}, {-6393, -12540, -18205}, {-12540, -23171, -30274}, {-18205, -30274, -32138},
                                              {-23171, -32767, -23171}, {-27246, -30274, -6393}, {-30274, -23171, 12540}, {-32138, -12540, 27246}};

// Gain of every band for denormalise_bands (ref src/celt.cpp:948), lane b < 21 -> band b: mantissa << 16 | shift & 0xff,
// 0 outside [start, end) (what `denormalise` in celt_synth.cuh keeps in its `gs` scratch).
__device__ __forceinline__ int32_t band_gain_packed(int band, int16_t logE, int start, int end) {
    if (band < start || band >= end) return 0;
    s16 lg = sat16((s32)logE + shl32((s32)kEMeans[band], 6));
    int shift = 16 - (lg >> 10);
    s16 g;
    if (shift > 31) { shift = 0; g = 0; }
    else g = (s16)exp2_frac((s16)(lg & 1023));
    if (shift <= -2) { g = 16384; shift = -2; }
    return (int32_t)(((u32)(uint16_t)g << 16) | (u32)(shift & 0xff));
}
__device__ __forceinline__ s32 denorm_bin(int16_t x, int32_t pk) {
    const s16 g = (s16)(pk >> 16);
    const int shift = (int)(int8_t)(pk & 0xff);
    const s32 v = mul16(x, g);
    return shift < 0 ? shl32(v, -shift) : v >> shift;
}

// Denormalisation and pre-rotation in one sweep, all 32 lanes (ref :948-1008 and :3219-3239): FFT input i needs the
// coefficients 2 i and 959 - 2 i, which are taken straight from the normalised spectrum Xs (int16, shared memory) and
// scaled with their bands' gains on the way; the pre-rotated value goes to its digit-reversed position p, row p / 32 of
// the transpose buffer T.  downmix: the frame codes two channels for a mono decoder, coefficient = (c0 >> 1) + (c1 >> 1).
__device__ __forceinline__ void imdct480_prerotate(int lane, const int16_t* Xs, const int32_t* gs, const int32_t* gs1, const int16_t* trig,
                                                   const int16_t* bitrev, Cpx* T, bool downmix) {
#pragma unroll 1
    for (int i = lane; i < 480; i += 32) {
        const int ja = 2 * i, jb = 959 - 2 * i;
        s32 x1 = denorm_bin(Xs[ja], gs[kBinBand[ja >> 3]]), x2 = denorm_bin(Xs[jb], gs[kBinBand[jb >> 3]]);
        if (downmix) {
            x1 = (x1 >> 1) + (denorm_bin(Xs[kMaxFrame + ja], gs1[kBinBand[ja >> 3]]) >> 1);
            x2 = (x2 >> 1) + (denorm_bin(Xs[kMaxFrame + jb], gs1[kBinBand[jb >> 3]]) >> 1);
        }
        const s16 t0 = trig[i], t1 = trig[480 + i];
        const s32 yr = add_w(smul(x2, t0), smul(x1, t1));
        const s32 yi = sub_w(smul(x1, t0), smul(x2, t1));
        const int p = bitrev[i];
        Cpx y;
        y.r = yi; y.i = yr;  // real and imaginary parts swapped (ref :3236-3237)
        T[p + (p >> 5)] = y;  // row stride kImdctRow = 33
    }
}

// Pass A for one lane: the 32 FFT positions of one block (a row of T) through the first three stages, in registers, and
// back to the same row.
__device__ __forceinline__ void imdct480_pass_a(Cpx* T_row) {
    Cpx v[32];
#pragma unroll
    for (int q = 0; q < 32; q++) v[q] = T_row[q];
#pragma unroll
    for (int a = 0; a < 8; a++) {  // radix 4, m = 1 (ref kf_bfly4 :2830, the twiddle-free case)
        Cpx a0 = v[4 * a], a1 = v[4 * a + 1], a2 = v[4 * a + 2], a3 = v[4 * a + 3];
        Cpx s0 = csub(a0, a2);
        a0 = cadd(a0, a2);
        Cpx s1 = cadd(a1, a3);
        v[4 * a + 2] = csub(a0, s1);
        v[4 * a] = cadd(a0, s1);
        s1 = csub(a1, a3);
        v[4 * a + 1].r = add_w(s0.r, s1.i); v[4 * a + 1].i = sub_w(s0.i, s1.r);
        v[4 * a + 3].r = sub_w(s0.r, s1.i); v[4 * a + 3].i = add_w(s0.i, s1.r);
    }
#pragma unroll
    for (int a = 0; a < 4; a++) {  // radix 2, m = 4 (ref kf_bfly2 :2794)
        const s16 tw = 23170;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const Cpx x = v[8 * a + k], b = v[8 * a + k + 4];
            Cpx tt;
            if (k == 0) tt = b;
            else if (k == 1) { tt.r = smul(add_w(b.r, b.i), tw); tt.i = smul(sub_w(b.i, b.r), tw); }
            else if (k == 2) { tt.r = b.i; tt.i = neg_w(b.r); }
            else { tt.r = smul(sub_w(b.i, b.r), tw); tt.i = smul(neg_w(add_w(b.i, b.r)), tw); }
            v[8 * a + k + 4] = csub(x, tt);
            v[8 * a + k] = cadd(x, tt);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {  // radix 4, m = 8, twiddle stride 15 (ref kf_bfly4 :2830)
        const Cpx c0 = cmul_c(v[j + 8], kTwS3Re[j][0], kTwS3Im[j][0]);
        const Cpx c1 = cmul_c(v[j + 16], kTwS3Re[j][1], kTwS3Im[j][1]);
        const Cpx c2 = cmul_c(v[j + 24], kTwS3Re[j][2], kTwS3Im[j][2]);
        Cpx a0 = v[j];
        const Cpx c5 = csub(a0, c1);
        a0 = cadd(a0, c1);
        const Cpx c3 = cadd(c0, c2);
        const Cpx c4 = csub(c0, c2);
        v[j + 16] = csub(a0, c3);
        v[j] = cadd(a0, c3);
        v[j + 8].r = add_w(c5.r, c4.i); v[j + 8].i = sub_w(c5.i, c4.r);
        v[j + 24].r = sub_w(c5.r, c4.i); v[j + 24].i = add_w(c5.i, c4.r);
    }
#pragma unroll
    for (int q = 0; q < 32; q++) T_row[q] = v[q];
}

// Pass B for one lane j: positions 32 k + j (column j of T) through radix 3 and radix 5, post-rotation, and the raw
// output y[2 p], y[2 p + 1] of every position to `out` (global memory, 960 values; 8-byte stores, coalesced).
__device__ __forceinline__ void imdct480_pass_b(const Cpx* T, const int16_t* trig, const int16_t* twr, const int16_t* twi, int j, int32_t* out) {
    Cpx u[15];
#pragma unroll
    for (int k = 0; k < 15; k++) u[k] = T[kImdctRow * k + j];
    {  // radix 3, m = 32, twiddle stride 5 (ref kf_bfly3 :2887): block i holds k = 3 i, 3 i + 1, 3 i + 2
        const s16 r1 = twr[5 * j], i1 = twi[5 * j], r2 = twr[10 * j], i2 = twi[10 * j];
#pragma unroll
        for (int i = 0; i < 5; i++) {
            const Cpx c1 = cmul_c(u[3 * i + 1], r1, i1);
            const Cpx c2 = cmul_c(u[3 * i + 2], r2, i2);
            const Cpx c3 = cadd(c1, c2);
            Cpx c0 = csub(c1, c2);
            const Cpx a0 = u[3 * i];
            Cpx fm;
            fm.r = sub_w(a0.r, c3.r >> 1);
            fm.i = sub_w(a0.i, c3.i >> 1);
            c0.r = smul(c0.r, -28378);
            c0.i = smul(c0.i, -28378);
            u[3 * i] = cadd(a0, c3);
            u[3 * i + 2].r = add_w(fm.r, c0.i); u[3 * i + 2].i = sub_w(fm.i, c0.r);
            u[3 * i + 1].r = sub_w(fm.r, c0.i); u[3 * i + 1].i = add_w(fm.i, c0.r);
        }
    }
    {  // radix 5, m = 96, twiddle stride 1 (ref kf_bfly5 :2930): butterfly u' = 32 k2 + j takes k = k2 + 3 t, t < 5
        const s16 yar = 10126, yai = -31164, ybr = -26510, ybi = -19261;
#pragma unroll
        for (int k2 = 0; k2 < 3; k2++) {
            const int w = 32 * k2 + j;
            const Cpx s0 = u[k2];
            const Cpx s1 = cmul_c(u[k2 + 3], twr[w], twi[w]);
            const Cpx s2 = cmul_c(u[k2 + 6], twr[2 * w], twi[2 * w]);
            const Cpx s3 = cmul_c(u[k2 + 9], twr[3 * w], twi[3 * w]);
            const Cpx s4 = cmul_c(u[k2 + 12], twr[4 * w], twi[4 * w]);
            const Cpx s7 = cadd(s1, s4), s10 = csub(s1, s4), s8 = cadd(s2, s3), s9 = csub(s2, s3);
            u[k2].r = add_w(s0.r, add_w(s7.r, s8.r));
            u[k2].i = add_w(s0.i, add_w(s7.i, s8.i));
            Cpx s5, s6, s11, s12;
            s5.r = add_w(s0.r, add_w(smul(s7.r, yar), smul(s8.r, ybr)));
            s5.i = add_w(s0.i, add_w(smul(s7.i, yar), smul(s8.i, ybr)));
            s6.r = add_w(smul(s10.i, yai), smul(s9.i, ybi));
            s6.i = neg_w(add_w(smul(s10.r, yai), smul(s9.r, ybi)));
            u[k2 + 3] = csub(s5, s6);
            u[k2 + 12] = cadd(s5, s6);
            s11.r = add_w(s0.r, add_w(smul(s7.r, ybr), smul(s8.r, yar)));
            s11.i = add_w(s0.i, add_w(smul(s7.i, ybr), smul(s8.i, yar)));
            s12.r = sub_w(smul(s9.i, yai), smul(s10.i, ybi));
            s12.i = sub_w(smul(s10.r, ybi), smul(s9.r, yai));
            u[k2 + 6] = cadd(s11, s12);
            u[k2 + 9] = csub(s11, s12);
        }
    }
    // Post-rotation (ref :3244-3277: in place from both ends, element i with element 479 - i).  Written per element it
    // is one rule: with e the FFT output at position p (.r the imaginary part, .i the real part, as above),
    //   y[2 p]       = e.i * trig[p] + e.r * trig[480 + p]        stays with p
    //   y[2 p' + 1]  = e.i * trig[480 + p] - e.r * trig[p]        goes to the partner p' = 479 - p, lane 31 - j, k' = 14 - k
    s32 mine[15], give[15];
#pragma unroll
    for (int k = 0; k < 15; k++) {
        const int p = 32 * k + j;
        const s16 t0 = trig[p], t1 = trig[480 + p];
        mine[k] = add_w(smul(u[k].i, t0), smul(u[k].r, t1));
        give[k] = sub_w(smul(u[k].i, t1), smul(u[k].r, t0));
    }
#pragma unroll
    for (int k = 0; k < 15; k++) {
        const s32 got = __shfl_sync(0xffffffffu, give[14 - k], 31 - j);
        *reinterpret_cast<int2*>(out + 2 * (32 * k + j)) = make_int2(mine[k], got);
    }
}

#endif  // __CUDACC__

}  // namespace opusb200
