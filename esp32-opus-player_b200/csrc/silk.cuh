// silk.cuh -- SILK layer of the decode path, 20 ms frames, 8/12/16 kHz: the entropy
// half, the per-stream state, and the parameter arithmetic.
//
//   silk_parse_frame   entropy half, one thread per (stream, frame): VAD/LBRR flags, stereo predictor,
//                      side-information indices, shell-coded pulses (ref src/silk.cpp:1565-1636, 708, 898)
//   silk_gains_dequant, silk_nlsf_decode, silk_nlsf2a ...   building blocks of the synthesis stages, which live in
//                      silk_stages.cuh (stage bodies) and silk_kernels.cuh (kernels)
//
// Parsing needs no decoder state because every packet carries one 20 ms frame, which the reference
// always codes independently (FrameIndex <= 0, src/silk.cpp:1678-1681).
// PLC/CNG bookkeeping (src/silk.cpp:1342, 2895) only feeds the loss-concealment code, which the reference
// never reaches; it is not reproduced.
#pragma once
#include "celt_script.cuh"
#include "rangedec.cuh"
#include "silk_tables.cuh"

namespace opusb200 {

// ---- fixed-point helpers with the reference macro each one reproduces (src/silk.h:52-156, 446-486) ----
OPUS_HD s32 smulbb(s32 a, s32 b) { return (s32)(s16)a * (s32)(s16)b; }                       // silk_SMULBB
OPUS_HD s32 smlabb(s32 a, s32 b, s32 c) { return (s32)((u32)a + (u32)smulbb(b, c)); }          // silk_SMLABB (/_ovflw)
// silk_SMULWB / silk_SMLAWB: floor(a * (int16)b / 2^16) is the high word of a * ((int16)b << 16), exactly, for every
// int32 a -- one multiply-high on the device (and with a constant b the shift folds away) instead of a 64-bit product
// and a funnel shift; the accumulate wraps like the reference's 32-bit add.
#if defined(__CUDA_ARCH__)
OPUS_HD s32 smulwb(s32 a, s32 b) { return __mulhi(a, (s32)((u32)(uint16_t)b << 16)); }
OPUS_HD s32 smlawb(s32 a, s32 b, s32 c) { return (s32)((u32)a + (u32)__mulhi(b, (s32)((u32)(uint16_t)c << 16))); }
#else
OPUS_HD s32 smulwb(s32 a, s32 b) { return (s32)(((s64)a * (s64)(s16)b) >> 16); }              // silk_SMULWB
OPUS_HD s32 smlawb(s32 a, s32 b, s32 c) { return (s32)((s64)a + (((s64)b * (s64)(s16)c) >> 16)); }  // silk_SMLAWB
#endif
OPUS_HD s32 smulww(s32 a, s32 b) { return (s32)(((s64)a * (s64)b) >> 16); }                   // silk_SMULWW
OPUS_HD s32 smlaww(s32 a, s32 b, s32 c) { return (s32)((s64)a + (((s64)b * (s64)c) >> 16)); } // silk_SMLAWW
OPUS_HD s32 smmul(s32 a, s32 b) { return (s32)(((s64)a * (s64)b) >> 32); }                    // silk_SMMUL
OPUS_HD s32 mul_w(s32 a, s32 b) { return (s32)((u32)a * (u32)b); }                            // silk_MUL (two's complement)
OPUS_HD s32 rshift_round(s32 a, int s) { return s == 1 ? (a >> 1) + (a & 1) : ((a >> (s - 1)) + 1) >> 1; }  // silk_RSHIFT_ROUND
// silk_RSHIFT_ROUND64 for |a| < 2^62 (every caller passes a product of two int32): one add and one shift
OPUS_HD s64 rshift_round64(s64 a, int s) { return (a + ((s64)1 << (s - 1))) >> s; }
OPUS_HD s32 limit32(s32 a, s32 lo, s32 hi) {                                                  // silk_LIMIT
    return lo > hi ? (a > lo ? lo : (a < hi ? hi : a)) : (a > hi ? hi : (a < lo ? lo : a));
}
OPUS_HD s32 add_sat32(s32 a, s32 b) {                                                         // silk_ADD_SAT32
#if defined(__CUDA_ARCH__)
    s32 r;
    asm("add.sat.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    s64 r = (s64)a + (s64)b;
    return r > 2147483647LL ? 2147483647 : (r < -2147483648LL ? (s32)(-2147483647 - 1) : (s32)r);
#endif
}
OPUS_HD s32 sub_sat32(s32 a, s32 b) {                                                         // silk_SUB_SAT32
#if defined(__CUDA_ARCH__)
    s32 r;
    asm("sub.sat.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    s64 r = (s64)a - (s64)b;
    return r > 2147483647LL ? 2147483647 : (r < -2147483648LL ? (s32)(-2147483647 - 1) : (s32)r);
#endif
}
OPUS_HD s32 lshift_sat32(s32 a, int s) {                                                      // silk_LSHIFT_SAT32
    return shl32(limit32(a, (s32)(-2147483647 - 1) >> s, 2147483647 >> s), s);
}
OPUS_HD int clz32(s32 x) { return x ? 32 - ilog32((u32)x) : 32; }                            // silk_CLZ32
OPUS_HD s32 iabs(s32 a) { return a > 0 ? a : (s32)(0u - (u32)a); }

// ref silk_INVERSE32_varQ / silk_DIV32_varQ (src/silk.h:913-996)
OPUS_HD s32 inverse32_varq(s32 b32, int Qres) {
    int b_headrm = clz32(iabs(b32)) - 1;
    s32 b32_nrm = shl32(b32, b_headrm);
    s32 b32_inv = (2147483647 >> 2) / (b32_nrm >> 16);
    s32 result = shl32(b32_inv, 16);
    s32 err_Q32 = shl32(((s32)1 << 29) - smulwb(b32_nrm, b32_inv), 3);
    result = smlaww(result, err_Q32, b32_inv);
    int lshift = 61 - b_headrm - Qres;
    if (lshift <= 0) return lshift_sat32(result, -lshift);
    return lshift < 32 ? result >> lshift : 0;
}
OPUS_HD s32 div32_varq(s32 a32, s32 b32, int Qres) {
    int a_headrm = clz32(iabs(a32)) - 1;
    s32 a32_nrm = shl32(a32, a_headrm);
    int b_headrm = clz32(iabs(b32)) - 1;
    s32 b32_nrm = shl32(b32, b_headrm);
    s32 b32_inv = (2147483647 >> 2) / (b32_nrm >> 16);
    s32 result = smulwb(a32_nrm, b32_inv);
    a32_nrm = sub_w(a32_nrm, shl32(smmul(b32_nrm, result), 3));
    result = smlawb(result, a32_nrm, b32_inv);
    int lshift = 29 + a_headrm - b_headrm - Qres;
    if (lshift < 0) return lshift_sat32(result, -lshift);
    return lshift < 32 ? result >> lshift : 0;
}
// ref silk_log2lin (src/silk.cpp:2248)
OPUS_HD s32 log2lin(s32 inLog_Q7) {
    if (inLog_Q7 < 0) return 0;
    if (inLog_Q7 >= 3967) return 2147483647;
    s32 out = shl32(1, inLog_Q7 >> 7);
    s32 frac_Q7 = inLog_Q7 & 0x7F;
    if (inLog_Q7 < 2048) out = out + (mul_w(out, smlawb(frac_Q7, smulbb(frac_Q7, 128 - frac_Q7), -174)) >> 7);
    else out = out + mul_w(out >> 7, smlawb(frac_Q7, smulbb(frac_Q7, 128 - frac_Q7), -174));
    return out;
}

// ------------------------------------------------------------------------------------------------
// Script of one SILK frame (entropy -> synthesis hand-off)
// ------------------------------------------------------------------------------------------------
enum { kSilkMaxFrame = 320, kSilkMaxOrder = 16 };

struct SilkIdx {  // ref SideInfoIndices_t (src/silk.h)
    int8_t gains[4];
    int8_t ltp[4];
    int8_t nlsf[kSilkMaxOrder + 1];
    int16_t lagIndex;
    int8_t contour, signalType, quantOffsetType, interpQ2, perIndex, ltpScaleIndex, seed;
};
struct alignas(16) SilkFrame {
    int16_t pulses[2][kSilkMaxFrame];  // first: every channel's row starts on a 16-byte boundary
    int32_t ms_pred_Q13[2];
    uint8_t fs_kHz, nch, mid_only, present;
    SilkIdx idx[2];
};

struct SilkCb {  // ref silk_NLSF_CB_struct
    int order, quantStepSize_Q16;
    const uint8_t *cb1_Q8, *cb1_icdf, *pred_Q8, *ec_sel, *ec_icdf;
    const int16_t *cb1_wght_Q9, *deltaMin_Q15;
};
OPUS_HD SilkCb silk_cb(int fs_kHz) {
    SilkCb c;
    if (fs_kHz == 16) {
        c.order = 16; c.quantStepSize_Q16 = 9830;
        c.cb1_Q8 = kNlsfCb1WbQ8; c.cb1_icdf = kNlsfCb1WbIcdf; c.pred_Q8 = kNlsfPredWb; c.ec_sel = kNlsfSelWb; c.ec_icdf = kNlsfCb2WbIcdf;
        c.cb1_wght_Q9 = kNlsfCb1WbWght; c.deltaMin_Q15 = kNlsfDeltaMinWb;
    } else {
        c.order = 10; c.quantStepSize_Q16 = 11796;
        c.cb1_Q8 = kNlsfCb1NbQ8; c.cb1_icdf = kNlsfCb1NbIcdf; c.pred_Q8 = kNlsfPredNb; c.ec_sel = kNlsfSelNb; c.ec_icdf = kNlsfCb2NbIcdf;
        c.cb1_wght_Q9 = kNlsfCb1NbWght; c.deltaMin_Q15 = kNlsfDeltaMinNb;
    }
    return c;
}
// ref silk_NLSF_unpack (src/silk.cpp:2762)
OPUS_HD void nlsf_unpack(int16_t* ec_ix, uint8_t* pred_Q8, const SilkCb& cb, int CB1_index) {
    const uint8_t* sel = cb.ec_sel + CB1_index * cb.order / 2;
    for (int i = 0; i < cb.order; i += 2) {
        uint8_t entry = *sel++;
        ec_ix[i] = (int16_t)smulbb((entry >> 1) & 7, 9);
        pred_Q8[i] = cb.pred_Q8[i + (entry & 1) * (cb.order - 1)];
        ec_ix[i + 1] = (int16_t)smulbb((entry >> 5) & 7, 9);
        pred_Q8[i + 1] = cb.pred_Q8[i + ((entry >> 4) & 1) * (cb.order - 1) + 1];
    }
}

// ---- entropy half ----
// The lanes of a warp that parse a SILK layer go through the same sequence of loops (side information, 20 shell
// blocks, signs); `m` names them and warp_converge(m) makes them meet at every loop head, `on` says whether this lane
// really has the unit in its bitstream (LBRR copies and the side channel are optional).  Without the meeting points the
// lanes drift apart for good, as they did in the CELT parser.
enum { kSilkMaxBlocks = kSilkMaxFrame / 16 };

// uniform and small iCDF tables of the side information (ref src/silk.cpp:37-142); at namespace scope, see rangedec.cuh
OPUS_LOCAL_TABLE uint8_t kUni3[3] = {171, 85, 0}, kUni4[4] = {192, 128, 64, 0}, kUni5[5] = {205, 154, 102, 51, 0},
                         kUni6[6] = {213, 171, 128, 85, 43, 0}, kUni8[8] = {224, 192, 160, 128, 96, 64, 32, 0};
OPUS_LOCAL_TABLE uint8_t kNlsfExtIcdf[7] = {100, 40, 16, 7, 3, 1, 0};
OPUS_LOCAL_TABLE uint8_t kTypeVadIcdf[4] = {232, 158, 10, 0};
OPUS_LOCAL_TABLE uint8_t kNlsfInterpIcdf[5] = {243, 221, 192, 181, 0};
OPUS_LOCAL_TABLE uint8_t kLtpPerIcdf[3] = {179, 99, 0};
OPUS_LOCAL_TABLE uint8_t kLtpScaleIcdf[3] = {128, 64, 0};
OPUS_LOCAL_TABLE uint8_t kNlsf2aOrder16[16] = {0, 15, 8, 7, 4, 11, 12, 3, 2, 13, 10, 5, 6, 9, 14, 1};
OPUS_LOCAL_TABLE uint8_t kNlsf2aOrder10[10] = {0, 9, 6, 3, 4, 5, 8, 1, 2, 7};

// ref silk_decode_indices (src/silk.cpp:708) with condCoding == CODE_INDEPENDENTLY
OPUS_HD void silk_parse_indices(RangeDec& rd, SilkIdx& ix, int fs_kHz, int vad_or_lbrr, bool on, unsigned m) {
    const SilkCb cb = silk_cb(fs_kHz);
    int16_t ec_ix[kSilkMaxOrder];
    uint8_t pred_Q8[kSilkMaxOrder];
    int Ix;
    if (on) {
        if (vad_or_lbrr) {
            Ix = rd_icdf(rd, kTypeVadIcdf, 8) + 2;
        } else {
            Ix = rd_icdf2(rd, 230, 8);
        }
        ix.signalType = (int8_t)(Ix >> 1);
        ix.quantOffsetType = (int8_t)(Ix & 1);
        ix.gains[0] = (int8_t)(rd_icdf_tab(rd, kGainIcdf + ix.signalType * 8, 8) << 3);
        ix.gains[0] += (int8_t)rd_icdf(rd, kUni8, 8);
        for (int i = 1; i < 4; i++) ix.gains[i] = (int8_t)rd_icdf_tab(rd, kDeltaGainIcdf, 8);
        ix.nlsf[0] = (int8_t)rd_icdf_tab(rd, cb.cb1_icdf + (ix.signalType >> 1) * 32, 8);
        nlsf_unpack(ec_ix, pred_Q8, cb, ix.nlsf[0]);
    }
#pragma unroll 1
    for (int i = 0; i < kSilkMaxOrder; i++) {
        warp_converge(m);
        if (on && i < cb.order) {
            Ix = rd_icdf_tab(rd, cb.ec_icdf + ec_ix[i], 8);
            if (Ix == 0) Ix -= rd_icdf(rd, kNlsfExtIcdf, 8);
            else if (Ix == 8) Ix += rd_icdf(rd, kNlsfExtIcdf, 8);
            ix.nlsf[i + 1] = (int8_t)(Ix - 4);
        }
    }
    warp_converge(m);
    if (on) {
        ix.interpQ2 = (int8_t)rd_icdf(rd, kNlsfInterpIcdf, 8);
        ix.lagIndex = 0; ix.contour = 0; ix.perIndex = 0; ix.ltpScaleIndex = 0;
        for (int k = 0; k < 4; k++) ix.ltp[k] = 0;
        if (ix.signalType == 2) {
            ix.lagIndex = (int16_t)((int16_t)rd_icdf_tab(rd, kPitchLagIcdf, 8) * (fs_kHz >> 1));
            ix.lagIndex += (int16_t)rd_icdf(rd, fs_kHz == 16 ? kUni8 : fs_kHz == 12 ? kUni6 : kUni4, 8);
            ix.contour = (int8_t)rd_icdf_tab(rd, fs_kHz == 8 ? kPitchContourNbIcdf : kPitchContourIcdf, 8);
            ix.perIndex = (int8_t)rd_icdf(rd, kLtpPerIcdf, 8);
            const uint8_t* g = ix.perIndex == 0 ? kLtpGainIcdf0 : ix.perIndex == 1 ? kLtpGainIcdf1 : kLtpGainIcdf2;
            for (int k = 0; k < 4; k++) ix.ltp[k] = (int8_t)rd_icdf_tab(rd, g, 8);
            ix.ltpScaleIndex = (int8_t)rd_icdf(rd, kLtpScaleIcdf, 8);
        }
        ix.seed = (int8_t)rd_icdf(rd, kUni4, 8);
    }
}

// ref silk_shell_decoder / decode_split (src/silk.cpp:1146, 1162)
OPUS_HD void shell_split(RangeDec& rd, int16_t* c1, int16_t* c2, int p, const uint8_t* table, bool on, unsigned m) {
    if (!on) return;
    if (p > 0) {
        *c1 = (int16_t)rd_icdf_tab(rd, table + kShellOffsets[p], 8);
        *c2 = (int16_t)(p - *c1);
    } else {
        *c1 = 0;
        *c2 = 0;
    }
}
// every lane of `m` walks the 15 splits of the block together; `on`: this lane has pulses to split in this block
OPUS_HD void shell_decode(RangeDec& rd, int16_t* p0, int pulses4, bool on, unsigned m) {
    int16_t p3[2], p2[4], p1[8];
    warp_converge(m);
    shell_split(rd, &p3[0], &p3[1], pulses4, kShell3, on, m);
#pragma unroll 1
    for (int a = 0; a < 2; a++) {
        shell_split(rd, &p2[2 * a], &p2[2 * a + 1], p3[a], kShell2, on, m);
#pragma unroll 1
        for (int b = 0; b < 2; b++) {
            shell_split(rd, &p1[4 * a + 2 * b], &p1[4 * a + 2 * b + 1], p2[2 * a + b], kShell1, on, m);
#pragma unroll 1
            for (int c = 0; c < 2; c++)
                shell_split(rd, &p0[8 * a + 4 * b + 2 * c], &p0[8 * a + 4 * b + 2 * c + 1], p1[4 * a + 2 * b + c], kShell0, on, m);
        }
    }
}

// ref silk_decode_pulses + silk_decode_signs (src/silk.cpp:898, 1436)
OPUS_HD void silk_parse_pulses(RangeDec& rd, int16_t* pulses, int signalType, int quantOffsetType, int frame_length, bool on, unsigned m) {
    int sum_pulses[kSilkMaxBlocks], nLshifts[kSilkMaxBlocks];
    int iter = frame_length >> 4;
    if (iter * 16 < frame_length) iter++;
    const uint8_t* cdf = kPulsesPerBlockIcdf;
    if (on) cdf += 18 * rd_icdf_tab(rd, kRateLevelsIcdf + (signalType >> 1) * 9, 8);
#pragma unroll 1
    for (int i = 0; i < kSilkMaxBlocks; i++) {
        warp_converge(m);
        if (on && i < iter) {
            nLshifts[i] = 0;
            sum_pulses[i] = rd_icdf_tab(rd, cdf, 8);
#pragma unroll 1
            while (sum_pulses[i] == 17) {
                nLshifts[i]++;
                sum_pulses[i] = rd_icdf_tab(rd, kPulsesPerBlockIcdf + 9 * 18 + (nLshifts[i] == 10), 8);
            }
        }
    }
#pragma unroll 1
    for (int i = 0; i < kSilkMaxBlocks; i++) {
        const bool blk = on && i < iter;
        const bool coded = blk && sum_pulses[i] > 0;
        alignas(16) int16_t p0[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        shell_decode(rd, p0, coded ? sum_pulses[i] : 0, coded, m);
        if (blk) ld_block16(pulses + 16 * i, p0);
    }
    if (on) {
#pragma unroll 1
        for (int i = 0; i < iter; i++) {
            if (nLshifts[i] > 0) {
                const int nLS = nLshifts[i];
                int16_t* pp = pulses + 16 * i;
#pragma unroll 1
                for (int k = 0; k < 16; k++) {
                    s32 abs_q = pp[k];
#pragma unroll 1
                    for (int j = 0; j < nLS; j++) {
                        abs_q = shl32(abs_q, 1);
                        abs_q += rd_icdf2(rd, 120, 8);
                    }
                    pp[k] = (int16_t)abs_q;
                }
                sum_pulses[i] |= nLS << 5;
            }
        }
    }
    const uint8_t* icdf_ptr = kSignIcdf + smulbb(7, quantOffsetType + (signalType << 1));
    const int nblk = (frame_length + 8) >> 4;
#pragma unroll 1
    for (int i = 0; i < kSilkMaxBlocks; i++) {
        warp_converge(m);
        if (on && i < nblk && sum_pulses[i] > 0) {
            const u32 p0 = icdf_ptr[imin(sum_pulses[i] & 0x1F, 6)];
            // the block's 16 magnitudes come in and go out as two 16-byte words
            alignas(16) int16_t q[16];
            ld_block16(q, pulses + 16 * i);
#pragma unroll
            for (int j = 0; j < 16; j++)
                if (q[j] > 0) q[j] = (int16_t)(q[j] * ((rd_icdf2(rd, p0, 8) << 1) - 1));
            ld_block16(pulses + 16 * i, q);
        }
    }
}

// ref silk_stereo_decode_pred (src/silk.cpp:592)
OPUS_HD void silk_parse_stereo_pred(RangeDec& rd, int32_t* pred_Q13) {
    int ix[2][3];
    int n = rd_icdf_tab(rd, kStereoJointIcdf, 8);
    ix[0][2] = n / 5;
    ix[1][2] = n - 5 * ix[0][2];
    for (n = 0; n < 2; n++) {
        ix[n][0] = rd_icdf(rd, kUni3, 8);
        ix[n][1] = rd_icdf(rd, kUni5, 8);
    }
    for (n = 0; n < 2; n++) {
        ix[n][0] += 3 * ix[n][2];
        s32 low_Q13 = kStereoPredQ13[ix[n][0]];
        s32 step_Q13 = smulwb(kStereoPredQ13[ix[n][0] + 1] - low_Q13, 6554);
        pred_Q13[n] = smlabb(low_Q13, step_Q13, 2 * ix[n][1] + 1);
    }
    pred_Q13[0] -= pred_Q13[1];
}

// The SILK part of one Opus frame as silk_Decode reads it (ref src/silk.cpp:1565-1697, lostFlag = 0, one
// 20 ms frame per packet).  Every lane named in `m` calls this; the four possible (side information + pulses) units of a
// frame -- the LBRR copy of either channel, which regular decoding skips by parsing it (ref :1590-1616), then the
// channels themselves -- go through one call site, in lock step.
OPUS_HD void silk_parse_frame(RangeDec& rd, SilkFrame& f, int fs_kHz, int nch, unsigned m) {
    const int frame_length = 20 * fs_kHz;
    int vad[2] = {0, 0}, lbrr[2] = {0, 0};
    f.fs_kHz = (uint8_t)fs_kHz; f.nch = (uint8_t)nch; f.mid_only = 0; f.present = 1;
    f.ms_pred_Q13[0] = f.ms_pred_Q13[1] = 0;
    for (int n = 0; n < nch; n++) {
        vad[n] = rd_bit_logp(rd, 1);
        lbrr[n] = rd_bit_logp(rd, 1);
    }
    int mid_only = 0;
#pragma unroll 1
    for (int slot = 0; slot < 4; slot++) {
        warp_converge(m);
        const int n = slot & 1, is_lbrr = slot < 2;
        if (n == 0 && nch == 2 && (!is_lbrr || lbrr[0])) {  // the stereo predictor (and mid-only flag) precede channel 0
            int32_t tmp[2];
            silk_parse_stereo_pred(rd, is_lbrr ? tmp : f.ms_pred_Q13);
            if ((is_lbrr ? lbrr[1] : vad[1]) == 0) {
                const int flag = rd_icdf2(rd, 64, 8);
                if (!is_lbrr) mid_only = flag;
            }
        }
        const bool on = n < nch && (is_lbrr ? lbrr[n] != 0 : (n == 0 || !mid_only));
        silk_parse_indices(rd, f.idx[n], fs_kHz, is_lbrr ? 1 : vad[n], on, m);
        silk_parse_pulses(rd, f.pulses[n], f.idx[n].signalType, f.idx[n].quantOffsetType, frame_length, on, m);
    }
    f.mid_only = (uint8_t)mid_only;
}

// ------------------------------------------------------------------------------------------------
// Synthesis half
// ------------------------------------------------------------------------------------------------
struct SilkResampler {  // ref silk_resampler_state_struct_t (src/silk.h:654-670), up-sampling to 48 kHz only
    int32_t sIIR[6];
    int16_t sFIR[8];
    int16_t delayBuf[48];
    int32_t inputDelay, invRatio_Q16, batchSize, Fs_in_kHz;
};
struct SilkChannel {    // ref silk_decoder_state_t (src/silk.h:705-741), live fields only
    int32_t prev_gain_Q16;
    int32_t sLPC_Q14_buf[kSilkMaxOrder];
    int16_t outBuf[480];
    int32_t lagPrev;
    int32_t fs_kHz, fs_API_hz, frame_length, subfr_length, ltp_mem_length, LPC_order;
    int16_t prevNLSF_Q15[kSilkMaxOrder];
    int8_t LastGainIndex, prevSignalType, first_frame_after_reset, pad;
    SilkResampler rs;
};
struct SilkStreamState {
    SilkChannel ch[2];
    int16_t pred_prev_Q13[2], sMid[2], sSide[2];  // ref stereo_dec_state (src/silk.h:672-676)
    int32_t prev_decode_only_middle;
    int32_t nChannelsInternal;                    // ref silk_decoder_t
    int32_t prev_mode;                            // Opus layer: 0 none, 1000 SILK, 1001 hybrid, 1002 CELT
};
// ref silk_init_decoder (src/silk.cpp:2192)
OPUS_HD void silk_channel_init(SilkChannel& c) {
    c.prev_gain_Q16 = 65536;
    for (int i = 0; i < kSilkMaxOrder; i++) { c.sLPC_Q14_buf[i] = 0; c.prevNLSF_Q15[i] = 0; }
    for (int i = 0; i < 480; i++) c.outBuf[i] = 0;
    c.lagPrev = 0; c.fs_kHz = 0; c.fs_API_hz = 0; c.frame_length = 0; c.subfr_length = 0; c.ltp_mem_length = 0; c.LPC_order = 0;
    c.LastGainIndex = 0; c.prevSignalType = 0; c.first_frame_after_reset = 1; c.pad = 0;
    for (int i = 0; i < 6; i++) c.rs.sIIR[i] = 0;
    for (int i = 0; i < 8; i++) c.rs.sFIR[i] = 0;
    for (int i = 0; i < 48; i++) c.rs.delayBuf[i] = 0;
    c.rs.inputDelay = 0; c.rs.invRatio_Q16 = 0; c.rs.batchSize = 0; c.rs.Fs_in_kHz = 0;
}
// ref silk_InitDecoder (src/silk.cpp:1792); the resampler singletons are not touched by it
OPUS_HD void silk_stream_reinit(SilkStreamState& s) {
    for (int n = 0; n < 2; n++) {
        SilkResampler keep = s.ch[n].rs;
        silk_channel_init(s.ch[n]);
        s.ch[n].rs = keep;
    }
    for (int i = 0; i < 2; i++) { s.pred_prev_Q13[i] = 0; s.sMid[i] = 0; s.sSide[i] = 0; }
    s.prev_decode_only_middle = 0;
}

// ref silk_gains_dequant (src/silk.cpp:2148), conditional = 0
OPUS_HD void silk_gains_dequant(s32* gain_Q16, const int8_t* ind, int8_t* prev_ind) {
    for (int k = 0; k < 4; k++) {
        if (k == 0) {
            *prev_ind = (int8_t)imax(ind[k], *prev_ind - 16);
        } else {
            int ind_tmp = ind[k] + (-4);
            int thr = 2 * 36 - 64 + *prev_ind;
            if (ind_tmp > thr) *prev_ind = (int8_t)(*prev_ind + (shl32(ind_tmp, 1) - thr));
            else *prev_ind = (int8_t)(*prev_ind + ind_tmp);
        }
        *prev_ind = (int8_t)limit32(*prev_ind, 0, 63);
        gain_Q16[k] = log2lin(imin(smulwb(1907825, *prev_ind) + 2090, 3967));
    }
}

// ref silk_NLSF_stabilize (src/silk.cpp:2676)
OPUS_HD void silk_nlsf_stabilize(int16_t* NLSF_Q15, const int16_t* NDeltaMin_Q15, int L) {
    int loops;
    for (loops = 0; loops < 20; loops++) {
        s32 min_diff = NLSF_Q15[0] - NDeltaMin_Q15[0];
        int I = 0;
        for (int i = 1; i <= L - 1; i++) {
            s32 d = NLSF_Q15[i] - (NLSF_Q15[i - 1] + NDeltaMin_Q15[i]);
            if (d < min_diff) { min_diff = d; I = i; }
        }
        s32 d = (1 << 15) - (NLSF_Q15[L - 1] + NDeltaMin_Q15[L]);
        if (d < min_diff) { min_diff = d; I = L; }
        if (min_diff >= 0) return;
        if (I == 0) NLSF_Q15[0] = NDeltaMin_Q15[0];
        else if (I == L) NLSF_Q15[L - 1] = (int16_t)((1 << 15) - NDeltaMin_Q15[L]);
        else {
            s32 min_center = 0;
            for (int k = 0; k < I; k++) min_center += NDeltaMin_Q15[k];
            min_center += NDeltaMin_Q15[I] >> 1;
            s32 max_center = 1 << 15;
            for (int k = L; k > I; k--) max_center -= NDeltaMin_Q15[k];
            max_center -= NDeltaMin_Q15[I] >> 1;
            int16_t center = (int16_t)limit32(rshift_round((s32)NLSF_Q15[I - 1] + (s32)NLSF_Q15[I], 1), min_center, max_center);
            NLSF_Q15[I - 1] = (int16_t)(center - (NDeltaMin_Q15[I] >> 1));
            NLSF_Q15[I] = (int16_t)(NLSF_Q15[I - 1] + NDeltaMin_Q15[I]);
        }
    }
    // fall-back: sort, then enforce the minimum distances from both ends (ref :2741-2758)
    for (int i = 1; i < L; i++) {
        int16_t v = NLSF_Q15[i];
        int j = i - 1;
        for (; j >= 0 && v < NLSF_Q15[j]; j--) NLSF_Q15[j + 1] = NLSF_Q15[j];
        NLSF_Q15[j + 1] = v;
    }
    NLSF_Q15[0] = (int16_t)imax(NLSF_Q15[0], NDeltaMin_Q15[0]);
    for (int i = 1; i < L; i++) NLSF_Q15[i] = (int16_t)imax(NLSF_Q15[i], sat16((s32)NLSF_Q15[i - 1] + NDeltaMin_Q15[i]));
    NLSF_Q15[L - 1] = (int16_t)imin(NLSF_Q15[L - 1], (1 << 15) - NDeltaMin_Q15[L]);
    for (int i = L - 2; i >= 0; i--) NLSF_Q15[i] = (int16_t)imin(NLSF_Q15[i], NLSF_Q15[i + 1] - NDeltaMin_Q15[i + 1]);
}

// ref silk_NLSF_decode (src/silk.cpp:2466) with silk_NLSF_residual_dequant (:2445)
OPUS_HD void silk_nlsf_decode(int16_t* pNLSF_Q15, const int8_t* NLSFIndices, const SilkCb& cb) {
    uint8_t pred_Q8[kSilkMaxOrder];
    int16_t ec_ix[kSilkMaxOrder], res_Q10[kSilkMaxOrder];
    nlsf_unpack(ec_ix, pred_Q8, cb, NLSFIndices[0]);
    s32 out_Q10 = 0;
    for (int i = cb.order - 1; i >= 0; i--) {
        s32 pred_Q10 = smulbb(out_Q10, (int16_t)pred_Q8[i]) >> 8;
        out_Q10 = shl32(NLSFIndices[1 + i], 10);
        if (out_Q10 > 0) out_Q10 = out_Q10 - 102;
        else if (out_Q10 < 0) out_Q10 = out_Q10 + 102;
        out_Q10 = smlawb(pred_Q10, out_Q10, cb.quantStepSize_Q16);
        res_Q10[i] = (int16_t)out_Q10;
    }
    const uint8_t* el = cb.cb1_Q8 + NLSFIndices[0] * cb.order;
    const int16_t* wg = cb.cb1_wght_Q9 + NLSFIndices[0] * cb.order;
    for (int i = 0; i < cb.order; i++) {
        s32 v = (shl32((s32)res_Q10[i], 14) / wg[i]) + shl32((s32)(int16_t)el[i], 7);
        pNLSF_Q15[i] = (int16_t)limit32(v, 0, 32767);
    }
    silk_nlsf_stabilize(pNLSF_Q15, cb.deltaMin_Q15, cb.order);
}

// ref silk_bwexpander_32 (src/silk.cpp:561)
OPUS_HD void silk_bwexpander_32(s32* ar, int d, s32 chirp_Q16) {
    const s32 chirp_minus_one_Q16 = chirp_Q16 - 65536;
    for (int i = 0; i < d - 1; i++) {
        ar[i] = smulww(chirp_Q16, ar[i]);
        chirp_Q16 += rshift_round(mul_w(chirp_Q16, chirp_minus_one_Q16), 16);
    }
    ar[d - 1] = smulww(chirp_Q16, ar[d - 1]);
}
// ref silk_LPC_fit (src/silk.cpp:2314) with QOUT = 12, QIN = 17
OPUS_HD void silk_lpc_fit(int16_t* a_QOUT, s32* a_QIN, int d) {
    int i, idx = 0;
    for (i = 0; i < 10; i++) {
        s32 maxabs = 0;
        for (int k = 0; k < d; k++) {
            s32 absval = iabs(a_QIN[k]);
            if (absval > maxabs) { maxabs = absval; idx = k; }
        }
        maxabs = rshift_round(maxabs, 5);
        if (maxabs > 32767) {
            maxabs = imin(maxabs, 163838);
            s32 chirp_Q16 = 65470 - shl32(maxabs - 32767, 14) / (mul_w(maxabs, idx + 1) >> 2);
            silk_bwexpander_32(a_QIN, d, chirp_Q16);
        } else {
            break;
        }
    }
    if (i == 10) {
        for (int k = 0; k < d; k++) {
            a_QOUT[k] = sat16(rshift_round(a_QIN[k], 5));
            a_QIN[k] = shl32((s32)a_QOUT[k], 5);
        }
    } else {
        for (int k = 0; k < d; k++) a_QOUT[k] = (int16_t)rshift_round(a_QIN[k], 5);
    }
}
// ref silk_LPC_inverse_pred_gain_c + LPC_inverse_pred_gain_QA_c (src/silk.cpp:2359, 2425): 0 = unstable
OPUS_HD s32 silk_lpc_inverse_pred_gain(const int16_t* A_Q12, int order) {
    s32 A_QA[kSilkMaxOrder];
    s32 DC_resp = 0;
    for (int k = 0; k < order; k++) {
        DC_resp += (s32)A_Q12[k];
        A_QA[k] = shl32((s32)A_Q12[k], 12);
    }
    if (DC_resp >= 4096) return 0;
    const s32 A_LIMIT = 16773022;
    s32 invGain_Q30 = 1 << 30;
    int k;
    for (k = order - 1; k > 0; k--) {
        if (A_QA[k] > A_LIMIT || A_QA[k] < -A_LIMIT) return 0;
        s32 rc_Q31 = (s32)(0u - (u32)shl32(A_QA[k], 7));
        s32 rc_mult1_Q30 = (1 << 30) - smmul(rc_Q31, rc_Q31);
        invGain_Q30 = shl32(smmul(invGain_Q30, rc_mult1_Q30), 2);
        if (invGain_Q30 < 107374) return 0;
        int mult2Q = 32 - clz32(iabs(rc_mult1_Q30));
        s32 rc_mult2 = inverse32_varq(rc_mult1_Q30, mult2Q + 30);
        for (int n = 0; n < (k + 1) >> 1; n++) {
            s32 tmp1 = A_QA[n], tmp2 = A_QA[k - n - 1];
            s64 t64 = rshift_round64((s64)sub_sat32(tmp1, (s32)rshift_round64((s64)tmp2 * rc_Q31, 31)) * rc_mult2, mult2Q);
            if (t64 != (s64)(s32)t64) return 0;
            A_QA[n] = (s32)t64;
            t64 = rshift_round64((s64)sub_sat32(tmp2, (s32)rshift_round64((s64)tmp1 * rc_Q31, 31)) * rc_mult2, mult2Q);
            if (t64 != (s64)(s32)t64) return 0;
            A_QA[k - n - 1] = (s32)t64;
        }
    }
    if (A_QA[k] > A_LIMIT || A_QA[k] < -A_LIMIT) return 0;
    s32 rc_Q31 = (s32)(0u - (u32)shl32(A_QA[0], 7));
    s32 rc_mult1_Q30 = (1 << 30) - smmul(rc_Q31, rc_Q31);
    invGain_Q30 = shl32(smmul(invGain_Q30, rc_mult1_Q30), 2);
    if (invGain_Q30 < 107374) return 0;
    return invGain_Q30;
}
// ref silk_NLSF2A_find_poly / silk_NLSF2A (src/silk.cpp:626, 642)
OPUS_HD void silk_nlsf2a_find_poly(s32* out, const s32* cLSF, int dd) {
    out[0] = 1 << 16;
    out[1] = -cLSF[0];
    for (int k = 1; k < dd; k++) {
        s32 ftmp = cLSF[2 * k];
        out[k + 1] = shl32(out[k - 1], 1) - (s32)rshift_round64((s64)ftmp * out[k], 16);
        for (int n = k; n > 1; n--) out[n] += out[n - 2] - (s32)rshift_round64((s64)ftmp * out[n - 1], 16);
        out[1] -= ftmp;
    }
}
// silk_NLSF2A in two halves: everything up to and including the first stability check, and the bandwidth-expansion
// loop that runs only for filters that fail it (ref :665-672).  Those take up to 16 rounds of expansion + re-check
// while the rest of their warp waits, so the params kernel queues them and a second kernel runs the loop one failed
// filter per thread (k_silk_fix).  Returns whether the filter is stable as it is.
OPUS_HD bool silk_nlsf2a_first(int16_t* a_Q12, s32* a32_QA1, const int16_t* NLSF, int d) {
    const uint8_t* ordering = d == 16 ? kNlsf2aOrder16 : kNlsf2aOrder10;
    s32 cos_LSF_QA[kSilkMaxOrder], P[kSilkMaxOrder / 2 + 1], Q[kSilkMaxOrder / 2 + 1];
    for (int k = 0; k < d; k++) {
        s32 f_int = NLSF[k] >> 8;
        s32 f_frac = NLSF[k] - shl32(f_int, 8);
        s32 cos_val = kLsfCos[f_int];
        s32 delta = kLsfCos[f_int + 1] - cos_val;
        cos_LSF_QA[ordering[k]] = rshift_round(shl32(cos_val, 8) + mul_w(delta, f_frac), 4);
    }
    const int dd = d >> 1;
    silk_nlsf2a_find_poly(P, &cos_LSF_QA[0], dd);
    silk_nlsf2a_find_poly(Q, &cos_LSF_QA[1], dd);
    for (int k = 0; k < dd; k++) {
        s32 Ptmp = P[k + 1] + P[k];
        s32 Qtmp = Q[k + 1] - Q[k];
        a32_QA1[k] = -Qtmp - Ptmp;
        a32_QA1[d - k - 1] = Qtmp - Ptmp;
    }
    silk_lpc_fit(a_Q12, a32_QA1, d);
    return silk_lpc_inverse_pred_gain(a_Q12, d) != 0;
}
// the loop for a filter that failed the check; a_Q12 is rebuilt from a32_QA1 in every round
OPUS_HD void silk_nlsf2a_fix(int16_t* a_Q12, s32* a32_QA1, int d) {
    for (int i = 0; i < 16; i++) {
        silk_bwexpander_32(a32_QA1, d, 65536 - shl32(2, i));
        for (int k = 0; k < d; k++) a_Q12[k] = (int16_t)rshift_round(a32_QA1[k], 5);
        if (silk_lpc_inverse_pred_gain(a_Q12, d) != 0) break;
    }
}
OPUS_HD void silk_nlsf2a(int16_t* a_Q12, const int16_t* NLSF, int d) {
    s32 a32_QA1[kSilkMaxOrder];
    if (!silk_nlsf2a_first(a_Q12, a32_QA1, NLSF, d)) silk_nlsf2a_fix(a_Q12, a32_QA1, d);
}

}  // namespace opusb200
