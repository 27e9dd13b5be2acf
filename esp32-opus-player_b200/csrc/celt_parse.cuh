// celt_parse.cuh -- CELT entropy stage: one thread decodes one (stream, frame).
//
// Covers SURVEY section 8 rows a1 (range decoder, rangedec.cuh), a2 (energy symbols), a3 (tf, caps,
// dynalloc, allocation), a4 (quant_all_bands / quant_band / quant_partition / compute_theta control
// flow) and the serial half of a5 (PVQ index -> pulses -> normalised, spread-rotated leaf vector).
// Reference order of symbols: celt_decode_with_ec, src/celt.cpp:2162-2367.
//
// Output per frame: a CeltFrameHdr, a list of BandOps for the warp-parallel band kernel, and the
// coded leaves of X (int16, unit-norm Q14) written straight to the frame's X buffer in HBM.
#pragma once
#include "celt_script.cuh"
#include "celt_tables.cuh"
#include "rangedec.cuh"

namespace opusb200 {

// The parser's per-band work arrays live in the thread's local memory, 24 warps of them per SM: their footprint decides
// whether they stay in L1 / L2.  They are kept in the narrowest type that holds their values (the arithmetic stays
// 32-bit); the host build of these bodies (tests/hostsim, -DOPUS_CHECK_NARROW) aborts if a stored value does not fit.
template <class T>
OPUS_HD T narrow(s32 v) {
#if !defined(__CUDA_ARCH__) && defined(OPUS_CHECK_NARROW)
    if ((s32)(T)v != v) { fprintf(stderr, "narrow: %d does not fit %d bytes\n", (int)v, (int)sizeof(T)); abort(); }
#endif
    return (T)v;
}

#ifndef PARSE_TF_T
#define PARSE_TF_T int8_t      // tf_res: -3 .. 3
#endif
#ifndef PARSE_CAP_T
#define PARSE_CAP_T int16_t    // cap, dynalloc offsets: <= 9792 + 352
#endif
#ifndef PARSE_FINE_T
#define PARSE_FINE_T int8_t    // fine_quant 0 .. 8, fine_priority 0 / 1
#endif
#ifndef PARSE_BITS_T
#define PARSE_BITS_T uint16_t  // bits1, bits2 of the allocation search: 0 .. 34784
#endif
#ifndef PARSE_THR_T
#define PARSE_THR_T int16_t    // thresh <= 264, trim_offset within +-7040
#endif
#ifndef PARSE_FACC_T
#define PARSE_FACC_T int16_t   // fine energy sums, kept mod 2^16
#endif
typedef PARSE_TF_T tf_t;
typedef PARSE_CAP_T cap_t;
typedef PARSE_FINE_T fineq_t;
typedef PARSE_BITS_T abits_t;
typedef PARSE_THR_T thr_t;
typedef PARSE_FACC_T facc_t;

struct ParseCtx {
    RangeDec rd;
    // frame constants
    int C, LM, start, end, disable_inv;
    // band loop state (ref band_ctx_t, src/celt.h:125)
    int band, intensity, spread, tf_change, remaining_bits;
    // outputs
    BandOp* ops;      // transform ops from the front, leaf records from the back
    int n_ops, n_leaves, ops_cap, overflow;
    uint32_t lcg_steps;
    // transforms of the folding source of the current band job, emitted only if a leaf really folds (see lowband_prep)
    int prep_on, prep_scratch, prep_src, prep_N, prep_recombine, prep_tf, prep_NB, prep_B, prep_long;
    unsigned wmask;   // lanes of this warp that parse a CELT frame together (warp_converge / warp_any)
};

OPUS_HD void emit_op(ParseCtx& p, uint8_t code, int a, int off, int off2, int n, int g, int g2) {
    if (p.n_ops + p.n_leaves >= p.ops_cap) { p.overflow = 1; return; }
    BandOp o;
    o.code = code; o.a = (uint8_t)a; o.off = (uint16_t)off; o.off2 = (uint16_t)off2; o.n = (uint16_t)n;
    o.g = (int16_t)g; o.g2 = (int16_t)g2; o.aux = 0;
    p.ops[p.n_ops++] = o;
}
OPUS_HD void emit_leaf(ParseCtx& p, int B, int off, int K, int n, int gain, u32 index) {
    if (p.n_ops + p.n_leaves >= p.ops_cap) { p.overflow = 1; return; }
    BandOp o;
    o.code = OP_LEAF; o.a = (uint8_t)B; o.off = (uint16_t)off; o.off2 = (uint16_t)K; o.n = (uint16_t)n;
    o.g = (int16_t)gain; o.g2 = 0; o.aux = index;
    p.ops[p.ops_cap - 1 - p.n_leaves++] = o;
}

// The reference reshapes the folding source of every band (copy to scratch, haar1 per tf step, deinterleave;
// src/celt.cpp:1553-1584) whether or not any leaf of the band ends up folding.  Here the reshaping is emitted on
// demand, in front of the first OP_FOLD of the band job; when every leaf is coded (the common case at 64 kb/s)
// none of it is emitted and the band kernel replays a quarter fewer ops.  The result is the same: nothing else
// reads the scratch copy, and the in-place variant (last band) only touches folding history nobody reads again.
OPUS_HD void lowband_prep(ParseCtx& p) {
    if (!p.prep_on) return;
    p.prep_on = 0;
    int lowband = p.prep_src;
    const int N = p.prep_N, recombine = p.prep_recombine;
    int tf_change = p.prep_tf, N_B = p.prep_NB, B = p.prep_B;
    const int B00 = B;
    if (p.prep_scratch >= 0 && (recombine || ((N_B & 1) == 0 && tf_change < 0) || B00 > 1)) {
        emit_op(p, OP_COPY, 0, p.prep_scratch, lowband, N, 0, 0);
        lowband = p.prep_scratch;
    }
#pragma unroll 1
    for (int k = 0; k < recombine; k++) emit_op(p, OP_HAAR, 1 << k, lowband, 0, N >> k, 0, 0);
    B >>= recombine;
    N_B <<= recombine;
#pragma unroll 1
    while ((N_B & 1) == 0 && tf_change < 0) {
        emit_op(p, OP_HAAR, B, lowband, 0, N_B, 0, 0);
        B <<= 1;
        N_B >>= 1;
        tf_change++;
    }
    if (B > 1) emit_op(p, OP_DEINT, B << recombine, lowband, 0, N_B >> recombine, p.prep_long, 0);
}

// ---- pulse cache helpers (ref src/celt.h:533-569) ----
OPUS_HD int get_pulses(int i) { return i < 8 ? i : (8 + (i & 7)) << ((i >> 3) - 1); }
OPUS_HD const uint8_t* pulse_cache(int band, int LM) { return kCacheBits + kCacheIndex[(LM + 1) * kNbEBands + band]; }
OPUS_HD int bits2pulses(int band, int LM, int bits) {
    const uint8_t* cache = pulse_cache(band, LM);
    int lo = 0, hi = cache[0];
    bits--;
#pragma unroll 1
    for (int i = 0; i < 6; i++) {
        int mid = (lo + hi + 1) >> 1;
        if ((int)cache[mid] >= bits) hi = mid; else lo = mid;
    }
    if (bits - (lo == 0 ? -1 : (int)cache[lo]) <= (int)cache[hi] - bits) return lo;
    return hi;
}
OPUS_HD int pulses2bits(int band, int LM, int pulses) {
    return pulses == 0 ? 0 : pulse_cache(band, LM)[pulses] + 1;
}

// ---- PVQ codebook size and index decode (ref src/celt.cpp:653-660, 2545) ----
OPUS_HD u32 pvq_u(int n, int k) { return n < k ? kPvqU[kPvqRow[n] + k] : kPvqU[kPvqRow[k] + n]; }
OPUS_HD u32 pvq_v(int n, int k) { return pvq_u(n, k) + pvq_u(n, k + 1); }

// Walks a PVQ codeword index (ref cwrsi, src/celt.cpp:2545) only far enough to know which of the B
// interleaved blocks receive a pulse: the collapse mask (ref extract_collapse_mask :760) steers later
// folding decisions, so the entropy thread needs it; the pulses themselves are expanded by the band kernel.
OPUS_HD u32 pvq_index_mask(int n, int k, u32 i, int B) {
    const int N0 = (int)((u32)n / (u32)B);
    const u32 full = (1u << B) - 1u;
    u32 p, cm = 0;
    int s, k0;
    u32 bit = 1;      // mask bit of the block the current coefficient lies in
    int left = N0;    // coefficients left in that block
#pragma unroll 1
    while (n > 2) {
        u32 q;
        int val;
        if (k >= n) {
            const u32* row = kPvqU + kPvqRow[n];
            p = row[k + 1];
            s = -(int)(i >= p);
            i -= p & (u32)s;
            k0 = k;
            q = row[n];
            if (q > i) {
                k = n;
#pragma unroll 1
                do p = kPvqU[kPvqRow[--k] + n]; while (p > i);
            } else {
#pragma unroll 1
                for (p = row[k]; p > i; p = row[k]) k--;
            }
            i -= p;
            val = k0 - k;
        } else {
            p = kPvqU[kPvqRow[k] + n];
            q = kPvqU[kPvqRow[k + 1] + n];
            if (p <= i && i < q) {
                i -= p;
                val = 0;
            } else {
                s = -(int)(i >= q);
                i -= q & (u32)s;
                k0 = k;
#pragma unroll 1
                do p = kPvqU[kPvqRow[--k] + n]; while (p > i);
                i -= p;
                val = k0 - k;
            }
        }
        if (val) cm |= bit;
        if (--left == 0) { left = N0; bit <<= 1; }
        n--;
        // no pulses left: the remaining coefficients are zero; every block hit: nothing more to learn
        if (k == 0 || cm == full) return cm;
    }
    p = 2 * (u32)k + 1;
    s = -(int)(i >= p);
    i -= p & (u32)s;
    k0 = k;
    k = (int)((i + 1) >> 1);
    if (k0 - k) cm |= bit;
    if (--left == 0) bit <<= 1;
    if (k) cm |= bit;
    return cm;
}

// ref alg_unquant (src/celt.cpp:782), entropy half: read the codeword index, record the leaf.
OPUS_HD u32 leaf_unquant(ParseCtx& p, int xoff, int N, int K, int B, s16 gain) {
    const u32 index = rd_uint(p.rd, pvq_v(N, K));
    emit_leaf(p, B, xoff, K, N, gain, index);
    return B <= 1 ? 1u : pvq_index_mask(N, K, index, B);
}

struct ThetaOut { int inv, imid, iside, delta, itheta, qalloc; };

OPUS_LOCAL_TABLE s16 kExp2Table8[8] = {16384, 17866, 19483, 21247, 23170, 25267, 27554, 30048};               // ref :1216
OPUS_LOCAL_TABLE uint8_t kBitInterleave[16] = {0, 1, 1, 1, 2, 3, 3, 3, 2, 3, 3, 3, 2, 3, 3, 3};              // ref :1560
OPUS_LOCAL_TABLE uint8_t kBitDeinterleave[16] = {0x00, 0x03, 0x0C, 0x0F, 0x30, 0x33, 0x3C, 0x3F,            // ref :1606
                                                 0xC0, 0xC3, 0xCC, 0xCF, 0xF0, 0xF3, 0xFC, 0xFF};

// ref compute_qn (src/celt.cpp:1215)
OPUS_HD int compute_qn(int N, int b, int offset, int pulse_cap, int stereo) {
    int N2 = 2 * N - 1;
    if (stereo && N == 2) N2--;
    int qb = (b + N2 * offset) / N2;
    qb = imin(b - pulse_cap - (4 << 3), qb);
    qb = imin(8 << 3, qb);
    int qn;
    if (qb < (1 << 3 >> 1)) qn = 1;
    else {
        qn = kExp2Table8[qb & 0x7] >> (14 - (qb >> 3));
        qn = (qn + 1) >> 1 << 1;
    }
    return qn;
}

// ref compute_theta (src/celt.cpp:1241), decoder side
OPUS_HD void decode_theta(ParseCtx& p, ThetaOut& o, int N, int* b, int B, int B0, int LM, int stereo, int* fill) {
    int itheta = 0, inv = 0;
    const int i = p.band;
    int pulse_cap = kLogN400[i] + LM * (1 << 3);
    int offset = (pulse_cap >> 1) - (stereo && N == 2 ? 16 : 4);
    int qn = compute_qn(N, *b, offset, pulse_cap, stereo);
    if (stereo && i >= p.intensity) qn = 1;
    int tell = (int)rd_tell_frac(p.rd);
    if (qn != 1) {
        if (stereo && N > 2) {
            const int p0 = 3;
            int x0 = qn / 2;
            u32 ft = (u32)(p0 * (x0 + 1) + x0);
            int fs = (int)rd_decode(p.rd, ft);
            int x;
            if (fs < (x0 + 1) * p0) x = fs / p0;
            else x = x0 + 1 + (fs - (x0 + 1) * p0);
            rd_update(p.rd, (u32)(x <= x0 ? p0 * x : (x - 1 - x0) + (x0 + 1) * p0),
                      (u32)(x <= x0 ? p0 * (x + 1) : (x - x0) + (x0 + 1) * p0), ft);
            itheta = x;
        } else if (B0 > 1 || stereo) {
            itheta = (int)rd_uint(p.rd, (u32)qn + 1);
        } else {
            int fs = 1, fl = 0;
            int ft = ((qn >> 1) + 1) * ((qn >> 1) + 1);
            int fm = (int)rd_decode(p.rd, (u32)ft);
            if (fm < ((qn >> 1) * ((qn >> 1) + 1) >> 1)) {
                itheta = (int)((isqrt32(8 * (u32)fm + 1) - 1) >> 1);
                fs = itheta + 1;
                fl = itheta * (itheta + 1) >> 1;
            } else {
                itheta = (int)((2 * (u32)(qn + 1) - isqrt32(8 * (u32)(ft - fm - 1) + 1)) >> 1);
                fs = qn + 1 - itheta;
                fl = ft - ((qn + 1 - itheta) * (qn + 2 - itheta) >> 1);
            }
            rd_update(p.rd, (u32)fl, (u32)(fl + fs), (u32)ft);
        }
        itheta = (int)((u32)(itheta * 16384) / (u32)qn);
    } else if (stereo) {
        if (*b > 2 << 3 && p.remaining_bits > 2 << 3) inv = rd_bit_logp(p.rd, 2);
        else inv = 0;
        if (p.disable_inv) inv = 0;
        itheta = 0;
    }
    int qalloc = (int)rd_tell_frac(p.rd) - tell;
    *b -= qalloc;
    int imid, iside, delta;
    if (itheta == 0) {
        imid = 32767; iside = 0;
        *fill &= (1 << B) - 1;
        delta = -16384;
    } else if (itheta == 16384) {
        imid = 0; iside = 32767;
        *fill &= ((1 << B) - 1) << B;
        delta = 16384;
    } else {
        imid = bitexact_cos((s16)itheta);
        iside = bitexact_cos((s16)(16384 - itheta));
        delta = frac_mul16((N - 1) << 7, bitexact_log2tan(iside, imid));
    }
    o.inv = inv; o.imid = imid; o.iside = iside; o.delta = delta; o.itheta = itheta; o.qalloc = qalloc;
}

// ref quant_band_n1 (src/celt.cpp:1357). xoff/yoff index X (yoff < 0: mono), out = lowband_out offset or -1.
OPUS_HD u32 band_n1(ParseCtx& p, int xoff, int yoff, int out) {
    int x0 = 0;
#pragma unroll 1
    for (int c = 0; c < 1 + (yoff >= 0); c++) {
        int sign = 0;
        if (p.remaining_bits >= 1 << 3) {
            sign = (int)rd_bits(p.rd, 1);
            p.remaining_bits -= 1 << 3;
        }
        int v = sign ? -16384 : 16384;
        emit_op(p, OP_SETW, 0, c ? yoff : xoff, 0, 1, v, 0);
        if (c == 0) x0 = v;
    }
    if (out >= 0) emit_op(p, OP_SETW, 0, out, 0, 1, x0 >> 4, 0);
    return 1;
}

// ref quant_partition (src/celt.cpp:1382). Offsets index the W workspace; lowband < 0 means NULL.
// The reference recurses (a band splits into up to 16 leaves); here the tree is walked with an explicit
// stack so that all lanes of a warp run the same loop whatever their depth -- recursion would leave
// every lane in its own call frame and the warp would never reconverge.
struct PartNode { s32 b, fill; uint16_t xoff, N; s16 lowband, gain; uint8_t B; int8_t LM; };
struct PartFrame {
    PartNode second;     // the child decoded second (its b is adjusted by the rebalance)
    s32 first_bits, rebalance;
    u32 cm;
    s16 itheta;
    uint8_t shift, first_is_mid, phase;
};

// Every lane of p.wmask calls this once per band job (`active` false: nothing to decode) and all of them step
// through the loop together: one tree node per step, so the lanes that split run decode_theta side by side
// and the lanes at a leaf read their PVQ index side by side.
OPUS_HD u32 partition(ParseCtx& p, bool active, int xoff, int N, int b, int B, int lowband, int LM, s16 gain, int fill) {
    const int i = p.band;
    PartFrame st[5];
    int sp = 0;
    PartNode cur;
    cur.xoff = narrow<uint16_t>(xoff); cur.N = narrow<uint16_t>(N); cur.b = b; cur.B = narrow<uint8_t>(B); cur.lowband = narrow<s16>(lowband);
    cur.LM = narrow<int8_t>(LM); cur.fill = fill; cur.gain = gain;
    u32 ret = 0;
    bool descend = true, done = !active;
#pragma unroll 1
    while (warp_any(p.wmask, !done)) {
        if (done) continue;
        if (descend) {
            const uint8_t* cache = pulse_cache(i, cur.LM);
            if (cur.LM != -1 && cur.b > cache[cache[0]] + 12 && cur.N > 2) {
                ThetaOut t;
                const int B0 = cur.B;
                int Nh = cur.N >> 1, LMh = cur.LM - 1, bb = cur.b, fl = cur.fill;
                if (cur.B == 1) fl = (fl & 1) | (fl << 1);
                const int Bh = (cur.B + 1) >> 1;
                decode_theta(p, t, Nh, &bb, Bh, B0, LMh, 0, &fl);
                int delta = t.delta;
                const int itheta = t.itheta;
                if (B0 > 1 && (itheta & 0x3fff)) {
                    if (itheta > 8192) delta -= delta >> (4 - LMh);
                    else delta = imin(0, delta + (Nh << 3 >> (5 - LMh)));
                }
                const int mbits = imax(0, imin(bb, (bb - delta) / 2));
                const int sbits = bb - mbits;
                p.remaining_bits -= t.qalloc;
                PartNode mid, side;
                mid.xoff = cur.xoff; mid.N = narrow<uint16_t>(Nh); mid.b = mbits; mid.B = narrow<uint8_t>(Bh); mid.lowband = cur.lowband;
                mid.LM = narrow<int8_t>(LMh); mid.fill = fl;
                mid.gain = (s16)mul16_p15(cur.gain, (s16)t.imid);
                side.xoff = narrow<uint16_t>(cur.xoff + Nh); side.N = narrow<uint16_t>(Nh); side.b = sbits; side.B = narrow<uint8_t>(Bh);
                side.lowband = cur.lowband >= 0 ? narrow<s16>(cur.lowband + Nh) : (s16)-1;
                side.LM = narrow<int8_t>(LMh); side.fill = fl >> Bh;
                side.gain = (s16)mul16_p15(cur.gain, (s16)t.iside);
                PartFrame& f = st[sp++];
                f.rebalance = p.remaining_bits; f.itheta = narrow<s16>(itheta); f.shift = narrow<uint8_t>(B0 >> 1); f.cm = 0; f.phase = 0;
                if (mbits >= sbits) { f.first_is_mid = 1; f.first_bits = mbits; f.second = side; cur = mid; }
                else { f.first_is_mid = 0; f.first_bits = sbits; f.second = mid; cur = side; }
            } else {
                int q = bits2pulses(i, cur.LM, cur.b);
                int curr_bits = pulses2bits(i, cur.LM, q);
                p.remaining_bits -= curr_bits;
#pragma unroll 1
                while (p.remaining_bits < 0 && q > 0) {
                    p.remaining_bits += curr_bits;
                    q--;
                    curr_bits = pulses2bits(i, cur.LM, q);
                    p.remaining_bits -= curr_bits;
                }
                ret = 0;
                if (q != 0) {
                    ret = leaf_unquant(p, cur.xoff, cur.N, get_pulses(q), cur.B, cur.gain);
                } else {
                    const u32 cm_mask = (u32)(1UL << cur.B) - 1;
                    const int fl = cur.fill & (int)cm_mask;
                    if (!fl) {
                        emit_op(p, OP_ZERO, 0, cur.xoff, 0, cur.N, 0, 0);
                    } else if (cur.lowband < 0) {
                        emit_op(p, OP_NOISE, 0, cur.xoff, 0, cur.N, cur.gain, 0);
                        p.lcg_steps += (u32)cur.N;
                        ret = cm_mask;
                    } else {
                        lowband_prep(p);
                        emit_op(p, OP_FOLD, 0, cur.xoff, cur.lowband, cur.N, cur.gain, 0);
                        p.lcg_steps += (u32)cur.N;
                        ret = (u32)fl;
                    }
                }
                descend = false;
            }
        } else if (sp == 0) {
            done = true;
        } else {
            PartFrame& f = st[sp - 1];
            if (f.phase == 0) {
                f.cm = f.first_is_mid ? ret : ret << f.shift;
                const int rebalance = f.first_bits - (f.rebalance - p.remaining_bits);
                if (rebalance > 3 << 3 && f.itheta != (f.first_is_mid ? 0 : 16384)) f.second.b += rebalance - (3 << 3);
                f.phase = 1;
                cur = f.second;
                descend = true;
            } else {
                ret = f.cm | (f.first_is_mid ? ret << f.shift : ret);
                sp--;
            }
        }
    }
    return ret;
}

// ref quant_band (src/celt.cpp:1526), split around the partition walk so that the walk has one call site that
// every lane reaches: band_mono_pre emits the transforms of the folding source and says whether there is a tree
// to walk, band_mono_post undoes the time/frequency reshaping on the decoded band.
struct MonoCtx { int xoff, N0, B, B0, N_B, N_B0, recombine, time_divide, longBlocks, lowband, lowband_out, fill; };

OPUS_HD bool band_mono_pre(ParseCtx& p, MonoCtx& m, int xoff, int N, int B, int lowband, int lowband_out, int scratch, int fill) {
    m.xoff = xoff; m.N0 = N; m.lowband_out = lowband_out;
    int N_B = (int)((u32)N / (u32)B);
    int B0 = B, time_divide = 0, recombine = 0;
    int tf_change = p.tf_change;
    m.longBlocks = B0 == 1;
    if (N == 1) {
        band_n1(p, xoff, -1, lowband_out);
        return false;
    }
    if (tf_change > 0) recombine = tf_change;
    p.prep_on = lowband >= 0;
    p.prep_scratch = scratch; p.prep_src = lowband; p.prep_N = N; p.prep_recombine = recombine; p.prep_tf = tf_change;
    p.prep_NB = N_B; p.prep_B = B; p.prep_long = m.longBlocks;
    if (scratch >= 0 && lowband >= 0 && (recombine || ((N_B & 1) == 0 && tf_change < 0) || B0 > 1)) lowband = scratch;
#pragma unroll 1
    for (int k = 0; k < recombine; k++) {
        fill = kBitInterleave[fill & 0xF] | kBitInterleave[fill >> 4] << 2;
    }
    B >>= recombine;
    N_B <<= recombine;
#pragma unroll 1
    while ((N_B & 1) == 0 && tf_change < 0) {
        fill |= fill << B;
        B <<= 1;
        N_B >>= 1;
        time_divide++;
        tf_change++;
    }
    B0 = B;
    m.B = B; m.B0 = B0; m.N_B = N_B; m.N_B0 = N_B; m.recombine = recombine; m.time_divide = time_divide; m.lowband = lowband; m.fill = fill;
    return true;
}

OPUS_HD u32 band_mono_post(ParseCtx& p, const MonoCtx& m, u32 cm) {
    const int xoff = m.xoff, recombine = m.recombine, B0 = m.B0;
    int N_B = m.N_B, B = m.B;
    if (B0 > 1) emit_op(p, OP_INT, B0 << recombine, xoff, 0, N_B >> recombine, m.longBlocks, 0);
    N_B = m.N_B0;
    B = B0;
#pragma unroll 1
    for (int k = 0; k < m.time_divide; k++) {
        B >>= 1;
        N_B <<= 1;
        cm |= cm >> B;
        emit_op(p, OP_HAAR, B, xoff, 0, N_B, 0, 0);
    }
#pragma unroll 1
    for (int k = 0; k < recombine; k++) {
        cm = kBitDeinterleave[cm];
        emit_op(p, OP_HAAR, 1 << k, xoff, 0, m.N0 >> k, 0, 0);
    }
    B <<= recombine;
    if (m.lowband_out >= 0) emit_op(p, OP_SCALE_OUT, 0, xoff, m.lowband_out, m.N0, sqrt_q(shl32(m.N0, 22)), 0);
    cm &= (u32)((1 << B) - 1);
    return cm;
}

// ref quant_all_bands (src/celt.cpp:1754) with quant_band_stereo (:1628) folded in.
// Every band is turned into at most two "mono jobs" (mid/side in either order, or the two channels of a
// dual-stereo band) that go through ONE call site of band_mono: which job runs first is data, not
// control flow, so the lanes of a warp keep executing the same code whatever their stereo decisions.
struct BandJob { int xoff, b, lowband, out, scratch, fill; s16 gain; };

OPUS_HD void all_bands(ParseCtx& p, uint8_t* cmask, const int* pulses, int shortBlocks, int dual_stereo, int intensity,
                       const tf_t* tf_res, int total_bits, int balance, int codedBands) {
    const int LM = p.LM, C = p.C, start = p.start, end = p.end;
    const int M = 1 << LM;
    const int N = M * kShortMdct;
    const int B = shortBlocks ? M : 1;
    const int norm_offset = M * kEBands[start];
    int lowband_offset = 0, update_lowband = 1;
    p.intensity = intensity;
    const int start0 = warp_min(p.wmask, start);  // lanes may start at band 0 or 17 (hybrid): one loop for all of them
#pragma unroll 1
    for (int i = start0; i < end; i++) {
        warp_converge(p.wmask);
        if (i < start) {
            // keep in step with the lanes that do have this band: two empty job slots
#pragma unroll 1
            for (int j = 0; j < 2; j++) partition(p, false, 0, 0, 0, 0, 0, 0, 0, 0);
            continue;
        }
        p.band = i;
        const int last = (i == end - 1);
        const int xoff = M * kEBands[i];
        const int yoff = C == 2 ? N + xoff : -1;
        const int Nb = M * kEBands[i + 1] - M * kEBands[i];
        const int tell = (int)rd_tell_frac(p.rd);
        if (i != start) balance -= tell;
        const int remaining = total_bits - tell - 1;
        p.remaining_bits = remaining;
        int b;
        if (i <= codedBands - 1) {
            int curr_balance = balance / imin(3, codedBands - i);
            b = imax(0, imin(16383, imin(remaining + 1, pulses[i] + curr_balance)));
        } else {
            b = 0;
        }
        if ((M * kEBands[i] - Nb >= M * kEBands[start] || i == start + 1) && (update_lowband || lowband_offset == 0))
            lowband_offset = i;
        if (i == start + 1) {  // special_hybrid_folding (ref :1743)
            int n1 = M * (kEBands[start + 1] - kEBands[start]);
            int n2 = M * (kEBands[start + 2] - kEBands[start + 1]);
            if (n2 > n1) {
                emit_op(p, OP_COPY, 0, W_NORM + n1, W_NORM + 2 * n1 - n2, n2 - n1, 0, 0);
                if (dual_stereo) emit_op(p, OP_COPY, 0, W_NORM2 + n1, W_NORM2 + 2 * n1 - n2, n2 - n1, 0, 0);
            }
        }
        const int tf_change = tf_res[i];
        p.tf_change = tf_change;
        const int scratch = last ? -1 : W_SCRATCH;
        int effective_lowband = -1;
        u32 x_cm, y_cm;
        if (lowband_offset != 0 && (p.spread != 3 || B > 1 || tf_change < 0)) {
            effective_lowband = imax(0, M * kEBands[lowband_offset] - norm_offset - Nb);
            int fold_start = lowband_offset;
#pragma unroll 1
            while (M * kEBands[--fold_start] > effective_lowband + norm_offset) {}
            int fold_end = lowband_offset - 1;
#pragma unroll 1
            while (++fold_end < i && M * kEBands[fold_end] < effective_lowband + norm_offset + Nb) {}
            x_cm = y_cm = 0;
            int fold_i = fold_start;
#pragma unroll 1
            do {
                x_cm |= cmask[fold_i * C + 0];
                y_cm |= cmask[fold_i * C + C - 1];
            } while (++fold_i < fold_end);
        } else {
            x_cm = y_cm = (u32)((1 << B) - 1);
        }
        if (dual_stereo && i == intensity) {
            dual_stereo = 0;
            int n = M * kEBands[i] - norm_offset;
            if (n > 0) emit_op(p, OP_AVG, 0, W_NORM, W_NORM2, n, 0, 0);
        }
        const int lb = effective_lowband != -1 ? W_NORM + effective_lowband : -1;
        const int lb2 = effective_lowband != -1 ? W_NORM2 + effective_lowband : -1;
        const int out = last ? -1 : W_NORM + M * kEBands[i] - norm_offset;
        const int out2 = last ? -1 : W_NORM2 + M * kEBands[i] - norm_offset;

        // ---- plan the jobs of this band
        BandJob job[2] = {};
        int njobs = 0;
        // mode: 0 = independent jobs (mono / dual stereo), 1 = mid/side with rebalance, 2 = N=2 stereo
        int mode = 0, first_is_mid = 1, itheta = 0, inv = 0, first_bits = 0, rebalance0 = 0, n2_arg = 0, n2_x2 = 0, n2_y2 = 0;
        s16 mid = 0, side = 0;
        u32 cm_direct = 0;
        if (dual_stereo) {
            job[0] = BandJob{xoff, b / 2, lb, out, scratch, (int)x_cm, 32767};
            job[1] = BandJob{yoff, b / 2, lb2, out2, scratch, (int)y_cm, 32767};
            njobs = 2;
        } else if (yoff < 0) {
            job[0] = BandJob{xoff, b, lb, out, scratch, (int)(x_cm | y_cm), 32767};
            njobs = 1;
        } else if (Nb == 1) {
            cm_direct = band_n1(p, xoff, yoff, out);
        } else {  // quant_band_stereo
            int fill = (int)(x_cm | y_cm);
            const int orig_fill = fill;
            ThetaOut t;
            int bb = b;
            decode_theta(p, t, Nb, &bb, B, B, LM, 1, &fill);
            mid = (s16)t.imid; side = (s16)t.iside; itheta = t.itheta; inv = t.inv;
            if (Nb == 2) {
                mode = 2;
                int mbits = bb, sbits = 0;
                if (itheta != 0 && itheta != 16384) sbits = 1 << 3;
                mbits -= sbits;
                const int c = itheta > 8192;
                p.remaining_bits -= t.qalloc + sbits;
                n2_x2 = c ? yoff : xoff;
                n2_y2 = c ? xoff : yoff;
                int sign = 0;
                if (sbits) sign = (int)rd_bits(p.rd, 1);
                n2_arg = sign | (c << 1);
                job[0] = BandJob{n2_x2, mbits, lb, out, scratch, orig_fill, 32767};
                njobs = 1;
            } else {
                mode = 1;
                const int mbits = imax(0, imin(bb, (bb - t.delta) / 2));
                const int sbits = bb - mbits;
                p.remaining_bits -= t.qalloc;
                rebalance0 = p.remaining_bits;
                const BandJob jm = BandJob{xoff, mbits, lb, out, scratch, fill, 32767};
                const BandJob js = BandJob{yoff, sbits, -1, -1, -1, fill >> B, side};
                first_is_mid = mbits >= sbits;
                first_bits = first_is_mid ? mbits : sbits;
                job[0] = first_is_mid ? jm : js;
                job[1] = first_is_mid ? js : jm;
                njobs = 2;
            }
        }
        // ---- run them through the single band_mono call site
        u32 cm_job[2] = {0, 0};
#pragma unroll 1
        for (int j = 0; j < 2; j++) {
            const BandJob& J = job[j < njobs ? j : 0];
            MonoCtx m = {};
            const bool act = j < njobs;
            const bool walk = act && band_mono_pre(p, m, J.xoff, Nb, B, J.lowband, J.out, J.scratch, J.fill);
            const u32 cm = partition(p, walk, J.xoff, Nb, J.b, m.B, m.lowband, LM, J.gain, m.fill);
            p.prep_on = 0;
            if (act) cm_job[j] = walk ? band_mono_post(p, m, cm) : 1u;
            if (mode == 1 && j == 0) {
                const int rebalance = first_bits - (rebalance0 - p.remaining_bits);
                if (rebalance > 3 << 3 && itheta != (first_is_mid ? 0 : 16384)) job[1].b += rebalance - (3 << 3);
            }
        }
        if (dual_stereo) {
            x_cm = cm_job[0];
            y_cm = cm_job[1];
        } else {
            if (mode == 2) emit_op(p, OP_N2, n2_arg, n2_x2, n2_y2, 2, mid, side);
            if (mode == 1) emit_op(p, OP_MERGE, 0, xoff, yoff, Nb, mid, 0);
            if (mode != 0 && inv) emit_op(p, OP_NEG, 0, yoff, 0, Nb, 0, 0);
            x_cm = y_cm = cm_job[0] | cm_job[1] | cm_direct;
        }
        cmask[i * C + 0] = (uint8_t)x_cm;
        cmask[i * C + C - 1] = (uint8_t)y_cm;
        balance += pulses[i] + tell;
        update_lowband = b > (Nb << 3);
    }
}

// ref interp_bits2pulses (src/celt.cpp:3298), decoder side
OPUS_HD int interp_bits2pulses(RangeDec& rd, int start, int end, int skip_start, const abits_t* bits1, const abits_t* bits2,
                               const thr_t* thresh, const cap_t* cap, int total, int* balance_out, int skip_rsv, int* intensity,
                               int intensity_rsv, int* dual_stereo, int dual_stereo_rsv, int* bits, fineq_t* ebits,
                               fineq_t* fine_priority, int C, int LM) {
    const int alloc_floor = C << 3;
    const int stereo = C > 1;
    const int logM = LM << 3;
    int lo = 0, hi = 1 << 6, psum, done, j;
#pragma unroll 1
    for (int it = 0; it < 6; it++) {
        int mid = (lo + hi) >> 1;
        psum = 0; done = 0;
#pragma unroll 1
        for (j = end; j-- > start;) {
            int tmp = bits1[j] + (mid * bits2[j] >> 6);
            if (tmp >= thresh[j] || done) { done = 1; psum += imin(tmp, cap[j]); }
            else if (tmp >= alloc_floor) psum += alloc_floor;
        }
        if (psum > total) hi = mid; else lo = mid;
    }
    psum = 0; done = 0;
#pragma unroll 1
    for (j = end; j-- > start;) {
        int tmp = bits1[j] + (lo * bits2[j] >> 6);
        if (tmp < thresh[j] && !done) tmp = tmp >= alloc_floor ? alloc_floor : 0;
        else done = 1;
        tmp = imin(tmp, cap[j]);
        bits[j] = tmp;
        psum += tmp;
    }
    int codedBands, left, percoeff;
#pragma unroll 1
    for (codedBands = end;; codedBands--) {
        j = codedBands - 1;
        if (j <= skip_start) { total += skip_rsv; break; }
        left = total - psum;
        percoeff = (int)((u32)left / (u32)(kEBands[codedBands] - kEBands[start]));
        left -= (kEBands[codedBands] - kEBands[start]) * percoeff;
        int rem = imax(left - (kEBands[j] - kEBands[start]), 0);
        int band_width = kEBands[codedBands] - kEBands[j];
        int band_bits = bits[j] + percoeff * band_width + rem;
        if (band_bits >= imax(thresh[j], alloc_floor + (1 << 3))) {
            if (rd_bit_logp(rd, 1)) break;
            psum += 1 << 3;
            band_bits -= 1 << 3;
        }
        psum -= bits[j] + intensity_rsv;
        if (intensity_rsv > 0) intensity_rsv = kLog2Frac[j - start];
        psum += intensity_rsv;
        if (band_bits >= alloc_floor) { psum += alloc_floor; bits[j] = alloc_floor; }
        else bits[j] = 0;
    }
    if (intensity_rsv > 0) *intensity = start + (int)rd_uint(rd, (u32)(codedBands + 1 - start));
    else *intensity = 0;
    if (*intensity <= start) { total += dual_stereo_rsv; dual_stereo_rsv = 0; }
    if (dual_stereo_rsv > 0) *dual_stereo = rd_bit_logp(rd, 1);
    else *dual_stereo = 0;

    left = total - psum;
    percoeff = (int)((u32)left / (u32)(kEBands[codedBands] - kEBands[start]));
    left -= (kEBands[codedBands] - kEBands[start]) * percoeff;
#pragma unroll 1
    for (j = start; j < codedBands; j++) bits[j] += percoeff * (kEBands[j + 1] - kEBands[j]);
#pragma unroll 1
    for (j = start; j < codedBands; j++) {
        int tmp = imin(left, kEBands[j + 1] - kEBands[j]);
        bits[j] += tmp;
        left -= tmp;
    }
    int balance = 0;
#pragma unroll 1
    for (j = start; j < codedBands; j++) {
        int N0 = kEBands[j + 1] - kEBands[j];
        int N = N0 << LM;
        int bit = bits[j] + balance;
        int excess;
        if (N > 1) {
            excess = imax(bit - cap[j], 0);
            bits[j] = bit - excess;
            int den = C * N + ((C == 2 && N > 2 && !*dual_stereo && j < *intensity) ? 1 : 0);
            int NClogN = den * (kLogN400[j] + logM);
            int offset = (NClogN >> 1) - den * 21;
            if (N == 2) offset += den << 3 >> 2;
            if (bits[j] + offset < den * 2 << 3) offset += NClogN >> 2;
            else if (bits[j] + offset < den * 3 << 3) offset += NClogN >> 3;
            int eb = imax(0, bits[j] + offset + (den << 2));
            eb = (int)((u32)eb / (u32)den) >> 3;
            if (C * eb > (bits[j] >> 3)) eb = bits[j] >> stereo >> 3;
            eb = imin(eb, 8);
            ebits[j] = narrow<fineq_t>(eb);
            fine_priority[j] = eb * (den << 3) >= bits[j] + offset;
            bits[j] -= C * eb << 3;
        } else {
            excess = imax(0, bit - (C << 3));
            bits[j] = bit - excess;
            ebits[j] = 0;
            fine_priority[j] = 1;
        }
        if (excess > 0) {
            int extra_fine = imin(excess >> (stereo + 3), 8 - ebits[j]);
            ebits[j] = narrow<fineq_t>(ebits[j] + extra_fine);
            int extra_bits = extra_fine * C << 3;
            fine_priority[j] = extra_bits >= excess - balance;
            excess -= extra_bits;
        }
        balance = excess;
    }
    *balance_out = balance;
#pragma unroll 1
    for (; j < end; j++) {
        const int eb = bits[j] >> stereo >> 3;
        ebits[j] = narrow<fineq_t>(eb);
        bits[j] = 0;
        fine_priority[j] = eb < 1;
    }
    return codedBands;
}

// ref clt_compute_allocation (src/celt.cpp:3523), decoder side
OPUS_HD int compute_allocation(RangeDec& rd, int start, int end, const cap_t* offsets, const cap_t* cap, int alloc_trim,
                               int* intensity, int* dual_stereo, int total, int* balance, int* pulses, fineq_t* ebits,
                               fineq_t* fine_priority, int C, int LM) {
    total = imax(total, 0);
    const int len = kNbEBands;
    int skip_start = start;
    int skip_rsv = total >= 1 << 3 ? 1 << 3 : 0;
    total -= skip_rsv;
    int intensity_rsv = 0, dual_stereo_rsv = 0;
    if (C == 2) {
        intensity_rsv = kLog2Frac[end - start];
        if (intensity_rsv > total) intensity_rsv = 0;
        else {
            total -= intensity_rsv;
            dual_stereo_rsv = total >= 1 << 3 ? 1 << 3 : 0;
            total -= dual_stereo_rsv;
        }
    }
    abits_t bits1[kNbEBands], bits2[kNbEBands];
    thr_t thresh[kNbEBands], trim_offset[kNbEBands];
#pragma unroll 1
    for (int j = start; j < end; j++) {
        int w = kEBands[j + 1] - kEBands[j];
        thresh[j] = narrow<thr_t>(imax(C << 3, (3 * w << LM << 3) >> 4));
        int tr = C * w * (alloc_trim - 5 - LM) * (end - j - 1) * (1 << (LM + 3)) >> 6;
        if (w << LM == 1) tr -= C << 3;
        trim_offset[j] = narrow<thr_t>(tr);
    }
    int lo = 1, hi = 11 - 1;
#pragma unroll 1
    do {
        int done = 0, psum = 0;
        int mid = (lo + hi) >> 1;
#pragma unroll 1
        for (int j = end; j-- > start;) {
            int w = kEBands[j + 1] - kEBands[j];
            int bitsj = C * w * kBandAlloc[mid * len + j] << LM >> 2;
            if (bitsj > 0) bitsj = imax(0, bitsj + trim_offset[j]);
            bitsj += offsets[j];
            if (bitsj >= thresh[j] || done) { done = 1; psum += imin(bitsj, cap[j]); }
            else if (bitsj >= C << 3) psum += C << 3;
        }
        if (psum > total) hi = mid - 1; else lo = mid + 1;
    } while (lo <= hi);
    hi = lo--;
#pragma unroll 1
    for (int j = start; j < end; j++) {
        int w = kEBands[j + 1] - kEBands[j];
        int bits1j = C * w * kBandAlloc[lo * len + j] << LM >> 2;
        int bits2j = hi >= 11 ? cap[j] : C * w * kBandAlloc[hi * len + j] << LM >> 2;
        if (bits1j > 0) bits1j = imax(0, bits1j + trim_offset[j]);
        if (bits2j > 0) bits2j = imax(0, bits2j + trim_offset[j]);
        if (lo > 0) bits1j += offsets[j];
        bits2j += offsets[j];
        if (offsets[j] > 0) skip_start = j;
        bits2j = imax(0, bits2j - bits1j);
        bits1[j] = narrow<abits_t>(bits1j);
        bits2[j] = narrow<abits_t>(bits2j);
    }
    return interp_bits2pulses(rd, start, end, skip_start, bits1, bits2, thresh, cap, total, balance, skip_rsv, intensity,
                              intensity_rsv, dual_stereo, dual_stereo_rsv, pulses, ebits, fine_priority, C, LM);
}

// The first symbols of a CELT frame: silence flag, post-filter parameters, transient flag (ref src/celt.cpp:2239-2276).
// Shared by the parser proper and by k_classify, which reads just these to sort the frames of a launch by kind.
struct CeltHead { int silence, pf_gain, pf_pitch, pf_tapset, postfilter, transient, tell; };
OPUS_HD CeltHead celt_parse_head(RangeDec& rd, int start, int LM) {
    CeltHead h = {0, 0, 0, 0, 0, 0, 0};
    const int total_bits = (int)rd.storage * 8;
    int tell = rd_tell(rd);
    if (tell >= total_bits) h.silence = 1;
    else if (tell == 1) h.silence = rd_bit_logp(rd, 15);
    if (h.silence) {
        tell = (int)rd.storage * 8;
        rd.nbits_total += tell - rd_tell(rd);
    }
    if (start == 0 && tell + 16 <= total_bits) {
        if (rd_bit_logp(rd, 1)) {
            int octave = (int)rd_uint(rd, 6);
            h.pf_pitch = (16 << octave) + (int)rd_bits(rd, 4 + (u32)octave) - 1;
            int qg = (int)rd_bits(rd, 3);
            if (rd_tell(rd) + 2 <= total_bits) {
                h.pf_tapset = rd_icdf(rd, kIcdf210, 2);
            }
            h.pf_gain = 3072 * (qg + 1);  // QCONST16(.09375, 15)
            h.postfilter = 1;
        }
        tell = rd_tell(rd);
    }
    if (LM > 0 && tell + 3 <= total_bits) {
        h.transient = rd_bit_logp(rd, 3);
        tell = rd_tell(rd);
    }
    h.tell = tell;
    return h;
}

// One CELT frame: every symbol of celt_decode_with_ec (ref src/celt.cpp:2239-2367) in bitstream order.
// `p.rd` must already be initialised on the frame (and, for hybrid, advanced past the SILK layer).
// frame_size = 120 << LM.  Returns the status also stored in hdr.
OPUS_HD int celt_parse_frame(ParseCtx& p, CeltFrameHdr& hdr) {
    RangeDec& rd = p.rd;
    const int C = p.C, LM = p.LM, start = p.start, end = p.end;
    const int M = 1 << LM;
    hdr.LM = (uint8_t)LM; hdr.C = (uint8_t)C; hdr.start = (uint8_t)start; hdr.end = (uint8_t)end;
    hdr.n_ops = 0; hdr.n_leaves = 0; hdr.spread = 2; hdr.lcg_steps = 0; hdr.flags = 0; hdr.pf_pitch = 0; hdr.pf_gain = 0; hdr.pf_tapset = 0;
    hdr.final_rng = 0;
    if (rd.storage > (u32)kMaxPacket || rd.storage <= 1) {  // ref :2216, :2225
        hdr.status = OPUSB_CELT_BAD_ARG;
        return hdr.status;
    }
    p.n_ops = 0; p.n_leaves = 0; p.overflow = 0; p.lcg_steps = 0;
    const int start0 = warp_min(p.wmask, start);  // common first band of the lanes that parse together

    int total_bits = (int)rd.storage * 8;
    const CeltHead head = celt_parse_head(rd, start, LM);
    const int silence = head.silence, pf_gain = head.pf_gain, pf_pitch = head.pf_pitch, pf_tapset = head.pf_tapset;
    const int isTransient = head.transient;
    int tell = head.tell;
    int flags = head.postfilter ? FF_POSTFILTER : 0;
    const int shortBlocks = isTransient ? M : 0;
    const int intra = tell + 3 <= total_bits ? rd_bit_logp(rd, 3) : 0;

    // ---- coarse energy symbols (ref unquant_coarse_energy :3613): only the bitstream-dependent part;
    // the prediction from the previous frame is applied by the energy kernel.
    {
        const uint8_t* prob = kEProb + (LM * 2 + intra) * 42;
        const s16 beta = intra ? (s16)4915 : (LM == 0 ? (s16)30147 : LM == 1 ? (s16)22282 : LM == 2 ? (s16)12124 : (s16)6554);
        s32 prev[2] = {0, 0};
        const int budget = (int)rd.storage * 8;
#pragma unroll 1
        for (int i = start0; i < end; i++) {
            warp_converge(p.wmask);
            if (i < start) continue;
#pragma unroll 1
            for (int c = 0; c < C; c++) {
                int qi;
                int t = rd_tell(rd);
                if (budget - t >= 15) {
                    int pi = 2 * imin(i, 20);
                    qi = rd_laplace(rd, (u32)prob[pi] << 7, (int)prob[pi + 1] << 6);
                } else if (budget - t >= 2) {
                    qi = rd_icdf(rd, kIcdf210, 2);
                    qi = (qi >> 1) ^ -(qi & 1);
                } else if (budget - t >= 1) {
                    qi = -rd_bit_logp(rd, 1);
                } else {
                    qi = -1;
                }
                s32 q = shl32(qi, 10);
                hdr.pq[c][i] = prev[c] + shl32(q, 7);
                prev[c] = prev[c] + shl32(q, 7) - mul16(beta, pshr32(q, 8));
            }
        }
    }

    // ---- tf_decode (ref :2128)
    tf_t tf_res[kNbEBands];
    {
        u32 budget = rd.storage * 8;
        u32 t = (u32)rd_tell(rd);
        int logp = isTransient ? 2 : 4;
        int tf_select_rsv = LM > 0 && t + (u32)logp + 1 <= budget;
        budget -= (u32)tf_select_rsv;
        int tf_changed = 0, curr = 0;
#pragma unroll 1
        for (int i = start; i < end; i++) {
            if (t + (u32)logp <= budget) {
                curr ^= rd_bit_logp(rd, (u32)logp);
                t = (u32)rd_tell(rd);
                tf_changed |= curr;
            }
            tf_res[i] = (tf_t)curr;
            logp = isTransient ? 4 : 5;
        }
        int tf_select = 0;
        if (tf_select_rsv && kTfSelect[LM * 8 + 4 * isTransient + 0 + tf_changed] != kTfSelect[LM * 8 + 4 * isTransient + 2 + tf_changed])
            tf_select = rd_bit_logp(rd, 1);
#pragma unroll 1
        for (int i = start; i < end; i++) tf_res[i] = kTfSelect[LM * 8 + 4 * isTransient + 2 * tf_select + tf_res[i]];
    }

    tell = rd_tell(rd);
    int spread = 2;  // SPREAD_NORMAL
    if (tell + 4 <= total_bits) spread = rd_icdf(rd, kSpreadIcdf, 5);
    p.spread = spread;

    // ---- caps + dynalloc (ref :911, :2301-2328)
    cap_t cap[kNbEBands], offsets[kNbEBands];
#pragma unroll 1
    for (int i = 0; i < kNbEBands; i++) {
        int N = (kEBands[i + 1] - kEBands[i]) << LM;
        cap[i] = narrow<cap_t>((kCacheCaps[kNbEBands * (2 * LM + C - 1) + i] + 64) * C * N >> 2);
        offsets[i] = 0;
    }
    int dynalloc_logp = 6;
    total_bits <<= 3;
    tell = (int)rd_tell_frac(rd);
#pragma unroll 1
    for (int i = start0; i < end; i++) {
        warp_converge(p.wmask);
        if (i < start) continue;
        int width = C * (kEBands[i + 1] - kEBands[i]) << LM;
        int quanta = imin(width << 3, imax(6 << 3, width));
        int loop_logp = dynalloc_logp;
        int boost = 0;
#pragma unroll 1
        while (tell + (loop_logp << 3) < total_bits && boost < cap[i]) {
            int flag = rd_bit_logp(rd, (u32)loop_logp);
            tell = (int)rd_tell_frac(rd);
            if (!flag) break;
            boost += quanta;
            total_bits -= quanta;
            loop_logp = 1;
        }
        offsets[i] = narrow<cap_t>(boost);
        if (boost > 0) dynalloc_logp = imax(2, dynalloc_logp - 1);
    }
    const int alloc_trim = tell + (6 << 3) <= total_bits ? rd_icdf(rd, kTrimIcdf, 7) : 5;

    int bits = (((int)rd.storage * 8) << 3) - (int)rd_tell_frac(rd) - 1;
    const int anti_collapse_rsv = isTransient && LM >= 2 && bits >= ((LM + 2) << 3) ? (1 << 3) : 0;
    bits -= anti_collapse_rsv;

    int pulses[kNbEBands];
    fineq_t fine_quant[kNbEBands], fine_priority[kNbEBands];
#pragma unroll 1
    for (int i = 0; i < kNbEBands; i++) { pulses[i] = 0; fine_quant[i] = 0; fine_priority[i] = 0; }
    int intensity = 0, dual_stereo = 0, balance = 0;
    const int codedBands = compute_allocation(rd, start, end, offsets, cap, alloc_trim, &intensity, &dual_stereo, bits, &balance,
                                              pulses, fine_quant, fine_priority, C, LM);

    // ---- fine energy (ref unquant_fine_energy :3664)
    facc_t fine_acc[2][kNbEBands];  // sums mod 2^16, which is all the header keeps
#pragma unroll 1
    for (int c = 0; c < 2; c++) for (int i = 0; i < kNbEBands; i++) fine_acc[c][i] = 0;
#pragma unroll 1
    for (int i = start0; i < end; i++) {
        warp_converge(p.wmask);
        if (i < start || fine_quant[i] <= 0) continue;
#pragma unroll 1
        for (int c = 0; c < C; c++) {
            s32 q2 = (s32)rd_bits(rd, (u32)fine_quant[i]);
            s32 offset = sub16((shl32(q2, 10) + 512) >> fine_quant[i], 512);
            fine_acc[c][i] = (facc_t)(fine_acc[c][i] + offset);
        }
    }

    // ---- PVQ / band structure (ref quant_all_bands :1754)
    uint8_t cmask[2 * kNbEBands];
#pragma unroll 1
    for (int i = 0; i < 2 * kNbEBands; i++) cmask[i] = 0;
    all_bands(p, cmask, pulses, shortBlocks, dual_stereo, intensity, tf_res,
              (int)rd.storage * (8 << 3) - anti_collapse_rsv, balance, codedBands);

    int anti_collapse_on = 0;
    if (anti_collapse_rsv > 0) anti_collapse_on = (int)rd_bits(rd, 1);

    // ---- final energy bits (ref unquant_energy_finalise :3681)
    {
        int bits_left = (int)rd.storage * 8 - rd_tell(rd);
#pragma unroll 1
        for (int prio = 0; prio < 2; prio++) {
#pragma unroll 1
            for (int i = start; i < end && bits_left >= C; i++) {
                if (fine_quant[i] >= 8 || fine_priority[i] != prio) continue;
#pragma unroll 1
                for (int c = 0; c < C; c++) {
                    s32 q2 = (s32)rd_bits(rd, 1);
                    s32 offset = ((s32)shl16(q2, 10) - 512) >> (fine_quant[i] + 1);
                    fine_acc[c][i] = (facc_t)(fine_acc[c][i] + offset);
                    bits_left--;
                }
            }
        }
    }

#pragma unroll 1
    for (int c = 0; c < 2; c++)
#pragma unroll 1
        for (int i = 0; i < kNbEBands; i++) {
            hdr.fine[c][i] = (int16_t)fine_acc[c][i];
            if (c >= C || i < start || i >= end) hdr.pq[c][i] = 0;
        }
#pragma unroll 1
    for (int i = 0; i < kNbEBands; i++) hdr.pulses[i] = (int16_t)pulses[i];
#pragma unroll 1
    for (int i = 0; i < 2 * kNbEBands; i++) hdr.cmask[i] = cmask[i];
    if (silence) flags |= FF_SILENCE;
    if (isTransient) flags |= FF_TRANSIENT;
    if (intra) flags |= FF_INTRA;
    if (anti_collapse_on) flags |= FF_ANTICOLLAPSE;
    if (rd.error) flags |= FF_EC_ERROR;
    hdr.flags = (uint8_t)flags;
    hdr.pf_pitch = (int16_t)pf_pitch;
    hdr.pf_gain = (int16_t)pf_gain;
    hdr.pf_tapset = (uint8_t)pf_tapset;
    hdr.n_ops = (uint16_t)p.n_ops;
    hdr.n_leaves = (uint16_t)p.n_leaves;
    hdr.spread = (uint8_t)spread;
    hdr.lcg_steps = p.lcg_steps;
    hdr.final_rng = rd.rng;
    int status = kShortMdct << LM;
    if (rd_tell(rd) > 8 * (int)rd.storage) status = OPUSB_INTERNAL_ERROR;  // ref :2441
    if (p.overflow) status = OPUSB_SCRIPT_OVERFLOW;
    hdr.status = status;
    return status;
}

}  // namespace opusb200
