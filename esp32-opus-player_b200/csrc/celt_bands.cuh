// celt_bands.cuh -- band reconstruction stage: one warp replays one frame's BandOp script.
//
// SURVEY section 8 rows a4 (folding / noise fill), a5 (renormalise_vector), a6 (stereo_merge, haar1,
// (de)interleave_hadamard) and a7 (anti_collapse).  All operands live in the warp's int16 workspace W
// in shared memory (layout in celt_script.cuh); lanes stride over the vector, reductions are warp
// reductions, the LCG noise sequence is generated with jump-ahead so that lane j owns draw j.
#pragma once
#include "celt_script.cuh"
#include "celt_tables.cuh"
#include "lanes.cuh"

namespace opusb200 {

// Output of the per-stream sequencer for one frame: everything a frame needs from its predecessors.
struct FrameEnergy {
    int16_t E[2][kNbEBands];     // band log-energy after this frame's coarse+fine+final (ref oldBandE)
    int16_t pmin[2][kNbEBands];  // min(oldLogE, oldLogE2) before this frame (anti_collapse input, ref :1044-1050)
    uint32_t seed;               // noise seed entering the frame = previous frame's final range (ref :2359, 2437)
    int16_t pf_period, pf_period_old;  // post-filter state entering the frame (ref :2381-2402)
    int16_t pf_gain, pf_gain_old;
    uint8_t pf_tapset, pf_tapset_old;
    uint8_t skip;                // bit0: no CELT layer this frame; bit1: SILK layer present; bit2: frame failed, no output; bit3: frame not decoded at all
    uint8_t pad;
    uint16_t celt_n;             // samples the CELT layer produces in this frame (120 << LM), 0 when inactive
    uint16_t pad2;
};
enum { FS_NO_CELT = 1, FS_HAS_SILK = 2, FS_FAILED = 4, FS_DROPPED = 8 };

// advance the LCG by cnt (1..kLanes) draws
OPUS_HD u32 lcg_jump(u32 seed, int cnt) { return kLcgMul[cnt - 1] * seed + kLcgAdd[cnt - 1]; }

// ref renormalise_vector (src/celt.cpp:797)
OPUS_HD void w_renormalise(const Lanes& L, int16_t* X, int N, s16 gain) {
    s32 acc = 0;
#pragma unroll 1
    for (int j = L.lane; j < N; j += kLanes) acc = add_w(acc, mul16(X[j], X[j]));
    s32 E = add_w(1, L.sum(acc));
    int k = ilog2p(E) >> 1;
    s32 t = vshr32(E, 2 * (k - 7));
    s16 g = (s16)mul16_p15(rsqrt_norm(t), gain);
#pragma unroll 1
    for (int j = L.lane; j < N; j += kLanes) X[j] = (s16)pshr32(mul16(g, X[j]), k + 1);
    L.sync();
}

// ref haar1 (src/celt.cpp:1202); stride is a power of two (1 << k or the block count B)
OPUS_HD void w_haar(const Lanes& L, int16_t* X, int N0, int stride) {
    const int pairs = (N0 >> 1) * stride;
    const int sh = ilog2p(stride);
#pragma unroll 1
    for (int t = L.lane; t < pairs; t += kLanes) {
        const int i = t & (stride - 1), j = t >> sh;
        const int ia = ((2 * j) << sh) + i, ib = ia + stride;
        s32 t1 = mul16(23170, X[ia]), t2 = mul16(23170, X[ib]);
        X[ia] = (s16)pshr32(t1 + t2, 15);
        X[ib] = (s16)pshr32(t1 - t2, 15);
    }
    L.sync();
}

// inverse of the ordery permutation (ref ordery_table, src/celt.cpp:1160), same packing as kOrdery
OPUS_TABLE uint8_t kOrderyInv[30] = {
    1, 0, 1, 3, 2, 0, 1, 5, 7, 3, 2, 6, 4, 0, 1, 9, 13, 5, 7, 15, 11, 3, 2, 10, 14, 6, 4, 12, 8, 0,
};

// ref deinterleave_hadamard / interleave_hadamard (src/celt.cpp:1162, 1183): a permutation of N0*stride values;
// stride is a power of two
OPUS_HD void w_hadamard_perm(const Lanes& L, int16_t* X, int N0, int stride, int hadamard, int inverse) {
    const int N = N0 * stride;
    const uint8_t* ordery = kOrdery + stride - 2;
    const uint8_t* orderyinv = kOrderyInv + stride - 2;
    const int sh = ilog2p(stride);
    const SmallDiv byN0(N0);
    enum { kPer = (kMaxLeaf + kLanes - 1) / kLanes };
    s16 tmp[kPer];
    // destination index d -> source index
    int q = 0;
#pragma unroll 1
    for (int d = L.lane; d < N; d += kLanes, q++) {
        int src;
        if (!inverse) {  // tmp[ord(i)*N0 + j] = X[j*stride + i]
            const int oi = byN0.div(d), j = d - oi * N0;
            const int i = hadamard ? orderyinv[oi] : oi;
            src = (j << sh) + i;
        } else {         // tmp[j*stride + i] = X[ord(i)*N0 + j]
            const int j = d >> sh, i = d & (stride - 1);
            src = (hadamard ? ordery[i] : i) * N0 + j;
        }
        tmp[q] = X[src];
    }
    L.sync();
    q = 0;
#pragma unroll 1
    for (int d = L.lane; d < N; d += kLanes, q++) X[d] = tmp[q];
    L.sync();
}

// ref stereo_merge (src/celt.cpp:1113)
OPUS_HD void w_stereo_merge(const Lanes& L, int16_t* X, int16_t* Y, s16 mid, int N) {
    s32 a = 0, b = 0;
#pragma unroll 1
    for (int j = L.lane; j < N; j += kLanes) {
        a = add_w(a, mul16(Y[j], X[j]));
        b = add_w(b, mul16(Y[j], Y[j]));
    }
    s32 xp = L.sum(a), side = L.sum(b);
    xp = mul16x32_q15(mid, xp);
    s16 mid2 = (s16)(mid >> 1);
    s32 El = mul16(mid2, mid2) + side - 2 * xp;
    s32 Er = mul16(mid2, mid2) + side + 2 * xp;
    if (Er < 161061 || El < 161061) {  // QCONST32(6e-4f, 28)
#pragma unroll 1
        for (int j = L.lane; j < N; j += kLanes) Y[j] = X[j];
        L.sync();
        return;
    }
    int kl = ilog2p(El) >> 1, kr = ilog2p(Er) >> 1;
    s32 lgain = rsqrt_norm(vshr32(El, (kl - 7) << 1));
    s32 rgain = rsqrt_norm(vshr32(Er, (kr - 7) << 1));
    if (kl < 7) kl = 7;
    if (kr < 7) kr = 7;
#pragma unroll 1
    for (int j = L.lane; j < N; j += kLanes) {
        s16 l = (s16)mul16_p15(mid, X[j]);
        s16 r = Y[j];
        X[j] = (s16)pshr32(mul16(lgain, sub16(l, r)), kl + 1);
        Y[j] = (s16)pshr32(mul16(rgain, add16(l, r)), kr + 1);
    }
    L.sync();
}

// ---- coded PVQ leaves: one lane expands one leaf (index -> pulses -> unit-norm rotated vector) ----
// ref cwrsi (src/celt.cpp:2545): codeword index -> signed pulses y[0..n), returns sum y^2
// One flat loop, one table probe per trip: at a fresh coefficient strip the sign (index >= U(n, k+1)) and start the
// search for the largest k' <= k with U(n, k') <= index; the coefficient holds k - k' pulses.  (The reference nests a
// search loop inside the coefficient loop; flat, the lanes of a warp that expand leaves of similar size differ only in
// their total trip count, not at every coefficient.)
OPUS_HD s32 pvq_expand(int n, int k, u32 i, int16_t* y) {
    s32 yy = 0;
    int s = 0, k0 = k;
    bool fresh = true;
#pragma unroll 1
    while (n > 2) {
        if (fresh) {
            const u32 p1 = n <= k ? kPvqU[kPvqRow[n] + k + 1] : kPvqU[kPvqRow[k + 1] + n];  // U(n, k + 1)
            s = -(int)(i >= p1);
            i -= p1 & (u32)s;
            k0 = k;
            if (k >= n && kPvqU[kPvqRow[n] + n] > i) k = n - 1;  // lots of pulses, but fewer than n remain: skip to U(n, n-1)
            fresh = false;
        }
        const u32 p = n <= k ? kPvqU[kPvqRow[n] + k] : kPvqU[kPvqRow[k] + n];  // U(n, k)
        if (p > i) { k--; continue; }
        i -= p;
        const s16 val = (s16)((k0 - k + s) ^ s);
        *y++ = val;
        yy += mul16(val, val);
        n--;
        fresh = true;
    }
    u32 p = 2 * (u32)k + 1;
    s = -(int)(i >= p);
    i -= p & (u32)s;
    k0 = k;
    k = (int)((i + 1) >> 1);
    if (k) i -= 2 * (u32)k - 1;
    s16 val = (s16)((k0 - k + s) ^ s);
    *y++ = val;
    yy += mul16(val, val);
    s = -(int)i;
    val = (s16)((k + s) ^ s);
    *y = val;
    yy += mul16(val, val);
    return yy;
}

// ref exp_rotation1 / exp_rotation (src/celt.cpp:684, 707), decoder direction (dir = -1).  Each pass is a
// recurrence with per-step rounding: serial along the vector, so a lane keeps the carried value in a register.
OPUS_HD void leaf_rotate1(int16_t* X, int len, int stride, s16 c, s16 s) {
    const s16 ms = (s16)(-(s32)s);
    if (stride == 1) {
        // the value a step hands to the next one stays in a register: one load and one store per step
        if (len < 2) return;
        s16 v = X[0];
#pragma unroll 1
        for (int i = 0; i < len - 1; i++) {
            const s16 x2 = X[i + 1];
            X[i] = (s16)pshr32(mul16(c, v) + mul16(ms, x2), 15);
            v = (s16)pshr32(mul16(c, x2) + mul16(s, v), 15);
        }
        X[len - 1] = v;
        if (len < 3) return;
        s16 w = X[len - 2];
#pragma unroll 1
        for (int i = len - 3; i >= 0; i--) {
            const s16 x1 = X[i];
            X[i + 1] = (s16)pshr32(mul16(c, w) + mul16(s, x1), 15);
            w = (s16)pshr32(mul16(c, x1) + mul16(ms, w), 15);
        }
        X[0] = w;
        return;
    }
#pragma unroll 1
    for (int i = 0; i < len - stride; i++) {
        s16 x1 = X[i], x2 = X[i + stride];
        X[i + stride] = (s16)pshr32(mul16(c, x2) + mul16(s, x1), 15);
        X[i] = (s16)pshr32(mul16(c, x1) + mul16(ms, x2), 15);
    }
#pragma unroll 1
    for (int i = len - 2 * stride - 1; i >= 0; i--) {
        s16 x1 = X[i], x2 = X[i + stride];
        X[i + stride] = (s16)pshr32(mul16(c, x2) + mul16(s, x1), 15);
        X[i] = (s16)pshr32(mul16(c, x1) + mul16(ms, x2), 15);
    }
}
OPUS_HD void leaf_rotate(int16_t* X, int len, int stride, int K, int spread) {
    if (2 * K >= len || spread == 0) return;
    const int factor = spread == 1 ? 15 : (spread == 2 ? 10 : 5);
    s16 gain = (s16)mul32_q31((s32)mul16(32767, len), rcp_q(len + factor * K));  // celt_div
    s16 theta = (s16)(mul16_q15(gain, gain) >> 1);
    s16 c = cos_norm((s32)theta);
    s16 s = cos_norm((s32)sub16(32767, theta));
    int stride2 = 0;
    if (len >= 8 * stride) {
        stride2 = 1;
        while ((stride2 * stride2 + stride2) * stride + (stride >> 2) < len) stride2++;
    }
    len = (int)((u32)len / (u32)stride);
#pragma unroll 1
    for (int i = 0; i < stride; i++) {
        if (stride2) leaf_rotate1(X + i * len, len, stride2, s, c);
        leaf_rotate1(X + i * len, len, 1, c, s);
    }
}

// ref alg_unquant (src/celt.cpp:782) for one coded leaf: codeword index -> pulses -> unit-norm, spread-rotated vector
// in y[0 .. o.n).
OPUS_HD void leaf_expand(const BandOp& o, int spread, int16_t* y) {
    const int N = o.n, K = o.off2, B = o.a;
    const s32 Ryy = pvq_expand(N, K, o.aux, y);
    // normalise_residual (ref :745)
    const int k = ilog2p(Ryy) >> 1;
    const s32 t = vshr32(Ryy, 2 * (k - 7));
    const s16 g = (s16)mul16_p15(rsqrt_norm(t), o.g);
#pragma unroll 1
    for (int i = 0; i < N; i++) y[i] = (s16)pshr32(mul16(g, y[i]), k + 1);
    leaf_rotate(y, N, B, K, spread);
}
// Cost class of leaf_expand, 0 .. kLeafClasses-1 in quarter-octaves of its approximate trip count (the codeword walk,
// the scaling, the rotation passes): leaves of one class take within ~1.2x of each other.
enum { kLeafClasses = 32 };
OPUS_HD int leaf_class(const BandOp& o, int spread) {
    const int N = o.n, K = o.off2;
    const bool rotates = !(2 * K >= N || spread == 0);
    const int cost = 2 * N + K + (rotates ? 4 * N : 0);  // >= 5
    const int lg = ilog2p(cost);                         // >= 2
    const int c = 4 * lg + ((cost >> (lg - 2)) & 3) - 10;
    return c < 0 ? 0 : (c >= kLeafClasses ? kLeafClasses - 1 : c);
}

// Every coded leaf of the frame, one lane per leaf, straight into the workspace (host simulation, and the GPU fallback
// when a step has more leaves than the sorted list holds).  `last` points at the first leaf record (leaves are stored
// downwards from the end of the frame's op buffer).
OPUS_HD void leaf_stage(const Lanes& L, int16_t* W, const BandOp* last, int n_leaves, int spread) {
#pragma unroll 1
    for (int l0 = 0; l0 < n_leaves; l0 += kLanes) {
        const int l = l0 + L.lane;
        if (l < n_leaves) {
            const BandOp o = last[-l];
            leaf_expand(o, spread, W + o.off);
        }
    }
    L.sync();
}

// Replays the script of one frame.  `seed` = noise seed at frame start (= previous frame's final range).
// Returns the seed after all fills (input of anti_collapse, ref src/celt.cpp:2371).
OPUS_HD u32 band_stage_ops(const Lanes& L, int16_t* W, const BandOp* ops, int n_ops, u32 seed) {
    // ops are fetched kLanes at a time (one 16-byte load per lane) and broadcast, so the replay loop
    // never waits on a dependent global load per op
    BandOp mine = BandOp();
#pragma unroll 1
    for (int k = 0; k < n_ops; k++) {
        if ((k % kLanes) == 0 && k + L.lane < n_ops) mine = ops[k + L.lane];
        const BandOp o = L.bcast_op(mine, k % kLanes);
        int16_t* X = W + o.off;
        const int n = o.n;
        switch (o.code) {
            case OP_NOISE:
#pragma unroll 1
                for (int j0 = 0; j0 < n; j0 += kLanes) {
                    int cnt = imin(kLanes, n - j0);
                    if (L.lane < cnt) {
                        u32 s = lcg_jump(seed, L.lane + 1);
                        X[j0 + L.lane] = (s16)((s32)s >> 20);
                    }
                    seed = lcg_jump(seed, cnt);
                }
                L.sync();
                w_renormalise(L, X, n, o.g);
                break;
            case OP_FOLD: {
                const int16_t* lb = W + o.off2;
#pragma unroll 1
                for (int j0 = 0; j0 < n; j0 += kLanes) {
                    int cnt = imin(kLanes, n - j0);
                    if (L.lane < cnt) {
                        u32 s = lcg_jump(seed, L.lane + 1);
                        s32 tmp = (s & 0x8000u) ? 4 : -4;  // QCONST16(1/256, 10)
                        X[j0 + L.lane] = (s16)(lb[j0 + L.lane] + tmp);
                    }
                    seed = lcg_jump(seed, cnt);
                }
                L.sync();
                w_renormalise(L, X, n, o.g);
            } break;
            case OP_COPY: {
                const int16_t* src = W + o.off2;
#pragma unroll 1
                for (int j = L.lane; j < n; j += kLanes) X[j] = src[j];
                L.sync();
            } break;
            case OP_HAAR: w_haar(L, X, n, o.a); break;
            case OP_DEINT: w_hadamard_perm(L, X, n, o.a, o.g, 0); break;
            case OP_INT: w_hadamard_perm(L, X, n, o.a, o.g, 1); break;
            case OP_SCALE_OUT: {
                int16_t* dst = W + o.off2;
#pragma unroll 1
                for (int j = L.lane; j < n; j += kLanes) dst[j] = (s16)mul16_q15(o.g, X[j]);
                L.sync();
            } break;
            case OP_SETW:
                if (L.lane == 0) X[0] = o.g;
                L.sync();
                break;
            case OP_MERGE: w_stereo_merge(L, X, W + o.off2, o.g, n); break;
            case OP_NEG:
#pragma unroll 1
                for (int j = L.lane; j < n; j += kLanes) X[j] = (s16)(-(s32)X[j]);
                L.sync();
                break;
            case OP_N2: {
                if (L.lane == 0) {
                    int16_t* x2 = X;
                    int16_t* y2 = W + o.off2;
                    const int sign = 1 - 2 * (o.a & 1);
                    y2[0] = (s16)(-sign * x2[1]);
                    y2[1] = (s16)(sign * x2[0]);
                    int16_t* Xc = (o.a & 2) ? y2 : x2;  // channel 0
                    int16_t* Yc = (o.a & 2) ? x2 : y2;  // channel 1
                    Xc[0] = (s16)mul16_q15(o.g, Xc[0]);
                    Xc[1] = (s16)mul16_q15(o.g, Xc[1]);
                    Yc[0] = (s16)mul16_q15(o.g2, Yc[0]);
                    Yc[1] = (s16)mul16_q15(o.g2, Yc[1]);
                    s16 tmp = Xc[0];
                    Xc[0] = (s16)sub16(tmp, Yc[0]);
                    Yc[0] = add16(tmp, Yc[0]);
                    tmp = Xc[1];
                    Xc[1] = (s16)sub16(tmp, Yc[1]);
                    Yc[1] = add16(tmp, Yc[1]);
                }
                L.sync();
            } break;
            case OP_ZERO:
#pragma unroll 1
                for (int j = L.lane; j < n; j += kLanes) X[j] = 0;
                L.sync();
                break;
            case OP_AVG: {
                const int16_t* src = W + o.off2;
#pragma unroll 1
                for (int j = L.lane; j < n; j += kLanes) X[j] = (s16)(((s32)X[j] + (s32)src[j]) >> 1);
                L.sync();
            } break;
            default: break;
        }
    }
    return seed;
}

// ref anti_collapse (src/celt.cpp:1010).  X = W + W_XY holds [C][N].
OPUS_HD void band_stage_anti_collapse(const Lanes& L, int16_t* W, const CeltFrameHdr& h, const FrameEnergy& fe, u32 seed) {
    const int LM = h.LM, C = h.C, size = kShortMdct << LM;
    for (int i = h.start; i < h.end; i++) {
        const int N0 = kEBands[i + 1] - kEBands[i];
        const int depth = (int)((u32)(1 + h.pulses[i]) / (u32)N0) >> LM;
        s32 thresh32 = exp2_q((s16)(-(s32)shl16(depth, 10 - 3))) >> 1;
        s16 thresh = (s16)mul16x32_q15(16384, imin(32767, thresh32));
        s32 t = N0 << LM;
        const int shift = ilog2p(t) >> 1;
        t = shl32(t, (7 - shift) << 1);
        const s16 sqrt_1 = rsqrt_norm(t);
#pragma unroll 1
        for (int c = 0; c < C; c++) {
            s32 Ediff = (s32)fe.E[c][i] - (s32)fe.pmin[c][i];
            Ediff = imax(0, Ediff);
            s16 r;
            if (Ediff < 16384) {
                s32 r32 = exp2_q((s16)(-(s32)(s16)Ediff)) >> 1;
                r = (s16)(2 * imin(16383, r32));
            } else {
                r = 0;
            }
            if (LM == 3) r = (s16)mul16_q14(23170, imin(23169, r));
            r = (s16)(imin(thresh, r) >> 1);
            r = (s16)(mul16_q15(sqrt_1, r) >> shift);
            int16_t* X = W + W_XY + c * size + (kEBands[i] << LM);
            int renorm = 0;
#pragma unroll 1
            for (int k = 0; k < (1 << LM); k++) {
                if (!(h.cmask[i * C + c] & (1 << k))) {
#pragma unroll 1
                    for (int j0 = 0; j0 < N0; j0 += kLanes) {
                        int cnt = imin(kLanes, N0 - j0);
                        if (L.lane < cnt) {
                            u32 s = lcg_jump(seed, L.lane + 1);
                            X[((j0 + L.lane) << LM) + k] = (s & 0x8000u) ? r : (s16)(-(s32)r);
                        }
                        seed = lcg_jump(seed, cnt);
                    }
                    renorm = 1;
                }
            }
            L.sync();
            if (renorm) w_renormalise(L, X, N0 << LM, 32767);
        }
    }
}

}  // namespace opusb200
