// fixmath.cuh -- integer arithmetic primitives of the decode path.
//
// Every helper states the reference macro/function whose *exact* integer result it reproduces
// (schreibfaul1/ESP32-Opus-Player, src/celt.h:252-378, 401-511; src/celt.cpp:926-945, 3086-3202).
// Bit-exactness rules: `>>` on signed values is arithmetic, 16-bit operands are
// truncated *before* multiplying, wrap-around adds go through unsigned arithmetic.
#pragma once
#include <stdint.h>

// Under nvcc the stage bodies are device code only (the product has no CPU decode path); a plain
// C++ compiler sees them as ordinary inline functions, which is how tests/hostsim checks their logic.
#if defined(__CUDACC__)
#define OPUS_HD __device__ __forceinline__
#else
#define OPUS_HD inline
#endif

// Small constant tables live at namespace scope, never as function-local arrays: a local `const T x[] = {...}` that is
// indexed at run time is rebuilt on the thread's stack at every call, and nvcc 12.9 for sm_100a was caught sharing that
// stack slot with a LIVE array of the caller (the parser's collapse masks came back holding the eight words of the
// ec_tell_frac correction table; found with dm_debug_frame_script + tools/debug_script.py).
#if defined(__CUDACC__)
#define OPUS_LOCAL_TABLE static __device__ const
#else
#define OPUS_LOCAL_TABLE static const
#endif

#if !defined(__CUDACC__)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
struct uint2 { uint32_t x, y; };
#endif

namespace opusb200 {

typedef int16_t s16;
typedef int32_t s32;
typedef int64_t s64;
typedef uint32_t u32;

OPUS_HD s32 imin(s32 a, s32 b) { return a < b ? a : b; }
OPUS_HD s32 imax(s32 a, s32 b) { return a > b ? a : b; }
OPUS_HD u32 umin(u32 a, u32 b) { return a < b ? a : b; }

// 16 int16 values (32 bytes, 16-byte aligned) between memory and a local block: two 128-bit accesses on the device
OPUS_HD void ld_block16(int16_t* dst, const int16_t* src) {
#if defined(__CUDA_ARCH__)
    reinterpret_cast<uint4*>(dst)[0] = reinterpret_cast<const uint4*>(src)[0];
    reinterpret_cast<uint4*>(dst)[1] = reinterpret_cast<const uint4*>(src)[1];
#else
    memcpy(dst, src, 32);
#endif
}

// ---- warp convergence points for thread-per-frame kernels ----
// The entropy stage runs one frame per lane through data-dependent control flow; left alone the lanes of a
// warp drift apart and never meet again (measured: 2.6 of 32 lanes active).  These helpers make the lanes named
// in `m` meet at loop heads whose trip count is the same for every lane.  Host build (tests/hostsim): one lane.
#if defined(__CUDA_ARCH__)
OPUS_HD void warp_converge(unsigned m) { __syncwarp(m); }
OPUS_HD bool warp_any(unsigned m, bool v) { return __any_sync(m, v) != 0; }
OPUS_HD int warp_min(unsigned m, int v) { return __reduce_min_sync(m, v); }
#else
OPUS_HD void warp_converge(unsigned) {}
OPUS_HD bool warp_any(unsigned, bool v) { return v; }
OPUS_HD int warp_min(unsigned, int v) { return v; }
#endif

// 32 - clz(x): ref EC_ILOG (src/celt.h:250); undefined for 0 in the reference, we return 0.
OPUS_HD int ilog32(u32 x) {
#if defined(__CUDACC__)
    return 32 - __clz((int)x);
#else
    return x ? 32 - __builtin_clz(x) : 0;
#endif
}
// ref celt_ilog2 (src/celt.h:469)
OPUS_HD int ilog2p(s32 x) { return ilog32((u32)x) - 1; }

// ref MULT16_16 (src/celt.h:338): both operands truncated to int16 first
OPUS_HD s32 mul16(s32 a, s32 b) { return (s32)(s16)a * (s32)(s16)b; }
OPUS_HD s32 mul16_q15(s32 a, s32 b) { return mul16(a, b) >> 15; }              // MULT16_16_Q15 :355
OPUS_HD s32 mul16_q14(s32 a, s32 b) { return mul16(a, b) >> 14; }              // MULT16_16_Q14 :354
OPUS_HD s32 mul16_p15(s32 a, s32 b) { return (16384 + mul16(a, b)) >> 15; }    // MULT16_16_P15 :359
OPUS_HD s32 frac_mul16(s32 a, s32 b) { return (16384 + mul16(a, b)) >> 15; }   // FRAC_MUL16 :378
// ref MULT16_32_Q15(a16, b32) == S_MUL(b32, a16) (src/celt.h:263, 192): 48-bit product >> 15
OPUS_HD s32 mul16x32_q15(s32 a16, s32 b32) { return (s32)(((s64)(s16)a16 * (s64)b32) >> 15); }
// ref MULT32_32_Q31 (src/celt.h:266)
OPUS_HD s32 mul32_q31(s32 a, s32 b) { return (s32)(((s64)a * (s64)b) >> 31); }

OPUS_HD s32 shl32(s32 a, int s) { return (s32)((u32)a << s); }                 // SHL32 :292
OPUS_HD s16 shl16(s32 a, int s) { return (s16)((u32)(uint16_t)a << s); }       // SHL16 :288
OPUS_HD s32 pshr32(s32 a, int s) { return (a + ((1 << s) >> 1)) >> s; }        // PSHR32 :295
OPUS_HD s32 vshr32(s32 a, int s) { return s > 0 ? (a >> s) : shl32(a, -s); }   // VSHR32 :297
OPUS_HD s16 add16(s32 a, s32 b) { return (s16)((s32)(s16)a + (s32)(s16)b); }   // ADD16 :317 (truncates)
OPUS_HD s32 sub16(s32 a, s32 b) { return (s32)(s16)a - (s32)(s16)b; }          // SUB16 :319 (does NOT truncate)
OPUS_HD s32 add_w(s32 a, s32 b) { return (s32)((u32)a + (u32)b); }             // ADD32_ovflw :326
OPUS_HD s32 sub_w(s32 a, s32 b) { return (s32)((u32)a - (u32)b); }             // SUB32_ovflw :328
OPUS_HD s32 neg_w(s32 a) { return (s32)(0u - (u32)a); }                        // NEG32_ovflw :331
OPUS_HD s32 sat_sym(s32 x, s32 a) { return x > a ? a : (x < -a ? -a : x); }    // SATURATE :303
OPUS_HD s16 sat16(s32 x) { return (s16)(x > 32767 ? 32767 : (x < -32768 ? -32768 : x)); }  // SAT16 :401 / SATURATE16 :305

// ref celt_rsqrt_norm (src/celt.cpp:3109): Q16 in [0.25,1) -> Q14 reciprocal square root
OPUS_HD s16 rsqrt_norm(s32 x) {
    s16 n = (s16)(x - 32768);
    s16 r = add16(23557, mul16_q15(n, add16(-13490, mul16_q15(n, 6713))));
    s16 r2 = (s16)mul16_q15(r, r);
    s16 y = shl16(sub16(add16(mul16_q15(r2, n), r2), 16384), 1);
    return add16(r, mul16_q15(r, mul16_q15(y, sub16(mul16_q15(y, 12288), 16384))));
}

// ref celt_sqrt (src/celt.cpp:3131)
OPUS_HD s32 sqrt_q(s32 x) {
    if (x == 0) return 0;
    if (x >= 1073741824) return 32767;
    int k = (ilog2p(x) >> 1) - 7;
    x = vshr32(x, 2 * k);
    s16 n = (s16)(x - 32768);
    s32 rt = add16(23175, mul16_q15(n, add16(11561, mul16_q15(n, add16(-3011, mul16_q15(n, add16(1699, mul16_q15(n, -664))))))));
    return vshr32(rt, 7 - k);
}

// ref _celt_cos_pi_2 / celt_cos_norm (src/celt.cpp:3151, 3161)
OPUS_HD s16 cos_pi_2(s16 x) {
    s16 x2 = (s16)mul16_p15(x, x);
    s32 inner = (s32)sub16(32767, x2) + mul16_p15(x2, -7651 + mul16_p15(x2, 8277 + mul16_p15(-626, x2)));
    return add16(1, imin(32766, inner));
}
OPUS_HD s16 cos_norm(s32 x) {
    x = x & 0x0001ffff;
    if (x > (1 << 16)) x = (1 << 17) - x;
    if (x & 0x00007fff) {
        if (x < (1 << 15)) return cos_pi_2((s16)x);
        return (s16)(-(s32)cos_pi_2((s16)(65536 - x)));
    }
    if (x & 0x0000ffff) return 0;
    if (x & 0x0001ffff) return -32767;
    return 32767;
}

// ref celt_rcp (src/celt.cpp:3181): Q15 in, Q16 out
OPUS_HD s32 rcp_q(s32 x) {
    int i = ilog2p(x);
    s16 n = (s16)(vshr32(x, i - 15) - 32768);
    s16 r = add16(30840, mul16_q15(-15420, n));
    r = (s16)sub16(r, mul16_q15(r, add16(mul16_q15(r, n), add16(r, -32768))));
    r = (s16)sub16(r, add16(1, mul16_q15(r, add16(mul16_q15(r, n), add16(r, -32768)))));
    return vshr32((s32)r, i - 16);
}

// ref celt_exp2_frac / celt_exp2 (src/celt.h:494, 501)
OPUS_HD s32 exp2_frac(s16 x) {
    s16 frac = shl16(x, 4);
    return add16(16383, mul16_q15(frac, add16(22804, mul16_q15(frac, add16(14819, mul16_q15(10204, frac))))));
}
OPUS_HD s32 exp2_q(s16 x) {
    s32 integer = x >> 10;
    if (integer > 14) return 0x7f000000;
    if (integer < -15) return 0;
    s16 frac = (s16)exp2_frac((s16)(x - shl16(integer, 10)));
    return vshr32((s32)frac, -integer - 2);
}

// ref bitexact_cos / bitexact_log2tan (src/celt.cpp:926, 937)
OPUS_HD s16 bitexact_cos(s16 x) {
    s32 tmp = (4096 + (s32)x * x) >> 13;
    s16 x2 = (s16)tmp;
    x2 = (s16)((32767 - x2) + frac_mul16(x2, (-7651 + frac_mul16(x2, (8277 + frac_mul16(-626, x2))))));
    return (s16)(1 + x2);
}
OPUS_HD s32 bitexact_log2tan(s32 isin, s32 icos) {
    int lc = ilog32((u32)icos), ls = ilog32((u32)isin);
    icos <<= 15 - lc;
    isin <<= 15 - ls;
    return (ls - lc) * (1 << 11) + frac_mul16(isin, frac_mul16(isin, -2597) + 7932) - frac_mul16(icos, frac_mul16(icos, -2597) + 7932);
}

// ref isqrt32 (src/celt.cpp:3086): floor(sqrt(v)), v > 0
OPUS_HD u32 isqrt32(u32 v) {
    u32 g = 0;
    int bshift = (ilog32(v) - 1) >> 1;
    u32 b = 1u << bshift;
    do {
        u32 t = ((g << 1) + b) << bshift;
        if (t <= v) { g += b; v -= t; }
        b >>= 1;
        bshift--;
    } while (bshift >= 0);
    return g;
}

// ref celt_lcg_rand (src/celt.cpp:921)
OPUS_HD u32 lcg_next(u32 seed) { return 1664525u * seed + 1013904223u; }

// d / n and d % n for small operands without the 20-instruction integer divide: rcp = ceil(2^32 / n) is exact
// for d * n < 2^32 (here d < 2048, n <= 176).  n == 1 has no 32-bit reciprocal.
struct SmallDiv {
    u32 n, rcp;
    OPUS_HD explicit SmallDiv(int n_) : n((u32)n_), rcp(n_ > 1 ? 0xFFFFFFFFu / (u32)n_ + 1u : 0u) {}
    OPUS_HD int div(int d) const {
#if defined(__CUDA_ARCH__)
        return n > 1 ? (int)__umulhi((u32)d, rcp) : d;
#else
        return (int)((u32)d / n);
#endif
    }
};

}  // namespace opusb200
