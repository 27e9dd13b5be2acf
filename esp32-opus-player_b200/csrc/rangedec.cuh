// rangedec.cuh -- Opus range decoder, one instance per (stream, frame) thread.
//
// Reproduces ec_dec_* of the reference bit for bit (src/celt.cpp:2627-2792, state src/celt.h:73-86):
// 32-bit rng/val, bytes consumed from the front, raw bits from the back, ec_tell in 1/8 bit.
// The decoder is a plain struct in registers/local memory; the packet lives in global memory (HBM/L2)
// and is only ever read.
#pragma once
#include "fixmath.cuh"

namespace opusb200 {

struct RangeDec {
    const uint8_t* buf;
    u32 sbase;        // shared-memory address of payload byte 0 when the packet was staged there (k_entropy), else unused
    u32 staged;       // payload bytes [0, staged) are in shared memory; 0: everything is read from global memory
    u32 storage;      // packet length in bytes
    u32 end_offs;     // bytes consumed from the back
    u32 end_window;   // raw-bit window
    int nend_bits;    // valid bits in end_window
    int nbits_total;  // bits consumed so far (whole bits)
    u32 offs;         // bytes consumed from the front
    u32 rng, val, ext;
    int rem;
    int error;
};

// Byte i of the payload: from the thread's staging window in shared memory if it holds it (k_entropy copies the first
// kStageBytes of every packet there with one bulk copy per packet), else from global memory.
OPUS_HD int rd_byte_at(const RangeDec& d, u32 i) {
#if defined(__CUDA_ARCH__)
    if (i < d.staged) {
        u32 v;
        asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(d.sbase + i));
        return (int)v;
    }
#endif
    return d.buf[i];
}
OPUS_HD int rd_read_byte(RangeDec& d) { return d.offs < d.storage ? rd_byte_at(d, d.offs++) : 0; }
OPUS_HD int rd_read_byte_end(RangeDec& d) { return d.end_offs < d.storage ? rd_byte_at(d, d.storage - ++d.end_offs) : 0; }

// ref ec_dec_normalize (src/celt.cpp:2649)
OPUS_HD void rd_normalize(RangeDec& d) {
#pragma unroll 1
    while (d.rng <= 0x800000u) {
        d.nbits_total += 8;
        d.rng <<= 8;
        int sym = d.rem;
        d.rem = rd_read_byte(d);
        sym = (sym << 8 | d.rem) >> 1;  // EC_SYM_BITS - EC_CODE_EXTRA = 8 - 7
        d.val = ((d.val << 8) + (255u & ~(u32)sym)) & 0x7fffffffu;
    }
}

// ref ec_dec_init (src/celt.cpp:2666)
OPUS_HD void rd_init(RangeDec& d, const uint8_t* buf, u32 len, u32 sbase = 0, u32 staged = 0) {
    d.buf = buf;
    d.sbase = sbase;
    d.staged = staged;
    d.storage = len;
    d.end_offs = 0;
    d.end_window = 0;
    d.nend_bits = 0;
    d.nbits_total = 32 + 1 - ((32 - 7) / 8) * 8;  // = 9
    d.offs = 0;
    d.rng = 1u << 7;
    d.rem = rd_read_byte(d);
    d.val = d.rng - 1 - (u32)(d.rem >> 1);
    d.ext = 0;
    d.error = 0;
    rd_normalize(d);
}

OPUS_LOCAL_TABLE u32 kTellFracCorr[8] = {35733, 38967, 42495, 46340, 50535, 55109, 60097, 65535};  // ref correction[8], :2628
OPUS_LOCAL_TABLE uint8_t kIcdf210[3] = {2, 1, 0};                                                   // tapset_icdf :587, small_energy_icdf :579

// ref ec_tell (src/celt.h:420), ec_tell_frac (src/celt.cpp:2627)
OPUS_HD int rd_tell(const RangeDec& d) { return d.nbits_total - ilog32(d.rng); }
OPUS_HD u32 rd_tell_frac(const RangeDec& d) {
    u32 nbits = (u32)d.nbits_total << 3;
    int l = ilog32(d.rng);
    u32 r = d.rng >> (l - 16);
    u32 b = (r >> 12) - 8;
    b += r > kTellFracCorr[b];
    return nbits - (u32)((l << 3) + (int)b);
}

// ref ec_decode (src/celt.cpp:2683), ec_decode_bin (:2690), ec_dec_update (:2697)
OPUS_HD u32 rd_decode(RangeDec& d, u32 ft) {
    d.ext = d.rng / ft;
    u32 s = d.val / d.ext;
    return ft - umin(s + 1, ft);
}
OPUS_HD u32 rd_decode_bin(RangeDec& d, u32 bits) {
    d.ext = d.rng >> bits;
    u32 s = d.val / d.ext;
    return (1u << bits) - umin(s + 1u, 1u << bits);
}
OPUS_HD void rd_update(RangeDec& d, u32 fl, u32 fh, u32 ft) {
    u32 s = d.ext * (ft - fh);
    d.val -= s;
    d.rng = fl > 0 ? d.ext * (fh - fl) : d.rng - s;
    rd_normalize(d);
}

// ref ec_dec_bit_logp (src/celt.cpp:2712)
OPUS_HD int rd_bit_logp(RangeDec& d, u32 logp) {
    u32 r = d.rng, v = d.val, s = r >> logp;
    int ret = v < s;
    if (!ret) d.val = v - s;
    d.rng = ret ? s : r - s;
    rd_normalize(d);
    return ret;
}

// ref ec_dec_icdf (src/celt.cpp:2727)
OPUS_HD int rd_icdf(RangeDec& d, const uint8_t* icdf, u32 ftb) {
    u32 s = d.rng, v = d.val, r = s >> ftb, t;
    int ret = -1;
#pragma unroll 1
    do {
        t = s;
        s = r * icdf[++ret];
    } while (v < s);
    d.val = v - s;
    d.rng = t - s;
    rd_normalize(d);
    return ret;
}
// Two-entry table {p0, 0}: one comparison
OPUS_HD int rd_icdf2(RangeDec& d, u32 p0, u32 ftb) {
    const u32 s = d.rng, v = d.val, s1 = (s >> ftb) * p0;
    const int ret = v < s1;
    d.val = ret ? v : v - s1;
    d.rng = ret ? s1 : s - s1;
    rd_normalize(d);
    return ret;
}
// Same for a table in the constant data of the module (never for a local array: it is read through 32-bit words).  The
// table is searched from the front like the reference does, but fetched one aligned word (four entries) at a time: the
// search is latency bound (each probe used to be a dependent byte load) and symbols mostly sit in the first few entries.
// A further word is only read when every entry so far was too large, i.e. while the terminating 0 has not been seen, so
// no byte outside the words of the table is touched.
OPUS_HD int rd_icdf_tab(RangeDec& d, const uint8_t* icdf, u32 ftb) {
    u32 s = d.rng, v = d.val, r = s >> ftb, t;
    const uintptr_t a = reinterpret_cast<uintptr_t>(icdf);
    const u32* wp = reinterpret_cast<const u32*>(a & ~(uintptr_t)3);
    int k = (int)(a & 3);  // byte of the current word that holds the next entry
    u32 w = *wp >> (8 * k);
    int ret = -1;
#pragma unroll 1
    for (;;) {
        t = s;
        s = r * (w & 0xffu);
        ret++;
        if (v >= s) break;
        w >>= 8;
        if (++k == 4) { k = 0; w = *++wp; }
    }
    d.val = v - s;
    d.rng = t - s;
    rd_normalize(d);
    return ret;
}

// ref ec_dec_bits (src/celt.cpp:2773)
OPUS_HD u32 rd_bits(RangeDec& d, u32 bits) {
    u32 window = d.end_window;
    int available = d.nend_bits;
    if ((u32)available < bits) {
#pragma unroll 1
        do {
            window |= (u32)rd_read_byte_end(d) << available;
            available += 8;
        } while (available <= 32 - 8);
    }
    u32 ret = window & ((1u << bits) - 1u);
    window >>= bits;
    available -= (int)bits;
    d.end_window = window;
    d.nend_bits = available;
    d.nbits_total += (int)bits;
    return ret;
}

// ref ec_dec_uint (src/celt.cpp:2747); ft > 1
OPUS_HD u32 rd_uint(RangeDec& d, u32 ft_in) {
    u32 ft = ft_in - 1;
    int ftb = ilog32(ft);
    if (ftb > 8) {
        ftb -= 8;
        u32 f = (ft >> ftb) + 1;
        u32 s = rd_decode(d, f);
        rd_update(d, s, s + 1, f);
        u32 t = s << ftb | rd_bits(d, (u32)ftb);
        if (t <= ft) return t;
        d.error = 1;
        return ft;
    }
    ft++;
    u32 s = rd_decode(d, ft);
    rd_update(d, s, s + 1, ft);
    return s;
}

// ref ec_laplace_decode (src/celt.cpp:3047) with LAPLACE_MINP=1, LAPLACE_NMIN=16
OPUS_HD int rd_laplace(RangeDec& d, u32 fs, int decay) {
    int val = 0;
    u32 fm = rd_decode_bin(d, 15);
    u32 fl = 0;
    if (fm >= fs) {
        val++;
        fl = fs;
        fs = (((32768u - 32u - fs) * (u32)(16384 - decay)) >> 15) + 1;  // ec_laplace_get_freq1 :3041
#pragma unroll 1
        while (fs > 1 && fm >= fl + 2 * fs) {
            fs *= 2;
            fl += fs;
            fs = ((fs - 2) * (u32)decay) >> 15;
            fs += 1;
            val++;
        }
        if (fs <= 1) {
            int di = (int)((fm - fl) >> 1);
            val += di;
            fl += 2u * (u32)di;
        }
        if (fm < fl + fs) val = -val;
        else fl += fs;
    }
    rd_update(d, fl, umin(fl + fs, 32768u), 32768u);
    return val;
}

}  // namespace opusb200
