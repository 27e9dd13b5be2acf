// ogg_demux.hpp -- host-side Ogg page framer and packet assembler.
//
// Replaces, for the batched player, what the reference does with its bundled libogg derivative:
//   page capture + CRC    ogg_sync_pageseek           src/ogg.cpp:839-923
//   page -> segments      ogg_stream_pagein           src/ogg.cpp:969-1096
//   segments -> packets   ogg_stream_packetout        src/ogg.cpp:1138-1190
// Written for throughput on a many-core host feeding a GPU: one OggFramer per file, files spread over
// threads, CRC by slicing-by-8, packets copied once into a per-file arena that the staging code reads.
// Behaviour the player depends on is kept: a page that fails capture or CRC makes the framer skip ahead to
// the next 'O' and report the bytes skipped; a gap in page sequence numbers is reported as a hole before the
// first packet after it; a continued page without its head is dropped up to the next packet boundary; the
// page's granule position is attached to the last packet that ends on it, the end-of-stream flag to the last
// segment of the page.
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

#if defined(__x86_64__) && defined(__GNUC__)
#define OPUSB_CRC_CLMUL 1  // carry-less multiplication for the page checksum (OggCrc::fold), picked at run time
#include <immintrin.h>
#endif

namespace opusb200 {

// CRC-32 of Ogg: polynomial 0x04C11DB7, MSB first, zero init, no final xor (RFC 3533).
//  * Tables, slicing by 16: t_[s][b] is the CRC of byte b followed by s zero bytes, so sixteen look-ups that do not depend on
//    each other advance the CRC by sixteen bytes (the reference's libogg derivative slices by 8, src/ogg.cpp:439-460).
//  * x86 with PCLMULQDQ: carry-less multiplication folds the message 64 bytes at a time.  A 128-bit register holds a piece
//    A(x) of the message (first byte = highest powers, hence the byte reversal on load); moving it D bits towards the end
//    multiplies it by x^D, and A x^D = A_hi (x^(D+64) mod P) + A_lo (x^D mod P) modulo P, two 64 x 32-bit products that fit
//    the register again and are added onto the piece D bits further on.  Four registers leap-frog with D = 512, are folded
//    into one with D = 128, and the last sixteen bytes plus the tail go through the tables.  A start value is the same as
//    that value added onto the first four bytes.  Checked against the bitwise definition in tests/test_ogg_crc.py.
class OggCrc {
public:
    OggCrc() {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t r = i << 24;
            for (int k = 0; k < 8; k++) r = (r & 0x80000000u) ? (r << 1) ^ 0x04C11DB7u : (r << 1);
            t_[0][i] = r;
        }
        for (int s = 1; s < 16; s++)
            for (uint32_t i = 0; i < 256; i++) t_[s][i] = (t_[s - 1][i] << 8) ^ t_[0][t_[s - 1][i] >> 24];
#ifdef OPUSB_CRC_CLMUL
        clmul_ = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("ssse3");
        k128_[0] = xpow(128); k128_[1] = xpow(128 + 64);
        k512_[0] = xpow(512); k512_[1] = xpow(512 + 64);
#endif
    }
    uint32_t update(uint32_t crc, const uint8_t* p, size_t n) const {
#ifdef OPUSB_CRC_CLMUL
        if (clmul_ && n >= 128) return fold(crc, p, n);
#endif
        return tables(crc, p, n);
    }
    // (tests) the table path alone, and whether the folding path is in use on this machine
    uint32_t update_tables(uint32_t crc, const uint8_t* p, size_t n) const { return tables(crc, p, n); }
    bool folding() const {
#ifdef OPUSB_CRC_CLMUL
        return clmul_;
#else
        return false;
#endif
    }

private:
    uint32_t tables(uint32_t crc, const uint8_t* p, size_t n) const {
        while (n >= 16) {
            uint32_t w0, w1, w2, w3;
            memcpy(&w0, p, 4); memcpy(&w1, p + 4, 4); memcpy(&w2, p + 8, 4); memcpy(&w3, p + 12, 4);
            w0 = __builtin_bswap32(w0) ^ crc; w1 = __builtin_bswap32(w1); w2 = __builtin_bswap32(w2); w3 = __builtin_bswap32(w3);
            crc = t_[15][w0 >> 24] ^ t_[14][(w0 >> 16) & 255] ^ t_[13][(w0 >> 8) & 255] ^ t_[12][w0 & 255] ^
                  t_[11][w1 >> 24] ^ t_[10][(w1 >> 16) & 255] ^ t_[9][(w1 >> 8) & 255] ^ t_[8][w1 & 255] ^
                  t_[7][w2 >> 24] ^ t_[6][(w2 >> 16) & 255] ^ t_[5][(w2 >> 8) & 255] ^ t_[4][w2 & 255] ^
                  t_[3][w3 >> 24] ^ t_[2][(w3 >> 16) & 255] ^ t_[1][(w3 >> 8) & 255] ^ t_[0][w3 & 255];
            p += 16;
            n -= 16;
        }
        while (n--) crc = (crc << 8) ^ t_[0][(crc >> 24) ^ *p++];
        return crc;
    }
#ifdef OPUSB_CRC_CLMUL
    static uint32_t xpow(int e) {  // x^e mod P, e >= 32
        uint32_t r = 0x04C11DB7u;  // x^32
        for (int i = 32; i < e; i++) r = (r & 0x80000000u) ? (r << 1) ^ 0x04C11DB7u : (r << 1);
        return r;
    }
    // (macros, not lambdas: a lambda does not inherit the target attribute of the function around it)
#define OPUSB_CRC_LOAD(q) _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)(q)), rev)
#define OPUSB_CRC_STEP(a, k, next) _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(a, k, 0x11), _mm_clmulepi64_si128(a, k, 0x00)), next)
    __attribute__((target("pclmul,ssse3"))) uint32_t fold(uint32_t crc, const uint8_t* p, size_t n) const {
        const __m128i rev = _mm_set_epi8(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
        const __m128i k512 = _mm_set_epi64x((long long)k512_[1], (long long)k512_[0]);
        const __m128i k128 = _mm_set_epi64x((long long)k128_[1], (long long)k128_[0]);
        // the start value goes onto the first four bytes = the top 32 bits of the first register
        __m128i a0 = _mm_xor_si128(OPUSB_CRC_LOAD(p), _mm_set_epi32((int)crc, 0, 0, 0));
        __m128i a1 = OPUSB_CRC_LOAD(p + 16), a2 = OPUSB_CRC_LOAD(p + 32), a3 = OPUSB_CRC_LOAD(p + 48);
        p += 64; n -= 64;
        while (n >= 64) {
            a0 = OPUSB_CRC_STEP(a0, k512, OPUSB_CRC_LOAD(p)); a1 = OPUSB_CRC_STEP(a1, k512, OPUSB_CRC_LOAD(p + 16));
            a2 = OPUSB_CRC_STEP(a2, k512, OPUSB_CRC_LOAD(p + 32)); a3 = OPUSB_CRC_STEP(a3, k512, OPUSB_CRC_LOAD(p + 48));
            p += 64; n -= 64;
        }
        a1 = OPUSB_CRC_STEP(a0, k128, a1); a2 = OPUSB_CRC_STEP(a1, k128, a2); a3 = OPUSB_CRC_STEP(a2, k128, a3);
        while (n >= 16) { a3 = OPUSB_CRC_STEP(a3, k128, OPUSB_CRC_LOAD(p)); p += 16; n -= 16; }
        uint8_t last[16];
        _mm_storeu_si128((__m128i*)last, _mm_shuffle_epi8(a3, rev));
        return tables(tables(0, last, 16), p, n);
    }
#undef OPUSB_CRC_LOAD
#undef OPUSB_CRC_STEP
    bool clmul_ = false;
    uint32_t k128_[2], k512_[2];
#endif
    uint32_t t_[16][256];
};
inline const OggCrc& ogg_crc() {
    static const OggCrc c;
    return c;
}

struct OggPage {  // view into the framer's buffer, valid until the next framer call
    const uint8_t* hdr;
    const uint8_t* body;
    int hdr_len, body_len;
    int segments() const { return hdr[26]; }
    bool continued() const { return hdr[5] & 1; }
    bool bos() const { return hdr[5] & 2; }
    bool eos() const { return hdr[5] & 4; }
    int version() const { return hdr[4]; }
    int64_t granulepos() const {
        uint64_t g = 0;
        for (int i = 7; i >= 0; i--) g = g << 8 | hdr[6 + i];
        return (int64_t)g;
    }
    uint32_t serial() const { return (uint32_t)hdr[14] | (uint32_t)hdr[15] << 8 | (uint32_t)hdr[16] << 16 | (uint32_t)hdr[17] << 24; }
    uint32_t seqno() const { return (uint32_t)hdr[18] | (uint32_t)hdr[19] << 8 | (uint32_t)hdr[20] << 16 | (uint32_t)hdr[21] << 24; }
};

// Byte stream -> pages.
class OggFramer {
public:
    // A source that is one block of memory is framed in place: the framer's buffer IS the image, reserve() hands out the place
    // the next bytes already are at (the caller copies nothing when source and destination coincide) and wrote() only moves
    // the end of what has been "read" so far -- every limit on how far a search may read keeps its meaning.
    void attach(const uint8_t* image, size_t len) { ext_ = image; ext_len_ = len; rd_ = wr_ = 0; }

    // room for n more bytes; write them there, then call wrote()
    uint8_t* reserve(size_t n) {
        if (ext_) return const_cast<uint8_t*>(ext_) + (wr_ < ext_len_ ? wr_ : ext_len_);
        if (rd_ && rd_ == wr_) rd_ = wr_ = 0;
        if (buf_.size() < wr_ + n) {
            if (rd_ > (buf_.size() >> 1)) compact();
            if (buf_.size() < wr_ + n) buf_.resize((wr_ + n) * 2 + 4096);
        }
        return buf_.data() + wr_;
    }
    void wrote(size_t n) { wr_ += n; }
    size_t buffered() const { return wr_ - rd_; }
    void reset() { rd_ = wr_ = 0; }

    // > 0: a page of that many bytes was framed (og filled); 0: need more data; < 0: skipped -ret bytes
    long seek(OggPage* og) {
        const uint8_t* p = (ext_ ? ext_ : buf_.data()) + rd_;
        const size_t have = wr_ - rd_;
        if (have < 27) return 0;
        if (memcmp(p, "OggS", 4) != 0) return resync(p, have);
        const int nseg = p[26];
        const size_t hl = 27 + (size_t)nseg;
        if (have < hl) return 0;
        size_t bl = 0;
        for (int i = 0; i < nseg; i++) bl += p[27 + i];
        if (have < hl + bl) return 0;
        // checksum over the page with the CRC field zeroed
        const OggCrc& C = ogg_crc();
        static const uint8_t zero4[4] = {0, 0, 0, 0};
        uint32_t crc = C.update(0, p, 22);
        crc = C.update(crc, zero4, 4);
        crc = C.update(crc, p + 26, hl + bl - 26);
        const uint32_t want = (uint32_t)p[22] | (uint32_t)p[23] << 8 | (uint32_t)p[24] << 16 | (uint32_t)p[25] << 24;
        if (crc != want) return resync(p, have);
        if (og) {
            og->hdr = p; og->hdr_len = (int)hl; og->body = p + hl; og->body_len = (int)bl;
        }
        rd_ += hl + bl;
        return (long)(hl + bl);
    }

private:
    long resync(const uint8_t* p, size_t have) {
        const uint8_t* nx = (const uint8_t*)memchr(p + 1, 'O', have - 1);
        const size_t skip = nx ? (size_t)(nx - p) : have;
        rd_ += skip;
        return -(long)skip;
    }
    void compact() {
        memmove(buf_.data(), buf_.data() + rd_, wr_ - rd_);
        wr_ -= rd_;
        rd_ = 0;
    }
    std::vector<uint8_t> buf_;
    const uint8_t* ext_ = nullptr;
    size_t ext_len_ = 0;
    size_t rd_ = 0, wr_ = 0;
};

struct OggPacket {
    const uint8_t* data;  // valid until the next page_in()
    long bytes;
    int64_t granulepos;   // -1 unless this is the last packet ending on its page
    bool bos, eos;
};

// Pages of one logical stream -> packets.
class OggPacketizer {
public:
    void reset(uint32_t serial) {
        serial_ = serial;
        segs_.clear(); body_.clear();
        seg_rd_ = 0; body_rd_ = 0; whole_ = 0; next_seq_ = -1;
    }
    uint32_t serial() const { return serial_; }

    // Returns 0, or -1 when the page is not for this stream / has an unknown version.
    int page_in(const OggPage& og) {
        // drop what packet_out() already handed out
        if (body_rd_) { body_.erase(body_.begin(), body_.begin() + (long)body_rd_); body_rd_ = 0; }
        if (seg_rd_) { segs_.erase(segs_.begin(), segs_.begin() + (long)seg_rd_); whole_ -= seg_rd_; seg_rd_ = 0; }
        if (og.serial() != serial_ || og.version() > 0) return -1;
        const int nseg = og.segments();
        const uint8_t* lace = og.hdr + 27;
        const uint8_t* body = og.body;
        long body_len = og.body_len;
        int s = 0;
        bool bos = og.bos();
        if ((int64_t)og.seqno() != next_seq_) {
            // out of sequence: forget the unfinished packet, and leave a hole marker unless this is the first page
            for (size_t i = whole_; i < segs_.size(); i++) body_.resize(body_.size() - (segs_[i].val & 0xff));
            segs_.resize(whole_);
            if (next_seq_ != -1) {
                segs_.push_back(Seg{kHole, -1});
                whole_++;
            }
        }
        if (og.continued()) {
            // nothing to continue (no open packet): skip the tail of the packet we never saw the start of
            if (segs_.empty() || (segs_.back().val & 0xff) < 255 || segs_.back().val == kHole) {
                bos = false;
                for (; s < nseg; s++) {
                    const int v = lace[s];
                    body += v;
                    body_len -= v;
                    if (v < 255) { s++; break; }
                }
            }
        }
        if (body_len > 0) body_.insert(body_.end(), body, body + body_len);
        long last_end = -1;
        for (; s < nseg; s++) {
            int v = lace[s];
            const bool ends = v < 255;
            if (bos) { v |= kBos; bos = false; }
            segs_.push_back(Seg{v, -1});
            if (ends) { last_end = (long)segs_.size() - 1; whole_ = segs_.size(); }
        }
        if (last_end >= 0) segs_[(size_t)last_end].gp = og.granulepos();
        if (og.eos() && !segs_.empty()) segs_.back().val |= kEos;
        next_seq_ = (int64_t)og.seqno() + 1;
        return 0;
    }

    // 1: packet, 0: no complete packet buffered, -1: hole (lost pages) before the next packet
    int packet_out(OggPacket* op) {
        size_t i = seg_rd_;
        if (whole_ <= i) return 0;
        if (segs_[i].val & kHole) {
            seg_rd_++;
            return -1;
        }
        int v = segs_[i].val;
        long bytes = v & 0xff;
        bool eos = v & kEos;
        const bool bos = v & kBos;
        while ((v & 0xff) == 255) {
            v = segs_[++i].val;
            if (v & kEos) eos = true;
            bytes += v & 0xff;
        }
        op->data = body_.data() + body_rd_;
        op->bytes = bytes;
        op->granulepos = segs_[i].gp;
        op->bos = bos;
        op->eos = eos;
        body_rd_ += (size_t)bytes;
        seg_rd_ = i + 1;
        return 1;
    }

private:
    enum { kBos = 0x100, kEos = 0x200, kHole = 0x400 };
    struct Seg {
        int val;
        int64_t gp;
    };
    std::vector<Seg> segs_;
    std::vector<uint8_t> body_;
    size_t seg_rd_ = 0, body_rd_ = 0, whole_ = 0;
    int64_t next_seq_ = -1;
    uint32_t serial_ = 0xFFFFFFFFu;
};

}  // namespace opusb200
