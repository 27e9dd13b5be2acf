// The page checksum of csrc/ogg_demux.hpp (tables sliced by 16, and the carry-less-multiplication folding where the CPU has
// it) against the bitwise definition of the Ogg CRC (RFC 3533: polynomial 0x04C11DB7, MSB first, no reflection, no final
// xor), over random offsets, lengths 0 .. 70 000 and start values.  Built and run by tests/test_ogg_crc.py.  Prints
// "folding <0|1> cases <n> mismatches <m>".
#include <stdio.h>

#include <vector>

#include "../../esp32-opus-player_b200/csrc/ogg_demux.hpp"

static uint32_t bitwise(uint32_t c, const uint8_t* p, size_t n) {
    for (size_t i = 0; i < n; i++) {
        c ^= (uint32_t)p[i] << 24;
        for (int k = 0; k < 8; k++) c = (c & 0x80000000u) ? (c << 1) ^ 0x04C11DB7u : (c << 1);
    }
    return c;
}

int main() {
    std::vector<uint8_t> b(1 << 17);
    uint64_t s = 0x9E3779B97F4A7C15ull;
    for (size_t i = 0; i < b.size(); i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; b[i] = (uint8_t)(s >> 11); }
    const opusb200::OggCrc& C = opusb200::ogg_crc();
    int bad = 0, cases = 0;
    for (int t = 0; t < 4000; t++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        const size_t o = (size_t)(s >> 20) % 4096;
        const size_t n = (t % 4 == 0) ? (size_t)(s >> 33) % 70000 : (size_t)(s >> 33) % 400;
        const uint32_t start = (t & 1) ? (uint32_t)(s >> 7) : 0u;
        const uint32_t want = bitwise(start, b.data() + o, n);
        if (C.update(start, b.data() + o, n) != want) bad++;
        if (C.update_tables(start, b.data() + o, n) != want) bad++;
        // a checksum continued over two calls (the framer: header, four zero bytes, rest of the page)
        const size_t cut = n ? (size_t)(s >> 45) % n : 0;
        if (C.update(C.update(start, b.data() + o, cut), b.data() + o + cut, n - cut) != want) bad++;
        cases += 3;
    }
    printf("folding %d cases %d mismatches %d\n", (int)C.folding(), cases, bad);
    return bad ? 1 : 0;
}
