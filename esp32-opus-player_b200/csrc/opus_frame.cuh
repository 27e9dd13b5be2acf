// opus_frame.cuh -- Opus frame layer of the entropy stage.
//
// One thread reads one coded frame exactly in the order opus_decode_frame does (ref src/opus_decoder.cpp:154-278):
// range decoder init, SILK layer (SILK-only / hybrid), the hybrid redundancy flag, then the CELT layer with
// the band range and frame size the reference uses, including its quirks: every frame is
// decoded as 20 ms (F1), CELT `end` is always 21 (F3), the redundancy bit is read and dropped (F4), and a
// SILK-only frame that follows a hybrid frame still decodes a 120-sample CELT frame from the live range
// decoder (D4).  Everything here depends only on this packet and on the previous packet's mode.
#pragma once
#include "celt_parse.cuh"
#include "silk.cuh"

namespace opusb200 {

enum { MODE_NONE = 0, MODE_SILK_ONLY = 1000, MODE_HYBRID = 1001, MODE_CELT_ONLY = 1002 };

enum OpusFrameFlags : uint8_t {
    OF_NO_CELT = 1,      // no CELT layer in this frame (SILK-only)
    OF_HAS_SILK = 2,     // SILK layer present: its 48 kHz output is added to the CELT output (ref :271-273)
    OF_RESET_CELT = 4,   // OPUS_RESET_STATE before the CELT layer (mode change, ref :255-256)
    OF_SILK_REINIT = 8,  // silk_InitDecoder before the SILK layer (previous frame was CELT-only, ref :183)
};

OPUS_HD bool opus_has_silk_layer(int toc, int CC) { return !(toc & 0x80) && ((toc & 0x4) ? 2 : 1) == CC; }
OPUS_HD int opus_toc_mode(int toc) { return (toc & 0x80) ? MODE_CELT_ONLY : ((toc & 0x60) == 0x60 ? MODE_HYBRID : MODE_SILK_ONLY); }

// Everything opus_decode_frame reads before the CELT layer.  Sets up `p` (range decoder, band range, frame size)
// for the single celt_parse_frame call site of the entropy kernel.  `silk_lanes`: the lanes of the warp that have a SILK
// layer to parse (opus_has_silk_layer), which step through it together.  Returns the Opus-layer flags; *celt says
// whether a CELT frame follows, *fade that it is the 120-sample hybrid -> SILK fade-out frame whose status
// the reference ignores (ref :263-268).
OPUS_HD int opus_parse_prologue(ParseCtx& p, const uint8_t* payload, int len, int toc, int prev_mode, int CC, CeltFrameHdr& hdr,
                                SilkFrame& sf, bool* celt, bool* fade, unsigned silk_lanes, u32 sbase = 0, u32 staged = 0) {
    const int mode = opus_toc_mode(toc);
    const int C = (toc & 0x4) ? 2 : 1;
    int oflags = 0;
    *celt = true; *fade = false;
    hdr.flags = 0; hdr.oflags = 0; hdr.n_ops = 0; hdr.n_leaves = 0; hdr.LM = 3; hdr.C = (uint8_t)C; hdr.start = 0; hdr.end = kNbEBands;
    hdr.final_rng = 0; hdr.spread = 2; hdr.pf_pitch = 0; hdr.pf_gain = 0; hdr.pf_tapset = 0; hdr.lcg_steps = 0;
    if (mode != MODE_CELT_ONLY && C != CC) {
        // the reference mixes a mono SILK layer into half of a stereo buffer here; not reproduced
        hdr.status = OPUSB_UNIMPLEMENTED;
        *celt = false;
        return OF_NO_CELT;
    }
    rd_init(p.rd, payload, (u32)len, sbase, staged);
    p.C = C; p.LM = 3; p.start = 0;
    if (mode != MODE_CELT_ONLY) {
        oflags |= OF_HAS_SILK;
        if (prev_mode == MODE_CELT_ONLY) oflags |= OF_SILK_REINIT;
        int fs_kHz = 16;
        if (mode == MODE_SILK_ONLY) {
            const int bw = (toc >> 5) & 3;  // NB, MB, WB (ref :188-197)
            fs_kHz = bw == 0 ? 8 : bw == 1 ? 12 : 16;
        }
        silk_parse_frame(p.rd, sf, fs_kHz, C, silk_lanes);
        if (mode == MODE_HYBRID && rd_tell(p.rd) + 17 + 20 <= 8 * len) rd_bit_logp(p.rd, 12);  // redundancy flag, dropped (ref :218-221)
        p.start = 17;
    }
    if (mode != MODE_SILK_ONLY) {
        if (mode != prev_mode && prev_mode > 0) oflags |= OF_RESET_CELT;
    } else if (prev_mode == MODE_HYBRID) {
        p.start = 0;
        p.LM = 0;
        *fade = true;
    } else {
        hdr.status = 960;
        *celt = false;
        oflags |= OF_NO_CELT;
    }
    return oflags;
}

// The same for a frame with a CELT-only TOC (ref src/opus_decoder.cpp:154-278 with mode == MODE_CELT_ONLY): no SILK
// layer, bands from 0, 20 ms; OPUS_RESET_STATE when the previous frame had another mode.
OPUS_HD int opus_parse_prologue_celt(ParseCtx& p, const uint8_t* payload, int len, int toc, int prev_mode, CeltFrameHdr& hdr, u32 sbase = 0,
                                     u32 staged = 0) {
    const int C = (toc & 0x4) ? 2 : 1;
    hdr.flags = 0; hdr.oflags = 0; hdr.n_ops = 0; hdr.n_leaves = 0; hdr.LM = 3; hdr.C = (uint8_t)C; hdr.start = 0; hdr.end = kNbEBands;
    hdr.final_rng = 0; hdr.spread = 2; hdr.pf_pitch = 0; hdr.pf_gain = 0; hdr.pf_tapset = 0; hdr.lcg_steps = 0;
    rd_init(p.rd, payload, (u32)len, sbase, staged);
    p.C = C; p.LM = 3; p.start = 0;
    return (prev_mode != MODE_CELT_ONLY && prev_mode > 0) ? OF_RESET_CELT : 0;
}

}  // namespace opusb200
