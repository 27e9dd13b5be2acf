// celt_script.cuh -- the per-frame "script" handed from the entropy kernel to the band / energy /
// synthesis kernels (parse stage -> reconstruct stage).
//
// The entropy kernel (one thread per (stream, frame)) does everything that is serial in the bitstream:
// range decoding, allocation, the quant_all_bands control flow, PVQ index -> pulses -> normalised
// rotated vector.  Everything that is a vector operation on more than one leaf (folding, Haar,
// Hadamard reordering, stereo merge, band energy, IMDCT ...) is recorded as a BandOp and executed
// later by a whole warp.  None of the recorded decisions depends on decoded signal values or on
// earlier frames, which is what makes the (stream, frame) grid legal.
#pragma once
#include <stdint.h>

namespace opusb200 {

enum {
    kNbEBands = 21,
    kOverlap = 120,
    kShortMdct = 120,
    kMaxFrame = 960,
    kMaxPacket = 1275,
    kNormMax = 624,  // 8 * eBands[20]
    kMaxLeaf = 176,  // largest band at LM=3
};

// int16 workspace of the band kernel (element offsets). X|Y live at [0, 2*960); the folding
// reference ("norm") and its second-channel copy follow; then the lowband scratch.
enum {
    W_XY = 0,
    W_NORM = 2 * kMaxFrame,
    W_NORM2 = W_NORM + kNormMax,
    W_SCRATCH = W_NORM2 + kNormMax,
    W_SIZE = W_SCRATCH + kMaxLeaf,
};

enum BandOpCode : uint8_t {
    OP_END = 0,
    OP_NOISE,      // W[off..+n) = LCG noise, renormalise to g            (ref src/celt.cpp:1496-1503,1516)
    OP_FOLD,       // W[off..+n) = W[off2..+n) +- 1/256, renormalise to g  (ref :1504-1516)
    OP_COPY,       // W[off..+n) = W[off2..+n)                             (ref :1555, 1749)
    OP_HAAR,       // haar1(W+off, N0=n, stride=a)                         (ref :1202)
    OP_DEINT,      // deinterleave_hadamard(W+off, N0=n, stride=a, hadamard=g)   (ref :1162)
    OP_INT,        // interleave_hadamard(W+off, N0=n, stride=a, hadamard=g)     (ref :1183)
    OP_SCALE_OUT,  // W[off2+j] = (g * W[off+j]) >> 15, j<n  (lowband_out) (ref :1615-1621)
    OP_SETW,       // W[off] = g                                           (ref :1376)
    OP_MERGE,      // stereo_merge(W+off, W+off2, mid=g, n)                (ref :1113)
    OP_NEG,        // W[off..+n) = -W[..]                                  (ref :1733-1738)
    OP_N2,         // N=2 stereo special case: x2=off, y2=off2, a=sign bit(0/1), g=mid, g2=side (ref :1682-1696)
    OP_AVG,        // W[off+j] = (W[off+j] + W[off2+j]) >> 1, j<n  (dual stereo -> intensity) (ref :1888-1889)
    OP_ZERO,       // W[off..+n) = 0                                       (ref :1493)
    OP_LEAF,       // coded PVQ leaf: W[off..+n) = alg_unquant(index=aux, N=n, K=off2, B=a, gain=g) (ref :782)
};

struct alignas(16) BandOp {  // 16 bytes, one 128-bit access
    uint8_t code;
    uint8_t a;
    uint16_t off;
    uint16_t off2;
    uint16_t n;
    int16_t g;
    int16_t g2;
    uint32_t aux;
};

// One op buffer per frame: transform ops grow from the front in bitstream order, coded-leaf records
// (OP_LEAF) grow from the back -- they are independent of everything else and are expanded first.
enum {
    kMaxOps = 1360,  // worst case: 21 bands x 2 channels x (16 leaves + ~16 transforms)
};

enum FrameFlags : uint8_t {
    FF_SILENCE = 1,
    FF_TRANSIENT = 2,
    FF_INTRA = 4,
    FF_ANTICOLLAPSE = 8,
    FF_EC_ERROR = 16,
    FF_POSTFILTER = 32,
};

// Error codes follow the reference numbering (src/opus_decoder.h:40-77).
enum {
    OPUSB_OK = 0,
    OPUSB_BAD_ARG = -1,
    OPUSB_BUFFER_TOO_SMALL = -2,
    OPUSB_INTERNAL_ERROR = -3,
    OPUSB_INVALID_PACKET = -4,
    OPUSB_UNIMPLEMENTED = -5,
    OPUSB_CELT_BAD_ARG = -18,       // ERR_OPUS_CELT_BAD_ARG
    OPUSB_FRAME_SKIPPED = -2004,       // ours: filler frame, or a frame behind a failed frame of the same packet
    OPUSB_SCRIPT_OVERFLOW = -1000,  // ours: script capacity exceeded (cannot happen for valid sizes)
};

struct CeltFrameHdr {
    uint32_t final_rng;  // range decoder rng at the end of the frame -> next frame's noise seed (ref :2437)
    int32_t status;      // >=0: samples per channel; <0 error
    uint8_t flags;
    uint8_t LM;
    uint8_t C;           // channels coded in this frame
    uint8_t start;
    uint8_t end;
    uint8_t pf_tapset;
    uint8_t spread;
    uint8_t oflags;      // OpusFrameFlags (opus_frame.cuh)
    uint16_t n_ops;
    uint16_t n_leaves;
    int16_t pf_pitch;
    int16_t pf_gain;
    uint32_t lcg_steps;  // LCG draws consumed by noise/fold fills (anti-collapse seed = seed advanced by this)
    int32_t pq[2][kNbEBands];     // coarse energy: prev[c] + (q << 7) per band (ref :3656)
    int16_t fine[2][kNbEBands];   // fine + final energy offsets, summed mod 2^16 (ref :3676, 3695)
    int16_t pulses[kNbEBands];    // PVQ bit allocation per band (anti_collapse depth, ref :1025)
    uint8_t cmask[2 * kNbEBands]; // collapse masks [band * C + c] (ref :1913-1914)
};

}  // namespace opusb200
