// tests/hostsim/hostsim.cpp -- TEST INFRASTRUCTURE: compiles the kernels' __host__ __device__ stage
// bodies with g++ so their logic can be checked against the reference oracle on a CPU-only box.
// Never linked into the product library; the product path only launches the CUDA kernels.
#include <string.h>
#include <stdlib.h>
#include "../../esp32-opus-player_b200/csrc/celt_parse.cuh"

using namespace opusb200;

extern "C" {

static int celt_parse_rd(const RangeDec& rd, int C, int LM, int start, int disable_inv, CeltFrameHdr* hdr, BandOp* ops) {
    ParseCtx p;
    memset(&p, 0, sizeof(p));
    p.rd = rd;
    p.C = C; p.LM = LM; p.start = start; p.end = kNbEBands; p.disable_inv = disable_inv;
    p.ops = ops; p.ops_cap = kMaxOps;
    return celt_parse_frame(p, *hdr);
}
int hs_celt_parse(const uint8_t* pkt, int len, int C, int LM, int start, int disable_inv, CeltFrameHdr* hdr, BandOp* ops,
                  int16_t* X) {
    (void)X;
    RangeDec rd;
    rd_init(rd, pkt, (u32)len);
    return celt_parse_rd(rd, C, LM, start, disable_inv, hdr, ops);
}
int hs_sizeof_hdr() { return (int)sizeof(CeltFrameHdr); }
int hs_sizeof_op() { return (int)sizeof(BandOp); }

}

// ---- whole CELT frame through the stage bodies, serially (lanes = 1) ----
#include "../../esp32-opus-player_b200/csrc/celt_synth.cuh"

struct HsCelt {
    CeltStreamState st;
    int CC;
};

extern "C" {

void* hs_celt_new(int channels) {
    HsCelt* d = (HsCelt*)calloc(1, sizeof(HsCelt));
    d->CC = channels;
    for (int i = 0; i < kNbEBands; i++)
        for (int c = 0; c < 2; c++) d->st.band[i].logE[c] = d->st.band[i].logE2[c] = -(28 << 10);
    return d;
}
void hs_celt_free(void* d) { free(d); }

static int celt_decode_rd(HsCelt* d, const RangeDec& rd, int C, int frame_size, int start, int16_t* pcm, int32_t* dbg_mem);

// decode one CELT frame; C = coded channels, frame_size in {120,240,480,960}; pcm gets frame_size*CC
int hs_celt_decode(void* dv, const uint8_t* pkt, int len, int C, int frame_size, int start, int16_t* pcm, int32_t* dbg_mem) {
    RangeDec rd;
    rd_init(rd, pkt, (u32)len);
    return celt_decode_rd((HsCelt*)dv, rd, C, frame_size, start, pcm, dbg_mem);
}

}  // extern "C"

static int celt_decode_rd(HsCelt* d, const RangeDec& rd, int C, int frame_size, int start, int16_t* pcm, int32_t* dbg_mem) {
    CeltStreamState& st = d->st;
    const int CC = d->CC;
    int LM = frame_size == 120 ? 0 : frame_size == 240 ? 1 : frame_size == 480 ? 2 : frame_size == 960 ? 3 : -1;
    if (LM < 0) return OPUSB_CELT_BAD_ARG;
    static CeltFrameHdr hdr;
    static BandOp ops[kMaxOps];
    static int16_t Xg[2 * kMaxFrame];
    static int16_t W[W_SIZE];
    (void)Xg;
    int status = celt_parse_rd(rd, C, LM, start, CC == 1, &hdr, ops);
    if (status == OPUSB_CELT_BAD_ARG) return status;
    const int N = frame_size, M = 1 << LM;
    Lanes L;
    FrameEnergy fe;
    for (int i = 0; i < kNbEBands; i++) energy_step(st.band[i], i, hdr, fe);
    leaf_stage(L, W, ops + kMaxOps - 1, hdr.n_leaves, hdr.spread);
    u32 seed = band_stage_ops(L, W, ops, hdr.n_ops, st.rng);
    if (hdr.flags & FF_ANTICOLLAPSE) band_stage_anti_collapse(L, W, hdr, fe, seed);
    const int transient = (hdr.flags & FF_TRANSIENT) != 0, silence = (hdr.flags & FF_SILENCE) != 0;
    const int nblk = transient ? M : 1, shift = transient ? 3 : 3 - LM, N2 = N / nblk;
    alignas(16) static int32_t freq[kMaxFrame], freq2[kMaxFrame], work[kMaxFrame], buf[kCombHist + kMaxFrame], gs[kNbEBands + 1];
    for (int c = 0; c < CC; c++) {
        if (CC == 2 && C == 1) {
            denormalise(L, W + W_XY, freq, fe.E[0], hdr.start, hdr.end, M, silence, gs);
        } else if (CC == 1 && C == 2) {
            denormalise(L, W + W_XY, freq, fe.E[0], hdr.start, hdr.end, M, silence, gs);
            denormalise(L, W + W_XY + N, freq2, fe.E[1], hdr.start, hdr.end, M, silence, gs);
            for (int i = 0; i < N; i++) freq[i] = (freq[i] >> 1) + (freq2[i] >> 1);
        } else {
            denormalise(L, W + W_XY + c * N, freq, fe.E[c], hdr.start, hdr.end, M, silence, gs);
        }
        imdct_raw(L, freq, work, shift, nblk);
        // the signal history lives in a ring like in the post-filter kernel; the ring origin moves from call to call so
        // that frames wrap around its end
        static int32_t ring[kRingSize];
        static int origin = 0;
        origin = (origin + 777) & (kRingSize - 1);
        RingView x{ring, origin + kCombHist};
        for (int i = 0; i < kCombHist; i++) x[i - kCombHist] = st.hist[c][i];
        tdac_frame(L, work, st.tail[c], x, N2, nblk);
        comb_filter_inplace(L, x, st.pf_period_old, st.pf_period, kShortMdct, (s16)st.pf_gain_old, (s16)st.pf_gain,
                            st.pf_tapset_old, st.pf_tapset);
        if (LM != 0)
            comb_filter_inplace(L, x + kShortMdct, st.pf_period, hdr.pf_pitch, N - kShortMdct, (s16)st.pf_gain,
                                hdr.pf_gain, st.pf_tapset, hdr.pf_tapset);
        for (int i = 0; i < kCombHist; i++) st.hist[c][i] = x[N - kCombHist + i];
        for (int i = 0; i < N; i++) buf[kCombHist + i] = x[i];
        if (dbg_mem) memcpy(dbg_mem + c * kMaxFrame, buf + kCombHist, sizeof(int32_t) * N);
        s32 mem = st.preemph_mem[c];
        for (int j = 0; j < N; j++) pcm[j * CC + c] = (int16_t)deemph_sample(buf[kCombHist + j], mem);
        st.preemph_mem[c] = mem;
    }
    st.pf_period_old = st.pf_period; st.pf_gain_old = st.pf_gain; st.pf_tapset_old = st.pf_tapset;
    st.pf_period = hdr.pf_pitch; st.pf_gain = hdr.pf_gain; st.pf_tapset = hdr.pf_tapset;
    if (LM != 0) { st.pf_period_old = st.pf_period; st.pf_gain_old = st.pf_gain; st.pf_tapset_old = st.pf_tapset; }
    st.rng = hdr.final_rng;
    return status;
}

// ---- Opus frame layer on top (ref opus_decode_frame, src/opus_decoder.cpp:154): SILK, hybrid, CELT ----
#include "../../esp32-opus-player_b200/csrc/opus_frame.cuh"
#include "../../esp32-opus-player_b200/csrc/silk_stages.cuh"

// The SILK stages of silk_stages.cuh run back to back for one stream (rings with stride 1).
struct HsSilk {
    SilkChainRegs R[2];
    SilkChainCtl K[2];
    int16_t prev_row[2][kSilkMaxFrame];  // the previous decoded frame of the chain (its output history)
    int prev_len[2];                      // 0: history is all zero
    int32_t ltp_ring[2][kSilkRing];
    int16_t prevNLSF[2][kSilkMaxOrder];
    int prev_mid_only;
    s32 m2, m1, s2, s1, p0, p1;  // stereo history (ref stereo_dec_state)
    SilkUp2Regs U[2];
    int16_t fh[2][8];
};
static void hs_silk_init(HsSilk& h) {
    memset(&h, 0, sizeof(h));
    for (int n = 0; n < 2; n++) { h.R[n].prev_gain_Q16 = 65536; h.K[n].first = 1; }
}
// one frame -> 960 * nch samples at 48 kHz, interleaved
static void hs_silk_frame(HsSilk& h, SilkFrame& F, bool reinit, int16_t* pcm48) {
    const int nch = F.nch, fs = F.fs_kHz, L = 20 * fs;
    static int16_t xq[2][kSilkMaxFrame], lr[2][kSilkMaxFrame], up[2][2 * kSilkMaxFrame + 8];
    if (reinit) { h.prev_mid_only = 0; h.m2 = h.m1 = h.s2 = h.s1 = h.p0 = h.p1 = 0; }
    bool rs_reset[2] = {false, false}, dec[2] = {false, false};
    for (int n = 0; n < nch; n++) {
        ChainRing<int32_t, 1> ltp{h.ltp_ring[n]};
        bool cleared;
        dec[n] = silk_chain_begin_frame(h.K[n], h.R[n], n, reinit, fs, F.mid_only, h.prev_mid_only, &rs_reset[n], &cleared);
        if (cleared) h.prev_len[n] = 0;
        if (dec[n]) {
            SilkCtl ctl;
            silk_params_item(F.idx[n], fs, h.K[n].first ? nullptr : h.prevNLSF[n], ctl, SerialFix{});
            s32 gains[4];
            int8_t lgi = (int8_t)h.K[n].LastGainIndex;
            silk_gains_dequant(gains, F.idx[n].gains, &lgi);
            h.K[n].LastGainIndex = lgi;
            const ChainHistory hist{h.prev_len[n] ? h.prev_row[n] + h.prev_len[n] : nullptr, xq[n]};
            silk_core_frame(h.R[n], hist, ltp, F.pulses[n], &ctl, gains, F.idx[n].signalType, F.idx[n].quantOffsetType, F.idx[n].seed, fs,
                            xq[n], true, SerialCoop{});
            memcpy(h.prev_row[n], xq[n], sizeof(int16_t) * (size_t)L);
            h.prev_len[n] = L;
            h.K[n].first = 0;
            memcpy(h.prevNLSF[n], ctl.nlsf_Q15, sizeof(ctl.nlsf_Q15));
        } else {
            for (int i = 0; i < L; i++) xq[n][i] = 0;
        }
    }
    if (nch == 1) {
        for (int i = 0; i < L; i++) lr[0][i] = i == 0 ? (int16_t)h.m1 : xq[0][i - 1];
    } else {
        SilkUnmix U;
        U.setup(F.ms_pred_Q13, h.p0, h.p1, fs);
        for (int i = 0; i < L; i++) {
            const s32 x1a = i >= 2 ? xq[0][i - 2] : (i == 1 ? h.m1 : h.m2);
            const s32 x1b = i >= 1 ? xq[0][i - 1] : h.m1;
            const s32 x2b = i >= 1 ? xq[1][i - 1] : h.s1;
            U.sample(i, x1a, x1b, xq[0][i], x2b, &lr[0][i], &lr[1][i]);
        }
        h.p0 = (int16_t)F.ms_pred_Q13[0]; h.p1 = (int16_t)F.ms_pred_Q13[1];
        h.s2 = xq[1][L - 2]; h.s1 = xq[1][L - 1];
    }
    h.m2 = xq[0][L - 2]; h.m1 = xq[0][L - 1];
    const s32 inv = silk_inv_ratio_Q16(fs);
    for (int n = 0; n < nch; n++) {
        if (rs_reset[n]) { memset(&h.U[n], 0, sizeof(h.U[n])); memset(h.fh[n], 0, sizeof(h.fh[n])); }
        const int D = silk_input_delay(fs);
        memcpy(up[n], h.fh[n], sizeof(h.fh[n]));
        for (int i = 0; i < L; i++) silk_up2_sample(h.U[n], lr[n][i], D, &up[n][8 + 2 * i], &up[n][8 + 2 * i + 1]);
        memcpy(h.fh[n], &up[n][2 * L], sizeof(h.fh[n]));
        for (int o = 0; o < 960; o++) pcm48[o * nch + n] = silk_fir_sample(up[n], o, fs, inv);
    }
    h.prev_mid_only = F.mid_only;
}

struct HsOpus {
    HsCelt celt;
    HsSilk silk;
    int channels, prev_mode;
};

extern "C" {

void* hs_opus_new(int channels) {
    HsOpus* d = (HsOpus*)calloc(1, sizeof(HsOpus));
    d->channels = channels;
    d->celt.CC = channels;
    for (int i = 0; i < kNbEBands; i++)
        for (int c = 0; c < 2; c++) d->celt.st.band[i].logE[c] = d->celt.st.band[i].logE2[c] = -(28 << 10);
    hs_silk_init(d->silk);
    return d;
}
void hs_opus_free(void* d) { free(d); }

// one code-0 packet: toc + payload; returns samples per channel (960) or a negative error
int hs_opus_decode(void* dv, int toc, const uint8_t* payload, int len, int16_t* pcm) {
    HsOpus* d = (HsOpus*)dv;
    const int CC = d->channels;
    static int16_t pcm_silk[960 * 2];
    static SilkFrame sf;
    static CeltFrameHdr hdr;
    ParseCtx p;
    memset(&p, 0, sizeof(p));
    bool celt, fade;
    const int oflags = opus_parse_prologue(p, payload, len, toc, d->prev_mode, CC, hdr, sf, &celt, &fade, 0u);
    if (!celt && !(oflags & OF_HAS_SILK) && opus_toc_mode(toc) != MODE_CELT_ONLY && ((toc & 4) ? 2 : 1) != CC) return OPUSB_UNIMPLEMENTED;
    if (oflags & OF_HAS_SILK) hs_silk_frame(d->silk, sf, (oflags & OF_SILK_REINIT) != 0, pcm_silk);
    if (oflags & OF_RESET_CELT) {  // OPUS_RESET_STATE (ref src/celt.cpp:2479)
        CeltStreamState& st = d->celt.st;
        st.rng = 0;
        st.pf_period = st.pf_period_old = st.pf_gain = st.pf_gain_old = st.pf_tapset = st.pf_tapset_old = 0;
        for (int i = 0; i < kNbEBands; i++)
            for (int c = 0; c < 2; c++) st.band[i].logE[c] = st.band[i].logE2[c] = -(28 << 10);
    }
    int ret = 960;
    for (int i = 0; i < 960 * CC; i++) pcm[i] = 0;
    if (celt) {
        const int r = celt_decode_rd(&d->celt, p.rd, p.C, kShortMdct << p.LM, p.start, pcm, nullptr);
        if (r < 0 && !fade) ret = r;
    }
    if (oflags & OF_HAS_SILK)
        for (int i = 0; i < 960 * CC; i++) pcm[i] = sat16((s32)pcm[i] + (s32)pcm_silk[i]);
    d->prev_mode = opus_toc_mode(toc);
    return ret;
}

}
