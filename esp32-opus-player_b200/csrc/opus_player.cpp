// opus_player.cpp -- host side of the player layer of libopusb200.so (include/opusfile_b200.h).
//
// What the reference does one file at a time in src/opusfile.cpp (header fetch :157-296, first-page time stamp
// validation :486-633, page processing with pre-skip / end-trim :835-1133, the read loop :1171-1331) is split here
// into three stages so that thousands of files keep one GPU busy:
//
//   OpusFileReader   per file, host thread: Ogg pages -> packets -> "player events" in the exact order the
//                    reference's read loop would meet them: a packet with its pre-skip / end-trim window, a return
//                    code (OP_HOLE, OP_EREAD, 0 at the end), or a decoder reset at a chained-link boundary
//   RoundScheduler   packs the next frames of every file into pinned staging rows and runs decode_many()
//   Track::read      serves op_read_stereo() calls from the decoded rows: same samples, same return values
//
// No audio is decoded on the CPU here.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <deque>
#include <thread>
#include <vector>

#include "../../include/opus_b200.h"
#include "../../include/opusfile_b200.h"
#include "ogg_demux.hpp"

namespace {
using namespace opusb200;

enum { kReadChunk = 2048, kHeaderWindow = 8192 };  // ref OP_READ_SIZE, OP_CHUNK_SIZE (src/opusfile.cpp:24-25)
enum ReadyState { kNotOpen = 0, kPartOpen = 1, kOpened = 2, kStreamSet = 3, kInitSet = 4 };
const int kPending = OPFB_PENDING;  // internal: the event exists but its round has not been decoded yet

// ---- granule positions: unsigned 64-bit with -1 invalid (ref src/opusfile.cpp:299-401) ----
bool gp_add(int64_t* dst, int64_t src, int32_t delta) {
    const uint64_t s = (uint64_t)src;
    if (delta > 0) {
        if (s > 0xFFFFFFFFFFFFFFFEull - (uint64_t)delta) return false;
    } else if (delta < 0) {
        if (s < (uint64_t)(-(int64_t)delta)) return false;
    }
    *dst = (int64_t)(s + (uint64_t)(int64_t)delta);
    return true;
}
bool gp_diff(int64_t* delta, int64_t a, int64_t b) {
    const uint64_t ua = (uint64_t)a, ub = (uint64_t)b;
    if (ua >= ub) {
        if (ua - ub > (uint64_t)INT64_MAX) return false;
        *delta = (int64_t)(ua - ub);
    } else {
        if (ub - ua > (uint64_t)1 << 63) return false;
        *delta = (int64_t)(0 - (ub - ua));
    }
    return true;
}
int gp_cmp(int64_t a, int64_t b) { return ((uint64_t)a > (uint64_t)b) - ((uint64_t)b > (uint64_t)a); }

// ref op_get_packet_duration (src/opusfile.cpp:406)
int packet_duration(const uint8_t* data, long len) {
    const int nframes = opusb_packet_get_nb_frames(data, (int32_t)len);
    if (nframes < 0) return OPB_EBADPACKET;
    const int n = nframes * opusb_packet_get_samples_per_frame(data, 48000);
    return n > 120 * 48 ? OPB_EBADPACKET : n;
}

struct Event {
    enum Kind { kPacket, kCode, kResetSoft, kResetHard } kind;
    int code;            // kCode: what op_read_stereo returns
    size_t off;          // kPacket: packet bytes in the track's arena
    int bytes, duration; // kPacket
    int od_pos, od_size; // kPacket: samples [od_pos, od_size) of the decoded packet are played (pre-skip, end trim)
    // filled in by the scheduler
    int status;          // kPacket: kPending, samples decoded (> 0) or OP_EBADPACKET
    const int16_t* pcm;  // kPacket: decoded sample od_pos (interleaved, decoder channels) -- in the scheduler's pinned round
                         // buffer, or in `own` once that buffer was needed for a later round
    int16_t* own;        // kPacket: private copy (malloc), or nullptr
    int round;           // kPacket: the round that decoded it
};

// One Ogg Opus file: the reference's open + page processing, producing events instead of decoding.
class OpusFileReader {
public:
    OpusHeadB200 head;
    std::vector<uint8_t> arena;  // packet bytes of queued events

    void set_callback(opusb_read_func fn, void* user) { read_ = fn; user_ = user; mem_ = nullptr; fr_.attach(nullptr, 0); }
    void set_memory(const uint8_t* p, size_t n, bool clean_eof) {
        mem_ = p; mem_len_ = n; mem_pos_ = 0; clean_eof_ = clean_eof; read_ = nullptr;
        fr_.attach(p, n);   // framed in place (ogg_demux.hpp): source_read() below then has nothing to copy
        arena.reserve(n);   // the packets of the whole file fit: no regrowth while events are queued
    }

    // ref opus_init_decoder -> op_open1 / op_open2 (src/opusfile.cpp:730-799)
    int open(std::deque<Event>* q) {
        q_ = q;
        ready_ = kNotOpen;
        offset_ = 0;
        memset(&head, 0, sizeof(head));
        OggPage og;
        bool have_page = false;
        int ret;
        for (;;) {
            ret = fetch_headers(have_page ? &og : nullptr, true);
            if (ret < 0) break;
            link_serial_ = pk_.serial();
            pcm_end_ = -1;
            ret = find_initial_pcm_offset(&og);
            if (ret <= 0) break;
            have_page = true;  // empty link: og holds the first page of the next one
        }
        if (ret < 0) return ret;
        ready_ = kStreamSet;
        return make_decode_ready();
    }

    // Advance until one more event is queued (ref op_read_native's loop, src/opusfile.cpp:1171-1290).
    void next() {
        for (;;) {
            if (ready_ >= kInitSet && op_pos_ < pend_.size()) {
                emit_packet(pend_[op_pos_++]);
                return;
            }
            const int ret = fetch_and_process_page();
            if (ret == OPB_EOF) { emit_code(0); return; }
            if (ret < 0) { emit_code(ret); return; }
        }
    }

private:
    struct Pending {
        const uint8_t* data;
        int bytes, duration;
        int64_t gp;
        bool eos;
    };

    int source_read(uint8_t* dst, int n) {
        if (read_) return read_(user_, dst, n);
        if (n == 0) return 0;
        if (mem_pos_ >= mem_len_) return clean_eof_ ? 0 : -1;
        const size_t k = std::min((size_t)n, mem_len_ - mem_pos_);
        if (dst != mem_ + mem_pos_) memcpy(dst, mem_ + mem_pos_, k);
        mem_pos_ += k;
        return (int)k;
    }

    // ref op_get_next_page (src/opusfile.cpp:63): boundary < 0 unbounded, else absolute offset limit for the page start
    int64_t get_next_page(OggPage* og, int64_t boundary) {
        while (boundary <= 0 || offset_ < boundary) {
            const long more = fr_.seek(og);
            if (more < 0) {
                offset_ -= more;
            } else if (more == 0) {
                if (!boundary) return OPB_FALSE;
                int want = kReadChunk;
                if (boundary > 0) {
                    const int64_t pos = offset_ + (int64_t)fr_.buffered();
                    if (pos >= boundary) return OPB_FALSE;
                    want = (int)std::min<int64_t>(boundary - pos, kReadChunk);
                } else if (!read_) {
                    want = 1 << 16;  // memory images: chunk size is not observable when the search is unbounded
                }
                const int got = source_read(fr_.reserve((size_t)want), want);
                if (got < 0) return OPB_EREAD;
                if (got == 0) return boundary < 0 ? OPB_FALSE : OPB_EBADLINK;
                fr_.wrote((size_t)got);
            } else {
                const int64_t at = offset_;
                offset_ += more;
                return at;
            }
        }
        return OPB_FALSE;
    }

    // ref op_fetch_headers / op_fetch_headers_impl (src/opusfile.cpp:157-296)
    int fetch_headers(OggPage* og_in, bool track_serials) {
        OggPage og;
        if (og_in) {
            og = *og_in;
        } else if (get_next_page(&og, offset_ + kHeaderWindow) < 0) {
            return OPB_ENOTFORMAT;
        }
        ready_ = kOpened;
        const int ret = fetch_headers_body(&og, track_serials);
        if (ret < 0) ready_ = kOpened;
        else if (og_in) *og_in = og;
        return ret;
    }
    int fetch_headers_body(OggPage* og, bool track_serials) {
        if (track_serials) serials_.clear();
        OggPacket op;
        while (og->bos()) {
            if (track_serials) {
                if (std::find(serials_.begin(), serials_.end(), og->serial()) != serials_.end()) return OPB_EBADHEADER;
                serials_.push_back(og->serial());
            }
            if (ready_ < kStreamSet) {
                pk_.reset(og->serial());
                pk_.page_in(*og);
                if (pk_.packet_out(&op) > 0) {
                    const int r = opusb_head_parse(&head, op.data, (size_t)op.bytes);
                    if (r >= 0) ready_ = kStreamSet;
                    else if (r != OPB_ENOTFORMAT) return r;
                }
            }
            if (get_next_page(og, offset_ + kHeaderWindow) < 0) return ready_ < kStreamSet ? OPB_ENOTFORMAT : OPB_EBADHEADER;
        }
        if (ready_ != kStreamSet) return OPB_ENOTFORMAT;
        if (pk_.serial() == og->serial()) pk_.page_in(*og);
        for (;;) {
            const int r = pk_.packet_out(&op);
            if (r < 0) return OPB_EBADHEADER;  // a hole inside the headers
            if (r > 0) {
                // the comment header (not parsed, like the reference); one more packet on the same page is dropped
                pk_.packet_out(&op);
                return 0;
            }
            for (;;) {
                if (get_next_page(og, offset_ + kHeaderWindow) < 0) return OPB_EBADHEADER;
                if (pk_.serial() == og->serial()) { pk_.page_in(*og); break; }
                if (og->bos()) return OPB_EBADHEADER;
            }
        }
    }

    // ref op_collect_audio_packets (src/opusfile.cpp:424)
    int32_t collect() {
        pend_.clear();
        op_pos_ = 0;
        int32_t total = 0;
        OggPacket op;
        for (;;) {
            const int r = pk_.packet_out(&op);
            if (!r) break;
            if (r < 0) return OPB_HOLE;
            const int d = packet_duration(op.data, op.bytes);
            if (d > 0) {
                pend_.push_back(Pending{op.data, (int)op.bytes, d, op.granulepos, op.eos});
                total += d;
            } else if (!pend_.empty()) {
                pend_.back().gp = op.granulepos;  // a packet with a bad TOC still passes its time stamp on
            }
        }
        return total;
    }

    void link_is_empty() { pcm_start_ = pcm_end_ = 0; }

    // ref op_find_initial_pcm_offset (src/opusfile.cpp:486).  0: ok, 1: *og is the first page of the next link.
    int find_initial_pcm_offset(OggPage* og) {
        const uint32_t serial = pk_.serial();
        int32_t total = 0;
        size_t count = 0;
        do {
            const int64_t at = get_next_page(og, -1);
            if (at < 0) {
                if (at < OPB_FALSE) return (int)at;
                if (head.pre_skip > 0) return OPB_EBADTIMESTAMP;
                link_is_empty();
                pend_.clear(); op_pos_ = 0;
                return 0;
            }
            if (og->bos()) {
                if (head.pre_skip > 0) return OPB_EBADTIMESTAMP;
                link_is_empty();
                pend_.clear(); op_pos_ = 0;
                return 1;
            }
            if (serial != og->serial()) continue;
            pk_.page_in(*og);
            do total = collect();
            while (total < 0);
            count = pend_.size();
        } while (count == 0);
        const int64_t page_gp = pend_.back().gp;
        if (page_gp == -1) return OPB_EBADTIMESTAMP;
        const bool last_page = pend_.back().eos;
        int64_t start = 0;
        if (!gp_add(&start, page_gp, -total)) {
            if (!last_page) return OPB_EBADTIMESTAMP;
            // less audio than the granule position claims on the only page: end trimming, start is 0 by definition
            start = 0;
            if (gp_cmp(page_gp, (int64_t)head.pre_skip) < 0) return OPB_EBADTIMESTAMP;
        }
        int64_t prev = start;
        size_t pi = 0;
        for (; pi < count; pi++) {
            Pending& p = pend_[pi];
            if (last_page) {
                int64_t played = 0;
                gp_diff(&played, page_gp, prev);
                const int64_t cut = p.duration - played;
                if (cut > 0) {
                    if (cut > p.duration) break;  // trimmed away completely
                    p.gp = prev = page_gp;
                    p.eos = true;
                    continue;
                }
            }
            gp_add(&p.gp, prev, p.duration);
            prev = p.gp;
        }
        pend_.resize(pi);
        discard_ = (int32_t)head.pre_skip;
        prev_packet_gp_ = pcm_start_ = start;
        return 0;
    }

    // ref op_make_decode_ready (src/opusfile.cpp:671): same layout -> OPUS_RESET_STATE, else a new decoder
    int make_decode_ready() {
        if (ready_ > kStreamSet) return 0;
        if (ready_ < kStreamSet) return OPB_EFAULT;
        if (head.stream_count != 1 || head.channel_count < 1 || head.channel_count > 2) return OPB_EIMPL;
        const bool same = have_dec_ && dec_streams_ == head.stream_count && dec_coupled_ == head.coupled_count &&
                          dec_channels_ == head.channel_count && memcmp(dec_map_, head.mapping, (size_t)head.channel_count) == 0;
        // A later link of a chain with another channel layout would need another decoder context (the reference creates a
        // new decoder, src/opusfile.cpp:683-697); the file's slot in a batch context has ONE channel count, and the samples
        // already queued are mapped with the first link's layout: such chains are refused.
        if (have_dec_ && !same) return OPB_EIMPL;
        Event e;
        memset(&e, 0, sizeof(e));
        e.kind = same ? Event::kResetSoft : Event::kResetHard;
        q_->push_back(e);
        have_dec_ = true;
        dec_streams_ = head.stream_count; dec_coupled_ = head.coupled_count; dec_channels_ = head.channel_count;
        memcpy(dec_map_, head.mapping, sizeof(dec_map_));
        ready_ = kInitSet;
        return 0;
    }

    // ref op_fetch_and_process_page, unseekable, _spanp = 1, _ignore_holes = 0 (src/opusfile.cpp:835-1133)
    int fetch_and_process_page() {
        for (;;) {
            OggPage og;
            const int64_t at = get_next_page(&og, -1);
            if (at < 0) return at < OPB_FALSE ? (int)at : OPB_EOF;
            if (ready_ >= kStreamSet && link_serial_ != og.serial()) {
                if (!og.bos()) continue;  // a page of some multiplexed stream
                if (ready_ >= kInitSet) {  // the next link of a chain begins (ref op_decode_clear :709)
                    pend_.clear(); op_pos_ = 0;
                    prev_packet_gp_ = -1;
                    ready_ = kOpened;
                }
            }
            if (ready_ < kStreamSet) {
                int ret;
                do {
                    ret = fetch_headers(&og, false);
                    if (ret < 0) return ret;
                    ret = find_initial_pcm_offset(&og);
                    if (ret < 0) return ret;
                    link_serial_ = pk_.serial();
                } while (ret > 0);
                if (pend_.empty()) continue;  // end trimming removed every packet of the first page
                return make_decode_ready();
            }
            // kStreamSet is only left over when a LATER link's headers parsed but its first audio page was refused
            // (find_initial_pcm_offset < 0 above): the reference then makes the decoder ready for that link's head on
            // the next page it sees and carries on (src/opusfile.cpp:949-952) -- found by tools/fuzz_demux.py
            if (ready_ == kStreamSet) {
                const int ret = make_decode_ready();
                if (ret < 0) return ret;
            }
            pk_.page_in(og);
            bool hole = false;
            int32_t total = collect();
            if (total < 0) {
                do total = collect();
                while (total < 0);
                hole = true;
                prev_packet_gp_ = -1;  // lost pages: time stamps restart from the next page's granule position
            }
            const size_t count = pend_.size();
            if (count > 0) {
                int64_t page_gp = pend_.back().gp;
                const bool last_page = pend_.back().eos;
                int64_t prev = prev_packet_gp_;
                int64_t diff = 0;
                if (prev == -1) {
                    // after a hole (the reference shares this with its raw-seek recovery, :974-1016)
                    if (last_page) {
                        pend_.clear(); op_pos_ = 0;
                        if (hole) return OPB_HOLE;
                        continue;
                    }
                    int32_t discard = 80 * 48;
                    prev = pcm_start_;
                    if (page_gp != -1) gp_add(&prev, page_gp, -total);
                    if (gp_diff(&diff, prev, pcm_start_)) {
                        const int32_t pre_skip = (int32_t)head.pre_skip;
                        if (diff >= 0 && diff <= std::max(0, pre_skip - 80 * 48)) discard = pre_skip - (int32_t)diff;
                    }
                    discard_ = discard;
                }
                if (page_gp == -1 && !gp_add(&page_gp, prev, total)) page_gp = pcm_end_;
                size_t pi = 0;
                if (last_page && gp_diff(&diff, page_gp, prev) && diff < total) {
                    // the last page carries fewer samples than its packets hold: trim from the end
                    int64_t gp = prev;
                    for (; pi < count; pi++) {
                        Pending& p = pend_[pi];
                        if (diff < 0 && INT64_MAX + diff < p.duration) diff = (int64_t)p.duration + 1;
                        else diff = p.duration - diff;
                        if (diff > 0) {
                            if (diff > p.duration) break;
                            gp = page_gp;
                            p.eos = true;
                        } else {
                            gp_add(&gp, gp, p.duration);
                        }
                        p.gp = gp;
                        gp_diff(&diff, page_gp, gp);
                    }
                } else {
                    // spread the page's granule position backwards over its packets
                    if (!gp_add(&prev, page_gp, -total)) prev = 0;
                    int32_t left = total;
                    for (; pi < count; pi++) {
                        Pending& p = pend_[pi];
                        int64_t gp = 0;
                        if (!gp_add(&gp, page_gp, -left)) gp = 0;
                        left -= p.duration;
                        gp_add(&gp, gp, p.duration);
                        p.gp = gp;
                    }
                }
                prev_packet_gp_ = prev;
                pend_.resize(pi);
            }
            if (hole) return OPB_HOLE;
            if (!pend_.empty()) return 0;
        }
    }

    // ref op_read_native's packet branch (src/opusfile.cpp:1204-1262): the play window of one packet
    void emit_packet(const Pending& p) {
        int trimmed = p.duration;
        if (p.eos) {
            int64_t diff = 0;
            if (gp_cmp(p.gp, prev_packet_gp_) <= 0) trimmed = 0;
            else if (gp_diff(&diff, p.gp, prev_packet_gp_)) trimmed = (int)std::min<int64_t>(diff, trimmed);
        }
        prev_packet_gp_ = p.gp;
        Event e;
        memset(&e, 0, sizeof(e));
        e.kind = Event::kPacket;
        e.off = arena.size();
        e.bytes = p.bytes;
        e.duration = p.duration;
        arena.insert(arena.end(), p.data, p.data + p.bytes);
        e.od_pos = std::min(trimmed, discard_);
        e.od_size = trimmed;
        e.status = kPending;
        // the reference only spends pre-skip once the packet decoded (:1234-1238); the scheduler gives it back on failure
        discard_ -= e.od_pos;
        q_->push_back(e);
    }
    void emit_code(int code) {
        Event e;
        memset(&e, 0, sizeof(e));
        e.kind = Event::kCode;
        e.code = code;
        q_->push_back(e);
    }

    opusb_read_func read_ = nullptr;
    void* user_ = nullptr;
    const uint8_t* mem_ = nullptr;
    size_t mem_len_ = 0, mem_pos_ = 0;
    bool clean_eof_ = false;

    OggFramer fr_;
    OggPacketizer pk_;
    std::deque<Event>* q_ = nullptr;
    std::vector<Pending> pend_;
    std::vector<uint32_t> serials_;
    size_t op_pos_ = 0;
    int ready_ = kNotOpen;
    int64_t offset_ = 0;
    uint32_t link_serial_ = 0;
    int64_t prev_packet_gp_ = -1, pcm_start_ = 0, pcm_end_ = -1;
    int32_t discard_ = 0;
    bool have_dec_ = false;
    int dec_streams_ = 0, dec_coupled_ = 0, dec_channels_ = 0;
    uint8_t dec_map_[2] = {0, 0};
};

// One file being played.
struct Track {
    OpusFileReader rd;
    std::deque<Event> q;        // events not yet consumed by read()
    size_t sched = 0;           // q[0..sched) went through the scheduler and have their results
    size_t planned = 0;         // q[sched..planned) belong to rounds that are in flight
    uint8_t last_toc = 0;       // TOC of the last frame handed to the decoder (filler frames repeat it)
    int open_result = OPB_EFAULT;
    bool terminal = false;      // the last queued event is the end of the data (keeps repeating)
    int terminal_code = 0;
    int cur_pos = 0;            // read cursor inside the front packet's play window
    bool cur_started = false;
    int discard_refund = 0;
    bool repoll = false;        // single-stream mode: ask the data source again after the end (the reference does)
    int64_t pcm_total = 0;

    ~Track() {
        for (Event& e : q) free(e.own);
    }
    int dec_channels() const { return rd.head.coupled_count ? 2 : 1; }

    // pull events until `frames` undecoded 20 ms frames are queued or the data ends
    void fill(int frames) {
        int have = 0;
        for (size_t i = sched; i < q.size(); i++)
            if (q[i].kind == Event::kPacket) have += std::max(1, q[i].duration / 960);
        while (!terminal && have < frames) {
            rd.next();
            const Event& e = q.back();
            if (e.kind == Event::kPacket) have += std::max(1, e.duration / 960);
            if (e.kind == Event::kCode && (e.code == 0 || e.code == OPB_EREAD)) {
                terminal = true;
                terminal_code = e.code;
            }
        }
    }

    // The samples of rounds <= `round` are about to lose their place in the scheduler's buffer: keep what is unread.
    void keep_unread(int round) {
        const int dc = dec_channels();
        for (size_t i = 0; i < sched && i < q.size(); i++) {
            Event& e = q[i];
            if (e.kind != Event::kPacket || e.status <= 0 || e.own || e.round > round) continue;
            const int n = e.od_size - e.od_pos;
            if (n <= 0) continue;
            e.own = (int16_t*)malloc(sizeof(int16_t) * (size_t)n * dc);
            if (e.own) {
                memcpy(e.own, e.pcm, sizeof(int16_t) * (size_t)n * dc);
                e.pcm = e.own;
            } else {  // out of host memory: the packet's samples are lost, read() reports it as a bad packet
                e.pcm = nullptr;
                e.status = OPB_EFAULT;
            }
        }
    }

    // ref op_read_stereo (src/opusfile.cpp:1293): one call
    int read(int16_t* out, int buf_size) {
        for (;;) {
            if (q.empty()) return terminal ? terminal_code : kPending;
            if (sched == 0) return kPending;  // the front event has not been through the scheduler yet
            Event& e = q.front();
            if (e.kind == Event::kResetSoft || e.kind == Event::kResetHard) {
                pop();
                continue;
            }
            if (e.kind == Event::kCode) {
                const int code = e.code;
                if (repoll || !(terminal && q.size() == 1)) pop();
                return code;
            }
            if (e.status == kPending) return kPending;
            if (e.status < 0) {
                pop();
                return OPB_EBADPACKET;
            }
            if (!cur_started) { cur_pos = e.od_pos; cur_started = true; }
            int n = e.od_size - cur_pos;
            if (n <= 0) {  // nothing of this packet is played (pre-skip or trimmed away): the reference moves on
                pop();
                continue;
            }
            n = std::min(n, buf_size >> 1);
            // decoder channels -> the file's channels (mapping, 255 = silence) -> stereo (mono is duplicated, :1310-1314)
            const int dc = dec_channels(), ch = rd.head.channel_count;
            const uint8_t* map = rd.head.mapping;
            const int16_t* src = e.pcm + (size_t)(cur_pos - e.od_pos) * dc;
            if (ch == 2 && dc == 2 && map[0] == 0 && map[1] == 1) {
                memcpy(out, src, sizeof(int16_t) * 2 * (size_t)n);
            } else if (ch == 1) {
                const int m = map[0];
                for (int i = 0; i < n; i++) out[2 * i] = out[2 * i + 1] = m == 255 ? (int16_t)0 : src[(size_t)i * dc + (m < dc ? m : 0)];
            } else {
                for (int i = 0; i < n; i++)
                    for (int c = 0; c < 2; c++) out[2 * i + c] = map[c] == 255 ? (int16_t)0 : src[(size_t)i * dc + (map[c] < dc ? map[c] : 0)];
            }
            cur_pos += n;
            // a fully read packet stays at the front until the next call finds it empty, exactly like od_buffer
            return n;
        }
    }

private:
    void pop() {
        free(q.front().own);
        q.pop_front();
        if (sched) sched--;
        if (planned) planned--;
        cur_started = false;
    }
};

// `fn(begin, end)` over [0, n) on up to `threads` host threads
template <class F>
void parallel_ranges(int n, int threads, F fn) {
    threads = std::max(1, std::min(threads, n / 64));
    if (threads == 1) { fn(0, n); return; }
    std::vector<std::thread> pool;
    for (int k = 1; k < threads; k++) pool.emplace_back([=]() { fn((int)((long long)n * k / threads), (int)((long long)n * (k + 1) / threads)); });
    fn(0, (int)((long long)n / threads));
    for (std::thread& th : pool) th.join();
}

// Packs the next frames of a group of tracks (same decoder channel count) into pinned rows and decodes them.
// Host work per round is spread over threads by track range: plan (which packets, how many bytes), copy the packet
// bytes into the pinned staging buffer, and after the GPU is done point every packet at its samples in the pinned PCM
// buffer -- the samples are not copied again until op_read_stereo hands them to the caller.
// Two rounds are kept in flight (decode_many_async): while the caller reads round r, round r+1 is crossing PCIe and
// round r+2 is in the kernels.  Staging is double buffered, PCM triple buffered; a packet still unread when its PCM
// buffer comes round again gets a private copy first.
class RoundScheduler {
public:
    enum { kDepth = 2, kPcmBufs = kDepth + 1 };
    ~RoundScheduler() { release(); }
    int init(int n_tracks, int channels, int frames, int device, int threads) {
        n_ = n_tracks; ch_ = channels; F_ = frames; threads_ = std::max(1, threads);
        int r = dm_create(&ctx_, n_, F_, ch_, device);
        if (r) return r;
        for (int k = 0; k < kPcmBufs; k++) {
            h_pcm_[k] = (int16_t*)dm_host_alloc(sizeof(int16_t) * (size_t)n_ * F_ * 960 * ch_);
            if (!h_pcm_[k]) return DM_ERR_NOMEM;
        }
        for (Stage& st : stage_) {
            st.frames = (dm_frame*)dm_host_alloc(sizeof(dm_frame) * (size_t)n_ * F_);
            st.status = (int32_t*)dm_host_alloc(sizeof(int32_t) * (size_t)n_ * F_);
            if (!st.frames || !st.status) return DM_ERR_NOMEM;
            st.plan.resize((size_t)n_);
        }
        return 0;
    }
    void release() {
        if (ctx_) dm_destroy(ctx_);  // waits for whatever is still in flight
        ctx_ = nullptr;
        for (int k = 0; k < kPcmBufs; k++) { dm_host_free(h_pcm_[k]); h_pcm_[k] = nullptr; }
        for (Stage& st : stage_) {
            dm_host_free(st.frames); dm_host_free(st.status); dm_host_free(st.bytes);
            st.frames = nullptr; st.status = nullptr; st.bytes = nullptr; st.bytes_cap = 0;
        }
    }
    long long launches() const { return ctx_ ? dm_launch_count(ctx_) : 0; }
    double t_keep = 0, t_plan = 0, t_reset = 0, t_stage = 0, t_wait = 0, t_finish = 0;  // host seconds per phase, summed over rounds

    // One round over tracks[0..n). Returns the number of frames decoded (0: nothing left), or < 0 (DM_ERR_*).
    long long round(Track** tracks) {
        long long r = top_up(tracks);
        if (r < 0) return r;
        if (submitted_ == completed_) return 0;
        const long long done = complete(tracks);
        if (done < 0) return done;
        r = top_up(tracks);
        if (r < 0) return r;
        return done;
    }

private:
    struct TrackPlan {
        size_t n_ev;               // events [sched, sched + n_ev) of the track's queue belong to this round once the rounds
                                   // before it are finished (reads only pop events in front of `sched`, which moves along)
        size_t bytes, byte0;       // packet bytes of this track, and where they start in the staging buffer
        int used;                  // frames of the row in use
        int packets;               // packets handled (decoded or refused)
        int reset;                 // 0 none, 1 OPUS_RESET_STATE, 2 new decoder -- before the round
    };
    struct Stage {
        dm_frame* frames = nullptr;
        int32_t* status = nullptr;
        uint8_t* bytes = nullptr;
        size_t bytes_cap = 0;
        std::vector<TrackPlan> plan;
        long long ticket = -1, nframes = 0;
    };

    // keep the pipeline full; a decoder reset needs it empty (it rewrites stream state on the device)
    long long top_up(Track** tracks) {
        while (submitted_ - completed_ < kDepth) {
            if (want_drain_ && submitted_ != completed_) break;
            const long long r = submit(tracks);
            if (r <= 0) return r;
        }
        return 1;
    }

    // keep, plan, reset, stage, enqueue.  Returns > 0 when a round was enqueued, 0 when nothing is left.
    long long submit(Track** tracks) {
        Stage& st = stage_[submitted_ % kDepth];
        const bool idle = submitted_ == completed_;
        if (idle) want_drain_ = false;
        // the PCM buffer of this round was last filled kPcmBufs rounds ago: whoever has not read those samples keeps a copy
        auto now = []() { return std::chrono::steady_clock::now(); };
        auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
        auto c0 = now();
        if (submitted_ >= kPcmBufs) {
            const int old = submitted_ - kPcmBufs;
            parallel_ranges(n_, threads_, [&](int s0, int s1) { for (int s = s0; s < s1; s++) tracks[s]->keep_unread(old); });
        }
        auto c1 = now();
        t_keep += secs(c0, c1);
        // pass 1: per track, choose the events of this round and lay out their frames (bytes relative to the track)
        parallel_ranges(n_, threads_, [&](int s0, int s1) {
            for (int s = s0; s < s1; s++) plan_track(st, *tracks[s], s, idle);
        });
        auto c2 = now();
        t_plan += secs(c1, c2);
        std::vector<int32_t> soft, hard;
        size_t nbytes = 0, nev = 0;
        long long nframes = 0;
        for (int s = 0; s < n_; s++) {
            TrackPlan& pl = st.plan[(size_t)s];
            pl.byte0 = nbytes;
            nbytes += pl.bytes;
            nframes += pl.used;
            nev += pl.n_ev;
            if (pl.reset == 1) soft.push_back(s);
            if (pl.reset == 2) hard.push_back(s);
        }
        if (nev == 0) return 0;  // every queue is planned to its end
        if (!soft.empty()) { int r = dm_reset_state(ctx_, soft.data(), (int)soft.size()); if (r) return r; }
        if (!hard.empty()) {
            const bool all = (int)hard.size() == n_;
            int r = dm_reset(ctx_, all ? nullptr : hard.data(), all ? 0 : (int)hard.size());
            if (r) return r;
        }
        auto c3 = now();
        t_reset += secs(c2, c3);
        st.ticket = -1;
        st.nframes = nframes;
        if (nframes > 0) {
            if (nbytes + 16 > st.bytes_cap) {
                dm_host_free(st.bytes);
                st.bytes_cap = nbytes * 2 + 4096;
                st.bytes = (uint8_t*)dm_host_alloc(st.bytes_cap);
                if (!st.bytes) return DM_ERR_NOMEM;
            }
            // pass 2: packet bytes into the pinned staging buffer, frame offsets made absolute
            parallel_ranges(n_, threads_, [&](int s0, int s1) {
                for (int s = s0; s < s1; s++) stage_track(st, *tracks[s], s);
            });
            dm_batch in;
            memset(&in, 0, sizeof(in));
            in.n_frames = F_; in.mode = DM_MODE_OPUS; in.bytes = st.bytes; in.n_bytes = nbytes; in.frames = st.frames;
            dm_out out = {h_pcm_[submitted_ % kPcmBufs], st.status};
            const int r = decode_many_async(ctx_, &in, &out, &st.ticket);
            if (r) return r;
        }
        t_stage += secs(c3, now());
        // the events of this round are spoken for
        for (int s = 0; s < n_; s++) tracks[s]->planned += st.plan[(size_t)s].n_ev;
        submitted_++;
        return nframes > 0 ? nframes : 1;
    }
    // wait for the oldest round in flight, results -> events
    long long complete(Track** tracks) {
        Stage& st = stage_[completed_ % kDepth];
        const auto w0 = std::chrono::steady_clock::now();
        if (st.ticket >= 0) {
            const int r = dm_wait(ctx_, st.ticket);
            if (r) return r;
        }
        const auto w1 = std::chrono::steady_clock::now();
        const int16_t* pcm = h_pcm_[completed_ % kPcmBufs];
        const int round = completed_;
        parallel_ranges(n_, threads_, [&](int s0, int s1) {
            for (int s = s0; s < s1; s++) finish_track(st, *tracks[s], s, pcm, round);
        });
        t_wait += std::chrono::duration<double>(w1 - w0).count();
        t_finish += std::chrono::duration<double>(std::chrono::steady_clock::now() - w1).count();
        completed_++;
        return st.nframes > 0 ? st.nframes : 1;
    }

    static bool packet_layout(const Track& t, const Event& e, int ch, uint8_t* toc, int16_t* size, int* off, int* count) {
        const uint8_t* pkt = t.rd.arena.data() + e.off;
        *count = opusb_packet_parse(pkt, e.bytes, toc, nullptr, size, off);
        // undecodable without undefined behaviour in the reference: reported, never decoded
        return !(*count < 0 || opusb_packet_get_samples_per_frame(pkt, 48000) != 960 || (!(*toc & 0x80) && ((*toc & 4) ? 2 : 1) != ch));
    }
    void plan_track(Stage& st, Track& t, int s, bool resets_allowed) {
        TrackPlan& pl = st.plan[(size_t)s];
        pl.bytes = 0; pl.used = 0; pl.reset = 0; pl.packets = 0;
        uint8_t last_toc = t.last_toc;
        size_t i = t.planned;
        for (; i < t.q.size(); i++) {
            Event& e = t.q[i];
            if (e.kind == Event::kCode) continue;
            if (e.kind == Event::kResetSoft || e.kind == Event::kResetHard) {
                if (pl.used || pl.reset) break;  // a reset sits at the start of a round
                if (!resets_allowed) { want_drain_ = true; break; }
                pl.reset = e.kind == Event::kResetSoft ? 1 : 2;
                continue;
            }
            uint8_t toc = 0;
            int16_t size[48];
            int off = 0, count = 0;
            if (!packet_layout(t, e, ch_, &toc, size, &off, &count)) { e.round = -1; pl.packets++; continue; }
            if (pl.used + count > F_) break;
            pl.packets++;
            e.round = pl.used;  // first frame of the packet in the row (the round number replaces it in finish_track)
            size_t pos = pl.bytes + (size_t)off;
            for (int k = 0; k < count; k++) {
                dm_frame& f = st.frames[(size_t)s * F_ + pl.used + k];
                f.offset = (uint32_t)pos;  // relative to the track's bytes until stage_track()
                f.len = (uint16_t)size[k];
                f.toc = toc;
                f.flags = k ? DMF_CONT : 0;
                pos += (size_t)size[k];
            }
            pl.bytes += (size_t)e.bytes;
            pl.used += count;
            last_toc = toc;
        }
        pl.n_ev = i - t.planned;
        t.last_toc = last_toc;
        for (int k = pl.used; k < F_; k++) {
            dm_frame& f = st.frames[(size_t)s * F_ + k];
            f.offset = 0; f.len = 0; f.toc = last_toc; f.flags = DMF_NULL;
        }
    }
    void stage_track(Stage& st, Track& t, int s) {
        const TrackPlan& pl = st.plan[(size_t)s];
        size_t pos = pl.byte0;
        for (size_t i = t.planned; i < t.planned + pl.n_ev; i++) {
            const Event& e = t.q[i];
            if (e.kind != Event::kPacket || e.round < 0) continue;
            memcpy(st.bytes + pos, t.rd.arena.data() + e.off, (size_t)e.bytes);
            pos += (size_t)e.bytes;
        }
        for (int k = 0; k < pl.used; k++) st.frames[(size_t)s * F_ + k].offset += (uint32_t)pl.byte0;
    }
    void finish_track(Stage& st, Track& t, int s, const int16_t* pcm, int round) {
        const TrackPlan& pl = st.plan[(size_t)s];
        for (size_t i = t.sched; i < t.sched + pl.n_ev; i++) {
            Event& e = t.q[i];
            if (e.kind != Event::kPacket) continue;
            const int first = e.round;
            int status = first < 0 ? OPB_EBADPACKET : 0;
            const int count = first < 0 ? 0 : e.duration / 960;
            for (int k = 0; k < count && status >= 0; k++) {
                const int32_t v = st.status[(size_t)s * F_ + first + k];
                if (v < 0) status = OPB_EBADPACKET;
                else status += v;
            }
            if (status >= 0 && status != e.duration) status = OPB_EBADPACKET;
            // a failed packet does not consume pre-skip (ref :1232-1238): give it to the packets behind it
            if (status < 0 && e.od_pos > 0) t.discard_refund += e.od_pos;
            else if (status >= 0 && t.discard_refund > 0) {
                const int extra = std::min(t.discard_refund, e.od_size - e.od_pos);
                e.od_pos += extra;
                t.discard_refund -= extra;
            }
            e.status = status;
            e.round = round;
            if (status > 0) {
                e.pcm = pcm + (((size_t)s * F_ + first) * 960 + (size_t)e.od_pos) * ch_;
                t.pcm_total += std::max(0, e.od_size - e.od_pos);
            }
        }
        t.sched += pl.n_ev;
    }

    dm_ctx* ctx_ = nullptr;
    int n_ = 0, ch_ = 0, F_ = 0, threads_ = 1;
    int submitted_ = 0, completed_ = 0;   // rounds handed to the GPU / rounds whose results are with the tracks
    std::atomic<bool> want_drain_{false}; // a track waits with a decoder reset until nothing is in flight
    int16_t* h_pcm_[kPcmBufs] = {};
    Stage stage_[kDepth];
};

}  // namespace

// ------------------------------------------------------------------------------------------------
struct OggOpusFileB200 {
    Track track;
    RoundScheduler sched;
    int frames;
};

struct opfb_set {
    int n_files, device, frames, n_threads;
    std::vector<Track*> tracks;
    std::vector<Track*> group[2];   // by decoder channel count (1, 2)
    RoundScheduler* sched[2];
    bool prepared;
    double demux_s, decode_s;
    long long packets, frames_done, ogg_bytes;
};

extern "C" {

// ref opus_head_parse (src/opusfile.cpp:1333)
int opusb_head_parse(OpusHeadB200* out, const uint8_t* d, size_t len) {
    OpusHeadB200 h;
    memset(&h, 0, sizeof(h));
    if (len < 8 || memcmp(d, "OpusHead", 8) != 0) return OPB_ENOTFORMAT;
    if (len < 9) return OPB_EBADHEADER;
    h.version = d[8];
    if (h.version > 15) return OPB_EVERSION;
    if (len < 19) return OPB_EBADHEADER;
    h.channel_count = d[9];
    h.pre_skip = (unsigned)(d[10] | d[11] << 8);
    h.input_sample_rate = (uint32_t)d[12] | (uint32_t)d[13] << 8 | (uint32_t)d[14] << 16 | (uint32_t)d[15] << 24;
    const int g = d[16] | d[17] << 8;
    h.output_gain = (g ^ 0x8000) - 0x8000;
    h.mapping_family = d[18];
    if (h.mapping_family == 0) {
        if (h.channel_count < 1 || h.channel_count > 2) return OPB_EBADHEADER;
        if (h.version <= 1 && len > 19) return OPB_EBADHEADER;
        h.stream_count = 1;
        h.coupled_count = h.channel_count - 1;
        h.mapping[0] = 0;
        h.mapping[1] = 1;
    } else if (h.mapping_family == 1) {
        if (h.channel_count < 1 || h.channel_count > 8) return OPB_EBADHEADER;
        const size_t size = 21 + (size_t)h.channel_count;
        if (len < size || (h.version <= 1 && len > size)) return OPB_EBADHEADER;
        h.stream_count = d[19];
        if (h.stream_count < 1) return OPB_EBADHEADER;
        h.coupled_count = d[20];
        if (h.coupled_count > h.stream_count) return OPB_EBADHEADER;
        for (int c = 0; c < h.channel_count; c++)
            if (d[21 + c] >= h.stream_count + h.coupled_count && d[21 + c] != 255) return OPB_EBADHEADER;
        for (int c = 0; c < h.channel_count && c < 2; c++) h.mapping[c] = d[21 + c];
    } else if (h.mapping_family == 255) {
        return OPB_EIMPL;
    } else {
        return OPB_EBADHEADER;
    }
    if (out) *out = h;
    return 0;
}

// ---- single stream ----
OggOpusFileB200* opusb_init_decoder(opusb_read_func read, void* user, int device, int* error) {
    int err = 0;
    OggOpusFileB200* of = nullptr;
    if (!read) {
        err = OPB_EINVAL;
    } else {
        of = new (std::nothrow) OggOpusFileB200();
        if (!of) err = OPB_EFAULT;
    }
    if (of) {
        of->frames = 48;
        of->track.rd.set_callback(read, user);
        of->track.repoll = true;
        err = of->track.rd.open(&of->track.q);
        of->track.open_result = err;
        if (!err) err = of->sched.init(1, of->track.dec_channels(), of->frames, device, 1);
        if (err) { delete of; of = nullptr; }
    }
    if (error) *error = err;
    return of;
}
int opusb_read_stereo(OggOpusFileB200* of, int16_t* pcm, int buf_size) {
    if (!of || !pcm) return OPB_EINVAL;
    for (;;) {
        const int r = of->track.read(pcm, buf_size);
        if (r != kPending) return r;
        // decode ahead: the next frames of the file go through decode_many() as one round
        if (of->track.q.empty()) of->track.terminal = false;  // ask the data source again, like the reference
        of->track.fill(of->frames - 6);
        Track* t = &of->track;
        const long long n = of->sched.round(&t);
        if (n < 0) return OPB_EFAULT;
    }
}
void opusb_free_decoder(OggOpusFileB200* of) { delete of; }
const OpusHeadB200* opusb_head(const OggOpusFileB200* of) { return of ? &of->track.rd.head : nullptr; }

// ---- many files ----
int opfb_set_create(opfb_set** out, int n_files, int device, int frames_per_round, int n_threads) {
    if (!out || n_files < 1 || frames_per_round < 6) return OPB_EINVAL;
    opfb_set* s = new (std::nothrow) opfb_set();
    if (!s) return OPB_EFAULT;
    s->n_files = n_files; s->device = device; s->frames = frames_per_round;
    s->n_threads = n_threads > 0 ? n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    s->tracks.assign((size_t)n_files, nullptr);
    s->sched[0] = s->sched[1] = nullptr;
    s->prepared = false;
    s->demux_s = s->decode_s = 0;
    s->packets = s->frames_done = s->ogg_bytes = 0;
    *out = s;
    return 0;
}
void opfb_set_destroy(opfb_set* s) {
    if (!s) return;
    for (Track* t : s->tracks) delete t;
    delete s->sched[0];
    delete s->sched[1];
    delete s;
}
int opfb_set_open_memory(opfb_set* s, int index, const uint8_t* data, size_t len, int clean_eof) {
    if (!s || index < 0 || index >= s->n_files || (!data && len) || s->prepared) return OPB_EINVAL;
    delete s->tracks[(size_t)index];
    Track* t = new (std::nothrow) Track();
    if (!t) return OPB_EFAULT;
    t->rd.set_memory(data, len, clean_eof != 0);
    s->tracks[(size_t)index] = t;
    s->ogg_bytes += (long long)len;
    return 0;
}
int opfb_set_prepare(opfb_set* s) {
    if (!s || s->prepared) return OPB_EINVAL;
    const auto t0 = std::chrono::steady_clock::now();
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= s->n_files) break;
            Track* t = s->tracks[(size_t)i];
            if (!t) continue;
            t->open_result = t->rd.open(&t->q);
            if (t->open_result == 0) t->fill(1 << 30);
        }
    };
    std::vector<std::thread> pool;
    const int nt = std::min(s->n_threads, s->n_files);
    for (int k = 1; k < nt; k++) pool.emplace_back(work);
    work();
    for (std::thread& th : pool) th.join();
    s->demux_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (Track* t : s->tracks) {
        if (!t || t->open_result) continue;
        s->group[t->dec_channels() - 1].push_back(t);
        for (const Event& e : t->q) s->packets += e.kind == Event::kPacket;
    }
    for (int g = 0; g < 2; g++) {
        if (s->group[g].empty()) continue;
        s->sched[g] = new (std::nothrow) RoundScheduler();
        if (!s->sched[g]) return OPB_EFAULT;
        const int r = s->sched[g]->init((int)s->group[g].size(), g + 1, s->frames, s->device, s->n_threads);
        if (r) return r;
    }
    s->prepared = true;
    return 0;
}
int opfb_set_open_result(const opfb_set* s, int index) {
    if (!s || index < 0 || index >= s->n_files || !s->tracks[(size_t)index]) return OPB_EINVAL;
    return s->tracks[(size_t)index]->open_result;
}
// One round for every group. Returns 1 when frames were decoded, 0 when every file is done, < 0 on failure.
int opfb_set_decode_round(opfb_set* s) {
    if (!s || !s->prepared) return OPB_EINVAL;
    const auto t0 = std::chrono::steady_clock::now();
    int any = 0;
    for (int g = 0; g < 2; g++) {
        if (!s->sched[g]) continue;
        const long long n = s->sched[g]->round(s->group[g].data());
        if (n < 0) return (int)n;
        if (n > 0) { any = 1; s->frames_done += n; }
    }
    s->decode_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return any;
}
int opfb_set_decode(opfb_set* s) {
    for (;;) {
        const int r = opfb_set_decode_round(s);
        if (r <= 0) return r;
    }
}
int opfb_set_read_stereo(opfb_set* s, int index, int16_t* pcm, int buf_size) {
    if (!s || index < 0 || index >= s->n_files || !s->tracks[(size_t)index] || !pcm) return OPB_EINVAL;
    Track* t = s->tracks[(size_t)index];
    if (t->open_result) return t->open_result;
    const int r = t->read(pcm, buf_size);
    return r;
}
// Benchmark consumer: every file is read with op_read_stereo calls of `buf_size` samples until it reports OPFB_PENDING or
// its end, on the set's host threads (files are independent).  Returns the stereo frames read.
int64_t opfb_set_drain(opfb_set* s, int buf_size) {
    if (!s || !s->prepared || buf_size < 2) return OPB_EINVAL;
    std::atomic<long long> total(0);
    parallel_ranges(s->n_files, s->n_threads, [&](int i0, int i1) {
        std::vector<int16_t> buf((size_t)buf_size);
        long long n_sum = 0;
        for (int i = i0; i < i1; i++) {
            Track* t = s->tracks[(size_t)i];
            if (!t || t->open_result) continue;
            for (;;) {
                const int n = t->read(buf.data(), buf_size);
                if (n <= 0) break;
                n_sum += n;
            }
        }
        total += n_sum;
    });
    return (int64_t)total.load();
}
int64_t opfb_set_pcm_total(const opfb_set* s, int index) {
    if (!s || index < 0 || index >= s->n_files || !s->tracks[(size_t)index]) return OPB_EINVAL;
    return s->tracks[(size_t)index]->pcm_total;
}
int opfb_set_stats(const opfb_set* s, double* out, int cap) {
    if (!s || !out) return 0;
    long long launches = 0;
    for (int g = 0; g < 2; g++) if (s->sched[g]) launches += s->sched[g]->launches();
    double ph[6] = {0, 0, 0, 0, 0, 0};
    for (int g = 0; g < 2; g++)
        if (s->sched[g]) {
            const RoundScheduler& r = *s->sched[g];
            ph[0] += r.t_keep; ph[1] += r.t_plan; ph[2] += r.t_reset; ph[3] += r.t_stage; ph[4] += r.t_wait; ph[5] += r.t_finish;
        }
    const double v[12] = {s->demux_s, s->decode_s, (double)s->packets, (double)s->frames_done, (double)s->ogg_bytes, (double)launches,
                          ph[0], ph[1], ph[2], ph[3], ph[4], ph[5]};
    const int n = cap < 12 ? cap : 12;
    for (int i = 0; i < n; i++) out[i] = v[i];
    return n;
}

// Synthetic Ogg Opus file for benchmarks and tests (SURVEY 8d config 1 layout): OpusHead (version 1, family 0), empty
// OpusTags, then `n_packets` code-0 packets of `packet_len` bytes (TOC + xorshift payload, the SURVEY 8c generator),
// `per_page` packets per page, granule positions 960 per packet, EOS on the last page.  Returns the bytes written, or
// the size needed when `out` is NULL / `cap` is too small.
size_t opfb_synth_ogg(uint8_t* out, size_t cap, int channels, int pre_skip, int toc, int packet_len, int n_packets, int per_page,
                      uint64_t seed, uint32_t serial) {
    if (per_page < 1) per_page = 1;
    if (per_page > 255 / ((packet_len / 255) + 1)) per_page = 255 / ((packet_len / 255) + 1);
    const OggCrc& C = ogg_crc();
    size_t pos = 0;
    uint32_t seq = 0;
    uint64_t rng = seed;
    auto page = [&](const uint8_t* const* pk, const int* len, int n, int64_t gp, int flags) {
        uint8_t hdr[27 + 255];
        int nseg = 0;
        size_t body = 0;
        for (int i = 0; i < n; i++) {
            int l = len[i];
            while (l >= 255) { hdr[27 + nseg++] = 255; l -= 255; }
            hdr[27 + nseg++] = (uint8_t)l;
            body += (size_t)len[i];
        }
        memcpy(hdr, "OggS", 4);
        hdr[4] = 0; hdr[5] = (uint8_t)flags;
        for (int i = 0; i < 8; i++) hdr[6 + i] = (uint8_t)((uint64_t)gp >> (8 * i));
        for (int i = 0; i < 4; i++) { hdr[14 + i] = (uint8_t)(serial >> (8 * i)); hdr[18 + i] = (uint8_t)(seq >> (8 * i)); hdr[22 + i] = 0; }
        hdr[26] = (uint8_t)nseg;
        const size_t total = 27 + (size_t)nseg + body;
        if (out && pos + total <= cap) {
            uint8_t* p = out + pos;
            memcpy(p, hdr, 27 + (size_t)nseg);
            size_t q = 27 + (size_t)nseg;
            for (int i = 0; i < n; i++) { memcpy(p + q, pk[i], (size_t)len[i]); q += (size_t)len[i]; }
            const uint32_t crc = C.update(0, p, total);
            for (int i = 0; i < 4; i++) p[22 + i] = (uint8_t)(crc >> (8 * i));
        }
        pos += total;
        seq++;
    };
    uint8_t head[19] = {'O', 'p', 'u', 's', 'H', 'e', 'a', 'd', 1, (uint8_t)channels, (uint8_t)(pre_skip & 255), (uint8_t)(pre_skip >> 8),
                        0x80, 0xBB, 0, 0, 0, 0, 0};
    uint8_t tags[8 + 4 + 4 + 4] = {'O', 'p', 'u', 's', 'T', 'a', 'g', 's', 4, 0, 0, 0, 'b', '2', '0', '0', 0, 0, 0, 0};
    const uint8_t* hp = head; int hl = (int)sizeof(head);
    page(&hp, &hl, 1, 0, 2);
    const uint8_t* tp = tags; int tl = (int)sizeof(tags);
    page(&tp, &tl, 1, 0, 0);
    std::vector<uint8_t> buf((size_t)per_page * (size_t)packet_len);
    std::vector<const uint8_t*> ptrs((size_t)per_page);
    std::vector<int> lens((size_t)per_page, packet_len);
    for (int i = 0; i < n_packets; i += per_page) {
        const int n = std::min(per_page, n_packets - i);
        for (int k = 0; k < n; k++) {
            uint8_t* p = buf.data() + (size_t)k * packet_len;
            p[0] = (uint8_t)toc;
            for (int j = 1; j < packet_len; j++) {
                rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
                p[j] = (uint8_t)(rng >> 11);
            }
            ptrs[(size_t)k] = p;
        }
        page(ptrs.data(), lens.data(), n, (int64_t)960 * (i + n), i + n >= n_packets ? 4 : 0);
    }
    return pos;
}

// Test access to the demux stage alone (no GPU): the events of one memory image, rows of
// (kind, code, bytes, duration, od_pos, od_size).  Returns the number of events, *open_result as opus_init_decoder.
int opfb_debug_events(const uint8_t* data, size_t len, int clean_eof, int32_t* out, int cap, int* open_result) {
    Track t;
    t.rd.set_memory(data, len, clean_eof != 0);
    t.open_result = t.rd.open(&t.q);
    if (open_result) *open_result = t.open_result;
    if (t.open_result == 0) t.fill(1 << 30);
    int n = 0;
    for (const Event& e : t.q) {
        if (n < cap && out) {
            int32_t* r = out + 6 * n;
            r[0] = (int32_t)e.kind; r[1] = e.code; r[2] = e.bytes; r[3] = e.duration; r[4] = e.od_pos; r[5] = e.od_size;
        }
        n++;
    }
    return n;
}

}  // extern "C"
