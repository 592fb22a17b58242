// seg_pack.cpp — the feeder's compact segment format ("packed16"): half the PCIe bytes of (int32 start, int32 end).
//
// Segments are grouped in blocks of 256 slots.  Block b has one int32 anchor; slot k holds a segment
// [anchor + off, anchor + off + len) with off,len as uint16 (len == 0: empty slot).  BAM order keeps consecutive
// starts within a few hundred bases, so almost every block is full; a new block is opened when the next start
// does not fit 16 bits from the anchor (sparse regions), and segments longer than 65535 are split (harmless for depth).
#include <stdint.h>
#include <string.h>
#include "../../../include/goleft_b200.h"

extern "C" {

int64_t gl_pack_segments16_bound(int64_t n) {
    // worst case: every segment opens its own block (and long ones split): callers size buffers with this for the
    // common case and retry with the exact count returned by gl_pack_segments16 when it does not fit
    return n / 256 + 2 + n / 64;
}

int gl_pack_segments16(const int32_t* start, const int32_t* end, int64_t n, int32_t* anchors, uint16_t* off, uint16_t* len,
                       int64_t cap_blocks, int64_t* n_blocks) {
    if (n < 0 || !n_blocks || (n > 0 && (!start || !end))) return GL_EINVAL;
    int64_t nb = 0;
    int cnt = 256;                                      // forces a new block for the first segment
    int32_t anchor = 0;
    bool fits = anchors && off && len;
    auto open_block = [&](int32_t first_start) {
        if (nb > 0 && fits && cnt < 256) {              // clear the unused tail of the previous block
            memset(off + (nb - 1) * 256 + cnt, 0, (size_t)(256 - cnt) * 2);
            memset(len + (nb - 1) * 256 + cnt, 0, (size_t)(256 - cnt) * 2);
        }
        // leave 16 kb of room below the first start: later segments of a BAM-ordered stream may start a little earlier
        anchor = first_start > INT32_MIN + 16384 ? first_start - 16384 : first_start;
        if (nb >= cap_blocks) fits = false;
        if (fits) anchors[nb] = anchor;
        nb++;
        cnt = 0;
    };
    for (int64_t i = 0; i < n; i++) {
        int64_t s = start[i];
        const int64_t e = end[i];
        if (e <= s) continue;                           // empty / inverted segments carry no depth
        while (s < e) {
            const int64_t piece = (e - s > 65535) ? 65535 : e - s;
            const int64_t o = s - anchor;
            if (cnt == 256 || o < 0 || o > 65535) open_block((int32_t)s);
            if (fits) {
                off[(nb - 1) * 256 + cnt] = (uint16_t)(s - anchor);
                len[(nb - 1) * 256 + cnt] = (uint16_t)piece;
            }
            cnt++;
            s += piece;
        }
    }
    if (nb > 0 && fits && cnt < 256) {
        memset(off + (nb - 1) * 256 + cnt, 0, (size_t)(256 - cnt) * 2);
        memset(len + (nb - 1) * 256 + cnt, 0, (size_t)(256 - cnt) * 2);
    }
    *n_blocks = nb;
    return fits || nb == 0 ? GL_OK : GL_ERANGE;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ "packed8"
// A quarter of the PCIe bytes of (int32 start, int32 end) for short-read data.  Blocks of 64 slots (kP8); block b has
// one int32 anchor (= start of its slot 0); slot k holds uint8 dstart (start[k] - start[k-1]; 0 for slot 0) and uint8
// len (0: filler / empty slot).  64 slots = about 320 bases at 30x: small enough that a kernel can fetch the blocks
// of one 4096-base tile by a search over the anchors alone (the anchors are sorted).  Depth does not depend on segment order, so the packer first puts the segments in start
// order (BAM order is nearly sorted: only the later blocks of spliced/deleted reads are out of place), cuts segments
// longer than 255 into pieces, bridges gaps of more than 255 bases with filler slots when that is cheaper than
// opening a new block.  The device rebuilds starts with one prefix sum per block (depth_unpack8_kernel).
#include <algorithm>
#include <utility>
#include <vector>

extern "C" {

static const int kP8 = 64;

int64_t gl_pack_segments8_bound(int64_t n) { return n / kP8 + 2 + n / 8; }

int gl_pack_segments8(const int32_t* start, const int32_t* end, int64_t n, int32_t* anchors, uint8_t* dstart, uint8_t* len,
                      int64_t cap_blocks, int64_t* n_blocks) {
    if (n < 0 || !n_blocks || (n > 0 && (!start || !end))) return GL_EINVAL;
    // pieces in start order
    std::vector<std::pair<int32_t, int32_t>> seg;       // (start, length <= 255)
    seg.reserve((size_t)n + 16);
    bool sorted = true;
    int64_t prev = INT64_MIN;
    for (int64_t i = 0; i < n; i++) {
        int64_t s = start[i];
        const int64_t e = end[i];
        if (e <= s) continue;
        if (s < prev) sorted = false;
        prev = s;
        if (e - s > 255) sorted = false;                // the pieces of a long segment interleave with its followers
        while (s < e) {
            const int64_t piece = e - s > 255 ? 255 : e - s;
            seg.emplace_back((int32_t)s, (int32_t)piece);
            s += piece;
        }
    }
    if (!sorted) {
        // nearly sorted input: insertion sort with a work bound, std::stable_sort when the disorder is not local
        const size_t m = seg.size();
        size_t budget = 64 * m + 1024;
        bool done = true;
        for (size_t i = 1; i < m; i++) {
            if (seg[i].first >= seg[i - 1].first) continue;
            const std::pair<int32_t, int32_t> v = seg[i];
            size_t j = i;
            while (j > 0 && seg[j - 1].first > v.first) {
                seg[j] = seg[j - 1];
                j--;
                if (--budget == 0) break;
            }
            seg[j] = v;
            if (budget == 0) { done = false; break; }
        }
        if (!done) std::stable_sort(seg.begin(), seg.end(), [](const std::pair<int32_t, int32_t>& a, const std::pair<int32_t, int32_t>& b) { return a.first < b.first; });
    }
    int64_t nb = 0;
    int cnt = kP8;
    int64_t last = 0;                                   // start of the previous slot
    bool fits = anchors && dstart && len;
    auto close_block = [&]() {
        if (nb > 0 && fits && cnt < kP8) {
            memset(dstart + (nb - 1) * kP8 + cnt, 0, (size_t)(kP8 - cnt));
            memset(len + (nb - 1) * kP8 + cnt, 0, (size_t)(kP8 - cnt));
        }
    };
    auto put = [&](int d, int l) {
        if (fits) { dstart[(nb - 1) * kP8 + cnt] = (uint8_t)d; len[(nb - 1) * kP8 + cnt] = (uint8_t)l; }
        cnt++;
    };
    for (const auto& sg : seg) {
        const int64_t s = sg.first;
        bool open = cnt == kP8;
        if (!open) {
            const int64_t gap = s - last;
            const int64_t fillers = gap > 255 ? (gap - 1) / 255 : 0;
            if (fillers + 1 > kP8 - cnt) open = true;   // bridging would not fit in this block: a new anchor is cheaper
            else {
                for (int64_t f = 0; f < fillers; f++) { put(255, 0); last += 255; }
            }
        }
        if (open) {
            close_block();
            if (nb >= cap_blocks) fits = false;
            if (fits) anchors[nb] = (int32_t)s;
            nb++;
            cnt = 0;
            last = s;
        }
        put((int)(s - last), sg.second);
        last = s;
    }
    close_block();
    *n_blocks = nb;
    return fits || nb == 0 ? GL_OK : GL_ERANGE;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ parallel packed8
// gl_pack_segments8_mt: the same format from `threads` workers.  The input (BAM order: starts sorted up to one read's
// span) is cut into P index chunks; x_k = the smallest start in chunks >= k, so the position ranges [x_k, x_{k+1}) are
// disjoint, increasing, and every piece of chunk k starts at or after x_k.  Worker k keeps the pieces of its chunk that
// start below x_{k+1} ("own", nearly sorted: bounded insertion sort) and sets aside the few that start later (second
// blocks of deletion / spliced reads at a chunk edge, tails of long segments); after every earlier chunk has published
// its set-aside list, worker k merges the ones that fall in its range, encodes its blocks privately, and copies them
// behind the blocks of worker k-1.  Anchors come out globally sorted, which is what K_tileidx8 needs.
#include <atomic>
#include <limits.h>
#include <memory>
#include <mutex>
#include "thread_pool.h"

namespace {

typedef std::pair<int32_t, int32_t> Piece;          // (start, length 1..255)

struct P8Out { std::vector<int32_t> anchors; std::vector<uint8_t> ds, len; int64_t nb = 0; };

void sort_nearly_sorted(Piece* seg, size_t m) {
    size_t budget = 64 * m + 1024;
    for (size_t i = 1; i < m; i++) {
        if (seg[i].first >= seg[i - 1].first) continue;
        const Piece v = seg[i];
        size_t j = i;
        while (j > 0 && seg[j - 1].first > v.first) {
            seg[j] = seg[j - 1];
            j--;
            if (--budget == 0) break;
        }
        seg[j] = v;
        if (budget == 0) {
            std::stable_sort(seg, seg + m, [](const Piece& a, const Piece& b) { return a.first < b.first; });
            return;
        }
    }
}
void sort_nearly_sorted(std::vector<Piece>& seg) { sort_nearly_sorted(seg.data(), seg.size()); }

// blocks of 64 slots from pieces in start order (same rules as gl_pack_segments8)
void encode_pieces(const Piece* seg, size_t m, P8Out& o) {
    o.anchors.clear(); o.ds.clear(); o.len.clear();
    o.anchors.reserve(m / 48 + 16);
    o.ds.resize((m / 48 + 16) * kP8); o.len.resize((m / 48 + 16) * kP8);
    int64_t nb = 0;
    int cnt = kP8;
    int64_t last = 0;
    uint8_t* ds = o.ds.data();
    uint8_t* ln = o.len.data();
    size_t cap = o.ds.size() / kP8;
    auto grow = [&]() {
        cap = cap * 2 + 16;
        o.ds.resize(cap * kP8); o.len.resize(cap * kP8);
        ds = o.ds.data(); ln = o.len.data();
    };
    auto close_block = [&]() {
        if (nb > 0 && cnt < kP8) {
            memset(ds + (nb - 1) * kP8 + cnt, 0, (size_t)(kP8 - cnt));
            memset(ln + (nb - 1) * kP8 + cnt, 0, (size_t)(kP8 - cnt));
        }
    };
    for (size_t i = 0; i < m; i++) {
        const int64_t s = seg[i].first;
        bool open = cnt == kP8;
        if (!open) {
            const int64_t gap = s - last;
            if (gap > 255) {
                const int64_t fillers = (gap - 1) / 255;
                if (fillers + 1 > kP8 - cnt) open = true;
                else {
                    uint8_t* d = ds + (nb - 1) * kP8 + cnt;
                    uint8_t* l = ln + (nb - 1) * kP8 + cnt;
                    for (int64_t f = 0; f < fillers; f++) { d[f] = 255; l[f] = 0; }
                    cnt += (int)fillers;
                    last += 255 * fillers;
                }
            }
        }
        if (open) {
            close_block();
            if ((size_t)nb >= cap) grow();
            o.anchors.push_back((int32_t)s);
            nb++;
            cnt = 0;
            last = s;
        }
        ds[(nb - 1) * kP8 + cnt] = (uint8_t)(s - last);
        ln[(nb - 1) * kP8 + cnt] = (uint8_t)seg[i].second;
        cnt++;
        last = s;
    }
    close_block();
    o.nb = nb;
}

struct MtScratch { std::vector<Piece> own, over, merged; P8Out out; };

}  // namespace

extern "C" int gl_pack_segments8_mt(const int32_t* start, const int32_t* end, int64_t n, int32_t threads, int32_t* anchors, uint8_t* dstart,
                                    uint8_t* len, int64_t cap_blocks, int64_t* n_blocks) {
    if (n < 0 || !n_blocks || (n > 0 && (!start || !end))) return GL_EINVAL;
    glhost::ThreadPool& pool = glhost::ThreadPool::global();
    int T = threads > 0 ? std::min<int>(threads, pool.size()) : pool.size();
    int64_t P = std::min<int64_t>((int64_t)T, n / 8192 + 1);
    if (P <= 1) return gl_pack_segments8(start, end, n, anchors, dstart, len, cap_blocks, n_blocks);
    std::vector<int64_t> cmin((size_t)P, INT64_MAX), x((size_t)P + 1);
    auto lo_of = [&](int64_t k) { return (int64_t)((__int128)n * k / P); };
    pool.run(P, [&](int64_t k, int) {
        int64_t m = INT64_MAX;
        const int64_t a = lo_of(k), b = lo_of(k + 1);
        for (int64_t i = a; i < b; i++) if (end[i] > start[i] && start[i] < m) m = start[i];
        cmin[(size_t)k] = m;
    }, T);
    x[(size_t)P] = INT64_MAX;
    for (int64_t k = P - 1; k >= 0; k--) x[(size_t)k] = std::min(cmin[(size_t)k], x[(size_t)k + 1]);
    x[0] = INT64_MIN;

    // scratch per task (a fast worker may run two tasks), kept between calls so the big lists are allocated and
    // page-faulted once; concurrent callers take turns (each call uses every pool thread anyway)
    static std::mutex scratch_mu;
    static std::vector<std::unique_ptr<MtScratch>> scratch;
    std::lock_guard<std::mutex> scratch_lk(scratch_mu);
    while ((int64_t)scratch.size() < P) scratch.emplace_back(new MtScratch());
    std::vector<MtScratch*> scr((size_t)P, nullptr);
    for (int64_t k = 0; k < P; k++) scr[(size_t)k] = scratch[(size_t)k].get();
    std::vector<std::atomic<int>> phase((size_t)P);          // 1: set-aside list published, 2: block count published
    std::vector<int64_t> off((size_t)P + 1, 0);
    for (auto& f : phase) f.store(0, std::memory_order_relaxed);
    const bool have_out = anchors && dstart && len;
    std::atomic<bool> fits(true);
    pool.run(P, [&](int64_t k, int) {
        MtScratch& S = *scr[(size_t)k];
        S.over.clear();
        const int64_t a = lo_of(k), b = lo_of(k + 1), hi = x[(size_t)k + 1];
        // one pass: split into pieces of <= 255, keep or set aside, and notice whether the kept ones are already sorted
        size_t own_cap = (size_t)(b - a) + 64, m_own = 0;
        if (S.own.size() < own_cap) S.own.resize(own_cap);
        Piece* own = S.own.data();
        bool own_sorted = true;
        int32_t prev = INT32_MIN;
        for (int64_t i = a; i < b; i++) {
            const int32_t s0 = start[i], e0 = end[i];
            const int64_t l0 = (int64_t)e0 - s0;
            if (l0 <= 0) continue;
            if (l0 <= 255 && s0 < hi) {                                  // the common case: one piece, kept
                if (m_own == own_cap) { S.own.resize(own_cap *= 2); own = S.own.data(); }
                own_sorted &= s0 >= prev;
                prev = s0;
                own[m_own++] = Piece(s0, (int32_t)l0);
                continue;
            }
            int64_t sx = s0;
            while (sx < e0) {
                const int64_t piece = e0 - sx > 255 ? 255 : e0 - sx;
                if (sx < hi) {
                    if (m_own == own_cap) { S.own.resize(own_cap *= 2); own = S.own.data(); }
                    own_sorted &= sx >= prev;
                    prev = (int32_t)sx;
                    own[m_own++] = Piece((int32_t)sx, (int32_t)piece);
                } else S.over.emplace_back((int32_t)sx, (int32_t)piece);
                sx += piece;
            }
        }
        sort_nearly_sorted(S.over);
        phase[(size_t)k].store(1, std::memory_order_release);
        if (!own_sorted) sort_nearly_sorted(own, m_own);
        // pieces the earlier chunks set aside that start in [x_k, x_{k+1})
        S.merged.clear();
        for (int64_t j = 0; j < k; j++) {
            while (phase[(size_t)j].load(std::memory_order_acquire) < 1) std::this_thread::yield();
            const std::vector<Piece>& ov = scr[(size_t)j]->over;
            auto lo = std::lower_bound(ov.begin(), ov.end(), x[(size_t)k], [](const Piece& p, int64_t v) { return (int64_t)p.first < v; });
            auto up = std::lower_bound(ov.begin(), ov.end(), hi, [](const Piece& p, int64_t v) { return (int64_t)p.first < v; });
            S.merged.insert(S.merged.end(), lo, up);
        }
        const Piece* src = own;
        size_t m = m_own;
        if (!S.merged.empty()) {
            std::stable_sort(S.merged.begin(), S.merged.end(), [](const Piece& p, const Piece& q) { return p.first < q.first; });
            const size_t mid = S.merged.size();
            S.merged.insert(S.merged.end(), own, own + m_own);
            std::inplace_merge(S.merged.begin(), S.merged.begin() + (long)mid, S.merged.end(), [](const Piece& p, const Piece& q) { return p.first < q.first; });
            src = S.merged.data();
            m = S.merged.size();
        }
        encode_pieces(src, m, S.out);
        if (k > 0) while (phase[(size_t)k - 1].load(std::memory_order_acquire) < 2) std::this_thread::yield();
        const int64_t o = off[(size_t)k];
        off[(size_t)k + 1] = o + S.out.nb;
        phase[(size_t)k].store(2, std::memory_order_release);
        if (have_out && o + S.out.nb <= cap_blocks) {
            memcpy(anchors + o, S.out.anchors.data(), (size_t)S.out.nb * 4);
            memcpy(dstart + o * kP8, S.out.ds.data(), (size_t)S.out.nb * kP8);
            memcpy(len + o * kP8, S.out.len.data(), (size_t)S.out.nb * kP8);
        } else if (S.out.nb > 0) fits.store(false);
    }, T);
    *n_blocks = off[(size_t)P];
    return (fits.load() && have_out) || off[(size_t)P] == 0 ? GL_OK : GL_ERANGE;
}
