// seg_pack.cpp — the feeder's compact segment format ("packed16"): half the PCIe bytes of (int32 start, int32 end).
//
// Segments are grouped in blocks of 256 slots.  Block b has one int32 anchor; slot k holds a segment
// [anchor + off, anchor + off + len) with off,len as uint16 (len == 0: empty slot).  BAM order keeps consecutive
// starts within a few hundred bases, so almost every block is full; a new block is opened when the next start
// does not fit 16 bits from the anchor (sparse regions), and segments longer than 65535 are split (harmless for depth).
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
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
// span) is cut into P index chunks; each worker turns its chunk into a start-sorted piece list (bounded insertion sort);
// assemble_p8 (seg_pack.h) then gives every worker a position range [x_k, x_{k+1}), x_k = the smallest start in chunks
// >= k: it takes its own pieces below x_{k+1} plus the few pieces earlier chunks hold beyond their range (second blocks
// of deletion / spliced reads at a chunk edge, tails of long segments), encodes its blocks privately, and the parts are
// concatenated.  Anchors come out globally sorted, which is what K_tileidx8 needs.
#include <atomic>
#include <limits.h>
#include <memory>
#include <mutex>
#include "seg_pack.h"
#include "thread_pool.h"

namespace glhost {

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

void split_long_pieces(std::vector<Piece>& list) {
    bool any = false;
    for (const Piece& p : list) if (p.second > 255) { any = true; break; }
    if (!any) return;
    std::vector<Piece> out;
    out.reserve(list.size() + list.size() / 4 + 16);
    for (const Piece& p : list) {
        int64_t s = p.first, left = p.second;
        while (left > 0) { const int64_t k = left > 255 ? 255 : left; out.emplace_back((int32_t)s, (int32_t)k); s += k; left -= k; }
    }
    sort_nearly_sorted(out.data(), out.size());
    list.swap(out);
}

// blocks of 64 slots from pieces in start order (same rules as gl_pack_segments8)
void encode_pieces(const Piece* seg, size_t m, P8Out& o) {
    o.anchors.clear();
    o.anchors.reserve(m / 48 + 16);
    size_t cap = m / 48 + 16;
    if (o.ds.size() < cap * kP8) { o.ds.resize(cap * kP8); o.len.resize(cap * kP8); }
    cap = o.ds.size() / kP8;
    int64_t nb = 0;
    int cnt = kP8;
    int64_t last = 0;
    uint8_t* ds = o.ds.data();
    uint8_t* ln = o.len.data();
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

namespace {

// x[k] = min start over lists >= k (x[0] = -inf, x[P] = +inf)
std::vector<int64_t> range_bounds(const std::vector<const std::vector<Piece>*>& lists) {
    const size_t P = lists.size();
    std::vector<int64_t> x(P + 1, INT64_MAX);
    for (size_t k = P; k-- > 0;) {
        x[k] = x[k + 1];
        if (!lists[k]->empty()) x[k] = std::min<int64_t>(x[k], lists[k]->front().first);
    }
    if (P) x[0] = INT64_MIN;
    return x;
}

// the pieces task k owns: list k below x[k+1] (a prefix) + what earlier lists hold in [x[k], x[k+1]).  Returns a pointer
// into list k when nothing has to be merged in, else fills `scratch`.
const Piece* gather_range(const std::vector<const std::vector<Piece>*>& lists, const std::vector<int64_t>& x, size_t k,
                          std::vector<Piece>& scratch, size_t* m) {
    auto cmp = [](const Piece& p, int64_t v) { return (int64_t)p.first < v; };
    const std::vector<Piece>& mine = *lists[k];
    const size_t own = (size_t)(std::lower_bound(mine.begin(), mine.end(), x[k + 1], cmp) - mine.begin());
    scratch.clear();
    for (size_t j = 0; j < k; j++) {
        const std::vector<Piece>& ov = *lists[j];
        if (ov.empty() || (int64_t)ov.back().first < x[k]) continue;
        auto lo = std::lower_bound(ov.begin(), ov.end(), x[k], cmp);
        auto up = std::lower_bound(lo, ov.end(), x[k + 1], cmp);
        scratch.insert(scratch.end(), lo, up);
    }
    if (scratch.empty()) { *m = own; return mine.data(); }
    auto less = [](const Piece& p, const Piece& q) { return p.first < q.first; };
    std::stable_sort(scratch.begin(), scratch.end(), less);
    const size_t mid = scratch.size();
    scratch.insert(scratch.end(), mine.begin(), mine.begin() + (long)own);
    std::inplace_merge(scratch.begin(), scratch.begin() + (long)mid, scratch.end(), less);
    *m = scratch.size();
    return scratch.data();
}

}  // namespace

void assemble_p8(const std::vector<const std::vector<Piece>*>& lists, int threads, Assembled& a) {
    const size_t P = lists.size();
    const std::vector<int64_t> x = range_bounds(lists);
    if (a.parts.size() < P) a.parts.resize(P);
    if (a.merged.size() < P) a.merged.resize(P);
    a.off.assign(P + 1, 0);
    ThreadPool::global().run((int64_t)P, [&](int64_t k, int) {
        size_t m = 0;
        const Piece* src = gather_range(lists, x, (size_t)k, a.merged[(size_t)k], &m);
        encode_pieces(src, m, a.parts[(size_t)k]);
    }, threads);
    for (size_t k = 0; k < P; k++) a.off[k + 1] = a.off[k] + a.parts[k].nb;
    a.total = a.off[P];
}

void assemble_i32(const std::vector<const std::vector<Piece>*>& lists, int threads, Assembled& a) {
    const size_t P = lists.size();
    const std::vector<int64_t> x = range_bounds(lists);
    if (a.merged.size() < P) a.merged.resize(P);
    a.off.assign(P + 1, 0);
    ThreadPool::global().run((int64_t)P, [&](int64_t k, int) {
        size_t m = 0;
        std::vector<Piece>& sc = a.merged[(size_t)k];
        const Piece* src = gather_range(lists, x, (size_t)k, sc, &m);
        if (src != sc.data()) sc.assign(src, src + m);           // keep a private copy: concat_i32 reads a.merged
    }, threads);
    for (size_t k = 0; k < P; k++) a.off[k + 1] = a.off[k] + (int64_t)a.merged[k].size();
    a.total = a.off[P];
}

void concat_p8(const Assembled& a, int threads, int32_t* anchors, uint8_t* dstart, uint8_t* len) {
    const size_t P = a.off.size() - 1;
    ThreadPool::global().run((int64_t)P, [&](int64_t k, int) {
        const P8Out& o = a.parts[(size_t)k];
        const int64_t at = a.off[(size_t)k];
        if (o.nb == 0) return;
        memcpy(anchors + at, o.anchors.data(), (size_t)o.nb * 4);
        memcpy(dstart + at * kP8, o.ds.data(), (size_t)o.nb * kP8);
        memcpy(len + at * kP8, o.len.data(), (size_t)o.nb * kP8);
    }, threads);
}

void concat_i32(const Assembled& a, int threads, int32_t* start, int32_t* end) {
    const size_t P = a.off.size() - 1;
    ThreadPool::global().run((int64_t)P, [&](int64_t k, int) {
        const std::vector<Piece>& v = a.merged[(size_t)k];
        int32_t* s = start + a.off[(size_t)k];
        int32_t* e = end + a.off[(size_t)k];
        for (size_t i = 0; i < v.size(); i++) { s[i] = v[i].first; e[i] = v[i].first + v[i].second; }
    }, threads);
}

}  // namespace glhost

extern "C" int gl_pack_segments8_mt(const int32_t* start, const int32_t* end, int64_t n, int32_t threads, int32_t* anchors, uint8_t* dstart,
                                    uint8_t* len, int64_t cap_blocks, int64_t* n_blocks) {
    using namespace glhost;
    if (n < 0 || !n_blocks || (n > 0 && (!start || !end))) return GL_EINVAL;
    ThreadPool& pool = ThreadPool::global();
    const int T = threads > 0 ? std::min<int>(threads, pool.size()) : pool.size();
    const int64_t P = std::min<int64_t>((int64_t)T, n / 8192 + 1);
    if (P <= 1) return gl_pack_segments8(start, end, n, anchors, dstart, len, cap_blocks, n_blocks);
    // scratch kept between calls so the big lists are allocated and page-faulted once; concurrent callers take turns
    // (each call uses every pool thread anyway)
    struct Scratch { std::vector<std::vector<Piece>> lists; Assembled a; };
    static std::mutex scratch_mu;
    static Scratch S;
    std::lock_guard<std::mutex> scratch_lk(scratch_mu);
    if ((int64_t)S.lists.size() < P) S.lists.resize((size_t)P);
    auto lo_of = [&](int64_t k) { return (int64_t)((__int128)n * k / P); };
    pool.run(P, [&](int64_t k, int) {
        std::vector<Piece>& L = S.lists[(size_t)k];
        const int64_t a = lo_of(k), b = lo_of(k + 1);
        size_t cap = (size_t)(b - a) + 64, m = 0;
        if (L.size() < cap) L.resize(cap);
        Piece* out = L.data();
        bool sorted = true;
        int32_t prev = INT32_MIN;
        for (int64_t i = a; i < b; i++) {
            const int32_t s0 = start[i], e0 = end[i];
            const int64_t l0 = (int64_t)e0 - s0;
            if (l0 <= 0) continue;
            int64_t sx = s0;
            do {                                                       // pieces of <= 255 (one for a short read's block)
                const int64_t piece = e0 - sx > 255 ? 255 : e0 - sx;
                if (m == cap) { L.resize(cap *= 2); out = L.data(); }
                sorted &= sx >= prev;
                prev = (int32_t)sx;
                out[m++] = Piece((int32_t)sx, (int32_t)piece);
                sx += piece;
            } while (sx < e0);
        }
        L.resize(m);
        if (!sorted) sort_nearly_sorted(L.data(), m);
    }, T);
    std::vector<const std::vector<Piece>*> lists((size_t)P);
    for (int64_t k = 0; k < P; k++) lists[(size_t)k] = &S.lists[(size_t)k];
    assemble_p8(lists, T, S.a);
    *n_blocks = S.a.total;
    if (S.a.total == 0) return GL_OK;
    if (!(anchors && dstart && len) || S.a.total > cap_blocks) return GL_ERANGE;
    concat_p8(S.a, T, anchors, dstart, len);
    return GL_OK;
}

// ------------------------------------------------------------------------------------------------ parallel packed16
// gl_pack_segments16_mt: packed16 keeps the caller's order and needs no sort, so the input is simply cut into index
// chunks; every worker first counts the blocks its chunk needs (a dry run of the same loop), the counts are prefix-summed,
// and a second pass writes the blocks in place.  Each chunk opens its own first block, so the result can differ from the
// single-threaded packer's by a few partly filled blocks; it decodes to the same segments in the same order.
namespace {
// one chunk: returns the number of blocks; writes when anchors != null (to block index b0 ...)
int64_t pack16_chunk(const int32_t* start, const int32_t* end, int64_t a, int64_t b, int32_t* anchors, uint16_t* off, uint16_t* len, int64_t b0) {
    int64_t nb = 0;
    int cnt = 256;
    int32_t anchor = 0;
    const bool wr = anchors != nullptr;
    for (int64_t i = a; i < b; i++) {
        int64_t s = start[i];
        const int64_t e = end[i];
        if (e <= s) continue;
        while (s < e) {
            const int64_t piece = (e - s > 65535) ? 65535 : e - s;
            const int64_t o = s - anchor;
            if (cnt == 256 || o < 0 || o > 65535) {
                if (wr && nb > 0 && cnt < 256) {
                    memset(off + (b0 + nb - 1) * 256 + cnt, 0, (size_t)(256 - cnt) * 2);
                    memset(len + (b0 + nb - 1) * 256 + cnt, 0, (size_t)(256 - cnt) * 2);
                }
                anchor = s > INT32_MIN + 16384 ? (int32_t)(s - 16384) : (int32_t)s;
                if (wr) anchors[b0 + nb] = anchor;
                nb++;
                cnt = 0;
            }
            if (wr) {
                off[(b0 + nb - 1) * 256 + cnt] = (uint16_t)(s - anchor);
                len[(b0 + nb - 1) * 256 + cnt] = (uint16_t)piece;
            }
            cnt++;
            s += piece;
        }
    }
    if (wr && nb > 0 && cnt < 256) {
        memset(off + (b0 + nb - 1) * 256 + cnt, 0, (size_t)(256 - cnt) * 2);
        memset(len + (b0 + nb - 1) * 256 + cnt, 0, (size_t)(256 - cnt) * 2);
    }
    return nb;
}
}  // namespace

extern "C" int gl_pack_segments16_mt(const int32_t* start, const int32_t* end, int64_t n, int32_t threads, int32_t* anchors, uint16_t* off,
                                     uint16_t* len, int64_t cap_blocks, int64_t* n_blocks) {
    using namespace glhost;
    if (n < 0 || !n_blocks || (n > 0 && (!start || !end))) return GL_EINVAL;
    ThreadPool& pool = ThreadPool::global();
    const int T = threads > 0 ? std::min<int>(threads, pool.size()) : pool.size();
    const int64_t P = std::min<int64_t>((int64_t)T, n / 16384 + 1);
    std::vector<int64_t> cnt((size_t)P + 1, 0);
    auto lo_of = [&](int64_t k) { return (int64_t)((__int128)n * k / P); };
    pool.run(P, [&](int64_t k, int) { cnt[(size_t)k + 1] = pack16_chunk(start, end, lo_of(k), lo_of(k + 1), nullptr, nullptr, nullptr, 0); }, T);
    for (int64_t k = 0; k < P; k++) cnt[(size_t)k + 1] += cnt[(size_t)k];
    *n_blocks = cnt[(size_t)P];
    if (cnt[(size_t)P] == 0) return GL_OK;
    if (!(anchors && off && len) || cnt[(size_t)P] > cap_blocks) return GL_ERANGE;
    pool.run(P, [&](int64_t k, int) { pack16_chunk(start, end, lo_of(k), lo_of(k + 1), anchors, off, len, cnt[(size_t)k]); }, T);
    return GL_OK;
}

// ------------------------------------------------------------------------------------------------ fixed-block packed16
// gl_pack_segments16_fixed*: the same device format (256 slots per block, int32 anchor, uint16 off/len) with the block
// boundaries FIXED: block b holds input segments [256 b, 256 b + 256) in the caller's order, anchor = their lowest start.
// No count pass, no prefix sum, no data-dependent block edges: one streaming pass (AVX2 when the CPU has it: min/max
// reduction + subtract + pack over 2 KB that stay in L1), every block independent.  A block whose starts span more than
// 65535 bases or that holds a segment longer than 65535 (a block straddling a centromere gap; long reads) is written as
// 256 empty slots and its segments are returned raw in the escape list (a few blocks per contig for short-read data).
// This is what the one-call int32 entry uses to halve its PCIe bytes: the pack runs at memory speed on the host pool,
// chunk c+1 is packed while chunk c is on the wire.
namespace {

inline bool block_scalar(const int32_t* s, const int32_t* e, int cnt, int32_t* anchor, uint16_t* off, uint16_t* len) {
    int32_t smin = s[0], smax = s[0];
    uint32_t lmax = 0;
    for (int i = 0; i < cnt; i++) {
        smin = s[i] < smin ? s[i] : smin;
        smax = s[i] > smax ? s[i] : smax;
        const uint32_t l = e[i] > s[i] ? (uint32_t)e[i] - (uint32_t)s[i] : 0u;
        lmax = l > lmax ? l : lmax;
    }
    *anchor = smin;
    if ((int64_t)smax - smin > 65535 || lmax > 65535u) return false;
    for (int i = 0; i < cnt; i++) {
        off[i] = (uint16_t)(s[i] - smin);
        len[i] = (uint16_t)(e[i] > s[i] ? e[i] - s[i] : 0);
    }
    for (int i = cnt; i < 256; i++) { off[i] = 0; len[i] = 0; }
    return true;
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) inline bool block256_avx2(const int32_t* s, const int32_t* e, int32_t* anchor, uint16_t* off, uint16_t* len) {
    __m256i vmin = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s)), vmax = vmin, lmax = _mm256_setzero_si256();
    for (int i = 0; i < 256; i += 8) {
        const __m256i S = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i));
        const __m256i E = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(e + i));
        vmin = _mm256_min_epi32(vmin, S);
        vmax = _mm256_max_epi32(vmax, S);
        lmax = _mm256_max_epu32(lmax, _mm256_and_si256(_mm256_sub_epi32(E, S), _mm256_cmpgt_epi32(E, S)));
    }
    alignas(32) int32_t a[8], b[8];
    alignas(32) uint32_t c[8];
    _mm256_store_si256(reinterpret_cast<__m256i*>(a), vmin);
    _mm256_store_si256(reinterpret_cast<__m256i*>(b), vmax);
    _mm256_store_si256(reinterpret_cast<__m256i*>(c), lmax);
    int32_t smin = a[0], smax = b[0];
    uint32_t lm = c[0];
    for (int k = 1; k < 8; k++) { smin = a[k] < smin ? a[k] : smin; smax = b[k] > smax ? b[k] : smax; lm = c[k] > lm ? c[k] : lm; }
    *anchor = smin;
    if ((int64_t)smax - smin > 65535 || lm > 65535u) return false;
    const __m256i A = _mm256_set1_epi32(smin);
    static const bool nt = [] { const char* e = getenv("GL_PACK_NT"); return !e || atoi(e) != 0; }();     // non-temporal stores (default on)
    const bool aligned = nt && ((reinterpret_cast<uintptr_t>(off) | reinterpret_cast<uintptr_t>(len)) & 31) == 0;
    for (int i = 0; i < 256; i += 16) {
        const __m256i S0 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i)), S1 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 8));
        const __m256i E0 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(e + i)), E1 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(e + i + 8));
        const __m256i O = _mm256_permute4x64_epi64(_mm256_packus_epi32(_mm256_sub_epi32(S0, A), _mm256_sub_epi32(S1, A)), 0xD8);
        const __m256i L0 = _mm256_and_si256(_mm256_sub_epi32(E0, S0), _mm256_cmpgt_epi32(E0, S0));
        const __m256i L1 = _mm256_and_si256(_mm256_sub_epi32(E1, S1), _mm256_cmpgt_epi32(E1, S1));
        const __m256i L = _mm256_permute4x64_epi64(_mm256_packus_epi32(L0, L1), 0xD8);
        if (aligned) {                                  // the output is written once and read by the DMA engine: keep it out of the caches
            _mm256_stream_si256(reinterpret_cast<__m256i*>(off + i), O);
            _mm256_stream_si256(reinterpret_cast<__m256i*>(len + i), L);
        } else {
            _mm256_storeu_si256(reinterpret_cast<__m256i*>(off + i), O);
            _mm256_storeu_si256(reinterpret_cast<__m256i*>(len + i), L);
        }
    }
    return true;
}
static const bool kHaveAvx2 = __builtin_cpu_supports("avx2");
#else
static const bool kHaveAvx2 = false;
#endif

// blocks [b0, b1) of the input; anchors/off/len are indexed by absolute block; escapes appended to es/ee
void pack16_fixed_range(const int32_t* start, const int32_t* end, int64_t n, int64_t b0, int64_t b1, int32_t* anchors, uint16_t* off,
                        uint16_t* len, std::vector<int32_t>& es, std::vector<int32_t>& ee) {
    for (int64_t b = b0; b < b1; b++) {
        const int64_t i0 = b * 256;
        const int cnt = (int)std::min<int64_t>(256, n - i0);
        bool ok;
#if defined(__x86_64__)
        if (cnt == 256 && kHaveAvx2) ok = block256_avx2(start + i0, end + i0, anchors + b, off + i0, len + i0);
        else
#endif
            ok = block_scalar(start + i0, end + i0, cnt, anchors + b, off + i0, len + i0);
        if (!ok) {
            memset(off + i0, 0, 512);
            memset(len + i0, 0, 512);
            for (int i = 0; i < cnt; i++)
                if (end[i0 + i] > start[i0 + i]) { es.push_back(start[i0 + i]); ee.push_back(end[i0 + i]); }
        }
    }
#if defined(__x86_64__)
    if (kHaveAvx2) _mm_sfence();
#endif
}

}  // namespace

extern "C" int gl_pack_segments16_fixed_range_mt(const int32_t* start, const int32_t* end, int64_t n, int64_t block_begin, int64_t block_end,
                                                 int32_t threads, int32_t* anchors, uint16_t* off, uint16_t* len, int32_t* esc_start,
                                                 int32_t* esc_end, int64_t esc_cap, int64_t* n_esc) {
    using namespace glhost;
    const int64_t nbk = (n + 255) / 256;
    if (n < 0 || block_begin < 0 || block_end < block_begin || block_end > nbk || !n_esc || *n_esc < 0) return GL_EINVAL;
    if (block_end > block_begin && (!start || !end || !anchors || !off || !len)) return GL_EINVAL;
    if (block_end == block_begin) return GL_OK;
    ThreadPool& pool = ThreadPool::global();
    const int T = threads > 0 ? std::min<int>(threads, pool.size()) : pool.size();
    const int64_t nb = block_end - block_begin;
    // tasks: GL_PACK_TASKS_PER_THREAD (default 2) per thread, at least GL_PACK_MIN_BLOCKS (default 64 = 16 K segments) each
    static const int tpt = [] { const char* e = getenv("GL_PACK_TASKS_PER_THREAD"); const int v = e ? atoi(e) : 2; return v < 1 ? 1 : v; }();
    static const int minb = [] { const char* e = getenv("GL_PACK_MIN_BLOCKS"); const int v = e ? atoi(e) : 64; return v < 1 ? 1 : v; }();
    const int64_t P = std::max<int64_t>(1, std::min<int64_t>((int64_t)T * tpt, nb / minb + 1));
    std::vector<std::vector<int32_t>> es((size_t)P), ee((size_t)P);
    auto lo_of = [&](int64_t k) { return block_begin + (int64_t)((__int128)nb * k / P); };
    pool.run(P, [&](int64_t k, int) { pack16_fixed_range(start, end, n, lo_of(k), lo_of(k + 1), anchors, off, len, es[(size_t)k], ee[(size_t)k]); }, T);
    int64_t m = *n_esc;
    for (int64_t k = 0; k < P; k++) {
        const int64_t c = (int64_t)es[(size_t)k].size();
        if (c == 0) continue;
        if (m + c > esc_cap || !esc_start || !esc_end) { *n_esc = m + c; return GL_ERANGE; }
        memcpy(esc_start + m, es[(size_t)k].data(), (size_t)c * 4);
        memcpy(esc_end + m, ee[(size_t)k].data(), (size_t)c * 4);
        m += c;
    }
    *n_esc = m;
    return GL_OK;
}

extern "C" int gl_pack_segments16_fixed_mt(const int32_t* start, const int32_t* end, int64_t n, int32_t threads, int32_t* anchors, uint16_t* off,
                                           uint16_t* len, int32_t* esc_start, int32_t* esc_end, int64_t esc_cap, int64_t* n_esc) {
    if (!n_esc) return GL_EINVAL;
    *n_esc = 0;
    return gl_pack_segments16_fixed_range_mt(start, end, n, 0, (n + 255) / 256, threads, anchors, off, len, esc_start, esc_end, esc_cap, n_esc);
}
