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
