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
