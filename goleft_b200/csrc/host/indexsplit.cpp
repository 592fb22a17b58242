// indexsplit.cpp — host side of `goleft indexsplit` (indexsplit/indexsplit.go:37-194): from the cohort's per-tile data
// sums (gl_indexsplit_accumulate, float64 in units of 1e9 bytes) to N regions of about equal data.
// Sequential by nature (a running sum with data-dependent cuts) and tiny (<= 200k tiles), so it stays on the host.
//
// Third-party arithmetic: gonum v0.14.0 (go.mod:18, not vendored) stat.MeanStdDev and floats.Sum.  MeanStdDev is the
// corrected two-pass formula; floats.Sum is a SIMD kernel on amd64 whose summation order depends on the slice's
// alignment, so its last bits cannot be pinned: plain left-to-right sums are used here (PARITY UNPINNED; the printed
// %.2f and the cut positions only differ when a sum lands within an ulp of a threshold).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>
#include "../../../include/goleft_b200.h"

namespace {

const int64_t kTileWidth = 16384;                      // indexcov.TileWidth

struct Interval { int64_t s, e; };

// depth.Overlaps (depth/intervals.go:25-40): half-open overlap with any interval of the chromosome
struct Problems {
    std::vector<Interval> iv;                          // sorted by start
    std::vector<int64_t> max_end;                      // running maximum of the ends
    void finish() {
        std::sort(iv.begin(), iv.end(), [](const Interval& a, const Interval& b) { return a.s < b.s; });
        max_end.resize(iv.size());
        int64_t m = INT64_MIN;
        for (size_t i = 0; i < iv.size(); i++) { m = std::max(m, iv[i].e); max_end[i] = m; }
    }
    bool overlaps(int64_t a, int64_t b) const {        // some interval with e > a and s < b
        size_t k = std::lower_bound(iv.begin(), iv.end(), b, [](const Interval& x, int64_t v) { return x.s < v; }) - iv.begin();
        return k > 0 && max_end[k - 1] > a;
    }
};

void emit(std::string& o, const char* chrom, int64_t s, int64_t e, double sum, int splits) {
    char b[96];
    int n = snprintf(b, sizeof b, "\t%lld\t%lld\t%.2f\t%d\n", (long long)s, (long long)e, sum, splits);   // Chunk.String, indexsplit.go:77-79
    o.append(chrom);
    o.append(b, (size_t)n);
}

}  // namespace

extern "C" {

int gl_indexsplit_chunks(const double* tile_sum, const int64_t* out_ptr, int32_t R, const char* const* ref_names, const int64_t* ref_lens,
                         const int32_t* ref_ids, int32_t n_refs, int32_t N, const int32_t* prob_ref, const int64_t* prob_start,
                         const int64_t* prob_end, int64_t n_prob, char** text, int64_t* text_len) {
    if (R < 0 || n_refs < 0 || !out_ptr || !text || !text_len || (n_refs > 0 && (!ref_names || !ref_lens || !ref_ids)) ||
        (n_prob > 0 && (!prob_ref || !prob_start || !prob_end)) || n_prob < 0)
        return GL_EINVAL;
    // chop (indexsplit.go:37-48): values above mean + 3 sd become 8 * mean
    std::vector<std::vector<double>> sizes((size_t)R);
    std::vector<double> sums((size_t)R, 0.0), pct((size_t)R, 0.0);
    double tot = 0.0;
    for (int32_t r = 0; r < R; r++) {
        const int64_t n = out_ptr[r + 1] - out_ptr[r];
        if (n < 0 || (n > 0 && !tile_sum)) return GL_EINVAL;
        std::vector<double>& v = sizes[(size_t)r];
        v.assign(tile_sum + out_ptr[r], tile_sum + out_ptr[r] + n);
        double s = 0.0;
        for (double x : v) s += x;
        const double mean = s / (double)n;             // NaN for an empty chromosome, like gonum's
        double ss = 0.0, comp = 0.0;
        for (double x : v) { const double d = x - mean; ss += d * d; comp += d; }
        const double sd = sqrt((ss - comp * comp / (double)n) / (double)(n - 1));
        const double mx = mean + 3 * sd;
        for (double& x : v) if (x > mx) x = 8 * mean;
        double t = 0.0;
        for (double x : v) t += x;
        sums[(size_t)r] = t;
        tot += t;
    }
    for (int32_t r = 0; r < R; r++) pct[(size_t)r] = sums[(size_t)r] / tot;

    std::vector<Problems> probs((size_t)std::max(n_refs, 0));
    for (int64_t k = 0; k < n_prob; k++)
        if (prob_ref[k] >= 0 && prob_ref[k] < n_refs) probs[(size_t)prob_ref[k]].iv.push_back({prob_start[k], prob_end[k]});
    for (Problems& p : probs) p.finish();

    std::string out;
    for (int32_t q = 0; q < n_refs; q++) {             // indexsplit.go:119-188
        const char* name = ref_names[q];
        const int64_t rlen = ref_lens[q];
        const int32_t ri = ref_ids[q];
        if (ri < 0 || ri >= R || sizes[(size_t)ri].empty()) { emit(out, name, 0, rlen, 0.0, 0); continue; }
        const double share = pct[(size_t)ri] * (double)N;
        int n = (share == share && share < 2147483647.0) ? (int)share : 0;       // NaN when the cohort has no data at all
        if (n == 0 && pct[(size_t)ri] > 0) n = 1;
        else if (n == 0) { emit(out, name, 0, rlen, 0.0, 0); continue; }
        const double chunk = sums[(size_t)ri] / (double)n;
        const std::vector<double>& size = sizes[(size_t)ri];
        const int64_t len = (int64_t)size.size();
        const Problems& tree = probs[(size_t)q];
        double sum = 0.0;
        int64_t lasti = 0;
        for (int64_t i = 0; i < len; i++) {
            const bool ovl = !tree.iv.empty() && tree.overlaps(i * kTileWidth, (i + 1) * kTileWidth);
            if (size[(size_t)i] > chunk || (size[(size_t)i] >= 0.05 * chunk && ovl)) {      // a heavy tile is cut into pieces
                if (i > lasti) emit(out, name, lasti * kTileWidth, i * kTileWidth, sum, 1);
                sum = size[(size_t)i];
                int nsplits = (int)(0.5 + (sum / (chunk / 2)));
                if (nsplits > 8) nsplits = 8;
                else if (nsplits < 1) { nsplits = 1; if (ovl) nsplits = 3; }
                int64_t start = i * kTileWidth;
                const int64_t l = (int64_t)((double)kTileWidth / (double)nsplits + 1);
                for (int k = 0; k < nsplits; k++) {
                    if (i + k == len + 1) emit(out, name, start, rlen, sum / (double)nsplits, nsplits);
                    else emit(out, name, start, std::min(start + l, (i + 1) * kTileWidth), sum / (double)nsplits, nsplits);
                    start += l;
                }
                lasti = i + 1;
                sum = 0.0;
                continue;
            }
            sum += size[(size_t)i];
            if (sum >= chunk || i == len - 1 || (sum >= 0.2 * chunk && ovl)) {
                if (i == len - 1) emit(out, name, lasti * kTileWidth, rlen, sum, 1);
                else emit(out, name, lasti * kTileWidth, (i + 1) * kTileWidth, sum, 1);
                lasti = i + 1;
                sum = 0.0;
            }
        }
    }
    char* p = (char*)malloc(out.size() + 1);
    if (!p) return GL_ENOMEM;
    memcpy(p, out.data(), out.size());
    p[out.size()] = 0;
    *text = p;
    *text_len = (int64_t)out.size();
    return GL_OK;
}

}  // extern "C"
