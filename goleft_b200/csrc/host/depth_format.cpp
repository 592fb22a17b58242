// depth_format.cpp — host-side text for `goleft depth`: turns the device results of one chunk
// (window sums + class runs) into the rows the reference callback writes
// (depth/depth.go:293-305 window rows, :310-327,:343-350 callable rows, :329-358 tail).
//
// The reference walks a per-base text stream; here the same rows are derived from aggregates:
//   * callable rows are exactly the run-length encoding of class(depth) over [rs,re) — the
//     reference emits (run) and (gap) rows alternately and never merges them, and since
//     `samtools depth` prints no zero-depth lines every NO_COVERAGE row is a gap row;
//   * window rows for windows before the last covered base are the aligned windows; the window
//     holding the last covered base and the trailing zero windows follow depth.go:329-358
//     literally, including the misaligned end / re-emitted window when that base lies in a
//     chunk's leading partial window (bed mode only).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include "../../../include/goleft_b200.h"

namespace {

const char* kClassName[4] = {"NO_COVERAGE", "LOW_COVERAGE", "CALLABLE", "EXCESSIVE_COVERAGE"};

inline void put_i64(std::string& o, long long v) {
    char b[24];
    int n = 0;
    bool neg = v < 0;
    unsigned long long u = neg ? 0ull - (unsigned long long)v : (unsigned long long)v;
    do { b[n++] = char('0' + u % 10); u /= 10; } while (u);
    if (neg) o.push_back('-');
    while (n) o.push_back(b[--n]);
}

inline void window_row(std::string& o, const char* chrom, size_t cl, long long s, long long e, long long sum, const double* st = nullptr) {
    o.append(chrom, cl);
    o.push_back('\t');
    put_i64(o, s);
    o.push_back('\t');
    put_i64(o, e);
    o.push_back('\t');
    // depth.go:181-189: float64 sum of the ints divided by float64(e-s); 0 when the cache is empty
    double mean = (sum == 0 || e == s) ? 0.0 : (double)sum / (double)(e - s);
    char b[40];
    int n = snprintf(b, sizeof b, "%.4g", mean);   // Go's %.4g and C's agree (shortest of %e/%f at 4 digits)
    o.append(b, (size_t)n);
    if (st) {                                      // depth.go:199 "\t%.3g\t%.3g\t%.3g" of Stats{GC, CpG, Masked}
        n = snprintf(b, sizeof b, "\t%.3g\t%.3g\t%.3g", st[0], st[1], st[2]);
        o.append(b, (size_t)n);
    }
    o.push_back('\n');
}

// The window rows of one chunk in the order the reference writes them: f(s, e, index into win_sum or -1 for a
// trailing zero window).  lastcov = last covered base (-1: none), derived from the class runs.
template <class F>
inline void for_each_window_row(long long rs, long long re, long long W, const int32_t* run_start, const uint8_t* run_class,
                                int64_t n_runs, F f) {
    const long long w0 = rs / W;
    long long lastcov;   // the chunk ends covered unless its last run is NO_COVERAGE
    if (run_class[n_runs - 1] == GL_NO_COVERAGE) lastcov = (run_start[n_runs - 1] == rs) ? -1 : (long long)run_start[n_runs - 1] - 1;
    else lastcov = re - 1;
    long long pos = 0;   // the reference's `pos` after the per-line loop (0 when no line was read)
    if (lastcov >= 0) {
        const long long wl = lastcov / W;
        for (long long iw = w0; iw < wl; iw++) {                       // depth.go:296-303
            long long s = iw * W < rs ? rs : iw * W, e = (iw + 1) * W > re ? re : (iw + 1) * W;
            f(s, e, iw - w0);
        }
        long long s = wl * W;                                           // depth.go:330
        if (s < rs) s = rs;                                             // :332
        long long e = s + W > re ? re : s + W;                          // :333
        f(s, e, wl - w0);
        pos = e;                                                        // :338
    }
    if (lastcov + 1 < re) {                                             // depth.go:343 (cache[1].start+1 < regionEnd)
        long long ds = (rs > pos ? rs : pos) / W * W;                   // :351
        for (; ds < re && pos < re; ds += W) {
            long long de = ds + W > re ? re : ds + W;
            long long s = ds < rs ? rs : ds;
            f(s, de, -1);
        }
    }
}

inline void class_row(std::string& o, const char* chrom, size_t cl, long long s, long long e, int cls) {
    o.append(chrom, cl);
    o.push_back('\t');
    put_i64(o, s);
    o.push_back('\t');
    put_i64(o, e);
    o.push_back('\t');
    o.append(kClassName[cls & 3]);
    o.push_back('\n');
}

char* dup_out(const std::string& s) {
    char* p = (char*)malloc(s.size() + 1);
    if (!p) return nullptr;
    memcpy(p, s.data(), s.size());
    p[s.size()] = 0;
    return p;
}

}  // namespace

extern "C" {

int gl_depth_chunk_rows(int64_t rs, int64_t re, int32_t W, const int32_t* run_start, const uint8_t* run_class, int64_t n_runs,
                        int64_t* row_s, int64_t* row_e, int64_t cap, int64_t* n_rows) {
    if (W <= 0 || re <= rs || n_runs < 1 || !run_start || !run_class || run_start[0] != rs || !n_rows) return GL_EINVAL;
    int64_t n = 0;
    for_each_window_row(rs, re, W, run_start, run_class, n_runs, [&](long long s, long long e, long long) {
        if (n < cap && row_s && row_e) { row_s[n] = s; row_e[n] = e; }
        n++;
    });
    *n_rows = n;
    return n > cap ? GL_ERANGE : GL_OK;
}

int gl_depth_format_chunk_stats(const char* chrom, int64_t rs, int64_t re, int32_t W, const int64_t* win_sum,
                                int64_t n_windows, const int32_t* run_start, const uint8_t* run_class, int64_t n_runs,
                                const double* stats3, int64_t n_stat_rows,
                                char** depth_bed, int64_t* depth_len, char** callable_bed, int64_t* callable_len) {
    if (!chrom || W <= 0 || re <= rs || !depth_bed || !callable_bed || !depth_len || !callable_len) return GL_EINVAL;
    const long long w0 = rs / W, w1 = (re - 1) / W;
    if (n_windows != w1 - w0 + 1 || n_runs < 1 || run_start[0] != rs) return GL_EINVAL;
    const size_t cl = strlen(chrom);
    std::string hd, ca;
    hd.reserve((size_t)n_windows * (cl + (stats3 ? 64 : 28)));
    ca.reserve((size_t)n_runs * (cl + 36));

    // ---- callable rows: one per run
    for (int64_t i = 0; i < n_runs; i++) {
        long long s = run_start[i], e = (i + 1 < n_runs) ? run_start[i + 1] : re;
        class_row(ca, chrom, cl, s, e, run_class[i]);
    }
    // ---- window rows
    int64_t row = 0;
    bool short_stats = false;
    for_each_window_row(rs, re, W, run_start, run_class, n_runs, [&](long long s, long long e, long long wi) {
        const double* st = nullptr;
        if (stats3) {
            if (row < n_stat_rows) st = stats3 + row * 3;
            else short_stats = true;
        }
        window_row(hd, chrom, cl, s, e, wi >= 0 ? win_sum[wi] : 0, st);
        row++;
    });
    if (short_stats || (stats3 && row != n_stat_rows)) return GL_EINVAL;   // stats must match gl_depth_chunk_rows
    *depth_bed = dup_out(hd);
    *callable_bed = dup_out(ca);
    if (!*depth_bed || !*callable_bed) return GL_ENOMEM;
    *depth_len = (int64_t)hd.size();
    *callable_len = (int64_t)ca.size();
    return GL_OK;
}

int gl_depth_format_chunk(const char* chrom, int64_t rs, int64_t re, int32_t W, const int64_t* win_sum,
                          int64_t n_windows, const int32_t* run_start, const uint8_t* run_class, int64_t n_runs,
                          char** depth_bed, int64_t* depth_len, char** callable_bed, int64_t* callable_len) {
    return gl_depth_format_chunk_stats(chrom, rs, re, W, win_sum, n_windows, run_start, run_class, n_runs, nullptr, 0,
                                       depth_bed, depth_len, callable_bed, callable_len);
}

void gl_free_text(char* p) { free(p); }

}  // extern "C"
