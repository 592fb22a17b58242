/*
 * oracle_indexcov.c — CPU restatement of goleft's indexcov / covstats / depthwed arithmetic.
 * TEST INFRASTRUCTURE ONLY (see oracle_depth.c): never linked into libgoleft_b200.so or the CLI.
 *
 * PARITY UNPINNED for the same reason as oracle_depth.c: no `go` here, BAI decoding lives in the
 * un-vendored github.com/biogo/hts v1.4.4, and the reference's tests hold no expected numbers for
 * this path that are checkable offline.  The one in-tree fixture
 * (indexcov/test-data/sample_issue_27_0001.bam.bai) is pinned to values derived from this restatement.
 *
 * Floating point: Go on amd64 (GOAMD64=v1) does not fuse a*b+c; compile with -ffp-contract=off.
 * Citations are relative to the reference checkout (brentp/goleft @ v0.2.6).
 */
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- I1: indexcov/types.go:45-82 (getSizes)
 * voff: a sample's linear-index virtual offsets, all refs concatenated; ref_ptr[n_refs+1] CSR.
 * Per ref with >= 2 intervals: n_intv-1 deltas vOffset(iv[k+1]) - vOffset(iv[k]) (indexcov.go:78-80:
 * File<<16 | Block is exactly the raw BAI u64).  Returns total tiles, or -1 on a negative delta
 * (the reference panics, types.go:75-77). */
int64_t orc_ic_sizes(const uint64_t* voff, const int64_t* ref_ptr, int32_t n_refs, int64_t* sizes, int64_t* size_ptr) {
    int64_t k = 0;
    size_ptr[0] = 0;
    for (int32_t r = 0; r < n_refs; r++) {
        int64_t a = ref_ptr[r], b = ref_ptr[r + 1];
        if (b - a >= 2) {                                  /* types.go:68-71 */
            for (int64_t i = a; i + 1 < b; i++) {
                int64_t d = (int64_t)voff[i + 1] - (int64_t)voff[i];
                if (d < 0) return -1;
                sizes[k++] = d;
            }
        }
        size_ptr[r + 1] = k;
    }
    return k;
}

static int cmp_i64(const void* a, const void* b) {
    int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
    return (x > y) - (x < y);
}

/* ---------------------------------------------------------------- I2: indexcov.go:83-125 (Index.init)
 * sort ascending; n98 = sorted[int(0.98*n)]; cumsum of min(s,n98); idx = first i with cumsum[i] > total/2
 * (integer division); clamp; median = sorted[idx].  n >= 1. */
int64_t orc_ic_median(const int64_t* sizes, int64_t n) {
    int64_t* s = (int64_t*)malloc((size_t)n * sizeof(int64_t));
    memcpy(s, sizes, (size_t)n * sizeof(int64_t));
    qsort(s, (size_t)n, sizeof(int64_t), cmp_i64);          /* :104 */
    int64_t n98 = s[(int64_t)(0.98 * (double)n)];           /* :111 */
    int64_t total = 0;
    int64_t* cum = (int64_t*)malloc((size_t)n * sizeof(int64_t));
    for (int64_t i = 0; i < n; i++) {
        int64_t v = s[i] > n98 ? n98 : s[i];
        total += v;
        cum[i] = total;
    }
    int64_t idx = n;                                        /* sort.Search: smallest i with cum[i] > total/2 */
    int64_t half = total / 2;
    for (int64_t i = 0; i < n; i++) if (cum[i] > half) { idx = i; break; }
    while (idx >= n) idx--;                                 /* :120-122 */
    int64_t med = s[idx];
    free(s); free(cum);
    return med;
}

/* ---------------------------------------------------------------- I3: indexcov.go:129-151 (NormalizedDepth) */
void orc_ic_normalize(const int64_t* sizes, int64_t n, double median, float* out) {
    for (int64_t i = 0; i < n; i++) {
        float d = (float)((double)sizes[i] / median);       /* float32(float64(o)/median) */
        if (d > 50000) d = 50000;
        out[i] = d;
    }
}

/* ---------------------------------------------------------------- I4: indexcov.go:153-177 (CountsAtDepth), :181-193 (CountsROC)
 * slots*float32(slotsMid) is a typed float32 constant: 70 * float32(2/3) rounded to float32 = 0x423AAAAB. */
static int tint(float f) {
    int v = (int)f;                                         /* truncation toward zero */
    if (v < 70) return v < 0 ? 0 : v;
    return 69;
}
void orc_ic_counts(const float* depths, int64_t n, int32_t* counts70) {
    const float K = 46.66666793823242f;
    for (int64_t i = 0; i < n; i++) {
        float v = depths[i] * K;
        v = v + 0.5f;
        counts70[tint(v)]++;
    }
}
void orc_ic_roc(const int32_t* counts70, float* roc70) {
    int32_t totals[70];
    totals[69] = counts70[69];
    for (int i = 68; i >= 0; i--) totals[i] = totals[i + 1] + counts70[i];
    float mx = (float)totals[0];
    for (int i = 0; i < 70; i++) roc70[i] = (float)totals[i] / mx;
}

/* ---------------------------------------------------------------- I5: indexcov.go:1050-1078 (counter.count) after the
 * MaxCN clip of :694-697.  out4 += {out, low, hi, in}. */
void orc_ic_bins(const float* depths, int64_t n, int64_t longest, int64_t* out4) {
    int64_t i = 0;
    for (; i < n; i++) {
        float d = depths[i];
        if (d > 8.0f) d = 8.0f;                            /* MaxCN = float32(8) */
        if (d < 0.85f || d > 1.15f) {
            out4[0]++;
            if (d > 1.15f) out4[2]++;
            else if (d < 0.15f) out4[1]++;
        } else {
            out4[3]++;
        }
    }
    out4[0] += longest - i;
    out4[1] += longest - i;
}

/* ---------------------------------------------------------------- I7: indexcov.go:549-597 (normalizeAcrossSamples)
 * depths: S rows of stride T; lens[i] valid entries in row i.  In place. */
void orc_ic_xnorm(float* depths, const int32_t* lens, int32_t S, int32_t T) {
    if (S < 5) return;                                      /* :551 */
    int32_t maxLen = 0;
    for (int32_t i = 0; i < S; i++) if (lens[i] > maxLen) maxLen = lens[i];
    for (int32_t j = 0; j < maxLen; j++) {
        double m = 0, n = 0;
        for (int32_t i = 0; i < S; i++) {
            float* d = depths + (size_t)i * T;
            if (lens[i] > j) {
                m += (double)d[j]; n++;
                if (j > 0) { m += (double)d[j - 1]; n++; }
                if (j < lens[i] - 1) { m += (double)d[j + 1]; n++; }
            }
        }
        if ((int)n < 3 * S - 4) continue;                   /* :577 */
        m /= n;
        if (m < 0.1) continue;
        float m32 = (float)m;
        for (int32_t i = 0; i < S; i++) {
            float* d = depths + (size_t)i * T;
            if (lens[i] > j) {
                d[j] /= m32;
                if (j > 2 && j < lens[i] - 3) {
                    float a = d[j - 3] + d[j - 2];
                    a = a + d[j - 1];
                    a = a + d[j];
                    a = a + d[j + 1] / m32;
                    a = a + d[j + 2] / m32;
                    a = a + d[j + 3] / m32;
                    d[j] = 0.14285714924335479736328125f * a;   /* float32(1.0/7.0) */
                }
            }
        }
    }
}

static int cmp_f32(const void* a, const void* b) {
    float x = *(const float*)a, y = *(const float*)b;
    return (x > y) - (x < y);
}

/* ---------------------------------------------------------------- I8: indexcov.go:957-991 (GetCN), one sample */
double orc_ic_getcn(const float* d, int64_t n) {
    float* tmp = (float*)malloc((size_t)(n > 0 ? n : 1) * sizeof(float));
    int64_t k = 0, lows = 0;
    for (int64_t i = 0; i < n; i++) {
        if (d[i] != 0) { tmp[k++] = d[i]; if (d[i] < 0.02f) lows++; }
    }
    double med = -0.1;
    if (k > 0) {
        qsort(tmp, (size_t)k, sizeof(float), cmp_f32);
        double pLo = (double)lows / (double)n;
        float* t = tmp; int64_t tk = k;
        if (pLo > 0.3) { t = tmp + lows; tk = k - lows; }
        med = 0;
        if (tk > 0) med = (double)(2.0f * t[(int64_t)((double)tk * 0.4)]);
    }
    free(tmp);
    return med;
}

/* ---------------------------------------------------------------- V2: covstats/covstats.go:57-89,175-218
 * Order statistics and moments of a sample of ints (already collected by the BAM walk of :138-173). */
static int cmp_i32(const void* a, const void* b) {
    int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
    return (x > y) - (x < y);
}
/* meanStd :78-89: mean += a/l sequentially, then std += (a-mean)^2/l, sqrt */
void orc_cs_mean_std(const int32_t* arr, int64_t n, double* mean, double* sd) {
    double l = (double)n, m = 0, s = 0;
    for (int64_t i = 0; i < n; i++) m += (double)arr[i] / l;
    for (int64_t i = 0; i < n; i++) s += pow((double)arr[i] - m, 2) / l;
    *mean = m; *sd = sqrt(s);
}
/* madFilter :57-76 on a SORTED array; returns the kept prefix length */
int64_t orc_cs_mad_filter(const int32_t* sorted, int64_t n, int nmads) {
    int32_t med = sorted[n / 2];
    int64_t nu = n - (n / 2 + 1);
    int32_t* um = (int32_t*)malloc((size_t)(nu > 0 ? nu : 1) * sizeof(int32_t));
    for (int64_t i = 0; i < nu; i++) um[i] = sorted[n / 2 + 1 + i] - med;
    qsort(um, (size_t)nu, sizeof(int32_t), cmp_i32);
    int32_t umad = um[nu / 2];
    free(um);
    int64_t upper = (int64_t)med + (int64_t)nmads * umad;
    int64_t i = 0;
    for (i = 0; i < n; i++) if (sorted[i] > upper) break;
    if (i == n) i = n - 1;          /* Go's `for i, a = range arr` leaves i at the last index when nothing breaks */
    return i;
}
/* the tail of BamStats :175-218 for insert sizes / template lengths.
 * out: [0]=pct5 [1]=pct95 [2]=insert mean [3]=insert sd [4]=template mean [5]=template sd; H (len returned). */
int64_t orc_cs_tail(int32_t* insert, int64_t n_ins, int32_t* tmpl, int64_t n_tmpl, int32_t max_read_len,
                    double* out6, double* H, int64_t H_cap) {
    qsort(insert, (size_t)n_ins, sizeof(int32_t), cmp_i32);
    double l = (double)(n_ins - 1);
    out6[0] = insert[(int64_t)(0.05 * l + 0.5)];
    out6[1] = insert[(int64_t)(0.95 * l + 0.5)];
    int64_t k = orc_cs_mad_filter(insert, n_ins, 10);
    orc_cs_mean_std(insert, k, &out6[2], &out6[3]);
    qsort(tmpl, (size_t)n_tmpl, sizeof(int32_t), cmp_i32);  /* madFilter sorts when unsorted (:58-60) */
    int64_t kt = orc_cs_mad_filter(tmpl, n_tmpl, 10);
    orc_cs_mean_std(tmpl, kt, &out6[4], &out6[5]);
    double start = (double)max_read_len, stop = out6[4] + out6[5] * 4;
    int64_t hn = (int64_t)(stop - start + 1);
    if (hn < 0) hn = 0;
    for (int64_t i = 0; i < hn && i < H_cap; i++) H[i] = 0;
    double cnt = 0;
    for (int64_t i = 0; i < kt; i++) {
        double x = (double)tmpl[i];
        if (x < start || x > stop) continue;
        int64_t j = (int64_t)(x - start);
        if (j < H_cap) H[j] += 1;
        cnt += 1;
    }
    for (int64_t i = 0; i < hn && i < H_cap; i++) H[i] /= cnt;
    return hn;
}
/* plain histogram of values in [lo,hi) */
void orc_bincount(const int32_t* v, int64_t n, int32_t lo, int32_t hi, uint64_t* hist) {
    memset(hist, 0, (size_t)(hi - lo) * sizeof(uint64_t));
    for (int64_t i = 0; i < n; i++) if (v[i] >= lo && v[i] < hi) hist[v[i] - lo]++;
}

/* ---------------------------------------------------------------- W1: depthwed/depthwed.go:93-157
 * means: S x R (row-major by sample) 4th-column values; starts/ends/chrom_id of the rows (file 0's).
 * depth = int(0.5+mean) (:103); rows of one chrom are merged until end-start >= size (:126).
 * out: n_out x S.  Returns n_out. */
int64_t orc_depthwed(const double* means, int32_t S, int64_t R, const int32_t* starts, const int32_t* ends,
                     const int32_t* chrom_id, int64_t size, int32_t* out_start, int32_t* out_end, int32_t* out_chrom,
                     int64_t* out, int64_t out_cap) {
    int64_t n_out = 0, r = 0;
    while (r < R) {
        int32_t chrom = chrom_id[r];
        int32_t st = starts[r], en = ends[r];
        int64_t r0 = r;
        r++;
        /* loop condition :126: keep reading while width < size and the next line is the same chrom */
        while (r < R && (int64_t)en - st < size && chrom_id[r] == chrom) { en = ends[r]; r++; }
        if (n_out < out_cap) {
            out_start[n_out] = st; out_end[n_out] = en; out_chrom[n_out] = chrom;
            for (int32_t s = 0; s < S; s++) {
                int64_t acc = 0;
                for (int64_t k = r0; k < r; k++) acc += (int64_t)(0.5 + means[(size_t)s * R + k]);   /* int(0.5+dep), :103,:151 */
                out[(size_t)n_out * S + s] = acc;
            }
        }
        n_out++;
    }
    return n_out;
}

/* ---------------------------------------------------------------- N4: indexcov/crai/crai.go:56-127 (makeSizes)
 * One reference's CRAM slices (alignment start, span, slice bytes) -> 16 KB pseudo-tile sizes.
 * Returns the number of tiles, -1 where the reference panics ("tilewidth logic error" / "logic error"). */
int64_t orc_crai_sizes(const int64_t* aln_start, const int64_t* aln_span, const int32_t* slice_len, int64_t n,
                       int64_t* sizes, int64_t cap) {
    const int64_t TW = 16384;
    int64_t k_out = 0, lastStart = 0, lastVal = 0;
    if (n == 0) return 0;
    for (int64_t s = 0; s < n; s++) {
        int64_t start = aln_start[s], span = aln_span[s];
        int k = 0;
        for (; lastStart < start - TW; lastStart += TW) {           /* :78-86 back fill gaps */
            if (k_out < cap) sizes[k_out] = (k == 0) ? lastVal : 0;
            k_out++;
            if (k == 0) lastVal = 0;
            k++;
        }
        int64_t overhang = start - lastStart;
        if (overhang > TW) return -1;                               /* :88-90 */
        while (overhang < -TW) {                                    /* :91-99 */
            start += TW; span -= TW;
            overhang = start - lastStart;
        }
        if (span <= 0) continue;                                    /* :100-104 */
        int64_t perBase = (int64_t)(100000 * (double)slice_len[s] / (double)span);   /* :106 */
        int64_t nTiles = (int64_t)((double)span / (double)TW);      /* :108 */
        if (nTiles == 0 && start - lastStart < TW) { lastVal = perBase; continue; }
        for (int64_t i = 0; i < nTiles; i++) { if (k_out < cap) sizes[k_out] = perBase; k_out++; }
        int64_t cmp = (start + span) / TW;
        if (k_out > cmp + 1 || cmp < k_out - 1) return -1;          /* :118-121 */
        lastStart += TW * nTiles;
        lastVal = perBase;
    }
    return k_out;
}

/* ------------------------------------------------------------------ indexsplit (indexsplit/indexsplit.go)
 * orc_indexsplit restates Split (:82-194) with getPercents/chop (:37-65) and Chunk.String (:77-79) for references whose
 * ID is their position (RefsFromBam).  sizes/ptr: as gl_indexsplit_accumulate (ptr[s*(R+1)+r]).  Problematic intervals
 * (-p): (reference number, start, end), half-open overlap like depth.Overlaps (depth/intervals.go:16-40).
 * Third-party: gonum v0.14.0 stat.MeanStdDev (corrected two-pass) and floats.Sum (amd64 SIMD kernel, alignment-dependent
 * order — restated as left-to-right; PARITY UNPINNED), biogo/store interval tree (overlap query only).
 * Returns the number of bytes written to out (0-terminated), or -1 if cap is too small. */
static int os_overlaps(const int32_t* pr, const int64_t* ps, const int64_t* pe, int64_t np, int ref, int64_t a, int64_t b) {
    for (int64_t k = 0; k < np; k++)
        if (pr[k] == ref && pe[k] > a && ps[k] < b) return 1;
    return 0;
}

int64_t orc_indexsplit(const int64_t* sizes_in, const int64_t* ptr, int S, int R, const char** names, const int64_t* ref_lens, int N,
                       const int32_t* pr, const int64_t* ps, const int64_t* pe, int64_t np, char* out, int64_t cap) {
    const int64_t TW = 16384;
    const double scalar = 1000000000.0;
    double** sizes = (double**)calloc((size_t)(R > 0 ? R : 1), sizeof(double*));
    int64_t* lens = (int64_t*)calloc((size_t)(R > 0 ? R : 1), sizeof(int64_t));
    for (int s = 0; s < S; s++) {                                    /* :92-115 */
        for (int i = 0; i < R; i++) {
            const int64_t a = ptr[(size_t)s * (R + 1) + i], b = ptr[(size_t)s * (R + 1) + i + 1];
            const int64_t olen = b - a;
            if (olen > lens[i]) {
                sizes[i] = (double*)realloc(sizes[i], (size_t)olen * sizeof(double));
                for (int64_t j = lens[i]; j < olen; j++) sizes[i][j] = 0.0;
            }
            const int64_t m = lens[i] < olen ? lens[i] : olen;
            int64_t j = 0;
            for (; j < m; j++) sizes[i][j] += (double)sizes_in[a + j] / scalar;
            for (; j < olen; j++) sizes[i][j] = (double)sizes_in[a + j] / scalar;   /* append */
            if (olen > lens[i]) lens[i] = olen;
        }
    }
    double* sums = (double*)calloc((size_t)(R > 0 ? R : 1), sizeof(double));
    double* pcts = (double*)calloc((size_t)(R > 0 ? R : 1), sizeof(double));
    double tot = 0.0;
    for (int k = 0; k < R; k++) {                                    /* chop :37-48, then sums :55-58 */
        const int64_t n = lens[k];
        double sum = 0.0;
        for (int64_t i = 0; i < n; i++) sum += sizes[k][i];
        const double m = sum / (double)n;
        double ss = 0.0, comp = 0.0;
        for (int64_t i = 0; i < n; i++) { double d = sizes[k][i] - m; ss += d * d; comp += d; }
        const double sd = sqrt((ss - comp * comp / (double)n) / (double)(n - 1));
        const double max = m + 3 * sd;
        for (int64_t i = 0; i < n; i++) if (sizes[k][i] > max) sizes[k][i] = 8 * m;
        double t = 0.0;
        for (int64_t i = 0; i < n; i++) t += sizes[k][i];
        sums[k] = t;
        tot += t;
    }
    for (int k = 0; k < R; k++) pcts[k] = sums[k] / tot;
    int64_t w = 0;
    int overflow = 0;
#define EMIT(nm, st, en, sm, sp) do { char b_[256]; int n_ = snprintf(b_, sizeof b_, "%s\t%lld\t%lld\t%.2f\t%d\n", nm, (long long)(st), (long long)(en), (double)(sm), (int)(sp)); \
        if (w + n_ + 1 > cap) overflow = 1; else { memcpy(out + w, b_, (size_t)n_); w += n_; } } while (0)
    for (int ri = 0; ri < R; ri++) {                                 /* :119-188 */
        const char* name = names[ri];
        const int64_t rlen = ref_lens[ri];
        if (lens[ri] == 0) { EMIT(name, 0, rlen, 0.0, 0); continue; }
        const double share = pcts[ri] * (double)N;
        int n = (share == share) ? (int)share : 0;
        if (n == 0 && pcts[ri] > 0) n = 1;
        else if (n == 0) { EMIT(name, 0, rlen, 0.0, 0); continue; }
        const double chunk = sums[ri] / (double)n;
        const double* size = sizes[ri];
        const int64_t len = lens[ri];
        double sum = 0.0;
        int64_t lasti = 0;
        for (int64_t i = 0; i < len; i++) {
            const int ovl = os_overlaps(pr, ps, pe, np, ri, i * TW, (i + 1) * TW);
            if (size[i] > chunk || (size[i] >= 0.05 * chunk && ovl)) {
                if (i > lasti) EMIT(name, lasti * TW, i * TW, sum, 1);
                sum = size[i];
                int nsplits = (int)(0.5 + (sum / (chunk / 2)));
                if (nsplits > 8) nsplits = 8;
                else if (nsplits < 1) { nsplits = 1; if (ovl) nsplits = 3; }
                int64_t start = i * TW;
                const int64_t l = (int64_t)((double)TW / (double)nsplits + 1);
                for (int k = 0; k < nsplits; k++) {
                    if (i + k == len + 1) EMIT(name, start, rlen, sum / (double)nsplits, nsplits);
                    else { int64_t en = start + l < (i + 1) * TW ? start + l : (i + 1) * TW; EMIT(name, start, en, sum / (double)nsplits, nsplits); }
                    start += l;
                }
                lasti = i + 1; sum = 0.0;
                continue;
            }
            sum += size[i];
            if (sum >= chunk || i == len - 1 || (sum >= 0.2 * chunk && ovl)) {
                if (i == len - 1) EMIT(name, lasti * TW, rlen, sum, 1);
                else EMIT(name, lasti * TW, (i + 1) * TW, sum, 1);
                lasti = i + 1; sum = 0.0;
            }
        }
    }
#undef EMIT
    for (int k = 0; k < R; k++) free(sizes[k]);
    free(sizes); free(lens); free(sums); free(pcts);
    if (overflow) return -1;
    out[w] = 0;
    return w;
}
