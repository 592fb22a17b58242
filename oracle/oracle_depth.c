/*
 * oracle_depth.c — CPU restatement of goleft's `depth` hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this; the product (libgoleft_b200.so, the goleft CLI) never links or calls it.
 *
 * PARITY UNPINNED: the reference is Go + an external `samtools depth` child; neither `go` nor
 * `samtools` exists in this image and the reference tree holds no expected-output files for
 * this path.  This file restates the cited lines; its only external pins are
 * the reference's own invariants (exact tiling, no duplicate rows; depth/functional-test.sh:10-39)
 * and an independent brute-force per-base counter (orc_pileup_brute) — see tests/test_oracle_depth.py.
 *
 * Citations are relative to the reference checkout (brentp/goleft @ v0.2.6).
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

/* ------------------------------------------------------------------ growable text buffer */
typedef struct { char* p; size_t len, cap; } obuf;
static void ob_reserve(obuf* b, size_t extra) {
    if (b->len + extra + 1 > b->cap) {
        size_t nc = b->cap ? b->cap * 2 : 4096;
        while (nc < b->len + extra + 1) nc *= 2;
        b->p = (char*)realloc(b->p, nc);
        b->cap = nc;
    }
}
static void ob_printf_row_f(obuf* b, const char* chrom, long long s, long long e, double v) {
    ob_reserve(b, strlen(chrom) + 96);
    /* depth.go:301  "%s\t%d\t%d\t%.4g%s\n" (stats tail empty without --stats) */
    b->len += (size_t)sprintf(b->p + b->len, "%s\t%lld\t%lld\t%.4g\n", chrom, s, e, v);
}
static void ob_printf_row_s(obuf* b, const char* chrom, long long s, long long e, const char* cls) {
    ob_reserve(b, strlen(chrom) + 96);
    /* depth.go:312  "%s\t%d\t%d\t%s\n" */
    b->len += (size_t)sprintf(b->p + b->len, "%s\t%lld\t%lld\t%s\n", chrom, s, e, cls);
}

void orc_free(void* p) { free(p); }

/* ------------------------------------------------------------------ D0: per-base counting
 * What the `samtools depth` child of depth/depth.go:45 computes for records that already passed
 * the flag (0x704) / MAPQ filter: one count per reference base covered by an M/=/X block.
 * Segments are [start,end) in contig coordinates; depth[] covers [rs,re).                      */

/* brute force: literally walk every covered base (independent of the difference-array trick) */
int64_t orc_pileup_brute(const int32_t* start, const int32_t* end, int64_t n,
                         int64_t rs, int64_t re, int32_t* depth) {
    memset(depth, 0, (size_t)(re - rs) * sizeof(int32_t));
    int64_t used = 0;
    for (int64_t i = 0; i < n; i++) {
        int64_t s = start[i] < rs ? rs : start[i];
        int64_t e = end[i] > re ? re : end[i];
        if (s >= e) continue;
        used++;
        for (int64_t x = s; x < e; x++) depth[x - rs]++;
    }
    return used;
}

/* best-effort CPU: difference array + running sum ("O" baseline of BASELINE.md §3) */
int64_t orc_pileup_diff(const int32_t* start, const int32_t* end, int64_t n,
                        int64_t rs, int64_t re, int32_t* depth) {
    int64_t len = re - rs;
    int32_t* diff = (int32_t*)calloc((size_t)len + 1, sizeof(int32_t));
    int64_t used = 0;
    for (int64_t i = 0; i < n; i++) {
        int64_t s = start[i] < rs ? rs : start[i];
        int64_t e = end[i] > re ? re : end[i];
        if (s >= e) continue;
        used++;
        diff[s - rs]++;
        diff[e - rs]--;
    }
    int32_t run = 0;
    for (int64_t x = 0; x < len; x++) { run += diff[x]; depth[x] = run; }
    free(diff);
    return used;
}

/* ------------------------------------------------------------------ D1: regions and chunks */

/* depth.go:73  regexp "(.+?)[:\t](\d+)([\-\t])(\d+).*?" + depth.go:75-94 chromStartEndFromLine.
 * Leftmost-first semantics with a lazy first group: the match starts at byte 0 and the chrom
 * is the SHORTEST prefix (>=1 byte, no '\n') that is followed by [:\t] digits+ [-\t] digits+.
 * Returns 0 on success; chrom_out must hold strlen(line)+1 bytes.                             */
int orc_chrom_start_end(const char* line, char* chrom_out, int64_t* start_out, int64_t* end_out) {
    size_t n = strlen(line);
    for (size_t m0 = 0; m0 < n; m0++) {           /* unanchored: leftmost match start */
        if (line[m0] == '\n') continue;
        for (size_t k = m0 + 1; k < n; k++) {     /* k = index of the [:\t] separator */
            if (line[k - 1] == '\n') break;       /* '.' does not match newline */
            if (line[k] != ':' && line[k] != '\t') continue;
            size_t a = k + 1, b = a;
            while (b < n && isdigit((unsigned char)line[b])) b++;
            if (b == a) continue;
            /* \d+ is greedy but may back off; the separator must be [-\t], which is not a digit,
             * so only the maximal digit run can be followed by it. */
            if (b >= n || (line[b] != '-' && line[b] != '\t')) continue;
            size_t c = b + 1, d = c;
            while (d < n && isdigit((unsigned char)line[d])) d++;
            if (d == c) continue;
            memcpy(chrom_out, line + m0, k - m0);
            chrom_out[k - m0] = 0;
            long long istart = atoll(line + a);
            if (line[b] == '-') istart--;          /* depth.go:86-88: region syntax is 1-based */
            long long iend = atoll(line + c);
            *start_out = istart < 0 ? 0 : istart;  /* depth.go:93 max(istart,0) */
            *end_out = iend;
            return 0;
        }
    }
    return -1;                                     /* depth.go:77-79 log.Fatal */
}

/* depth.go:132,150-151: step = max(1, 10_000_000/W)*W ; chunks [i, min(i+step,len)).
 * Returns the number of chunks; fills starts/ends up to cap.                                   */
int64_t orc_gen_chunks(int64_t length, int64_t W, int64_t* starts, int64_t* ends, int64_t cap) {
    int64_t step = 10000000 / W;
    if (step < 1) step = 1;
    step *= W;
    int64_t k = 0;
    for (int64_t i = 0; i < length; i += step) {
        if (k < cap) { starts[k] = i; ends[k] = (i + step < length) ? i + step : length; }
        k++;
    }
    return k;
}

/* ------------------------------------------------------------------ D2/D3: the chunk walker
 * Faithful restatement of the callback in depth/depth.go:238-364.  It consumes the stream the
 * `samtools depth` child prints (one "chrom\tpos1\tdepth" line per position with depth>0 inside
 * the region) through a line source, so the same walker runs over a per-base array (tests) and
 * over literal text (the reference-shaped CPU baseline "R").                                    */

typedef struct line_src {
    /* returns 1 and sets (*pos0,*depth) for the next line, 0 at EOF */
    int (*next)(struct line_src*, int64_t* pos0, int64_t* depth);
    /* array source */
    const int32_t* depth; int64_t rs, re, x;
    /* text source */
    const char* txt; size_t tlen, tpos;
} line_src;

static int next_array(line_src* s, int64_t* pos0, int64_t* depth) {
    while (s->x < s->re) {
        int32_t d = s->depth[s->x - s->rs];
        if (d > 0) { *pos0 = s->x; *depth = d; s->x++; return 1; }   /* samtools>=1.13: no zero lines */
        s->x++;
    }
    return 0;
}

/* depth.go:282-291 + getPosDepth :202-221: ReadString('\n'); chrom = line[:Index(line,"\t")];
 * SplitN(rest,"\t",2); Atoi(pos)-1; strip '\n'; Atoi(depth).                                   */
static int next_text(line_src* s, int64_t* pos0, int64_t* depth) {
    if (s->tpos >= s->tlen) return 0;
    const char* line = s->txt + s->tpos;
    const char* nl = (const char*)memchr(line, '\n', s->tlen - s->tpos);
    if (!nl) return 0;                                 /* err != nil -> loop ends (:282) */
    const char* t1 = (const char*)memchr(line, '\t', (size_t)(nl - line));
    if (!t1) return 0;
    const char* p = t1 + 1;
    long long pos = 0;
    if (p >= nl || !isdigit((unsigned char)*p)) return 0;   /* Atoi error -> break (:288-290) */
    while (p < nl && isdigit((unsigned char)*p)) { pos = pos * 10 + (*p - '0'); p++; }
    if (p >= nl || *p != '\t') return 0;
    p++;
    long long dp = 0;
    if (p >= nl || !isdigit((unsigned char)*p)) return 0;
    while (p < nl && isdigit((unsigned char)*p)) { dp = dp * 10 + (*p - '0'); p++; }
    if (p != nl) return 0;
    s->tpos = (size_t)(nl - s->txt) + 1;
    *pos0 = pos - 1;                                    /* :209 pos-- */
    *depth = dp;
    return 1;
}

/* depth.go:223-234 */
static const char* cov_class(int64_t depth, int64_t minCov, int64_t maxMeanDepth) {
    if (depth == 0) return "NO_COVERAGE";
    if (depth < minCov) return "LOW_COVERAGE";
    if (maxMeanDepth > 0 && depth >= maxMeanDepth) return "EXCESSIVE_COVERAGE";
    return "CALLABLE";
}

static int64_t imax(int64_t a, int64_t b) { return a > b ? a : b; }
static int64_t imin(int64_t a, int64_t b) { return a < b ? a : b; }

/* depth.go:181-189: float64 accumulation of ints, then / float64(l) */
static double mean_cache(const int64_t* sl, size_t n, int64_t l) {
    if (n == 0 || l == 0) return 0;
    double avg = 0;
    for (size_t i = 0; i < n; i++) avg += (double)sl[i];
    return avg / (double)l;
}

static void walk(line_src* src, const char* chrom, int64_t regionStart, int64_t regionEnd,
                 int64_t W, int64_t minCov, int64_t maxMeanDepth, obuf* fhHD, obuf* fhCA) {
    size_t cache_cap = 1024, cache_n = 0;
    int64_t* depthCache = (int64_t*)malloc(cache_cap * sizeof(int64_t));
    int64_t depth = 0, pos = 0;                                          /* :255 */
    int64_t lastWindow = imax(0, regionStart / W);                       /* :263 */
    int64_t cache0 = regionStart - 1, cache1 = regionStart - 1;          /* :264-266 */
    const char* lastCovClass = "";                                       /* :267 */

    while (src->next(src, &pos, &depth)) {                               /* :282-325 */
        if (pos / W != lastWindow) {                                     /* :293 */
            int64_t thisWindow = pos / W;
            for (int64_t iw = lastWindow; iw < thisWindow; iw++) {       /* :296 */
                int64_t s = imax(regionStart, iw * W);
                int64_t e = imin(regionEnd, (iw + 1) * W);
                ob_printf_row_f(fhHD, chrom, s, e, mean_cache(depthCache, cache_n, e - s));
                cache_n = 0;                                             /* :302 */
            }
            lastWindow = thisWindow;                                     /* :304 */
        }
        if (cache_n == cache_cap) { cache_cap *= 2; depthCache = (int64_t*)realloc(depthCache, cache_cap * sizeof(int64_t)); }
        depthCache[cache_n++] = depth;                                   /* :306 */
        const char* covClass = cov_class(depth, minCov, maxMeanDepth);   /* :307 */
        if (strcmp(covClass, lastCovClass) != 0 || pos != cache1 + 1) {  /* :310 */
            if (lastCovClass[0] != 0)
                ob_printf_row_s(fhCA, chrom, cache0, cache1 + 1, lastCovClass);        /* :312 */
            if (pos != cache1 + 1)
                ob_printf_row_s(fhCA, chrom, cache1 + 1, pos, "NO_COVERAGE");          /* :316 */
            lastCovClass = covClass;
            cache0 = pos; cache1 = pos;                                  /* :319-320 */
        } else {
            cache1 = pos;                                                /* :322 */
        }
    }
    if (cache0 != -1 && lastCovClass[0] != 0)                            /* :326 */
        ob_printf_row_s(fhCA, chrom, cache0, cache1 + 1, lastCovClass);
    if (cache_n > 0) {                                                   /* :329 */
        int64_t s = pos / W * W;
        if (s < regionEnd) {
            int64_t s2 = imax(s, regionStart);                           /* :332 (shadowed s) */
            int64_t e = imin(regionEnd, s2 + W);                         /* :333  (quirk Q2) */
            ob_printf_row_f(fhHD, chrom, s2, e, mean_cache(depthCache, cache_n, e - s2));
            cache_n = 0;
            pos = e;                                                     /* :338 */
        }
    }
    if (cache1 + 1 < regionEnd) {                                        /* :343 */
        if (cache1 != -1)
            ob_printf_row_s(fhCA, chrom, cache1 + 1, regionEnd, "NO_COVERAGE");        /* :346 */
        else
            ob_printf_row_s(fhCA, chrom, regionStart, regionEnd, "NO_COVERAGE");       /* :349 */
        for (int64_t ds = imax(regionStart, pos) / W * W; ds < regionEnd && pos < regionEnd; ds += W) {  /* :351 */
            int64_t de = imin(regionEnd, ds + W);
            int64_t s = imax(ds, regionStart);
            ob_printf_row_f(fhHD, chrom, s, de, mean_cache(depthCache, cache_n, de - s));
            cache_n = 0;
        }
    }
    free(depthCache);
}

/* array source: depth[] holds per-base depth for [rs,re) */
int orc_walk_chunk(const char* chrom, int64_t rs, int64_t re, int64_t W, int64_t mincov, int64_t maxmean,
                   const int32_t* depth,
                   char** depth_bed, int64_t* depth_len, char** callable_bed, int64_t* callable_len) {
    if (W <= 0 || re <= rs) return -1;
    obuf hd = {0}, ca = {0};
    ob_reserve(&hd, 1); ob_reserve(&ca, 1);
    line_src s; memset(&s, 0, sizeof s);
    s.next = next_array; s.depth = depth; s.rs = rs; s.re = re; s.x = rs;
    walk(&s, chrom, rs, re, W, mincov, maxmean, &hd, &ca);
    hd.p[hd.len] = 0; ca.p[ca.len] = 0;
    *depth_bed = hd.p; *depth_len = (int64_t)hd.len;
    *callable_bed = ca.p; *callable_len = (int64_t)ca.len;
    return 0;
}

/* the text the `echo region; samtools depth` child writes for one chunk (depth.go:45):
 * first the echoed region "chrom:rs+1-re", then one line per covered base.                  */
int orc_samtools_text(const char* chrom, int64_t rs, int64_t re, const int32_t* depth,
                      char** out, int64_t* out_len) {
    obuf b = {0};
    ob_reserve(&b, strlen(chrom) + 64);
    b.len += (size_t)sprintf(b.p + b.len, "%s:%lld-%lld\n", chrom, (long long)rs + 1, (long long)re);
    size_t cl = strlen(chrom);
    for (int64_t x = rs; x < re; x++) {
        int32_t d = depth[x - rs];
        if (d <= 0) continue;
        ob_reserve(&b, cl + 40);
        b.len += (size_t)sprintf(b.p + b.len, "%s\t%lld\t%d\n", chrom, (long long)x + 1, d);
    }
    b.p[b.len] = 0;
    *out = b.p; *out_len = (int64_t)b.len;
    return 0;
}

/* text source: exactly what the Go callback does with its io.Reader (depth.go:256-262 reads the
 * echoed region line first, then the per-line loop).                                         */
int orc_walk_text(const char* txt, int64_t txt_len, int64_t W, int64_t mincov, int64_t maxmean,
                  char** depth_bed, int64_t* depth_len, char** callable_bed, int64_t* callable_len) {
    const char* nl = (const char*)memchr(txt, '\n', (size_t)txt_len);
    if (!nl || W <= 0) return -1;
    size_t l0 = (size_t)(nl - txt);
    char* region = (char*)malloc(l0 + 1);
    char* chrom = (char*)malloc(l0 + 1);
    memcpy(region, txt, l0); region[l0] = 0;
    int64_t rs, re;
    if (orc_chrom_start_end(region, chrom, &rs, &re) != 0) { free(region); free(chrom); return -1; }
    obuf hd = {0}, ca = {0};
    ob_reserve(&hd, 1); ob_reserve(&ca, 1);
    line_src s; memset(&s, 0, sizeof s);
    s.next = next_text; s.txt = txt; s.tlen = (size_t)txt_len; s.tpos = l0 + 1;
    walk(&s, chrom, rs, re, W, mincov, maxmean, &hd, &ca);
    hd.p[hd.len] = 0; ca.p[ca.len] = 0;
    *depth_bed = hd.p; *depth_len = (int64_t)hd.len;
    *callable_bed = ca.p; *callable_len = (int64_t)ca.len;
    free(region); free(chrom);
    return 0;
}

/* ------------------------------------------------------------------ array-level summaries
 * (what the GPU returns before formatting) derived from the per-base array, for direct
 * integer comparison in the parity tests.                                                    */

/* windows of [rs,re): k-th covers [max(rs,(rs/W+k)W), min(re,(rs/W+k+1)W)) */
int64_t orc_window_sums(const int32_t* depth, int64_t rs, int64_t re, int64_t W,
                        int64_t* sum_out, int32_t* min_out, int64_t cap) {
    int64_t w0 = rs / W, w1 = (re - 1) / W, k = 0;
    for (int64_t iw = w0; iw <= w1; iw++, k++) {
        int64_t s = imax(rs, iw * W), e = imin(re, (iw + 1) * W);
        int64_t sum = 0; int32_t mn = INT32_MAX;
        for (int64_t x = s; x < e; x++) { int32_t d = depth[x - rs]; sum += d; if (d < mn) mn = d; }
        if (k < cap) { sum_out[k] = sum; if (min_out) min_out[k] = mn; }
    }
    return k;
}

/* run-length encode classes over [rs,re), breaking at multiples of run_break (0 = never) */
int64_t orc_class_runs(const int32_t* depth, int64_t rs, int64_t re, int64_t mincov, int64_t maxmean,
                       int64_t run_break, int32_t* run_start, uint8_t* run_class, int64_t cap) {
    int64_t k = 0; int prev = -1;
    for (int64_t x = rs; x < re; x++) {
        int64_t d = depth[x - rs];
        int c = d == 0 ? 0 : d < mincov ? 1 : (maxmean > 0 && d >= maxmean) ? 3 : 2;
        if (c != prev || (run_break > 0 && x % run_break == 0)) {
            if (k < cap) { run_start[k] = (int32_t)x; run_class[k] = (uint8_t)c; }
            k++;
        }
        prev = c;
    }
    return k;
}

/* ------------------------------------------------------------------ --stats (depth/depth.go:191-200)
 * getStats -> github.com/brentp/faidx Faidx.Stats (go.mod:11 v0.0.0-20200301150453-c39eb85760d8, NOT under the
 * reference tree; restated from that package's published source — PARITY UNPINNED, nothing in the reference's
 * tests asserts a --stats value).  `position` maps a 0-based base to a file offset through the .fai columns;
 * Stats walks the raw bytes mmap[position(start) : position(end) (+1 if inside the file)), skipping the slice's last
 * byte, so newlines are stepped over and a C whose next BYTE is G/g counts as CpG.
 * out3 = {GC, CpG, Masked} (the order depth.go:199 prints).  Returns -1 where faidx panics (coordinate > length). */
static int64_t orc_fa_position(int64_t rec_start, int64_t lbases, int64_t lbytes, int64_t p) {
    return rec_start + p / lbases * lbytes + p % lbases;
}

int orc_faidx_stats(const uint8_t* map, int64_t map_len, int64_t rec_start, int64_t rec_len, int64_t lbases, int64_t lbytes,
                    int64_t start, int64_t end, double* out3) {
    out3[0] = out3[1] = out3[2] = 0.0;
    if (start < 0 || end < 0 || start > rec_len || end > rec_len || lbases <= 0) return -1;
    int64_t pstart = orc_fa_position(rec_start, lbases, lbytes, start);
    int64_t pend = orc_fa_position(rec_start, lbases, lbytes, end);
    int64_t oend = pend;
    if (oend < map_len) oend++;
    if (pstart > oend || oend > map_len) return -1;              /* Go slice bounds panic */
    int64_t gc_up = 0, gc_lo = 0, at_up = 0, at_lo = 0, cpg = 0;
    int64_t n = oend - pstart;
    const uint8_t* buf = map + pstart;
    for (int64_t i = 0; i < n; i++) {
        if (i == n - 1) break;                                   /* "we added 1 to do the GC content" */
        uint8_t v = buf[i];
        if (v == 'G' || v == 'C') {
            if (v == 'C' && (buf[i + 1] == 'G' || buf[i + 1] == 'g')) cpg++;
            gc_up++;
        } else if (v == 'g' || v == 'c') {
            if (v == 'c' && (buf[i + 1] == 'G' || buf[i + 1] == 'g')) cpg++;
            gc_lo++;
        } else if (v == 'A' || v == 'T') {
            at_up++;
        } else if (v == 'a' || v == 't') {
            at_lo++;
        }
    }
    double tot = (double)(gc_up + gc_lo + at_up + at_lo);
    if (tot == 0.0) return 0;
    out3[0] = (double)(gc_lo + gc_up) / tot;
    out3[1] = (double)(2 * cpg) / tot;
    out3[2] = (double)(at_lo + gc_lo) / tot;
    return 0;
}

/* the "%s" tail of a window row (depth.go:199) */
int orc_stats_text(const double* st3, char* out, int64_t cap) {
    return snprintf(out, (size_t)cap, "\t%.3g\t%.3g\t%.3g", st3[0], st3[1], st3[2]);
}

/* ------------------------------------------------------------------ chunk-parallel driver (bench only)
 * `goleft depth -p N` runs one `samtools depth` child + callback per 10 Mb chunk, N at a time (depth.go:132,150-154,
 * 392-394).  orc_depth_jobs_mt is that loop for the bench's CPU arm: `threads` workers take chunk jobs from a shared
 * counter; per job: per-base counting (orc_pileup_diff, the child) + the callback walk + BED text (walk()), all in C with
 * per-worker buffers, so the Python driver's allocator / GIL do not throttle the baseline.  A job's segments are the
 * slice of its contig's start-sorted arrays that can reach the chunk (binary search with `slack` bases of look-back).
 * Returns the number of jobs done; *text_bytes = total bytes of both BED texts.                                        */
#include <pthread.h>
typedef struct {
    const char* chrom; const int32_t* start; const int32_t* end; int64_t n;   /* contig segments, sorted by start */
    int64_t rs, re;
} orc_chunk_job;

typedef struct {
    const orc_chunk_job* jobs; int64_t n_jobs; int64_t W, mincov, maxmean, slack;
    int64_t next; pthread_mutex_t mu; int64_t text_bytes, done;
} orc_jobs_ctx;

static int64_t lower_bound_i32(const int32_t* a, int64_t n, int64_t v) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if ((int64_t)a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}

static void* orc_jobs_worker(void* arg) {
    orc_jobs_ctx* c = (orc_jobs_ctx*)arg;
    int32_t* depth = NULL;
    int64_t depth_cap = 0, bytes = 0, done = 0;
    obuf hd = {0}, ca = {0};
    for (;;) {
        pthread_mutex_lock(&c->mu);
        const int64_t j = c->next++;
        pthread_mutex_unlock(&c->mu);
        if (j >= c->n_jobs) break;
        const orc_chunk_job* jb = &c->jobs[j];
        const int64_t len = jb->re - jb->rs;
        if (len > depth_cap) { free(depth); depth = (int32_t*)malloc((size_t)len * 4); depth_cap = len; }
        const int64_t lo = lower_bound_i32(jb->start, jb->n, jb->rs - c->slack), hi = lower_bound_i32(jb->start, jb->n, jb->re);
        orc_pileup_diff(jb->start + lo, jb->end + lo, hi - lo, jb->rs, jb->re, depth);
        hd.len = ca.len = 0;
        ob_reserve(&hd, 1); ob_reserve(&ca, 1);
        line_src s; memset(&s, 0, sizeof s);
        s.next = next_array; s.depth = depth; s.rs = jb->rs; s.re = jb->re; s.x = jb->rs;
        walk(&s, jb->chrom, jb->rs, jb->re, c->W, c->mincov, c->maxmean, &hd, &ca);
        bytes += (int64_t)(hd.len + ca.len);
        done++;
    }
    free(depth); free(hd.p); free(ca.p);
    pthread_mutex_lock(&c->mu);
    c->text_bytes += bytes; c->done += done;
    pthread_mutex_unlock(&c->mu);
    return NULL;
}

int64_t orc_depth_jobs_mt(const orc_chunk_job* jobs, int64_t n_jobs, int64_t W, int64_t mincov, int64_t maxmean, int64_t slack,
                          int threads, int64_t* text_bytes) {
    orc_jobs_ctx c;
    memset(&c, 0, sizeof c);
    c.jobs = jobs; c.n_jobs = n_jobs; c.W = W; c.mincov = mincov; c.maxmean = maxmean; c.slack = slack;
    pthread_mutex_init(&c.mu, NULL);
    if (threads < 1) threads = 1;
    if (threads > 1024) threads = 1024;
    if ((int64_t)threads > n_jobs) threads = (int)(n_jobs > 0 ? n_jobs : 1);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 1; t < threads; t++) pthread_create(&th[t], NULL, orc_jobs_worker, &c);
    orc_jobs_worker(&c);
    for (int t = 1; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    pthread_mutex_destroy(&c.mu);
    if (text_bytes) *text_bytes = c.text_bytes;
    return c.done;
}
