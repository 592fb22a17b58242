/* glsynth.c — seeded synthetic alignment segments shaped like a 30x short-read WGS BAM, fast enough for
 * the whole 3.1 Gb genome (618 M segments in a few seconds on all cores).  WORKLOAD GENERATOR for bench.py (both arms
 * load it, so they time the same input); not part of the depth engine and not linked into libgoleft_b200.so.
 * Same recipe as goleft_b200/synth.py (the tests' numpy generator): reads of read_len bases at uniform positions, a
 * 4.66 % gap at 40 % of the contig, a 200x pile-up at 70 %, CIGAR mix 97 % M / 1 % kM dD M / 1 % kM iI M / 1 % sS M,
 * flag mix (6 % dup, .2 % qcfail, .5 % secondary, .3 % supplementary, .1 % unmapped), MAPQ 0 for 5 %, filtered like
 * `samtools depth -Q min_mapq` (depth/depth.go:45): (flag & 0x704) == 0 && mapq >= min_mapq.  Output: the M blocks
 * [start,end) in BAM record order (a deletion read yields two consecutive blocks).
 * Positions: the contig is cut into 32 KB cells; cell c gets its share of the reads (uniform inside the cell, sorted),
 * every value comes from a counter-based hash of (seed, cell, read, field), so cells are independent -> parallel and
 * reproducible for any thread count. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define CELL 32768

static inline uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static inline uint64_t rnd(uint64_t seed, uint64_t cell, uint64_t i, uint64_t field) {
    return mix(mix(seed ^ (cell * 0xD1B54A32D192ED03ull)) ^ (i * 4 + field));
}
static inline double u01(uint64_t r) { return (double)(r >> 11) * (1.0 / 9007199254740992.0); }

typedef struct {
    int64_t length; double coverage; int read_len; uint64_t seed; int min_mapq; int gap, pileup;
    int64_t n_cells; int64_t* cell_off;            /* [n_cells+1] segment offsets after the counting pass */
    int32_t* start; int32_t* end;
    int pass;                                       /* 0 count, 1 fill */
    int64_t next; pthread_mutex_t mu;
    int64_t g0, g1, p0, plen; double extra_per_base;
} Job;

static int gls_cmp64(const void* a, const void* b) { const uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b; return (x > y) - (x < y); }

/* reads whose position falls in cell c -> their segments; returns the count, writes when start != NULL */
static int64_t do_cell(const Job* J, int64_t c, int32_t* start, int32_t* end, int32_t* posbuf, uint32_t* idbuf) {
    const int64_t lo = c * CELL, hi_all = J->length - J->read_len > 1 ? J->length - J->read_len : 1;
    int64_t hi = lo + CELL;
    if (lo >= hi_all) return 0;
    if (hi > hi_all) hi = hi_all;
    /* reads in this cell: base coverage + the pile-up's extra density where it overlaps */
    double want = J->coverage * (double)(hi - lo) / (double)J->read_len;
    if (J->pileup) {
        const int64_t a = lo > J->p0 - J->read_len + 1 ? lo : J->p0 - J->read_len + 1, b = hi < J->p0 + J->plen ? hi : J->p0 + J->plen;
        if (b > a) want += J->extra_per_base * (double)(b - a);
    }
    /* deterministic rounding with a per-cell dither so the total is right on average */
    int64_t n = (int64_t)(want + u01(rnd(J->seed, (uint64_t)c, 0xFFFFFFFFull, 3)));
    if (n > (1 << 22)) n = 1 << 22;
    int64_t m = 0;
    for (int64_t i = 0; i < n; i++) {
        int64_t p;
        const double u = u01(rnd(J->seed, (uint64_t)c, (uint64_t)i, 0));
        if (J->pileup) {
            /* choose between the base layer and the pile-up layer in proportion to their read counts in the cell */
            const int64_t a = lo > J->p0 - J->read_len + 1 ? lo : J->p0 - J->read_len + 1, b = hi < J->p0 + J->plen ? hi : J->p0 + J->plen;
            const double base = J->coverage * (double)(hi - lo) / (double)J->read_len;
            const double ex = b > a ? J->extra_per_base * (double)(b - a) : 0.0;
            const double v = u * (base + ex);
            if (v < base) p = lo + (int64_t)(v / base * (double)(hi - lo));
            else p = a + (int64_t)((v - base) / ex * (double)(b - a));
        } else p = lo + (int64_t)(u * (double)(hi - lo));
        if (p >= hi) p = hi - 1;
        if (J->gap && !(p + J->read_len + 16 <= J->g0 || p >= J->g1)) continue;
        posbuf[m] = (int32_t)p; idbuf[m] = (uint32_t)i; m++;
    }
    /* sort by position (ids follow: sort pairs packed in 64 bits) */
    uint64_t* pk = (uint64_t*)malloc((size_t)(m ? m : 1) * 8);
    for (int64_t i = 0; i < m; i++) pk[i] = ((uint64_t)(uint32_t)posbuf[i] << 32) | idbuf[i];
    /* positions are < 2^31 so unsigned order == signed order */
    qsort(pk, (size_t)m, 8, gls_cmp64);
    int64_t k = 0;
    const int L = J->read_len;
    for (int64_t j = 0; j < m; j++) {
        const int64_t p = (int64_t)(pk[j] >> 32);
        const uint64_t i = (uint32_t)pk[j];
        const uint64_t r1 = rnd(J->seed, (uint64_t)c, i, 1), r2 = rnd(J->seed, (uint64_t)c, i, 2);
        const double uf = u01(r1), um = u01(r1 * 0x9E3779B97F4A7C15ull + 1), uk = u01(r2);
        /* flags: dup .06 | qcfail .002 | secondary .005 | supplementary .003 (counted) | unmapped .001 */
        int pass = !(uf < 0.06 || (uf >= 0.06 && uf < 0.062) || (uf >= 0.062 && uf < 0.067) || (uf >= 0.070 && uf < 0.071));
        int mapq = um < 0.05 ? 0 : (um < 0.10 ? 1 + (int)((um - 0.05) / 0.05 * 59.0) : 60);
        if (!pass || mapq < J->min_mapq) continue;
        const int kind = uk >= 0.99 ? 3 : (uk >= 0.98 ? 2 : (uk >= 0.97 ? 1 : 0));
        const uint64_t r3 = mix(r2);
        const int kk = 1 + (int)(r3 % (uint64_t)(L - 13)), x = 1 + (int)((r3 >> 20) % 10), sc = 1 + (int)((r3 >> 40) % 50);
        if (kind == 1) {                                   /* kM xD (L-k)M */
            if (start) { start[k] = (int32_t)p; end[k] = (int32_t)(p + kk); start[k + 1] = (int32_t)(p + kk + x); end[k + 1] = (int32_t)(p + L + x); }
            k += 2;
        } else {
            const int len = kind == 2 ? L - x : (kind == 3 ? L - sc : L);     /* kM xI (L-k-x)M: the M blocks abut; sS (L-s)M */
            if (start) { start[k] = (int32_t)p; end[k] = (int32_t)(p + len); }
            k += 1;
        }
    }
    free(pk);
    return k;
}

static void* worker(void* arg) {
    Job* J = (Job*)arg;
    const size_t cap = (size_t)1 << 22;
    int32_t* posbuf = (int32_t*)malloc(cap * 4);
    uint32_t* idbuf = (uint32_t*)malloc(cap * 4);
    for (;;) {
        pthread_mutex_lock(&J->mu);
        const int64_t c0 = J->next;
        J->next += 16;
        pthread_mutex_unlock(&J->mu);
        if (c0 >= J->n_cells) break;
        for (int64_t c = c0; c < c0 + 16 && c < J->n_cells; c++) {
            if (J->pass == 0) J->cell_off[c + 1] = do_cell(J, c, NULL, NULL, posbuf, idbuf);
            else do_cell(J, c, J->start + J->cell_off[c], J->end + J->cell_off[c], posbuf, idbuf);
        }
    }
    free(posbuf); free(idbuf);
    return NULL;
}

static void run_pass(Job* J, int threads) {
    pthread_t th[256];
    if (threads > 256) threads = 256;
    if (threads < 1) threads = 1;
    J->next = 0;
    for (int t = 1; t < threads; t++) pthread_create(&th[t], NULL, worker, J);
    worker(J);
    for (int t = 1; t < threads; t++) pthread_join(th[t], NULL);
}

/* Call once with start == NULL to get *n (the segment count), then with arrays of that capacity. */
int gls_segments(int64_t length, double coverage, int read_len, uint64_t seed, int min_mapq, int gap, int pileup, int threads,
                 int32_t* start, int32_t* end, int64_t cap, int64_t* n) {
    if (length <= 0 || length > 2147483000ll || read_len < 20 || !n) return -1;
    Job J;
    memset(&J, 0, sizeof J);
    J.length = length; J.coverage = coverage; J.read_len = read_len; J.seed = seed; J.min_mapq = min_mapq;
    J.gap = gap && length > 50000; J.pileup = pileup && length > 50000;
    J.g0 = (int64_t)(0.40 * (double)length); J.g1 = J.g0 + (int64_t)(0.0466 * (double)length) + 1;
    J.p0 = (int64_t)(0.70 * (double)length); J.plen = length / 100 < 10000 ? length / 100 : 10000;
    J.extra_per_base = (200.0 - coverage) / (double)read_len * (double)J.plen / (double)(J.plen + read_len - 1);
    if (J.extra_per_base < 0) J.extra_per_base = 0;
    J.n_cells = (length + CELL - 1) / CELL;
    J.cell_off = (int64_t*)calloc((size_t)J.n_cells + 1, 8);
    pthread_mutex_init(&J.mu, NULL);
    J.pass = 0;
    run_pass(&J, threads);
    for (int64_t c = 0; c < J.n_cells; c++) J.cell_off[c + 1] += J.cell_off[c];
    *n = J.cell_off[J.n_cells];
    int rc = 0;
    if (start && end) {
        if (cap < *n) rc = -5;
        else { J.start = start; J.end = end; J.pass = 1; run_pass(&J, threads); }
    }
    pthread_mutex_destroy(&J.mu);
    free(J.cell_off);
    return rc;
}
