/* glsynth_core.h — the read recipe shared by glsynth.c (segments for bench.py) and bamsynth.c (a real BAM + BAI of the
 * same reads for the CLI wall-clock leg).  WORKLOAD GENERATOR, not part of the depth engine.
 * Recipe (same as goleft_b200/synth.py, the tests' numpy generator): reads of read_len bases at uniform positions, a
 * 4.66 % gap at 40 % of the contig, a 200x pile-up at 70 %, CIGAR mix 97 % M / 1 % kM dD M / 1 % kM iI M / 1 % sS M,
 * flag mix (6 % dup, .2 % qcfail, .5 % secondary, .3 % supplementary, .1 % unmapped), MAPQ 0 for 5 %, 1..59 for 5 %.
 * The contig is cut into 32 KB cells; every value comes from a counter-based hash of (seed, cell, read, field), so cells
 * are independent -> parallel and reproducible for any thread count. */
#ifndef GLSYNTH_CORE_H
#define GLSYNTH_CORE_H
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define GLS_CELL 32768
#define GLS_MAX_CELL_READS (1 << 22)

static inline uint64_t gls_mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static inline uint64_t gls_rnd(uint64_t seed, uint64_t cell, uint64_t i, uint64_t field) {
    return gls_mix(gls_mix(seed ^ (cell * 0xD1B54A32D192ED03ull)) ^ (i * 4 + field));
}
static inline double gls_u01(uint64_t r) { return (double)(r >> 11) * (1.0 / 9007199254740992.0); }

typedef struct {
    int64_t length; double coverage; int read_len; uint64_t seed; int gap, pileup;
    int64_t n_cells; int64_t g0, g1, p0, plen; double extra_per_base;
} gls_contig;

typedef struct {
    int32_t pos; uint32_t id; uint16_t flag; uint8_t mapq, kind;   /* kind 0: LM, 1: kM xD (L-k)M, 2: kM xI (L-k-x)M, 3: sS (L-s)M */
    int16_t k, x;                                                  /* k (kinds 1,2) or the soft clip s (kind 3); x = d or i */
} gls_read;

static void gls_contig_init(gls_contig* J, int64_t length, double coverage, int read_len, uint64_t seed, int gap, int pileup) {
    memset(J, 0, sizeof *J);
    J->length = length; J->coverage = coverage; J->read_len = read_len; J->seed = seed;
    J->gap = gap && length > 50000; J->pileup = pileup && length > 50000;
    J->g0 = (int64_t)(0.40 * (double)length); J->g1 = J->g0 + (int64_t)(0.0466 * (double)length) + 1;
    J->p0 = (int64_t)(0.70 * (double)length); J->plen = length / 100 < 10000 ? length / 100 : 10000;
    J->extra_per_base = (200.0 - coverage) / (double)read_len * (double)J->plen / (double)(J->plen + read_len - 1);
    if (J->extra_per_base < 0) J->extra_per_base = 0;
    J->n_cells = (length + GLS_CELL - 1) / GLS_CELL;
}

static int gls_cmp64(const void* a, const void* b) { const uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b; return (x > y) - (x < y); }

/* every read whose position falls in cell c, sorted by (position, id); out must hold GLS_MAX_CELL_READS; returns the count */
static int64_t gls_cell_reads(const gls_contig* J, int64_t c, gls_read* out, uint64_t* pk /* scratch, same capacity */) {
    const int64_t lo = c * GLS_CELL, hi_all = J->length - J->read_len > 1 ? J->length - J->read_len : 1;
    int64_t hi = lo + GLS_CELL;
    if (lo >= hi_all) return 0;
    if (hi > hi_all) hi = hi_all;
    const double base = J->coverage * (double)(hi - lo) / (double)J->read_len;
    int64_t a = 0, b = 0;
    double ex = 0.0;
    if (J->pileup) {
        a = lo > J->p0 - J->read_len + 1 ? lo : J->p0 - J->read_len + 1;
        b = hi < J->p0 + J->plen ? hi : J->p0 + J->plen;
        if (b > a) ex = J->extra_per_base * (double)(b - a);
    }
    /* deterministic rounding with a per-cell dither so the total is right on average */
    int64_t n = (int64_t)(base + ex + gls_u01(gls_rnd(J->seed, (uint64_t)c, 0xFFFFFFFFull, 3)));
    if (n > GLS_MAX_CELL_READS) n = GLS_MAX_CELL_READS;
    int64_t m = 0;
    for (int64_t i = 0; i < n; i++) {
        const double v = gls_u01(gls_rnd(J->seed, (uint64_t)c, (uint64_t)i, 0)) * (base + ex);
        int64_t p = v < base || ex <= 0 ? lo + (int64_t)(v / base * (double)(hi - lo)) : a + (int64_t)((v - base) / ex * (double)(b - a));
        if (p >= hi) p = hi - 1;
        if (p < lo) p = lo;
        if (J->gap && !(p + J->read_len + 16 <= J->g0 || p >= J->g1)) continue;
        pk[m++] = ((uint64_t)(uint32_t)p << 32) | (uint32_t)i;
    }
    qsort(pk, (size_t)m, 8, gls_cmp64);                 /* positions < 2^31: unsigned order == signed order */
    const int L = J->read_len;
    for (int64_t j = 0; j < m; j++) {
        gls_read* r = &out[j];
        r->pos = (int32_t)(pk[j] >> 32);
        r->id = (uint32_t)pk[j];
        const uint64_t r1 = gls_rnd(J->seed, (uint64_t)c, r->id, 1), r2 = gls_rnd(J->seed, (uint64_t)c, r->id, 2);
        const double uf = gls_u01(r1), um = gls_u01(r1 * 0x9E3779B97F4A7C15ull + 1), uk = gls_u01(r2);
        r->flag = uf < 0.06 ? 0x400 : uf < 0.062 ? 0x200 : uf < 0.067 ? 0x100 : uf < 0.070 ? 0x800 : uf < 0.071 ? 0x4 : 0;
        r->mapq = (uint8_t)(um < 0.05 ? 0 : (um < 0.10 ? 1 + (int)((um - 0.05) / 0.05 * 59.0) : 60));
        r->kind = (uint8_t)(uk >= 0.99 ? 3 : (uk >= 0.98 ? 2 : (uk >= 0.97 ? 1 : 0)));
        const uint64_t r3 = gls_mix(r2);
        r->x = (int16_t)(1 + (int)((r3 >> 20) % 10));
        r->k = (int16_t)(r->kind == 3 ? 1 + (int)((r3 >> 40) % 50) : 1 + (int)(r3 % (uint64_t)(L - 13)));
    }
    return m;
}

/* the `samtools depth -Q q` filter of depth/depth.go:45 */
static inline int gls_passes(const gls_read* r, int min_mapq) { return (r->flag & 0x704) == 0 && (int)r->mapq >= min_mapq; }
#endif
