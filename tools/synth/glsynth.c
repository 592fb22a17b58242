/* glsynth.c — seeded synthetic alignment segments shaped like a 30x short-read WGS BAM, fast enough for
 * the whole 3.1 Gb genome (a few seconds on all cores).  WORKLOAD GENERATOR for bench.py (both arms load it, so they time
 * the same input); not part of the depth engine and not linked into libgoleft_b200.so.  Recipe: glsynth_core.h.
 * Output: the M blocks [start,end) of the reads that pass the `samtools depth -Q min_mapq` filter, in BAM record order
 * (a deletion read yields two consecutive blocks). */
#include <pthread.h>
#include "glsynth_core.h"

typedef struct {
    gls_contig C; int min_mapq;
    int64_t* cell_off;            /* [n_cells+1] segment offsets after the counting pass */
    int32_t* start; int32_t* end;
    int pass;                     /* 0 count, 1 fill */
    int64_t next; pthread_mutex_t mu;
} Job;

static int64_t do_cell(const Job* J, int64_t c, int32_t* start, int32_t* end, gls_read* rd, uint64_t* pk) {
    const int64_t m = gls_cell_reads(&J->C, c, rd, pk);
    const int L = J->C.read_len;
    int64_t k = 0;
    for (int64_t j = 0; j < m; j++) {
        const gls_read* r = &rd[j];
        if (!gls_passes(r, J->min_mapq)) continue;
        const int64_t p = r->pos;
        if (r->kind == 1) {                                   /* kM xD (L-k)M */
            if (start) { start[k] = (int32_t)p; end[k] = (int32_t)(p + r->k); start[k + 1] = (int32_t)(p + r->k + r->x); end[k + 1] = (int32_t)(p + L + r->x); }
            k += 2;
        } else {
            const int len = r->kind == 2 ? L - r->x : (r->kind == 3 ? L - r->k : L);   /* kM xI (L-k-x)M: the M blocks abut; sS (L-s)M */
            if (start) { start[k] = (int32_t)p; end[k] = (int32_t)(p + len); }
            k += 1;
        }
    }
    return k;
}

static void* worker(void* arg) {
    Job* J = (Job*)arg;
    gls_read* rd = (gls_read*)malloc(sizeof(gls_read) * GLS_MAX_CELL_READS);
    uint64_t* pk = (uint64_t*)malloc(8 * (size_t)GLS_MAX_CELL_READS);
    for (;;) {
        pthread_mutex_lock(&J->mu);
        const int64_t c0 = J->next;
        J->next += 16;
        pthread_mutex_unlock(&J->mu);
        if (c0 >= J->C.n_cells) break;
        for (int64_t c = c0; c < c0 + 16 && c < J->C.n_cells; c++) {
            if (J->pass == 0) J->cell_off[c + 1] = do_cell(J, c, NULL, NULL, rd, pk);
            else do_cell(J, c, J->start + J->cell_off[c], J->end + J->cell_off[c], rd, pk);
        }
    }
    free(rd); free(pk);
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
    gls_contig_init(&J.C, length, coverage, read_len, seed, gap, pileup);
    J.min_mapq = min_mapq;
    J.cell_off = (int64_t*)calloc((size_t)J.C.n_cells + 1, 8);
    pthread_mutex_init(&J.mu, NULL);
    J.pass = 0;
    run_pass(&J, threads);
    for (int64_t c = 0; c < J.C.n_cells; c++) J.cell_off[c + 1] += J.cell_off[c];
    *n = J.cell_off[J.C.n_cells];
    int rc = 0;
    if (start && end) {
        if (cap < *n) rc = -5;
        else { J.start = start; J.end = end; J.pass = 1; run_pass(&J, threads); }
    }
    pthread_mutex_destroy(&J.mu);
    free(J.cell_off);
    return rc;
}
