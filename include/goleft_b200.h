/*
 * goleft_b200.h — C ABI of libgoleft_b200.so, the B200 (sm_100a) engine behind
 * goleft's windowed-depth hot path (`depth`, `indexcov`, `covstats`, `depthwed`).
 *
 * The reference (brentp/goleft, pure Go) has no FFI for this path; each entry point
 * below cites the reference code whose work it replaces (paths relative to the
 * reference checkout).  INTEGRATION.md shows the cgo binding a maintainer would add.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes only.  No torch / C++ types.
 *  - every call returns int: 0 = ok, <0 = error (GL_E*); text via gl_last_error().
 *  - the caller owns every host buffer; the library owns device memory inside an
 *    opaque gl_ctx (one ctx per GPU, one host thread at a time per ctx).
 *  - coordinates are 0-based half-open, like the reference's internal ints.
 *  - pointers named d_* are DEVICE pointers (from gl_dev_alloc); all others are
 *    host pointers (pageable is fine, the library stages through pinned memory).
 *  - calls are synchronous on the ctx's stream unless stated otherwise.
 *  - there is NO CPU fallback: without a CUDA device every call fails with GL_ECUDA.
 */
#ifndef GOLEFT_B200_H
#define GOLEFT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GL_OK        0
#define GL_EINVAL   -1   /* bad argument */
#define GL_ECUDA    -2   /* CUDA runtime error (message has the cudaError string) */
#define GL_ENOMEM   -3   /* device or pinned-host allocation failed */
#define GL_ESTATE   -4   /* call out of order (e.g. add_segments before begin) */
#define GL_ERANGE   -5   /* output capacity too small / value out of supported range */
#define GL_ENCCL    -6   /* NCCL error */

/* coverage classes of depth/depth.go:223-234 (getCovClass) */
#define GL_NO_COVERAGE        0
#define GL_LOW_COVERAGE       1
#define GL_CALLABLE           2
#define GL_EXCESSIVE_COVERAGE 3

/* indexcov constants: indexcov/indexcov.go:153 (slots), indexcov/types.go:15 (TileWidth) */
#define GL_INDEXCOV_SLOTS 70
#define GL_TILE_WIDTH     16384

typedef struct gl_ctx gl_ctx;
typedef struct gl_bam gl_bam;            /* an open BAM + its index (gl_bam_open) */

/* ---------------------------------------------------------------- context */
const char* gl_version(void);                       /* "goleft_b200 <semver> sm_100a" */
int  gl_device_count(int* n);
int  gl_ctx_create(int device, gl_ctx** out);
int  gl_ctx_destroy(gl_ctx* ctx);
const char* gl_last_error(gl_ctx* ctx);             /* ctx may be NULL: last global error */
int  gl_sync(gl_ctx* ctx);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches claim) */
int  gl_launch_count(gl_ctx* ctx, int64_t* n);
/* raw cudaStream_t of the ctx, as an integer, so a caller can record events on it */
int  gl_stream_handle(gl_ctx* ctx, uint64_t* stream);

/* Pins the calling thread, and the threads created after it (the library's host pool, feeders), to the CPUs of the NUMA
 * node of `device`; with share_count > 1 to the share_index-th slice of that node's cores (ranks sharing a socket).
 * Call it before the first pinned allocation / pool use.  *node = -1 when the topology is unreadable (nothing changed). */
int  gl_device_numa_node(int device, int* node);     /* -1: unknown */
int  gl_bind_numa_for_device(int device, int share_index, int share_count, int* node, int* n_cpus);
/* Longest-processing-time-first placement of n work items (contigs, weight = length) on `bins` GPUs: the unit of
 * independence of depth/depth.go:129-159.  bin_of[i] = GPU of item i, bin_load[b] = total weight (may be NULL). */
int  gl_lpt_assign(const int64_t* weight, int32_t n, int32_t bins, int32_t* bin_of, int64_t* bin_load);

/* device memory owned by the ctx's device (plumbing for callers that keep inputs resident) */
int  gl_dev_alloc(gl_ctx* ctx, int64_t bytes, void** d_ptr);
int  gl_dev_free(gl_ctx* ctx, void* d_ptr);
int  gl_host_alloc_pinned(gl_ctx* ctx, int64_t bytes, void** h_ptr);
int  gl_host_free_pinned(gl_ctx* ctx, void* h_ptr);
int  gl_memcpy_h2d(gl_ctx* ctx, void* d_dst, const void* h_src, int64_t bytes);
int  gl_memcpy_d2h(gl_ctx* ctx, void* h_dst, const void* d_src, int64_t bytes);
/* device-timed interval helpers (CUDA events on the ctx stream) */
int  gl_timer_start(gl_ctx* ctx);
int  gl_timer_stop_ms(gl_ctx* ctx, float* ms);      /* records, synchronizes, returns elapsed */
/* per-kernel CUDA-event timing on the ctx stream: enable, run, then read "name ms\n" lines
 * (one per kernel launched since the last read; synchronizes).  Off by default. */
int  gl_profile_enable(gl_ctx* ctx, int on);
int  gl_profile_read(gl_ctx* ctx, char* buf, int64_t cap, int64_t* needed);
/* write `bytes` of junk to a scratch buffer (> L2) so the next launch starts cold */
int  gl_flush_l2(gl_ctx* ctx);

/* ------------------------------------------------------------------ depth
 * Replaces, per region [region_start, region_end) of one contig:
 *   - the `samtools depth -Q q -d D -r chr:b-e` child of depth/depth.go:45,116,152
 *     (per-base counting of the M/=/X blocks of records that pass the flag/MAPQ filter;
 *      the filter itself is applied by the segment feeder, see gl_bam_* below),
 *   - the per-line window accumulation of depth/depth.go:282-306,329-341 (Σ depth per
 *     genome-aligned window, clipped to the region),
 *   - the per-line class run-length encoding of depth/depth.go:307-327,343-350.
 * Text formatting (%.4g of sum/len, BED rows, the Q2/Q3 chunk-edge quirks) stays on the
 * host: gl_depth_format_chunk().
 */

/* Start a region [region_start, region_end) of one contig.  0 <= start < end, end-start < 2^30. */
int  gl_depth_begin(gl_ctx* ctx, int64_t region_start, int64_t region_end);

/* Add n segments [start[i], end[i]) (absolute contig coordinates, end>start; clipped to the region
 * by the kernels; segments outside are ignored).  May be called repeatedly between begin and reduce.
 * Host-pointer variant copies into a ctx-owned device store on a copy stream (pinned sources DMA
 * directly, pageable ones are staged); _device variant references HBM-resident arrays, which must
 * stay valid until the results have been fetched.
 * Any order is accepted.  Coordinate-sorted segments (non-decreasing start, what a BAM yields) of at
 * most 16384 bases take the fused path: each 4096-base tile's difference array is built in shared
 * memory straight from the segments.  Anything else takes the general path (bucketed events, see
 * gl_depth_last_path).  The device checks eligibility itself; results are identical. */
int  gl_depth_add_segments(gl_ctx* ctx, const int32_t* start, const int32_t* end, int64_t n);
int  gl_depth_add_segments_device(gl_ctx* ctx, const int32_t* d_start, const int32_t* d_end, int64_t n);

/* The feeder's compact format ("packed16", half the PCIe bytes): blocks of 256 slots, one int32 anchor per
 * block, slot k = [anchor+off[k], anchor+off[k]+len[k]) with uint16 off/len (len 0 = empty slot).
 * gl_pack_segments16 builds it on the host from plain arrays (splitting segments longer than 65535, opening
 * a new block when a start does not fit 16 bits); returns GL_ERANGE with the needed *n_blocks when
 * cap_blocks is too small (pass NULL outputs to only count).  The GPU unpacks into its segment store. */
int64_t gl_pack_segments16_bound(int64_t n);
int  gl_pack_segments16(const int32_t* start, const int32_t* end, int64_t n, int32_t* anchors, uint16_t* off,
                        uint16_t* len, int64_t cap_blocks, int64_t* n_blocks);
/* the same from `threads` host threads (0 = the whole pool): order-preserving, no sort; index chunks are packed independently */
int  gl_pack_segments16_mt(const int32_t* start, const int32_t* end, int64_t n, int32_t threads, int32_t* anchors, uint16_t* off,
                           uint16_t* len, int64_t cap_blocks, int64_t* n_blocks);
int  gl_depth_add_segments_packed16(gl_ctx* ctx, const int32_t* anchors, const uint16_t* off, const uint16_t* len,
                                    int64_t n_blocks);
/* The same device format with FIXED block boundaries: block b = input segments [256 b, 256 b + 256) in the caller's order,
 * anchor = their lowest start; n_blocks = ceil(n / 256) (caller-sized outputs: anchors[n_blocks], off/len[256 n_blocks]).
 * One streaming pass on the host pool, no sort, no count pass.  A block that cannot be expressed (starts more than 65535
 * bases apart, or a segment longer than 65535) is written as 256 empty slots and its segments are appended raw to
 * esc_start/esc_end (*n_esc in/out for the _range form); GL_ERANGE when they exceed esc_cap (*n_esc = needed).  Adding the
 * blocks (gl_depth_add_segments_packed16) and the escapes (gl_depth_add_segments) gives exactly the input's segments.
 * This is how gl_depth_bed_region moves decoder-native int32 arrays at 4 B/segment. */
int  gl_pack_segments16_fixed_mt(const int32_t* start, const int32_t* end, int64_t n, int32_t threads, int32_t* anchors, uint16_t* off,
                                 uint16_t* len, int32_t* esc_start, int32_t* esc_end, int64_t esc_cap, int64_t* n_esc);
int  gl_pack_segments16_fixed_range_mt(const int32_t* start, const int32_t* end, int64_t n, int64_t block_begin, int64_t block_end,
                                       int32_t threads, int32_t* anchors, uint16_t* off, uint16_t* len, int32_t* esc_start,
                                       int32_t* esc_end, int64_t esc_cap, int64_t* n_esc);
/* "packed8": a quarter of the bytes, for short-read data.  Blocks of 64 slots: int32 anchor (start of slot 0) +
 * per slot uint8 dstart (start - previous slot's start) and uint8 len (0 = empty/filler).  gl_pack_segments8 puts the
 * segments in start order (depth is order-independent), cuts segments longer than 255 into pieces and bridges gaps
 * > 255 bases with filler slots or a new block, whichever is smaller; same return convention as gl_pack_segments16.
 * Use it when 132 * n_blocks < 8 * n (always for 100-250 bp reads at >= 1x). */
int64_t gl_pack_segments8_bound(int64_t n);
int  gl_pack_segments8(const int32_t* start, const int32_t* end, int64_t n, int32_t* anchors, uint8_t* dstart, uint8_t* len,
                       int64_t cap_blocks, int64_t* n_blocks);
/* The same format from `threads` host threads (0 = all of the library's pool; GL_THREADS sets its size): the input is cut
 * into position ranges that are packed independently and concatenated, so the blocks differ from gl_pack_segments8's at the
 * range edges (a block may end early) but decode to the same multiset of pieces.  Host-only. */
int  gl_pack_segments8_mt(const int32_t* start, const int32_t* end, int64_t n, int32_t threads, int32_t* anchors, uint8_t* dstart,
                          uint8_t* len, int64_t cap_blocks, int64_t* n_blocks);
int  gl_depth_add_segments_packed8(gl_ctx* ctx, const int32_t* anchors, const uint8_t* dstart, const uint8_t* len, int64_t n_blocks);
/* same with the three arrays already in device memory (caller-owned, 8-byte aligned, must stay valid until the region's
 * last reduce): nothing is copied; a region whose only batch is packed8 is reduced straight from these words. */
int  gl_depth_add_segments_packed8_device(gl_ctx* ctx, const int32_t* d_anchors, const uint8_t* d_dstart, const uint8_t* d_len,
                                          int64_t n_blocks);

/* One fused pass: prefix scan -> per-base depth (never written to HBM) ->
 *   (a) per-window int64 sums for genome-aligned windows of size W clipped to the region:
 *       window k covers [max(rs,(rs/W+k)*W), min(re,(rs/W+k+1)*W)),
 *   (b) class run starts: position x starts a run when x==rs, class(x)!=class(x-1), or
 *       run_break>0 and x%run_break==0 (the reference never merges runs across its 10 Mb
 *       chunks, depth/depth.go:132,150: pass run_break=step to reproduce that).
 * mincov/maxmean as depth/depth.go:223-234.  Synchronous; results stay on the device until fetched. */
int  gl_depth_reduce(gl_ctx* ctx, int32_t W, int32_t mincov, int32_t maxmean, int64_t run_break);

/* Sizes of the results of the last gl_depth_reduce. */
int  gl_depth_result_sizes(gl_ctx* ctx, int64_t* n_windows, int64_t* n_runs, int32_t* max_depth);
/* Which path the last reduce took: 1 = fused (nearly sorted int32 segments), 2 = HBM difference array (the north-star
 * pipeline: red.global scatter + scan; on request only), 3 = packed8 (the region's only batch is packed8: read as it is, no
 * int32 copy, no per-segment index), 4 = bucketed events (the general path: any order, any length; two 16-bit events per
 * segment bucketed by 4096-base tile, difference arrays built in shared memory). */
int  gl_depth_last_path(gl_ctx* ctx, int32_t* path);
/* How the last gl_depth_bed_region / gl_depth_bed_contig call moved its int32 arrays: transport 0 = as they are (8 B/segment),
 * 16 = fixed-block packed16 made by the host pool chunk by chunk (4 B/segment); host seconds spent packing (the calling
 * thread's wall time inside the pool runs), bytes sent host->device, segments that went up raw as escapes. */
int  gl_depth_transport_stats(gl_ctx* ctx, int32_t* transport, double* pack_s, int64_t* h2d_bytes, int64_t* n_escaped);
/* host wall-clock phases of the same call: [0] setup (buffers, gl_depth_begin), [1] the pack + enqueue loop, [2] reduce + text + D2H */
int  gl_depth_transport_phases(gl_ctx* ctx, double phases_s[3]);
/* 0 = choose automatically (default: packed8 -> fused -> bucketed events), 1 = never the packed8 kernel, 2 = always the HBM
 * difference array, 4 = always bucketed events (tests, comparison runs). */
int  gl_depth_set_path(gl_ctx* ctx, int32_t path);

/* Fetch results (host buffers).  run_end may be NULL (run i ends where run i+1 starts; the last
 * ends at region_end). */
int  gl_depth_get_windows(gl_ctx* ctx, int64_t* sum_out, int64_t cap);
int  gl_depth_get_runs(gl_ctx* ctx, int32_t* run_start, int32_t* run_end, uint8_t* run_class, int64_t cap);

/* Forms named in SURVEY.md §8(b): run the pass for one output only.  gl_depth_windows also gives the
 * per-window minimum depth when min_out != NULL (an extra: the reference prints means only). */
int  gl_depth_windows(gl_ctx* ctx, int32_t W, int64_t* sum_out, int32_t* min_out, int64_t n_windows);
int  gl_depth_classes(gl_ctx* ctx, int32_t mincov, int32_t maxmean, int32_t* run_start, int32_t* run_end,
                      uint8_t* run_class, int64_t cap, int64_t* n_runs);
/* Sum of per-base depth over n arbitrary intervals [a[i], b[i]) of the open region (clipped to it), computed straight
 * from the segments (sum of overlaps); uses the cell index of the last gl_depth_reduce when that took the fused path.
 * This is what BED mode needs for the clipped first/last window of every region (depth/depth.go:293-305,329-341). */
int  gl_depth_interval_sums(gl_ctx* ctx, const int32_t* a, const int32_t* b, int64_t n, int64_t* sums);
/* Per-base depth of the region (debug / parity): depth_out has region_end-region_start entries. */
int  gl_depth_perbase(gl_ctx* ctx, int32_t* depth_out);

/* One-call form used by the CLI and the end-to-end benchmark: host segments in, host
 * results out, for one region; uploads are chunked and overlapped with the scatter. */
int  gl_depth_region(gl_ctx* ctx, int64_t region_start, int64_t region_end,
                     const int32_t* start, const int32_t* end, int64_t n,
                     int32_t W, int32_t mincov, int32_t maxmean, int64_t run_break,
                     int64_t* sum_out, int64_t win_cap, int64_t* n_windows,
                     int32_t* run_start, uint8_t* run_class, int64_t run_cap, int64_t* n_runs);

int  gl_depth_region_packed16(gl_ctx* ctx, int64_t region_start, int64_t region_end, const int32_t* anchors,
                              const uint16_t* off, const uint16_t* len, int64_t n_blocks,
                              int32_t W, int32_t mincov, int32_t maxmean, int64_t run_break,
                              int64_t* sum_out, int64_t win_cap, int64_t* n_windows,
                              int32_t* run_start, uint8_t* run_class, int64_t run_cap, int64_t* n_runs);
int  gl_depth_region_packed8(gl_ctx* ctx, int64_t region_start, int64_t region_end, const int32_t* anchors,
                             const uint8_t* dstart, const uint8_t* len, int64_t n_blocks, int32_t W, int32_t mincov, int32_t maxmean,
                             int64_t run_break, int64_t* sum_out, int64_t win_cap, int64_t* n_windows, int32_t* run_start,
                             uint8_t* run_class, int64_t run_cap, int64_t* n_runs);

/* BED text of the last gl_depth_reduce, formatted ON THE DEVICE (depth_text.cu): only finished bytes cross PCIe.
 *   depth_bed    : "chrom\tstart\tend\t%.4g\n" per window (depth/depth.go:301,337,357; mean = float64(sum)/float64(e-s),
 *                  depth.go:181-189; the four digits are rounded half-even on the exact binary value, like strconv),
 *   callable_bed : "chrom\tstart\tend\tCLASS\n" per class run (depth.go:312-349).
 * Needs region_start % W == 0 and run_break % W == 0: then the reference's per-chunk rows (incl. its tail rows,
 * depth.go:329-358) are exactly the genome-aligned windows of the region, each once.  BED-mode regions with unaligned
 * starts go through gl_depth_chunk_rows + gl_depth_format_rows (explicit rows) or gl_depth_format_chunk (host).
 * GL_ERANGE with both lengths set when a buffer is too small; gl_depth_text_bound() is always enough for depth_bed. */
int  gl_depth_text(gl_ctx* ctx, const char* chrom, char* depth_bed, int64_t depth_cap, int64_t* depth_len,
                   char* callable_bed, int64_t callable_cap, int64_t* callable_len);
int64_t gl_depth_text_bound(const char* chrom, int64_t n_windows);
/* n explicit window rows (s, e, sum) -> the same row text (the misaligned / re-emitted rows of depth.go:329-358) */
int  gl_depth_format_rows(gl_ctx* ctx, const char* chrom, const int32_t* row_s, const int32_t* row_e, const int64_t* row_sum,
                          int64_t n, char* out, int64_t cap, int64_t* len);
/* One contig end to end, the call `goleft depth` makes per reference sequence in .fai mode (depth.go:129-159 + the
 * callback 238-364 for each of its chunks, concatenated in order): host segments in -> BED bytes out.  step = the chunk
 * length (depth.go:132: a multiple of W).  The int32 arrays go over PCIe as fixed-block packed16 (4 B/segment), rewritten
 * chunk by chunk by the library's host pool (`threads` of it, 0 = the default 16) while the previous chunk is on the wire,
 * when there are >= 2^20 segments and the pool has >= 48 threads; otherwise, and for long reads, as they are (GL_BED_PACK,
 * gl_depth_transport_stats).  Everything is inside the call; it is synchronous. */
/* The same for [region_start, region_end) of a contig: region_start must be a multiple of step (whole chunks), so that
 * contigs longer than 2^30-1 bases, or one GPU's share of a contig, can be done in pieces whose texts concatenate. */
int  gl_depth_bed_region(gl_ctx* ctx, const char* chrom, int64_t region_start, int64_t region_end, const int32_t* start, const int32_t* end,
                         int64_t n, int32_t W, int32_t mincov, int32_t maxmean, int64_t step, int32_t threads,
                         char* depth_bed, int64_t depth_cap, int64_t* depth_len, char* callable_bed, int64_t callable_cap,
                         int64_t* callable_len);
int  gl_depth_bed_region_packed8(gl_ctx* ctx, const char* chrom, int64_t region_start, int64_t region_end, const int32_t* anchors,
                                 const uint8_t* dstart, const uint8_t* len, int64_t n_blocks, int32_t W, int32_t mincov, int32_t maxmean,
                                 int64_t step, char* depth_bed, int64_t depth_cap, int64_t* depth_len, char* callable_bed,
                                 int64_t callable_cap, int64_t* callable_len);
int  gl_depth_bed_contig(gl_ctx* ctx, const char* chrom, int64_t contig_len, const int32_t* start, const int32_t* end, int64_t n,
                         int32_t W, int32_t mincov, int32_t maxmean, int64_t step, int32_t threads,
                         char* depth_bed, int64_t depth_cap, int64_t* depth_len, char* callable_bed, int64_t callable_cap,
                         int64_t* callable_len);
int  gl_depth_bed_contig_packed8(gl_ctx* ctx, const char* chrom, int64_t contig_len, const int32_t* anchors, const uint8_t* dstart,
                                 const uint8_t* len, int64_t n_blocks, int32_t W, int32_t mincov, int32_t maxmean, int64_t step,
                                 char* depth_bed, int64_t depth_cap, int64_t* depth_len, char* callable_bed, int64_t callable_cap,
                                 int64_t* callable_len);

/* Host-side text: reproduces the rows the reference callback writes for ONE chunk
 * [rs,re) (depth/depth.go:293-305,326-358 incl. the chunk-edge quirks) from window sums
 * and runs restricted to that chunk.  Appends to malloc'd buffers; free with gl_free_text. */
int  gl_depth_format_chunk(const char* chrom, int64_t rs, int64_t re, int32_t W,
                           const int64_t* win_sum /* windows of [rs,re) */, int64_t n_windows,
                           const int32_t* run_start, const uint8_t* run_class, int64_t n_runs,
                           char** depth_bed, int64_t* depth_len, char** callable_bed, int64_t* callable_len);
void gl_free_text(char* p);

/* ---- `goleft depth --stats` (depth/depth.go:191-200 getStats, :246-252): GC / CpG / masked fraction per window row.
 * The arithmetic is github.com/brentp/faidx Faidx.Stats (go.mod:11, not vendored; parity unpinned), which scans the raw
 * bytes of the FASTA record, newlines included.
 * gl_depth_chunk_rows lists the (s,e) of the window rows gl_depth_format_chunk writes for a chunk, in order (including
 *   the misaligned / re-emitted rows of depth.go:329-358); returns GL_ERANGE with *n_rows set when cap is too small.
 * gl_fasta_load uploads the bytes of one FASTA record (from the .fai offset to the end of its last line, plus the one
 *   byte after it when the file has one), gl_fasta_stats counts, for row r, the slice [byte_start[r], byte_end[r]) of it
 *   the way Faidx.Stats does (every byte but the slice's last is classified, the next byte decides CpG):
 *   counts4[r] = {G+C, lower-case, A+C+G+T, CpG} and/or stats3[r] = {GC, CpG, Masked} as float64 (0 when no ACGT).
 *   byte_start = position(s), byte_end = min(position(e)+1, file length) relative to the loaded record.
 * gl_depth_format_chunk_stats = gl_depth_format_chunk with the three "%.3g" columns appended to every window row. */
int  gl_depth_chunk_rows(int64_t rs, int64_t re, int32_t W, const int32_t* run_start, const uint8_t* run_class, int64_t n_runs,
                         int64_t* row_s, int64_t* row_e, int64_t cap, int64_t* n_rows);
int  gl_fasta_load(gl_ctx* ctx, const uint8_t* bytes, int64_t n);
int  gl_fasta_stats(gl_ctx* ctx, const int64_t* byte_start, const int64_t* byte_end, int64_t n_rows, int64_t* counts4, double* stats3);
int  gl_depth_format_chunk_stats(const char* chrom, int64_t rs, int64_t re, int32_t W, const int64_t* win_sum, int64_t n_windows,
                                 const int32_t* run_start, const uint8_t* run_class, int64_t n_runs,
                                 const double* stats3, int64_t n_stat_rows,
                                 char** depth_bed, int64_t* depth_len, char** callable_bed, int64_t* callable_len);

/* ---------------------------------------------------------------- feeder (host-only, no GPU needed)
 * BGZF (multi-threaded inflate) + BAM decode + the `samtools depth` record filter: records with
 * (flag & 0x704) == 0 and MAPQ >= min_mapq contribute their M/=/X blocks (D and N advance without
 * counting; blocks separated only by I/S/H/P are merged), per reference, in record order.
 * Replaces the decode half of the `samtools depth` child (depth/depth.go:45).  only_tid < 0: all refs. */
typedef struct gl_segset gl_segset;
int  gl_bam_decode_segments(const char* path, int32_t min_mapq, int32_t threads, int32_t only_tid, gl_segset** out,
                            char* err, int64_t err_cap);
int  gl_segset_n_refs(const gl_segset* s, int32_t* n_refs, int64_t* n_records, int64_t* n_pass);
int  gl_segset_ref(const gl_segset* s, int32_t tid, const char** name, int64_t* length, const int32_t** start,
                   const int32_t** end, int64_t* n);
void gl_segset_free(gl_segset* s);
/* Index-guided parallel feeder: what `samtools depth -r chr:b-e` gives the reference (depth/depth.go:116,152) — only
 * the BGZF blocks of the requested reference range are read (BAI linear index -> record-aligned virtual offsets); the
 * range is cut into units at those offsets and every host-pool thread inflates and parses its own units.
 * gl_bam_open also loads path + ".bai" (or .bam -> .bai); decoding needs that index.
 * gl_bam_decode: the M/=/X blocks of the records of `tid` that pass the filter and can overlap [beg,end), as
 *   format 8  : packed8 words  (a0 = int32 anchors[n], a1 = uint8 dstart[64 n], a2 = uint8 len[64 n]; n blocks), or
 *   format 32 : (a0 = int32 start[n], a1 = int32 end[n]) sorted by start,
 * want_format 0 chooses (packed8 unless the blocks average more than 400 bases: long reads).  The arrays belong to the
 * handle and stay valid until its next gl_bam_decode / gl_bam_close; one handle per decoding thread. */
typedef struct {
    int32_t format, units, max_len, _pad;
    int64_t n;
    const int32_t* a0; const void* a1; const void* a2;
    int64_t n_records, n_pass, bytes_in, bytes_out;      /* records seen / passing; compressed bytes read / bytes inflated */
    double inflate_s, parse_s, wall_s;                   /* summed over the workers; wall clock of the call */
} gl_bam_segments;
int  gl_bam_open(const char* path, gl_bam** out, char* err, int64_t err_cap);
void gl_bam_close(gl_bam* b);
int  gl_bam_info(const gl_bam* b, int32_t* n_refs, int32_t* has_index);
int  gl_bam_ref(const gl_bam* b, int32_t tid, const char** name, int64_t* length, int64_t* n_mapped /* -1: no stats bin */);
int  gl_bam_decode(gl_bam* b, int32_t tid, int64_t beg, int64_t end, int32_t min_mapq, int32_t threads, int32_t want_format,
                   gl_bam_segments* out, char* err, int64_t err_cap);
/* The whole feeder on the GPU for one reference of an indexed BAM: its compressed BGZF range is read into pinned memory and
 * uploaded, every member is inflated by a warp (gl_bgzf_inflate_device), the records between consecutive linear-index
 * offsets are walked by one thread each (filter + CIGAR, as gl_bam_decode), and the M/=/X blocks land in ctx-owned device
 * arrays *d_start / *d_end (*n of them, BAM order; valid until the next call on this ctx) — feed them to
 * gl_depth_add_segments_device.  Only compressed bytes cross PCIe.  GL_ESTATE: not possible for this reference on the device
 * (no usable index, a member or record the device code rejects): use gl_bam_decode. */
int  gl_bam_decode_device(gl_ctx* ctx, gl_bam* b, int32_t tid, int32_t min_mapq, const int32_t** d_start, const int32_t** d_end, int64_t* n,
                          gl_bam_segments* stats);
/* BGZF inflate ON THE GPU (inflate.cu): one warp per BGZF member, thousands in flight — the stage the reference's samtools
 * children (and this library's host feeder) spend their time in.  Block b = d_comp[comp_off[b], comp_off[b+1]) (the whole
 * member) -> d_out[out_off[b], out_off[b+1]), out_off from the members' ISIZE fields; d_status[b] != 0: malformed or
 * unsupported member (fall back to zlib).  Device pointers; asynchronous on the ctx stream. */
int  gl_bgzf_inflate_device(gl_ctx* ctx, const uint8_t* d_comp, const int64_t* d_comp_off, const int64_t* d_out_off, int64_t n_blocks,
                            uint8_t* d_out, int32_t* d_status);
/* BAI: per-reference linear index (16 KB tiles) + the 0x924a stats bin (indexcov/types.go:19,45-58),
 * what indexcov takes from biogo's bam.ReadIndex (indexcov/indexcov.go:514). */
typedef struct gl_bai gl_bai;
int  gl_bai_read(const char* path, gl_bai** out, char* err, int64_t err_cap);
int  gl_bai_n_refs(const gl_bai* b, int32_t* n_refs, uint64_t* n_no_coor);
int  gl_bai_ref(const gl_bai* b, int32_t tid, const uint64_t** ioffsets, int64_t* n_intv, uint64_t* mapped,
                uint64_t* unmapped, int32_t* has_stats);
void gl_bai_free(gl_bai* b);

/* CRAM index: one reference's slices (alignment start, span, slice bytes) -> 16 KB pseudo-tile sizes
 * (indexcov/crai/crai.go:56-127).  GL_ERANGE where the reference panics, or when cap is too small
 * (*n_sizes then holds the needed count). */
int  gl_crai_make_sizes(const int64_t* aln_start, const int64_t* aln_span, const int32_t* slice_len, int64_t n,
                        int64_t* sizes, int64_t cap, int64_t* n_sizes);

/* ---------------------------------------------------------------- indexcov
 * Replaces indexcov/types.go:45-82 (getSizes), indexcov/indexcov.go:83-125 (Index.init),
 * :129-151 (NormalizedDepth), :170-177 (CountsAtDepth), :1050-1078 (counter.count),
 * :549-597 (normalizeAcrossSamples).  float32/float64 results are bit-exact (no FMA contraction,
 * same rounding points as Go on amd64). */

/* I1: tile sizes from BAI linear-index virtual offsets.  voff: concatenated per-ref linear
 * index entries of one sample; ref_ptr: CSR offsets (n_refs+1).  sizes gets, per ref with
 * >=2 intervals, n_intv-1 deltas; size_ptr (n_refs+1) their CSR offsets.  Returns GL_ERANGE
 * if any delta is negative (the reference panics, types.go:75-77). */
int  gl_indexcov_sizes(gl_ctx* ctx, const uint64_t* voff, const int64_t* ref_ptr, int32_t n_refs,
                       int64_t* sizes, int64_t* size_ptr);
/* I2: the capped weighted median of Index.init: sorted[first i with cumsum(min(s,n98))[i] > total/2]. */
int  gl_indexcov_scale(gl_ctx* ctx, const int64_t* sizes, int64_t n, int64_t* median_out);
/* I3: depth[i] = min(float32(float64(sizes[i])/median), 50000). */
int  gl_indexcov_normalize(gl_ctx* ctx, const int64_t* sizes, int64_t n, double median, float* depth_out);
/* I2+I3 for a cohort in one kernel (one CTA per sample, radix selects instead of a sort):
 * S samples, tile sizes in CSR layout (sample_ptr, S+1); medians[S]; depths in the same layout
 * (depth_out may be NULL).  A sample with median 0 gets zero depths (the reference returns none). */
int  gl_indexcov_cohort(gl_ctx* ctx, const int64_t* sizes, const int64_t* sample_ptr, int32_t S,
                        double* medians, float* depth_out);
int  gl_indexcov_cohort_device(gl_ctx* ctx, const int64_t* d_sizes, const int64_t* d_sample_ptr, int32_t S,
                               double* d_medians, float* d_depth_out);
/* how many samples of the last gl_indexcov_cohort* call needed the exact fallback select (heavy ties / overlapping
 * brackets) instead of the two-pass bracketed path; the results are identical either way */
int  gl_indexcov_cohort_fallbacks(gl_ctx* ctx, int32_t* n);
/* I1 for a whole cohort in ONE launch, on device-resident data: d_voff = every sample's linear-index virtual offsets
 * concatenated; descriptor d (one per sample x reference with >= 2 entries) = its n_intv[d] entries start at voff_off[d] and
 * its n_intv[d]-1 sizes go to size_off[d].  GL_ERANGE on a negative delta (types.go:75-77). */
int  gl_indexcov_sizes_batch_device(gl_ctx* ctx, const uint64_t* d_voff, const int64_t* d_voff_off, const int32_t* d_n_intv,
                                    const int64_t* d_size_off, int64_t n_desc, int64_t* d_sizes);
/* I4: 70-slot histogram of one depth array (counts += ..., like the reference's counts[...]++). */
int  gl_indexcov_counts(gl_ctx* ctx, const float* depth, int64_t n, int32_t counts[GL_INDEXCOV_SLOTS]);
/* I5: counter.count over depth[0..n) with `longest` tiles expected; out4 += {out, low, hi, in}. */
int  gl_indexcov_bins(gl_ctx* ctx, const float* depth, int64_t n, int64_t longest, int64_t out4[4]);
/* I4+I5 for many (sample, chromosome) segments at once: seg_ptr CSR (n_seg+1) into depth,
 * longest[n_seg] (may be NULL); counts70[n_seg*70] and bins4[n_seg*4] are overwritten. */
int  gl_indexcov_counts_batch(gl_ctx* ctx, const float* depth, const int64_t* seg_ptr, const int64_t* longest,
                              int32_t n_seg, int32_t* counts70, int64_t* bins4);
int  gl_indexcov_counts_batch_device(gl_ctx* ctx, const float* d_depth, const int64_t* d_seg_ptr,
                                     const int64_t* d_longest, int32_t n_seg, int32_t* d_counts70, int64_t* d_bins4);
/* the same for arbitrary slices [seg_start[k], seg_start[k] + seg_len[k]) of a device-resident depth array: every
 * (chromosome, sample) pair of a cohort in one launch, straight from gl_indexcov_cohort_device's output */
int  gl_indexcov_counts_segs_device(gl_ctx* ctx, const float* d_depth, const int64_t* d_seg_start, const int64_t* d_seg_len,
                                    const int64_t* d_longest, int32_t n_seg, int32_t* d_counts70, int64_t* d_bins4);
/* I6: the "%.3g" of every value (the bed.gz row writer, indexcov.go:678-680,1038-1048): 10-byte tokens,
 * bytes [0..len) = text, byte 9 = len; len 0 means the magnitude is outside [1e-15,1e15) and the host formats it. */
int  gl_format_g3(gl_ctx* ctx, const float* vals, int64_t n, uint8_t* tokens);
int  gl_format_g3_device(gl_ctx* ctx, const float* d_vals, int64_t n, uint8_t* d_tokens);
/* I7: in-place cross-sample normalisation of one chromosome; depths is S rows of stride T,
 * lens[i] valid entries in row i.  No-op for S < 5 (indexcov.go:551). */
int  gl_indexcov_xnorm(gl_ctx* ctx, float* depths, const int32_t* lens, int32_t S, int32_t T);

/* indexsplit (indexsplit/indexsplit.go:92-115): amount of cohort data per 16 KB tile = sum over the S samples, in path
 * order, of float64(size)/1e9 — bit-identical to the reference's sequential adds.  sizes: every sample's tile sizes
 * (gl_indexcov_sizes / gl_crai_make_sizes) concatenated; ptr[s*(R+1)+r] = offset of sample s's tiles of reference r
 * (a sample whose index has fewer references repeats its end offset); out_ptr[r] = prefix sum of max_s(len) per
 * reference; out[out_ptr[r] + j] = the tile's sum. */
int  gl_indexsplit_accumulate(gl_ctx* ctx, const int64_t* sizes, const int64_t* ptr, int32_t S, int32_t R, const int64_t* out_ptr,
                              double* out);
/* Host-only: the rest of indexsplit.Split (indexsplit.go:37-65,119-188): chop outliers, share N regions over the
 * references by their data, walk the tiles and cut.  tile_sum/out_ptr as filled by gl_indexsplit_accumulate (R index
 * references); the n_refs listed references (name, length, id into the index references or -1) are reported in order;
 * problematic intervals (-p) are (listed-reference number, start, end).  Returns the lines "chrom\tstart\tend\t%.2f\t%d\n"
 * in *text (release with gl_free_text). */
int  gl_indexsplit_chunks(const double* tile_sum, const int64_t* out_ptr, int32_t R, const char* const* ref_names, const int64_t* ref_lens,
                          const int32_t* ref_ids, int32_t n_refs, int32_t N, const int32_t* prob_ref, const int64_t* prob_start,
                          const int64_t* prob_end, int64_t n_prob, char** text, int64_t* text_len);

/* ---------------------------------------------------------------- covstats
 * V2: histogram of int32 values in [lo,hi) -> hist[v-lo].  `goleft covstats` builds it over the sampled insert sizes,
 * template lengths and read lengths and reads every order statistic of covstats/covstats.go:57-76,175-199 off its running
 * count (k-th smallest = a search, madFilter's cut = a count) and feeds meanStd (:78-89) the values in ascending order
 * with multiplicity — no host sort (cli/goleft.cpp SortedCounts). */
int  gl_bincount_i32(gl_ctx* ctx, const int32_t* v, int64_t n, int32_t lo, int32_t hi, uint64_t* hist);

/* ---------------------------------------------------------------- depthwed
 * W1: depthwed/depthwed.go:93-157.  means: S samples x R rows (sample-major) of the float64 4th BED
 * column; starts/ends/chrom_id[R] of the rows.  Computes int(0.5+mean) and sums consecutive rows of
 * one chrom until end-start >= size.  out: n_out rows x S (row-major by OUTPUT ROW, like the
 * reference's TSV lines). */
int  gl_depthwed_aggregate(gl_ctx* ctx, const double* means, int32_t S, int64_t R,
                           const int32_t* starts, const int32_t* ends, const int32_t* chrom_id, int64_t size,
                           int32_t* out_start, int32_t* out_end, int32_t* out_chrom, int64_t* out, int64_t out_cap,
                           int64_t* n_out);
/* device form: d_grp[n_out+1] row groups (NULL: one row per output line); d_out n_out x S */
int  gl_depthwed_aggregate_device(gl_ctx* ctx, const double* d_means, int32_t S, int64_t R, const int64_t* d_grp,
                                  int64_t n_out, int64_t* d_out);

/* int32 forms — what BASELINE budgets (4 B in + 4 B out per cell): the reference rounds when it parses a line
 * (depthwed.go:103 d.depth = int(0.5 + dep)), so the caller hands over int32 depths and the matrix that travels over NVLink is
 * int32.  Group sums are taken in 64 bits; one that does not fit int32 makes the host form return GL_ERANGE with *n_out = -1
 * (rerun with gl_depthwed_aggregate).  The device form does the groups [g_begin, g_end) only (so the all-gather of one row
 * chunk can overlap the aggregation of the next), is ASYNCHRONOUS on the ctx stream, and sets *d_overflow (a device int the
 * caller zeroed) instead. */
int  gl_depthwed_aggregate_i32(gl_ctx* ctx, const int32_t* depth, int32_t S, int64_t R, const int32_t* starts, const int32_t* ends,
                               const int32_t* chrom_id, int64_t size, int32_t* out_start, int32_t* out_end, int32_t* out_chrom,
                               int32_t* out, int64_t out_cap, int64_t* n_out);
int  gl_depthwed_aggregate_i32_device(gl_ctx* ctx, const int32_t* d_depth, int32_t S, int64_t R, const int64_t* d_grp, int64_t g_begin,
                                      int64_t g_end, int32_t* d_out, int32_t* d_overflow);

/* W1 fused with its collective: the aggregation kernel stores every output row straight into the row-major n-sites x
 * n-samples matrix of EVERY GPU of the box (its own + NVLink peer mappings of the others): d_dst[d] = base of rank d's matrix
 * as a pointer valid on this ctx's device (gl_ipc_open between processes; plain peer pointers inside one process), row stride
 * `row_stride` ints, this rank's samples at columns [col_off, col_off + S).  No separate all-gather, no re-assembly.
 * ASYNCHRONOUS on the ctx stream; after gl_sync on every rank and a host barrier the matrix is complete everywhere. */
int  gl_depthwed_aggregate_i32_p2p(gl_ctx* ctx, const int32_t* d_depth, int32_t S, int64_t R, const int64_t* d_grp, int64_t g_begin,
                                   int64_t g_end, int32_t* const* d_dst, int32_t world, int64_t row_stride, int32_t col_off,
                                   int32_t* d_overflow);
/* peer memory between the per-GPU processes of one box: export a gl_dev_alloc allocation as a 64-byte handle, open another
 * rank's handle as a device pointer usable by kernels of this ctx (cudaIpc*; the link is NVLink / NVSwitch) */
int  gl_ipc_export(gl_ctx* ctx, void* d_ptr, uint8_t handle64[64]);
int  gl_ipc_open(gl_ctx* ctx, const uint8_t handle64[64], void** d_peer_ptr);
int  gl_ipc_close(gl_ctx* ctx, void* d_peer_ptr);

/* ------------------------------------------------------------- multi-GPU
 * One process per GPU.  The caller distributes a 128-byte NCCL unique id (rank 0 creates it).
 * The one collective on this path: an all-gather over NVLink that assembles the depthwed
 * n-sites x n-samples matrix (or the indexcov cohort) from per-GPU sample shards. */
int  gl_comm_unique_id(uint8_t id128[128]);
int  gl_comm_init(gl_ctx* ctx, const uint8_t id128[128], int rank, int world);
int  gl_comm_destroy(gl_ctx* ctx);
/* all-gather equal-sized blocks that live on the device: d_recv holds world*bytes */
int  gl_allgather_device(gl_ctx* ctx, const void* d_send, void* d_recv, int64_t bytes);
/* asynchronous on the ctx's communication stream: starts when the work queued on the compute stream so far is done, runs
 * beside later kernels; gl_comm_wait() joins both streams */
int  gl_allgather_device_async(gl_ctx* ctx, const void* d_send, void* d_recv, int64_t bytes);
int  gl_comm_wait(gl_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* GOLEFT_B200_H */
