// gl_common.cuh — context object and error plumbing shared by every translation unit of
// libgoleft_b200.so.  sm_100a only; there is no CPU fallback anywhere in this library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <algorithm>
#include "../../include/goleft_b200.h"

struct gl_buf {            // grow-only device buffer
    void* p = nullptr;
    size_t cap = 0;
};

struct gl_seg_batch {      // one gl_depth_add_segments* call
    const int32_t* s = nullptr;   // device pointers (caller-owned), or null when the batch lives in the ctx store
    const int32_t* e = nullptr;
    int64_t off = 0;              // offset into the ctx store when s == nullptr
    int64_t n = 0;
};

struct gl_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;       // compute stream
    cudaStream_t copy_stream = nullptr;  // H2D staging stream
    cudaStream_t d2h_stream = nullptr;   // streamed packed8 reduce: finished window sums leave while later chunks arrive
    std::string err;
    int64_t launches = 0;
    int sm_count = 148;

    cudaEvent_t ev0 = nullptr, ev1 = nullptr;         // gl_timer_*
    cudaEvent_t ev_copy[2] = {nullptr, nullptr};      // staging ring
    cudaEvent_t ev_used[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> ev_chunk;                // streamed packed8 upload: one per chunk

    // ---- depth state
    bool depth_active = false, depth_reduced = false;
    int64_t rs = 0, re = 0;
    int32_t red_W = 0, red_mincov = 0, red_maxmean = 0; int64_t red_break = 0;
    int64_t n_windows = 0, n_runs = 0; int32_t max_depth = 0;
    std::vector<gl_seg_batch> batches;      // segments added since gl_depth_begin
    gl_buf store_s, store_e;                // host-uploaded segments, contiguous
    gl_buf packed;                          // staging for packed16 uploads
    int64_t store_n = 0;
    bool copies_pending = false;            // H2D into the store still in flight on copy_stream
    bool g_valid = false;                   // the HBM difference array holds every batch (general path)
    bool ev_valid = false;                  // the per-tile event buckets hold every batch (general path, version 2)
    int last_path = 0;                      // 1 = fused sorted path, 2 = general scatter path
    // index tables of the last fused reduce (valid until the next reduce): used by gl_depth_interval_sums
    const int* idx_flags = nullptr; const unsigned* idx_cells = nullptr; int idx_origin = 0, idx_ncells = 0, idx_maxlen = -1;
    int force_path = 0;                     // 0 = auto, 1 = no packed8 kernel (int32 fused, else general), 2 = always general
    // a packed8 batch that is the region's only batch stays packed: K_fused8 reads it as it is (last_path 3); it is
    // unpacked into the store (its batch descriptor already points there) only when something needs int32 arrays
    struct { const int* anchors = nullptr; const void* ds = nullptr; const void* len = nullptr; int64_t n_blocks = 0, store_off = 0;
             bool pending = false;
             // upload deferred to the reduce (one-call entry with pinned host words): streamed in chunks, each chunk's
             // tiles reduced while the next chunk is on the wire
             const int* h_anchors = nullptr; const void* h_ds = nullptr; const void* h_len = nullptr; bool upload_pending = false; } p8;
    gl_buf diff;          // int32[len+1 (+pad)] + tile sums (general path only)
    void* win_sum_p = nullptr;   // u64[n_windows], inside `scratch`
    int64_t* prefetch_sums = nullptr; int64_t prefetch_cap = 0; bool prefetched = false;   // one-call entries: window sums go home while the header is read
    int32_t* prefetch_run_start = nullptr; uint8_t* prefetch_run_class = nullptr; int64_t prefetch_run_cap = 0, prefetched_runs = 0;   // and the first runs
    gl_buf win_min;       // i32[n_windows]
    gl_buf run_start;     // i32[run_cap]
    gl_buf run_class;     // u8[run_cap]
    gl_buf run_tmp_start; // claim-order staging of the runs
    gl_buf run_tmp_class;
    gl_buf scratch;       // header | window sums | index flags + cell tables | per-tile run table
    gl_buf seg[2];        // device staging for host segments: int32 start|end
    void* pinned[2] = {nullptr, nullptr};
    size_t pinned_bytes = 0;
    void* pack_pinned = nullptr; size_t pack_pinned_bytes = 0;   // gl_depth_bed_contig: pinned home of the packed8 words it makes
    bool cohort_attr_set = false;                                // ic_cohort2_kernel's dynamic shared-memory limit raised on this ctx's device
    double tr_phase[3] = {0, 0, 0};
    int tr_kind = 0; double tr_pack_s = 0; int64_t tr_bytes = 0, tr_esc = 0;   // gl_depth_transport_stats
    int32_t* esc_buf = nullptr; size_t esc_cap = 0;              // ... and of the raw segments of blocks that do not fit packed16 (2 x esc_cap ints, malloc)
    gl_buf flush;         // L2 flush scratch
    gl_buf misc;          // small outputs of other subsystems
    gl_buf bam_comp, bam_out, bam_seg;   // bamgpu.cu: compressed bytes + tables, inflated stream, decoded segments
    gl_buf text_slots, text_blk, text_len, text_out[2];   // depth_text.cu: row slots, block offsets, lengths, compact BED text (depth, callable)
    gl_buf fasta;         // raw FASTA bytes of one record (gl_fasta_load), newlines included
    int64_t fasta_n = 0;

    // per-kernel CUDA-event profiling (gl_profile_*): (name, start, stop) per launch while enabled
    bool prof_on = false;
    struct prof_rec { const char* name; cudaEvent_t a, b; };
    std::vector<prof_rec> prof;
    std::vector<cudaEvent_t> prof_pool;

    unsigned cohort_fallbacks = 0;   // ic_cohort2_kernel: samples that took the exact fallback select in the last call
    cudaStream_t comm_stream = nullptr; cudaEvent_t ev_comm = nullptr;   // gl_allgather_device_async
    void* nccl = nullptr; // ncclComm_t when gl_comm_init was called
    int rank = 0, world = 1;
};

extern thread_local std::string g_gl_err;

int gl_fail(gl_ctx* ctx, int code, const char* fmt, ...);
int gl_buf_reserve(gl_ctx* ctx, gl_buf& b, size_t bytes);

#define GL_CUDA(ctx, expr)                                                             \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess)                                                         \
            return gl_fail((ctx), GL_ECUDA, "%s failed: %s (%s:%d)", #expr,            \
                           cudaGetErrorString(_e), __FILE__, __LINE__);                \
    } while (0)

#define GL_CHECK(expr)                 \
    do {                               \
        int _rc = (expr);              \
        if (_rc != GL_OK) return _rc;  \
    } while (0)

#define GL_LAUNCHED(ctx, n)                                                            \
    do {                                                                               \
        (ctx)->launches += (n);                                                        \
        GL_CUDA((ctx), cudaGetLastError());                                            \
    } while (0)

// Bracket one kernel launch with events on the ctx stream when profiling is on.
struct gl_prof_scope {
    gl_ctx* ctx; cudaEvent_t b = nullptr;
    gl_prof_scope(gl_ctx* c, const char* name) : ctx(c) {
        if (!c->prof_on) return;
        cudaEvent_t ev[2];
        for (int i = 0; i < 2; i++) {
            if (!c->prof_pool.empty()) { ev[i] = c->prof_pool.back(); c->prof_pool.pop_back(); }
            else cudaEventCreate(&ev[i]);
        }
        cudaEventRecord(ev[0], c->stream);
        b = ev[1];
        c->prof.push_back({name, ev[0], ev[1]});
    }
    ~gl_prof_scope() { if (b) cudaEventRecord(b, ctx->stream); }
};

static inline int gl_use(gl_ctx* ctx) {
    if (!ctx) return gl_fail(nullptr, GL_EINVAL, "null ctx");
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return gl_fail(ctx, GL_ECUDA, "cudaSetDevice(%d): %s", ctx->device, cudaGetErrorString(e));
    return GL_OK;
}
