// depth.cu — the `goleft depth` hot path on sm_100a.
//
//   segments (start,end) --scatter (int32 red.global)--> difference array in HBM
//   difference array --ONE fused pass--> per-base depth (registers/smem only)
//                                        -> per-window int64 sum + int32 min   (depth/depth.go:293-306)
//                                        -> coverage-class run starts          (depth/depth.go:307-327)
//
// The fused pass is a single-pass chained scan (decoupled look-back over 4096-base tiles handed
// out by an atomic ticket), so the difference array is read exactly once and per-base depth never
// goes back to HBM.  A second look-back chain orders the variable-length run output.
// All arithmetic is integer; results are bit-exact against oracle/oracle_depth.c.
#include "gl_common.cuh"
#include <string.h>

namespace {

constexpr int kScanThreads = 256;
constexpr int kWarps = kScanThreads / 32;
constexpr int kRounds = 4;                        // int4 per lane per round
constexpr int kWarpElems = 32 * 4 * kRounds;      // 512 bases per warp
constexpr int kTile = kWarps * kWarpElems;        // 4096 bases per tile
constexpr int kHeaderWords = 8;                   // u64 words in front of the status arrays
constexpr unsigned kFull = 0xffffffffu;

constexpr uint64_t kFlagAgg = 1, kFlagPrefix = 2;

__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ int4 ld_stream_int4(const int4* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// ------------------------------------------------------------------------------------------------
// K1: scatter.  One thread = 4 segments (two 128-bit loads), 8 fire-and-forget int32 reductions.
// Coordinates are clipped to the region exactly as `samtools depth -r` clips its output
// (depth/depth.go:150-152): a read spanning a chunk edge counts on both sides.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void scatter_one(int s, int e, int rs, int re, int* __restrict__ diff) {
    s = max(s, rs);
    e = min(e, re);
    if (s < e) {
        atomicAdd(diff + (s - rs), 1);     // result unused -> RED.E.ADD
        atomicAdd(diff + (e - rs), -1);
    }
}

template <bool kVec>
__global__ void __launch_bounds__(256) depth_scatter_kernel(const int* __restrict__ start, const int* __restrict__ end,
                                                            long long n, int rs, int re, int* __restrict__ diff) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (kVec) {
        long long n4 = n >> 2;
        if (i < n4) {
            int4 s = ld_stream_int4(reinterpret_cast<const int4*>(start) + i);
            int4 e = ld_stream_int4(reinterpret_cast<const int4*>(end) + i);
            scatter_one(s.x, e.x, rs, re, diff);
            scatter_one(s.y, e.y, rs, re, diff);
            scatter_one(s.z, e.z, rs, re, diff);
            scatter_one(s.w, e.w, rs, re, diff);
        } else {
            long long k = (n4 << 2) + (i - n4);
            if (k < n) scatter_one(start[k], end[k], rs, re, diff);
        }
    } else {
        if (i < n) scatter_one(start[i], end[i], rs, re, diff);
    }
}

// ------------------------------------------------------------------------------------------------
// K2: fused scan + window reduce + class runs.
// ------------------------------------------------------------------------------------------------
struct ScanParams {
    const int* diff;              // padded to a whole number of tiles, zero beyond len
    int len;                      // bases in the region
    int rs;                       // absolute start of the region
    int W;                        // window size
    long long w0;                 // rs / W  (index of the first window)
    int mincov, maxmean;
    unsigned run_break;           // 0 = never (host clamps values >= 2^32 to 0: no multiple in range)
    unsigned long long* win_sum;  // [n_windows], zero-initialised
    int* win_min;                 // [n_windows], initialised to 0x7f7f7f7f (or null)
    int* run_start;               // [run_cap]
    unsigned char* run_class;     // [run_cap]
    long long run_cap;
    uint64_t* header;             // [0]=ticket [1]=n_runs [2]=max_depth
    uint64_t* status_depth;       // [num_tiles]
    uint64_t* status_runs;        // [num_tiles]
    int* depth_out;               // optional per-base output (debug/parity), else null
    int num_tiles;
    int do_windows, do_runs;
};

__device__ __forceinline__ int cov_class(int d, int mincov, int maxmean) {
    // depth/depth.go:223-234
    return d == 0 ? GL_NO_COVERAGE
                  : (d < mincov ? GL_LOW_COVERAGE : ((maxmean > 0 && d >= maxmean) ? GL_EXCESSIVE_COVERAGE : GL_CALLABLE));
}

// Decoupled look-back run by one full warp: returns the sum of the aggregates of all tiles < tile.
__device__ __forceinline__ int lookback(const uint64_t* status, int tile, int lane) {
    int excl = 0;
    int look = tile - 1;
    while (true) {
        int idx = look - lane;
        uint64_t w = (kFlagPrefix << 32);                // virtual tile -1: prefix 0
        if (idx >= 0) {
            do { w = ld_relaxed_u64(status + idx); } while ((w >> 32) == 0);
        }
        unsigned flag = (unsigned)(w >> 32);
        int val = (int)(unsigned)w;
        unsigned pm = __ballot_sync(kFull, flag == (unsigned)kFlagPrefix);
        int take = val;
        if (pm) {
            int first = __ffs(pm) - 1;
            take = lane <= first ? val : 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) take += __shfl_xor_sync(kFull, take, o);
        excl += take;
        if (pm) break;
        look -= 32;
    }
    return excl;
}

__global__ void __launch_bounds__(kScanThreads, 4) depth_scan_kernel(const ScanParams p) {
    __shared__ __align__(16) int s_depth[kTile];
    __shared__ int s_warp_tot[kWarps];
    __shared__ int s_warp_cnt[kWarps];
    __shared__ int s_tile, s_tile_excl, s_run_base;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) s_tile = (int)atomicAdd(reinterpret_cast<unsigned long long*>(p.header), 1ull);
    __syncthreads();
    const int tile = s_tile;
    const int tile_base = tile * kTile;                       // relative position of the tile
    const int warp_base = tile_base + warp * kWarpElems;

    // ---- load 16 diffs per thread, warp-coalesced 512 B per instruction
    int4 v[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; r++)
        v[r] = ld_stream_int4(reinterpret_cast<const int4*>(p.diff + warp_base + r * 128) + lane);

    // ---- warp-local exclusive prefix of each quad
    int pre[kRounds];
    int running = 0;
#pragma unroll
    for (int r = 0; r < kRounds; r++) {
        int q = v[r].x + v[r].y + v[r].z + v[r].w;
        int inc = q;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += t;
        }
        pre[r] = running + inc - q;
        running += __shfl_sync(kFull, inc, 31);
    }
    if (lane == 0) s_warp_tot[warp] = running;
    __syncthreads();

    // ---- chain 1: depth carried into the tile
    if (warp == 0) {
        int agg = 0;
#pragma unroll
        for (int w = 0; w < kWarps; w++) agg += s_warp_tot[w];
        if (tile == 0) {
            if (lane == 0) { st_relaxed_u64(p.status_depth, (kFlagPrefix << 32) | (unsigned)agg); s_tile_excl = 0; }
        } else {
            if (lane == 0) st_relaxed_u64(p.status_depth + tile, (kFlagAgg << 32) | (unsigned)agg);
            int excl = lookback(p.status_depth, tile, lane);
            if (lane == 0) {
                st_relaxed_u64(p.status_depth + tile, (kFlagPrefix << 32) | (unsigned)(excl + agg));
                s_tile_excl = excl;
            }
        }
    }
    __syncthreads();

    int wbase = s_tile_excl;
#pragma unroll
    for (int w = 0; w < kWarps; w++) wbase += (w < warp) ? s_warp_tot[w] : 0;

    // does a forced run break (multiple of run_break) fall inside this tile?
    bool tile_has_break = false;
    if (p.run_break > 0) {
        const unsigned a = (unsigned)p.rs + (unsigned)tile_base;      // absolute positions are < 2^32
        const unsigned rem = a % p.run_break;
        tile_has_break = (rem ? p.run_break - rem : 0u) < (unsigned)kTile;
    }

    // ---- per-base depth -> smem, class-change masks, max
    unsigned masks = 0;
    int maxd = 0;
    int* sw = s_depth + warp * kWarpElems;
#pragma unroll
    for (int r = 0; r < kRounds; r++) {
        const int idx = warp_base + r * 128 + lane * 4;       // relative position of element 0
        const int prev = wbase + pre[r];
        int4 d;
        d.x = prev + v[r].x;
        d.y = d.x + v[r].y;
        d.z = d.y + v[r].z;
        d.w = d.z + v[r].w;
        reinterpret_cast<int4*>(sw + r * 128)[lane] = d;
        if (p.depth_out) {
            if (idx + 3 < p.len) reinterpret_cast<int4*>(p.depth_out + idx)[0] = d;
            else {
                if (idx < p.len) p.depth_out[idx] = d.x;
                if (idx + 1 < p.len) p.depth_out[idx + 1] = d.y;
                if (idx + 2 < p.len) p.depth_out[idx + 2] = d.z;
            }
        }
        if (idx < p.len) {
            int nv = min(4, p.len - idx);
            maxd = max(maxd, d.x);
            if (nv > 1) maxd = max(maxd, d.y);
            if (nv > 2) maxd = max(maxd, d.z);
            if (nv > 3) maxd = max(maxd, d.w);
            if (p.do_runs) {
                int cp = cov_class(prev, p.mincov, p.maxmean);
                int c0 = cov_class(d.x, p.mincov, p.maxmean);
                int c1 = cov_class(d.y, p.mincov, p.maxmean);
                int c2 = cov_class(d.z, p.mincov, p.maxmean);
                int c3 = cov_class(d.w, p.mincov, p.maxmean);
                unsigned m = (c0 != cp ? 1u : 0u) | (c1 != c0 ? 2u : 0u) | (c2 != c1 ? 4u : 0u) | (c3 != c2 ? 8u : 0u);
                if (idx == 0) m |= 1u;                        // the region's first base always starts a run
                if (tile_has_break) {
                    const unsigned a = (unsigned)p.rs + (unsigned)idx;
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if ((a + j) % p.run_break == 0) m |= (1u << j);
                }
                m &= (1u << nv) - 1u;
                masks |= m << (4 * r);
            }
        }
    }
    __syncwarp();

    // ---- rank of each run start inside the warp, in position order (round, lane, j)
    int lane_rank[kRounds];
    int round_tot[kRounds];
    int warp_cnt = 0;
    if (p.do_runs) {
        const unsigned lt = (1u << lane) - 1u;
#pragma unroll
        for (int r = 0; r < kRounds; r++) {
            unsigned m = (masks >> (4 * r)) & 15u;
            lane_rank[r] = 0;
            round_tot[r] = 0;
            if (__any_sync(kFull, m != 0)) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    unsigned b = __ballot_sync(kFull, (m >> j) & 1u);
                    lane_rank[r] += __popc(b & lt);
                    round_tot[r] += __popc(b);
                }
            }
            warp_cnt += round_tot[r];
        }
        if (lane == 0) s_warp_cnt[warp] = warp_cnt;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxd = max(maxd, __shfl_xor_sync(kFull, maxd, o));
    if (lane == 0 && maxd > 0) {
        int cur = (int)ld_relaxed_u64(p.header + 2);
        if (maxd > cur) atomicMax(reinterpret_cast<unsigned long long*>(p.header + 2), (unsigned long long)maxd);
    }
    __syncthreads();

    // ---- chain 2: number of runs that start before the tile
    if (p.do_runs && warp == 0) {
        int agg = 0;
#pragma unroll
        for (int w = 0; w < kWarps; w++) agg += s_warp_cnt[w];
        int excl = 0;
        if (tile == 0) {
            if (lane == 0) st_relaxed_u64(p.status_runs, (kFlagPrefix << 32) | (unsigned)agg);
        } else {
            if (lane == 0) st_relaxed_u64(p.status_runs + tile, (kFlagAgg << 32) | (unsigned)agg);
            excl = lookback(p.status_runs, tile, lane);
            if (lane == 0) st_relaxed_u64(p.status_runs + tile, (kFlagPrefix << 32) | (unsigned)(excl + agg));
        }
        if (lane == 0) {
            s_run_base = excl;
            if (tile == p.num_tiles - 1) st_relaxed_u64(p.header + 1, (uint64_t)(unsigned)(excl + agg));
        }
    }

    // ---- window partial sums of this warp's 512 bases (reads only this warp's smem)
    if (p.do_windows) {
        const unsigned a0 = (unsigned)p.rs + (unsigned)warp_base;                        // absolute, < 2^32
        const unsigned a1 = (unsigned)p.rs + (unsigned)min(warp_base + kWarpElems, p.len);
        if (a0 < a1) {
            const unsigned uW = (unsigned)p.W;
            const unsigned iw0 = a0 / uW, iw1 = (a1 - 1) / uW;
            if (p.W > 16) {
                for (unsigned iw = iw0; iw <= iw1; iw++) {
                    const long long ws = (long long)iw * uW, we = ws + uW;
                    const int s = (int)(max((long long)a0, ws) - a0), e = (int)(min((long long)a1, we) - a0);
                    unsigned long long sum = 0;
                    int mn = 0x7fffffff;
                    for (int x = s + lane; x < e; x += 32) {
                        int d = sw[x];
                        sum += (unsigned long long)(unsigned)d;
                        mn = min(mn, d);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        sum += __shfl_xor_sync(kFull, sum, o);
                        mn = min(mn, __shfl_xor_sync(kFull, mn, o));
                    }
                    if (lane == 0) {
                        atomicAdd(p.win_sum + ((long long)iw - p.w0), sum);
                        if (p.win_min) atomicMin(p.win_min + ((long long)iw - p.w0), mn);
                    }
                }
            } else {
                for (unsigned iw = iw0 + lane; iw <= iw1; iw += 32) {
                    const long long ws = (long long)iw * uW, we = ws + uW;
                    const int s = (int)(max((long long)a0, ws) - a0), e = (int)(min((long long)a1, we) - a0);
                    unsigned long long sum = 0;
                    int mn = 0x7fffffff;
                    for (int x = s; x < e; x++) {
                        int d = sw[x];
                        sum += (unsigned long long)(unsigned)d;
                        mn = min(mn, d);
                    }
                    atomicAdd(p.win_sum + ((long long)iw - p.w0), sum);
                    if (p.win_min) atomicMin(p.win_min + ((long long)iw - p.w0), mn);
                }
            }
        }
    }
    if (!p.do_runs) return;
    __syncthreads();

    // ---- emit run starts in position order
    if (masks) {
        long long base = s_run_base;
#pragma unroll
        for (int w = 0; w < kWarps; w++) base += (w < warp) ? s_warp_cnt[w] : 0;
#pragma unroll
        for (int r = 0; r < kRounds; r++) {
            unsigned m = (masks >> (4 * r)) & 15u;
            long long rank = base + lane_rank[r];
            const int off = r * 128 + lane * 4;
            while (m) {
                int j = __ffs(m) - 1;
                m &= m - 1;
                if (rank < p.run_cap) {
                    p.run_start[rank] = p.rs + warp_base + off + j;
                    p.run_class[rank] = (unsigned char)cov_class(sw[off + j], p.mincov, p.maxmean);
                }
                rank++;
            }
            base += round_tot[r];
        }
    }
}

__global__ void fill_i32_kernel(int* p, long long n, int v) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void run_ends_kernel(const int* run_start, int* run_end, long long n, int re) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) run_end[i] = (i + 1 < n) ? run_start[i + 1] : re;
}

inline int64_t num_tiles_for(int64_t len) { return (len + kTile - 1) / kTile; }

int launch_scatter(gl_ctx* ctx, const int32_t* d_start, const int32_t* d_end, int64_t n) {
    if (n <= 0) return GL_OK;
    bool vec = ((reinterpret_cast<uintptr_t>(d_start) | reinterpret_cast<uintptr_t>(d_end)) & 15) == 0;
    int* diff = static_cast<int*>(ctx->diff.p);
    if (vec) {
        long long threads = (n >> 2) + (n & 3);
        unsigned grid = (unsigned)((threads + 255) / 256);
        depth_scatter_kernel<true><<<grid, 256, 0, ctx->stream>>>(d_start, d_end, n, (int)ctx->rs, (int)ctx->re, diff);
    } else {
        unsigned grid = (unsigned)((n + 255) / 256);
        depth_scatter_kernel<false><<<grid, 256, 0, ctx->stream>>>(d_start, d_end, n, (int)ctx->rs, (int)ctx->re, diff);
    }
    GL_LAUNCHED(ctx, 1);
    return GL_OK;
}

// runs the fused pass; on run-capacity overflow grows the buffers and runs again
int run_reduce(gl_ctx* ctx, int32_t W, int32_t mincov, int32_t maxmean, int64_t run_break, bool do_windows,
               bool do_runs, int32_t* d_depth_out, bool want_min) {
    const int64_t len = ctx->re - ctx->rs;
    const int64_t tiles = num_tiles_for(len);
    const int64_t w0 = ctx->rs / W;
    const int64_t n_windows = (ctx->re - 1) / W - w0 + 1;

    if (do_windows) {
        GL_CHECK(gl_buf_reserve(ctx, ctx->win_sum, (size_t)n_windows * 8));
        GL_CHECK(gl_buf_reserve(ctx, ctx->win_min, (size_t)n_windows * 4));
    }
    GL_CHECK(gl_buf_reserve(ctx, ctx->scratch, (size_t)(kHeaderWords + 2 * tiles) * 8));
    if (do_runs && ctx->run_start.cap == 0) {
        size_t cap = (size_t)(len / 16 + 4096);
        GL_CHECK(gl_buf_reserve(ctx, ctx->run_start, cap * 4));
        GL_CHECK(gl_buf_reserve(ctx, ctx->run_class, cap));
    }

    for (int attempt = 0; attempt < 2; attempt++) {
        ScanParams p;
        memset(&p, 0, sizeof p);
        p.diff = static_cast<const int*>(ctx->diff.p);
        p.len = (int)len;
        p.rs = (int)ctx->rs;
        p.W = W;
        p.w0 = w0;
        p.mincov = mincov;
        p.maxmean = maxmean;
        p.run_break = (run_break >= (int64_t(1) << 32)) ? 0u : (unsigned)run_break;
        p.win_sum = static_cast<unsigned long long*>(ctx->win_sum.p);
        p.win_min = want_min ? static_cast<int*>(ctx->win_min.p) : nullptr;
        p.run_start = static_cast<int*>(ctx->run_start.p);
        p.run_class = static_cast<unsigned char*>(ctx->run_class.p);
        p.run_cap = do_runs ? (long long)ctx->run_class.cap : 0;
        if (do_runs && (long long)(ctx->run_start.cap / 4) < p.run_cap) p.run_cap = (long long)(ctx->run_start.cap / 4);
        p.header = static_cast<uint64_t*>(ctx->scratch.p);
        p.status_depth = p.header + kHeaderWords;
        p.status_runs = p.status_depth + tiles;
        p.depth_out = d_depth_out;
        p.num_tiles = (int)tiles;
        p.do_windows = do_windows ? 1 : 0;
        p.do_runs = do_runs ? 1 : 0;

        GL_CUDA(ctx, cudaMemsetAsync(ctx->scratch.p, 0, (size_t)(kHeaderWords + 2 * tiles) * 8, ctx->stream));
        if (do_windows) {
            GL_CUDA(ctx, cudaMemsetAsync(ctx->win_sum.p, 0, (size_t)n_windows * 8, ctx->stream));
            if (want_min) GL_CUDA(ctx, cudaMemsetAsync(ctx->win_min.p, 0x7f, (size_t)n_windows * 4, ctx->stream));
        }
        depth_scan_kernel<<<(unsigned)tiles, kScanThreads, 0, ctx->stream>>>(p);
        GL_LAUNCHED(ctx, 1);

        ctx->red_W = W; ctx->red_mincov = mincov; ctx->red_maxmean = maxmean; ctx->red_break = run_break;
        ctx->n_windows = do_windows ? n_windows : 0;
        ctx->n_runs = -1;            // fetched lazily
        ctx->max_depth = -1;
        if (!do_runs) { ctx->n_runs = 0; break; }
        if (attempt == 0) {
            // The run count is needed to know whether the output fit; peek at it only when
            // the capacity could plausibly be exceeded (cap < len), else defer the sync.
            if (p.run_cap >= len) break;
            uint64_t hdr[3];
            GL_CUDA(ctx, cudaMemcpyAsync(hdr, ctx->scratch.p, sizeof hdr, cudaMemcpyDeviceToHost, ctx->stream));
            GL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            ctx->n_runs = (int64_t)hdr[1];
            ctx->max_depth = (int32_t)hdr[2];
            if ((long long)hdr[1] <= p.run_cap) break;
            size_t cap = (size_t)hdr[1] + 1024;
            GL_CHECK(gl_buf_reserve(ctx, ctx->run_start, cap * 4));
            GL_CHECK(gl_buf_reserve(ctx, ctx->run_class, cap));
        }
    }
    ctx->depth_reduced = true;
    return GL_OK;
}

int fetch_header(gl_ctx* ctx) {
    if (ctx->n_runs >= 0 && ctx->max_depth >= 0) return GL_OK;
    uint64_t hdr[3];
    GL_CUDA(ctx, cudaMemcpyAsync(hdr, ctx->scratch.p, sizeof hdr, cudaMemcpyDeviceToHost, ctx->stream));
    GL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->n_runs < 0) ctx->n_runs = (int64_t)hdr[1];
    ctx->max_depth = (int32_t)hdr[2];
    return GL_OK;
}

constexpr int64_t kStageSegs = int64_t(1) << 22;      // segments per staging slot (2 x 16 MiB)

bool is_pinned_host(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

}  // namespace

extern "C" {

int gl_depth_begin(gl_ctx* ctx, int64_t region_start, int64_t region_end) {
    GL_CHECK(gl_use(ctx));
    if (region_start < 0 || region_end <= region_start || region_end > INT32_MAX)
        return gl_fail(ctx, GL_EINVAL, "gl_depth_begin: bad region [%lld,%lld)", (long long)region_start, (long long)region_end);
    const int64_t len = region_end - region_start;
    if (len >= (int64_t(1) << 30)) return gl_fail(ctx, GL_ERANGE, "gl_depth_begin: region longer than 2^30-1 bases");
    const size_t entries = (size_t)num_tiles_for(len) * kTile + 4;
    GL_CHECK(gl_buf_reserve(ctx, ctx->diff, entries * 4));
    GL_CUDA(ctx, cudaMemsetAsync(ctx->diff.p, 0, entries * 4, ctx->stream));
    ctx->rs = region_start;
    ctx->re = region_end;
    ctx->depth_active = true;
    ctx->depth_reduced = false;
    ctx->n_windows = 0;
    ctx->n_runs = 0;
    ctx->max_depth = 0;
    return GL_OK;
}

int gl_depth_add_segments_device(gl_ctx* ctx, const int32_t* d_start, const int32_t* d_end, int64_t n) {
    GL_CHECK(gl_use(ctx));
    if (!ctx->depth_active) return gl_fail(ctx, GL_ESTATE, "gl_depth_add_segments: no region open (call gl_depth_begin)");
    if (n < 0 || (n > 0 && (!d_start || !d_end))) return gl_fail(ctx, GL_EINVAL, "gl_depth_add_segments: bad argument");
    ctx->depth_reduced = false;
    return launch_scatter(ctx, d_start, d_end, n);
}

int gl_depth_add_segments(gl_ctx* ctx, const int32_t* start, const int32_t* end, int64_t n) {
    GL_CHECK(gl_use(ctx));
    if (!ctx->depth_active) return gl_fail(ctx, GL_ESTATE, "gl_depth_add_segments: no region open (call gl_depth_begin)");
    if (n < 0 || (n > 0 && (!start || !end))) return gl_fail(ctx, GL_EINVAL, "gl_depth_add_segments: bad argument");
    if (n == 0) return GL_OK;
    ctx->depth_reduced = false;
    const bool pinned = is_pinned_host(start) && is_pinned_host(end);
    const size_t slot_bytes = (size_t)kStageSegs * 8;
    for (int i = 0; i < 2; i++) {
        GL_CHECK(gl_buf_reserve(ctx, ctx->seg[i], slot_bytes));
        if (!pinned && !ctx->pinned[i]) {
            cudaError_t e = cudaHostAlloc(&ctx->pinned[i], slot_bytes, cudaHostAllocDefault);
            if (e != cudaSuccess) { ctx->pinned[i] = nullptr; cudaGetLastError(); return gl_fail(ctx, GL_ENOMEM, "cudaHostAlloc staging: %s", cudaGetErrorString(e)); }
        }
    }
    int64_t done = 0;
    int slot = 0;
    while (done < n) {
        const int64_t m = (n - done < kStageSegs) ? n - done : kStageSegs;
        int32_t* d_s = static_cast<int32_t*>(ctx->seg[slot].p);
        int32_t* d_e = d_s + kStageSegs;
        // the scatter that last read this device slot must be finished before it is overwritten
        GL_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_used[slot], 0));
        if (pinned) {
            GL_CUDA(ctx, cudaMemcpyAsync(d_s, start + done, (size_t)m * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
            GL_CUDA(ctx, cudaMemcpyAsync(d_e, end + done, (size_t)m * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
        } else {
            // the previous H2D out of this pinned slot must have drained before we refill it
            GL_CUDA(ctx, cudaEventSynchronize(ctx->ev_copy[slot]));
            int32_t* h_s = static_cast<int32_t*>(ctx->pinned[slot]);
            int32_t* h_e = h_s + kStageSegs;
            memcpy(h_s, start + done, (size_t)m * 4);
            memcpy(h_e, end + done, (size_t)m * 4);
            GL_CUDA(ctx, cudaMemcpyAsync(d_s, h_s, (size_t)m * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
            GL_CUDA(ctx, cudaMemcpyAsync(d_e, h_e, (size_t)m * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
        }
        GL_CUDA(ctx, cudaEventRecord(ctx->ev_copy[slot], ctx->copy_stream));
        GL_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[slot], 0));
        GL_CHECK(launch_scatter(ctx, d_s, d_e, m));
        GL_CUDA(ctx, cudaEventRecord(ctx->ev_used[slot], ctx->stream));
        done += m;
        slot ^= 1;
    }
    return GL_OK;
}

int gl_depth_reduce(gl_ctx* ctx, int32_t W, int32_t mincov, int32_t maxmean, int64_t run_break) {
    GL_CHECK(gl_use(ctx));
    if (!ctx->depth_active) return gl_fail(ctx, GL_ESTATE, "gl_depth_reduce: no region open");
    if (W <= 0 || run_break < 0) return gl_fail(ctx, GL_EINVAL, "gl_depth_reduce: W must be > 0 and run_break >= 0");
    return run_reduce(ctx, W, mincov, maxmean, run_break, true, true, nullptr, true);
}

int gl_depth_result_sizes(gl_ctx* ctx, int64_t* n_windows, int64_t* n_runs, int32_t* max_depth) {
    GL_CHECK(gl_use(ctx));
    if (!ctx->depth_reduced) return gl_fail(ctx, GL_ESTATE, "gl_depth_result_sizes: call gl_depth_reduce first");
    GL_CHECK(fetch_header(ctx));
    if (n_windows) *n_windows = ctx->n_windows;
    if (n_runs) *n_runs = ctx->n_runs;
    if (max_depth) *max_depth = ctx->max_depth;
    return GL_OK;
}

int gl_depth_get_windows(gl_ctx* ctx, int64_t* sum_out, int32_t* min_out, int64_t cap) {
    GL_CHECK(gl_use(ctx));
    if (!ctx->depth_reduced || ctx->n_windows == 0) return gl_fail(ctx, GL_ESTATE, "gl_depth_get_windows: no window results");
    if (!sum_out) return gl_fail(ctx, GL_EINVAL, "gl_depth_get_windows: null sum_out");
    if (cap < ctx->n_windows) return gl_fail(ctx, GL_ERANGE, "gl_depth_get_windows: cap %lld < %lld windows", (long long)cap, (long long)ctx->n_windows);
    GL_CUDA(ctx, cudaMemcpyAsync(sum_out, ctx->win_sum.p, (size_t)ctx->n_windows * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (min_out) GL_CUDA(ctx, cudaMemcpyAsync(min_out, ctx->win_min.p, (size_t)ctx->n_windows * 4, cudaMemcpyDeviceToHost, ctx->stream));
    GL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GL_OK;
}

int gl_depth_get_runs(gl_ctx* ctx, int32_t* run_start, int32_t* run_end, uint8_t* run_class, int64_t cap) {
    GL_CHECK(gl_use(ctx));
    if (!ctx->depth_reduced) return gl_fail(ctx, GL_ESTATE, "gl_depth_get_runs: call gl_depth_reduce first");
    GL_CHECK(fetch_header(ctx));
    const int64_t n = ctx->n_runs;
    if (cap < n) return gl_fail(ctx, GL_ERANGE, "gl_depth_get_runs: cap %lld < %lld runs", (long long)cap, (long long)n);
    if (n == 0) return GL_OK;
    if (run_start) GL_CUDA(ctx, cudaMemcpyAsync(run_start, ctx->run_start.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (run_class) GL_CUDA(ctx, cudaMemcpyAsync(run_class, ctx->run_class.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (run_end) {
        GL_CHECK(gl_buf_reserve(ctx, ctx->misc, (size_t)n * 4));
        run_ends_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(static_cast<const int*>(ctx->run_start.p),
                                                                          static_cast<int*>(ctx->misc.p), n, (int)ctx->re);
        GL_LAUNCHED(ctx, 1);
        GL_CUDA(ctx, cudaMemcpyAsync(run_end, ctx->misc.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    GL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GL_OK;
}

int gl_depth_windows(gl_ctx* ctx, int32_t W, int64_t* sum_out, int32_t* min_out, int64_t n_windows) {
    GL_CHECK(gl_use(ctx));
    if (!ctx->depth_active) return gl_fail(ctx, GL_ESTATE, "gl_depth_windows: no region open");
    if (W <= 0) return gl_fail(ctx, GL_EINVAL, "gl_depth_windows: W must be > 0");
    GL_CHECK(run_reduce(ctx, W, 0, 0, 0, true, false, nullptr, min_out != nullptr));
    return gl_depth_get_windows(ctx, sum_out, min_out, n_windows);
}

int gl_depth_classes(gl_ctx* ctx, int32_t mincov, int32_t maxmean, int32_t* run_start, int32_t* run_end,
                     uint8_t* run_class, int64_t cap, int64_t* n_runs) {
    GL_CHECK(gl_use(ctx));
    if (!ctx->depth_active) return gl_fail(ctx, GL_ESTATE, "gl_depth_classes: no region open");
    GL_CHECK(run_reduce(ctx, 1 << 30, mincov, maxmean, 0, false, true, nullptr, false));
    GL_CHECK(fetch_header(ctx));
    if (n_runs) *n_runs = ctx->n_runs;
    return gl_depth_get_runs(ctx, run_start, run_end, run_class, cap);
}

int gl_depth_perbase(gl_ctx* ctx, int32_t* depth_out) {
    GL_CHECK(gl_use(ctx));
    if (!ctx->depth_active) return gl_fail(ctx, GL_ESTATE, "gl_depth_perbase: no region open");
    if (!depth_out) return gl_fail(ctx, GL_EINVAL, "gl_depth_perbase: null output");
    const int64_t len = ctx->re - ctx->rs;
    GL_CHECK(gl_buf_reserve(ctx, ctx->misc, (size_t)(len + 4) * 4));
    GL_CHECK(run_reduce(ctx, 1 << 30, 0, 0, 0, false, false, static_cast<int32_t*>(ctx->misc.p), false));
    GL_CUDA(ctx, cudaMemcpyAsync(depth_out, ctx->misc.p, (size_t)len * 4, cudaMemcpyDeviceToHost, ctx->stream));
    GL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->depth_reduced = false;
    return GL_OK;
}

int gl_depth_region(gl_ctx* ctx, int64_t region_start, int64_t region_end, const int32_t* start, const int32_t* end,
                    int64_t n, int32_t W, int32_t mincov, int32_t maxmean, int64_t run_break, int64_t* sum_out,
                    int64_t win_cap, int64_t* n_windows, int32_t* run_start, uint8_t* run_class, int64_t run_cap,
                    int64_t* n_runs) {
    GL_CHECK(gl_depth_begin(ctx, region_start, region_end));
    GL_CHECK(gl_depth_add_segments(ctx, start, end, n));
    GL_CHECK(gl_depth_reduce(ctx, W, mincov, maxmean, run_break));
    int64_t nw = 0, nr = 0;
    GL_CHECK(gl_depth_result_sizes(ctx, &nw, &nr, nullptr));
    if (n_windows) *n_windows = nw;
    if (n_runs) *n_runs = nr;
    if (nw > win_cap) return gl_fail(ctx, GL_ERANGE, "gl_depth_region: %lld windows > cap %lld", (long long)nw, (long long)win_cap);
    if (nr > run_cap) return gl_fail(ctx, GL_ERANGE, "gl_depth_region: %lld runs > cap %lld", (long long)nr, (long long)run_cap);
    GL_CUDA(ctx, cudaMemcpyAsync(sum_out, ctx->win_sum.p, (size_t)nw * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (nr > 0) {
        GL_CUDA(ctx, cudaMemcpyAsync(run_start, ctx->run_start.p, (size_t)nr * 4, cudaMemcpyDeviceToHost, ctx->stream));
        GL_CUDA(ctx, cudaMemcpyAsync(run_class, ctx->run_class.p, (size_t)nr, cudaMemcpyDeviceToHost, ctx->stream));
    }
    GL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GL_OK;
}

}  // extern "C"
