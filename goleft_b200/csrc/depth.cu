// depth.cu — the `goleft depth` hot path on sm_100a.
//
//   segments (start,end) --K1 scatter (int32 red.global)--> difference array in HBM
//                          + warp-aggregated net sums per 4096-base tile and per 64-tile super-tile
//   difference array    --K2 ONE fused streaming pass----> per-base depth (registers/smem only)
//                                        -> per-window int64 sum + int32 min   (depth/depth.go:293-306)
//                                        -> coverage-class run starts          (depth/depth.go:307-327)
//   run starts          --K3 gather----------------------> runs in position order
//
// Because the scatter already knows which tile every +1/-1 lands in, the depth carried into a tile
// is the sum of <= tiles/64 super-tile sums plus <= 63 tile sums — ~300 coalesced L2 loads that
// overlap the tile's own HBM loads — instead of a serial dependency through the 64M-element array.
// K2 therefore has no inter-CTA dependency at all (no look-back chain, no ticket), reads the
// difference array exactly once, and per-base depth never goes back to HBM.  Variable-length run
// output is claimed per tile with one atomic and put into position order by K3 the same way.
// All arithmetic is integer; results are bit-exact against oracle/oracle_depth.c.
#include "gl_common.cuh"
#include <string.h>

namespace {

constexpr int kScanThreads = 256;
constexpr int kWarps = kScanThreads / 32;
constexpr int kRounds = 4;                        // int4 per lane per round
constexpr int kWarpElems = 32 * 4 * kRounds;      // 512 bases per warp
constexpr int kTile = kWarps * kWarpElems;        // 4096 bases per tile
constexpr int kTileShift = 12;
static_assert((1 << kTileShift) == kTile, "tile shift");
constexpr int kHeaderWords = 8;                   // u64 words in front of the per-tile tables
constexpr unsigned kFull = 0xffffffffu;
constexpr int kSuperShift = 6;                    // 64 tiles per super-tile

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
// The same events are summed per 4096-base tile; coordinate-sorted input puts a whole warp in one
// or two tiles, so the warp reduces with REDUX and issues one red per distinct tile.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_tile_add(int* __restrict__ tile_sum, int t, int c, int sign, int lane) {
    unsigned rem = __ballot_sync(kFull, c != 0);
#pragma unroll 1
    for (int it = 0; rem != 0 && it < 3; it++) {
        const int leader = __ffs(rem) - 1;
        const int key = __shfl_sync(kFull, t, leader);
        const bool m = (c != 0) && (t == key);
        const int tot = __reduce_add_sync(kFull, m ? c : 0);
        if (lane == leader) atomicAdd(tile_sum + key, sign * tot);
        if (m) c = 0;
        rem = __ballot_sync(kFull, c != 0);
    }
    if (c != 0) atomicAdd(tile_sum + t, sign * c);     // unsorted input: fall back to one red per lane
}

template <bool kVec>
__global__ void __launch_bounds__(256) depth_scatter_kernel(const int* __restrict__ start, const int* __restrict__ end,
                                                            long long n, int rs, int re, int* __restrict__ diff,
                                                            int* __restrict__ tile_sum) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int s[4], e[4];
    int cnt = 0;
    if (kVec) {
        const long long n4 = n >> 2;
        if (i < n4) {
            int4 a = ld_stream_int4(reinterpret_cast<const int4*>(start) + i);
            int4 b = ld_stream_int4(reinterpret_cast<const int4*>(end) + i);
            s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w;
            e[0] = b.x; e[1] = b.y; e[2] = b.z; e[3] = b.w;
            cnt = 4;
        } else {
            const long long k = (n4 << 2) + (i - n4);
            if (k < n) { s[0] = start[k]; e[0] = end[k]; cnt = 1; }
        }
    } else if (i < n) {
        s[0] = start[i]; e[0] = end[i]; cnt = 1;
    }
    int tS = 0, cS = 0, tE = 0, cE = 0;
    bool uni = true;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (j < cnt) {
            const int a = max(s[j], rs) - rs, b = min(e[j], re) - rs;
            if (a < b) {
                atomicAdd(diff + a, 1);          // result unused -> RED.E.ADD
                atomicAdd(diff + b, -1);
                const int ta = a >> kTileShift, tb = b >> kTileShift;
                if (cS == 0) { tS = ta; tE = tb; }
                else if (ta != tS || tb != tE) uni = false;
                s[j] = ta; e[j] = tb;            // keep the tile ids for the slow path
                cS++; cE++;
            } else {
                s[j] = -1;
            }
        } else {
            s[j] = -1;
        }
    }
    if (!uni) {                                  // this thread straddles a tile edge: per-event reds
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (s[j] >= 0) { atomicAdd(tile_sum + s[j], 1); atomicAdd(tile_sum + e[j], -1); }
        cS = 0; cE = 0;
    }
    warp_tile_add(tile_sum, tS, cS, 1, lane);
    warp_tile_add(tile_sum, tE, cE, -1, lane);
}

// K1b: super_sum[st] = sum of the 64 tile sums of super-tile st (one warp each).  Doing this in the
// scatter itself would put every warp's red on the same handful of addresses — measured 7x slower.
__global__ void __launch_bounds__(256) depth_super_sum_kernel(const int* __restrict__ tile_sum, int* __restrict__ super_sum,
                                                             int tiles) {
    const int st = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int t0 = st << kSuperShift;
    if (t0 >= tiles) return;
    int v = 0;
    if (t0 + lane < tiles) v += tile_sum[t0 + lane];
    if (t0 + 32 + lane < tiles) v += tile_sum[t0 + 32 + lane];
    v = __reduce_add_sync(kFull, v);
    if (lane == 0) super_sum[st] = v;
}

// ------------------------------------------------------------------------------------------------
// K2: fused scan + window reduce + class runs.  One CTA per 4096-base tile, no inter-CTA dependency.
// ------------------------------------------------------------------------------------------------
struct ScanParams {
    const int* diff;              // padded to a whole number of tiles, zero beyond len
    const int* tile_sum;          // [num_tiles+1] net (+starts -ends) inside each tile
    const int* super_sum;         // [num_tiles/64+1] the same per 64 tiles
    int len;                      // bases in the region
    int rs;                       // absolute start of the region
    int W;                        // window size
    long long w0;                 // rs / W  (index of the first window)
    int mincov, maxmean;
    unsigned run_break;           // 0 = never (host clamps values >= 2^32 to 0: no multiple in range)
    unsigned long long* win_sum;  // [n_windows], zero-initialised
    int* win_min;                 // [n_windows], initialised to 0x7f7f7f7f (or null)
    int* tmp_start;               // [run_cap] run starts in claim order
    unsigned char* tmp_class;     // [run_cap]
    long long run_cap;
    uint64_t* header;             // [0]=runs claimed (= n_runs) [2]=max_depth
    uint64_t* tile_runs;          // [num_tiles] (claim offset << 32) | count
    unsigned* super_cnt;          // [num_tiles/64+1] run starts per super-tile
    int* depth_out;               // optional per-base output (debug/parity), else null
    int num_tiles;
    int do_windows, do_runs;
};

__device__ __forceinline__ int cov_class(int d, int mincov, int maxmean) {
    // depth/depth.go:223-234
    return d == 0 ? GL_NO_COVERAGE
                  : (d < mincov ? GL_LOW_COVERAGE : ((maxmean > 0 && d >= maxmean) ? GL_EXCESSIVE_COVERAGE : GL_CALLABLE));
}

// class(d) is monotone in d, so a set of depths is single-class iff class(min) == class(max).
__global__ void __launch_bounds__(kScanThreads, 4) depth_scan_kernel(const ScanParams p) {
    __shared__ __align__(16) int s_depth[kTile];
    __shared__ int s_warp_tot[kWarps];
    __shared__ int s_warp_cnt[kWarps];
    __shared__ int s_warp_max[kWarps];
    __shared__ int s_warp_carry[kWarps];
    __shared__ long long s_run_base;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int tile_base = tile * kTile;                       // relative position of the tile
    const int warp_base = tile_base + warp * kWarpElems;

    // ---- load 16 diffs per thread, warp-coalesced 512 B per instruction
    int4 v[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; r++)
        v[r] = ld_stream_int4(reinterpret_cast<const int4*>(p.diff + warp_base + r * 128) + lane);

    // ---- depth carried into the tile = sum of everything before it (two-level, overlaps the loads above)
    {
        const int st = tile >> kSuperShift;
        int part = 0;
        for (int k = tid; k < st; k += kScanThreads) part += p.super_sum[k];
        const int k2 = (st << kSuperShift) + tid;
        if (tid < (1 << kSuperShift) && k2 < tile) part += p.tile_sum[k2];
        part = __reduce_add_sync(kFull, part);
        if (lane == 0) s_warp_carry[warp] = part;
    }

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

    int wbase = 0;                                            // depth at the base before this warp's first
#pragma unroll
    for (int w = 0; w < kWarps; w++) wbase += s_warp_carry[w] + ((w < warp) ? s_warp_tot[w] : 0);

    // ---- per-base depth (registers + this warp's smem slice), per-round min / max / sum
    int* sw = s_depth + warp * kWarpElems;
    int4 d[kRounds];
    int mn = 0x7fffffff, mx = 0;
#pragma unroll
    for (int r = 0; r < kRounds; r++) {
        const int prev = wbase + pre[r];
        d[r].x = prev + v[r].x;
        d[r].y = d[r].x + v[r].y;
        d[r].z = d[r].y + v[r].z;
        d[r].w = d[r].z + v[r].w;
        reinterpret_cast<int4*>(sw + r * 128)[lane] = d[r];
        mn = min(mn, min(min(d[r].x, d[r].y), min(d[r].z, d[r].w)));
        mx = max(mx, max(max(d[r].x, d[r].y), max(d[r].z, d[r].w)));
        if (p.depth_out) {
            const int idx = warp_base + r * 128 + lane * 4;
            if (idx + 3 < p.len) reinterpret_cast<int4*>(p.depth_out + idx)[0] = d[r];
            else {
                if (idx < p.len) p.depth_out[idx] = d[r].x;
                if (idx + 1 < p.len) p.depth_out[idx + 1] = d[r].y;
                if (idx + 2 < p.len) p.depth_out[idx + 2] = d[r].z;
            }
        }
    }
    const bool full_warp = warp_base + kWarpElems <= p.len;   // false only at the region's ragged end
    const int wmin = __reduce_min_sync(kFull, mn), wmax = __reduce_max_sync(kFull, mx);
    __syncwarp();

    // does a forced run break (multiple of run_break) fall inside this tile?
    bool tile_has_break = false;
    if (p.do_runs && p.run_break > 0) {
        const unsigned a = (unsigned)p.rs + (unsigned)tile_base;      // absolute positions are < 2^32
        const unsigned rem = a % p.run_break;
        tile_has_break = (rem ? p.run_break - rem : 0u) < (unsigned)kTile;
    }

    // ---- class-change masks.  Fast path: the warp's 512 bases and the base before them are one class.
    unsigned masks = 0;
    int maxd = 0;
    bool slow_runs = false;
    if (full_warp) maxd = wmax;
    if (p.do_runs) {
        slow_runs = !full_warp || tile_has_break || warp_base == 0 ||
                    cov_class(min(wmin, wbase), p.mincov, p.maxmean) != cov_class(max(wmax, wbase), p.mincov, p.maxmean);
    }
    if (slow_runs || !full_warp) {
#pragma unroll
        for (int r = 0; r < kRounds; r++) {
            const int idx = warp_base + r * 128 + lane * 4;   // relative position of element 0
            if (idx < p.len) {
                const int nv = min(4, p.len - idx);
                maxd = max(maxd, d[r].x);
                if (nv > 1) maxd = max(maxd, d[r].y);
                if (nv > 2) maxd = max(maxd, d[r].z);
                if (nv > 3) maxd = max(maxd, d[r].w);
                if (slow_runs) {
                    const int cp = cov_class(wbase + pre[r], p.mincov, p.maxmean);
                    const int c0 = cov_class(d[r].x, p.mincov, p.maxmean);
                    const int c1 = cov_class(d[r].y, p.mincov, p.maxmean);
                    const int c2 = cov_class(d[r].z, p.mincov, p.maxmean);
                    const int c3 = cov_class(d[r].w, p.mincov, p.maxmean);
                    unsigned m = (c0 != cp ? 1u : 0u) | (c1 != c0 ? 2u : 0u) | (c2 != c1 ? 4u : 0u) | (c3 != c2 ? 8u : 0u);
                    if (idx == 0) m |= 1u;                    // the region's first base always starts a run
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
        if (!full_warp) maxd = __reduce_max_sync(kFull, maxd);
    }

    // ---- rank of each run start inside the warp, in position order (round, lane, j)
    int lane_rank[kRounds];
    int round_tot[kRounds];
    int warp_cnt = 0;
#pragma unroll
    for (int r = 0; r < kRounds; r++) { lane_rank[r] = 0; round_tot[r] = 0; }
    if (slow_runs) {
        const unsigned lt = (1u << lane) - 1u;
#pragma unroll
        for (int r = 0; r < kRounds; r++) {
            unsigned m = (masks >> (4 * r)) & 15u;
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
    }
    if (lane == 0) { s_warp_cnt[warp] = warp_cnt; s_warp_max[warp] = maxd; }

    // ---- window partial sums of this warp's 512 bases
    if (p.do_windows) {
        const unsigned a0 = (unsigned)p.rs + (unsigned)warp_base;                        // absolute, < 2^32
        const unsigned a1 = (unsigned)p.rs + (unsigned)min(warp_base + kWarpElems, p.len);
        const unsigned uW = (unsigned)p.W;
        if (full_warp && p.W >= 128 && wmax < (1 << 24)) {
            // Fast path from registers: a 128-base round holds at most one window edge; whole rounds are
            // summed with one REDUX (sums < 2^31 because depth < 2^24), the edge round is split by lane.
            unsigned iw = a0 / uW;
            long long next_b = (long long)(iw + 1) * uW - a0;         // offset of the next window edge in the warp
            unsigned long long acc = 0;
            int acc_min = 0x7fffffff;
#pragma unroll
            for (int r = 0; r < kRounds; r++) {
                const int q = d[r].x + d[r].y + d[r].z + d[r].w;
                const int qm = min(min(d[r].x, d[r].y), min(d[r].z, d[r].w));
                if (next_b >= (r + 1) * 128) {
                    acc += (unsigned)__reduce_add_sync(kFull, q);
                    if (p.win_min) acc_min = min(acc_min, __reduce_min_sync(kFull, qm));
                } else {
                    const int b = (int)next_b - r * 128 - lane * 4;   // elements j < b of this lane lie left of the edge
                    const int l = (b > 0 ? d[r].x : 0) + (b > 1 ? d[r].y : 0) + (b > 2 ? d[r].z : 0) + (b > 3 ? d[r].w : 0);
                    const unsigned left = (unsigned)__reduce_add_sync(kFull, l);
                    const unsigned right = (unsigned)__reduce_add_sync(kFull, q - l);
                    int lmin = 0x7fffffff, rmin = 0x7fffffff;
                    if (p.win_min) {
                        lmin = min(min(b > 0 ? d[r].x : lmin, b > 1 ? d[r].y : lmin), min(b > 2 ? d[r].z : lmin, b > 3 ? d[r].w : lmin));
                        rmin = min(min(b > 0 ? rmin : d[r].x, b > 1 ? rmin : d[r].y), min(b > 2 ? rmin : d[r].z, b > 3 ? rmin : d[r].w));
                        lmin = __reduce_min_sync(kFull, lmin);
                        rmin = __reduce_min_sync(kFull, rmin);
                    }
                    acc += left;
                    acc_min = min(acc_min, lmin);
                    if (lane == 0) {
                        atomicAdd(p.win_sum + ((long long)iw - p.w0), acc);
                        if (p.win_min) atomicMin(p.win_min + ((long long)iw - p.w0), acc_min);
                    }
                    iw++;
                    next_b += uW;
                    acc = right;
                    acc_min = rmin;
                }
            }
            if (lane == 0) {
                atomicAdd(p.win_sum + ((long long)iw - p.w0), acc);
                if (p.win_min) atomicMin(p.win_min + ((long long)iw - p.w0), acc_min);
            }
        } else if (a0 < a1) {
            // General path from this warp's smem slice: any W, ragged end, depth up to 2^31.
            const unsigned iw0 = a0 / uW, iw1 = (a1 - 1) / uW;
            if (p.W > 16) {
                for (unsigned iw = iw0; iw <= iw1; iw++) {
                    const long long ws = (long long)iw * uW, we = ws + uW;
                    const int s = (int)(max((long long)a0, ws) - a0), e = (int)(min((long long)a1, we) - a0);
                    unsigned long long sum = 0;
                    int m = 0x7fffffff;
                    for (int x = s + lane; x < e; x += 32) {
                        int dd = sw[x];
                        sum += (unsigned long long)(unsigned)dd;
                        m = min(m, dd);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(kFull, sum, o);
                    m = __reduce_min_sync(kFull, m);
                    if (lane == 0) {
                        atomicAdd(p.win_sum + ((long long)iw - p.w0), sum);
                        if (p.win_min) atomicMin(p.win_min + ((long long)iw - p.w0), m);
                    }
                }
            } else {
                for (unsigned iw = iw0 + lane; iw <= iw1; iw += 32) {
                    const long long ws = (long long)iw * uW, we = ws + uW;
                    const int s = (int)(max((long long)a0, ws) - a0), e = (int)(min((long long)a1, we) - a0);
                    unsigned long long sum = 0;
                    int m = 0x7fffffff;
                    for (int x = s; x < e; x++) {
                        int dd = sw[x];
                        sum += (unsigned long long)(unsigned)dd;
                        m = min(m, dd);
                    }
                    atomicAdd(p.win_sum + ((long long)iw - p.w0), sum);
                    if (p.win_min) atomicMin(p.win_min + ((long long)iw - p.w0), m);
                }
            }
        }
    }
    __syncthreads();

    // ---- per-tile summary; claim output slots for this tile's run starts
    int tile_cnt = 0, tile_mx = 0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) { tile_cnt += s_warp_cnt[w]; tile_mx = max(tile_mx, s_warp_max[w]); }
    if (tid == 0) {
        if (tile_mx > 0 && (unsigned long long)tile_mx > *reinterpret_cast<volatile unsigned long long*>(p.header + 2))
            atomicMax(reinterpret_cast<unsigned long long*>(p.header + 2), (unsigned long long)tile_mx);
        if (p.do_runs) {
            unsigned long long off = 0;
            if (tile_cnt) {
                off = atomicAdd(reinterpret_cast<unsigned long long*>(p.header), (unsigned long long)tile_cnt);
                atomicAdd(p.super_cnt + (tile >> kSuperShift), (unsigned)tile_cnt);
            }
            p.tile_runs[tile] = ((uint64_t)off << 32) | (unsigned)tile_cnt;
            s_run_base = (long long)off;
        }
    }
    if (!p.do_runs || tile_cnt == 0) return;
    __syncthreads();

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
                    p.tmp_start[rank] = p.rs + warp_base + off + j;
                    p.tmp_class[rank] = (unsigned char)cov_class(sw[off + j], p.mincov, p.maxmean);
                }
                rank++;
            }
            base += round_tot[r];
        }
    }
}

// K3: one warp per tile moves its runs from claim order to position order.  The ordered offset of a
// tile is the number of runs that start before it: super-tile counts + the tile counts of its own
// super-tile (only tiles that have runs pay for the sum).
__global__ void __launch_bounds__(256) depth_gather_runs_kernel(const uint64_t* __restrict__ tile_runs,
                                                               const unsigned* __restrict__ super_cnt, int tiles,
                                                               const int* __restrict__ tmp_start,
                                                               const unsigned char* __restrict__ tmp_class,
                                                               int* __restrict__ run_start, unsigned char* __restrict__ run_class,
                                                               long long cap) {
    const int t = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (t >= tiles) return;
    const uint64_t w = tile_runs[t];
    const unsigned cnt = (unsigned)w;
    if (cnt == 0) return;
    const int st = t >> kSuperShift;
    unsigned part = 0;
    for (int k = lane; k < st; k += 32) part += super_cnt[k];
    for (int k = (st << kSuperShift) + lane; k < t; k += 32) part += (unsigned)tile_runs[k];
    const long long dst = (long long)__reduce_add_sync(kFull, part);
    const long long src = (long long)(w >> 32);
    for (unsigned i = lane; i < cnt; i += 32) {
        if (src + i < cap && dst + i < cap) {
            run_start[dst + i] = tmp_start[src + i];
            run_class[dst + i] = tmp_class[src + i];
        }
    }
}

__global__ void run_ends_kernel(const int* run_start, int* run_end, long long n, int re) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) run_end[i] = (i + 1 < n) ? run_start[i + 1] : re;
}

inline int64_t num_tiles_for(int64_t len) { return (len + kTile - 1) / kTile; }
inline size_t diff_entries_for(int64_t len) { return (size_t)num_tiles_for(len) * kTile + 4; }
inline size_t tile_entries_for(int64_t len) { return ((size_t)num_tiles_for(len) + 1 + 3) & ~size_t(3); }
inline size_t super_entries_for(int64_t len) { return (((size_t)num_tiles_for(len) >> kSuperShift) + 2 + 3) & ~size_t(3); }
inline int* tile_sum_ptr(gl_ctx* ctx) {
    return static_cast<int*>(ctx->diff.p) + diff_entries_for(ctx->re - ctx->rs);
}
inline int* super_sum_ptr(gl_ctx* ctx) { return tile_sum_ptr(ctx) + tile_entries_for(ctx->re - ctx->rs); }

int launch_scatter(gl_ctx* ctx, const int32_t* d_start, const int32_t* d_end, int64_t n) {
    if (n <= 0) return GL_OK;
    bool vec = ((reinterpret_cast<uintptr_t>(d_start) | reinterpret_cast<uintptr_t>(d_end)) & 15) == 0;
    int* diff = static_cast<int*>(ctx->diff.p);
    int* tsum = tile_sum_ptr(ctx);
    if (vec) {
        long long threads = (n >> 2) + (n & 3);
        unsigned grid = (unsigned)((threads + 255) / 256);
        depth_scatter_kernel<true><<<grid, 256, 0, ctx->stream>>>(d_start, d_end, n, (int)ctx->rs, (int)ctx->re, diff, tsum);
    } else {
        unsigned grid = (unsigned)((n + 255) / 256);
        depth_scatter_kernel<false><<<grid, 256, 0, ctx->stream>>>(d_start, d_end, n, (int)ctx->rs, (int)ctx->re, diff, tsum);
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
    // scratch: header u64[8] | super_cnt u32[supers] | tile_runs u64[tiles]
    const size_t supers = super_entries_for(len);
    const size_t head_bytes = kHeaderWords * 8 + ((supers * 4 + 7) & ~size_t(7));
    GL_CHECK(gl_buf_reserve(ctx, ctx->scratch, head_bytes + (size_t)tiles * 8));
    if (do_runs && ctx->run_start.cap == 0) {
        size_t cap = (size_t)(len / 16 + 4096);
        GL_CHECK(gl_buf_reserve(ctx, ctx->run_start, cap * 4));
        GL_CHECK(gl_buf_reserve(ctx, ctx->run_class, cap));
        GL_CHECK(gl_buf_reserve(ctx, ctx->run_tmp_start, cap * 4));
        GL_CHECK(gl_buf_reserve(ctx, ctx->run_tmp_class, cap));
    }
    uint64_t* header = static_cast<uint64_t*>(ctx->scratch.p);
    unsigned* super_cnt = reinterpret_cast<unsigned*>(header + kHeaderWords);
    uint64_t* tile_runs = reinterpret_cast<uint64_t*>(static_cast<char*>(ctx->scratch.p) + head_bytes);

    {
        const int n_super = (int)((tiles + (1 << kSuperShift) - 1) >> kSuperShift);
        depth_super_sum_kernel<<<(unsigned)((n_super + 7) / 8), 256, 0, ctx->stream>>>(tile_sum_ptr(ctx), super_sum_ptr(ctx), (int)tiles);
        GL_LAUNCHED(ctx, 1);
    }

    for (int attempt = 0; attempt < 2; attempt++) {
        ScanParams p;
        memset(&p, 0, sizeof p);
        p.diff = static_cast<const int*>(ctx->diff.p);
        p.tile_sum = tile_sum_ptr(ctx);
        p.super_sum = super_sum_ptr(ctx);
        p.len = (int)len;
        p.rs = (int)ctx->rs;
        p.W = W;
        p.w0 = w0;
        p.mincov = mincov;
        p.maxmean = maxmean;
        p.run_break = (run_break >= (int64_t(1) << 32)) ? 0u : (unsigned)run_break;
        p.win_sum = static_cast<unsigned long long*>(ctx->win_sum.p);
        p.win_min = want_min ? static_cast<int*>(ctx->win_min.p) : nullptr;
        p.tmp_start = static_cast<int*>(ctx->run_tmp_start.p);
        p.tmp_class = static_cast<unsigned char*>(ctx->run_tmp_class.p);
        long long cap = 0;
        if (do_runs) {
            cap = (long long)ctx->run_class.cap;
            cap = std::min<long long>(cap, (long long)(ctx->run_start.cap / 4));
            cap = std::min<long long>(cap, (long long)ctx->run_tmp_class.cap);
            cap = std::min<long long>(cap, (long long)(ctx->run_tmp_start.cap / 4));
        }
        p.run_cap = cap;
        p.header = header;
        p.tile_runs = tile_runs;
        p.super_cnt = super_cnt;
        p.depth_out = d_depth_out;
        p.num_tiles = (int)tiles;
        p.do_windows = do_windows ? 1 : 0;
        p.do_runs = do_runs ? 1 : 0;

        GL_CUDA(ctx, cudaMemsetAsync(header, 0, head_bytes, ctx->stream));
        if (do_windows) {
            GL_CUDA(ctx, cudaMemsetAsync(ctx->win_sum.p, 0, (size_t)n_windows * 8, ctx->stream));
            if (want_min) GL_CUDA(ctx, cudaMemsetAsync(ctx->win_min.p, 0x7f, (size_t)n_windows * 4, ctx->stream));
        }
        depth_scan_kernel<<<(unsigned)tiles, kScanThreads, 0, ctx->stream>>>(p);
        GL_LAUNCHED(ctx, 1);
        if (do_runs) {
            depth_gather_runs_kernel<<<(unsigned)((tiles + 7) / 8), 256, 0, ctx->stream>>>(
                tile_runs, super_cnt, (int)tiles, p.tmp_start, p.tmp_class, static_cast<int*>(ctx->run_start.p),
                static_cast<unsigned char*>(ctx->run_class.p), cap);
            GL_LAUNCHED(ctx, 1);
        }

        ctx->red_W = W; ctx->red_mincov = mincov; ctx->red_maxmean = maxmean; ctx->red_break = run_break;
        ctx->n_windows = do_windows ? n_windows : 0;
        ctx->n_runs = -1;            // fetched lazily
        ctx->max_depth = -1;
        if (!do_runs) { ctx->n_runs = 0; break; }
        if (attempt == 0) {
            // The run count decides whether the output fit; peek at it only when the capacity
            // could be exceeded at all (cap < len), else defer the sync to the first getter.
            if (cap >= len) break;
            uint64_t hdr[3];
            GL_CUDA(ctx, cudaMemcpyAsync(hdr, ctx->scratch.p, sizeof hdr, cudaMemcpyDeviceToHost, ctx->stream));
            GL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            ctx->n_runs = (int64_t)hdr[0];
            ctx->max_depth = (int32_t)hdr[2];
            if ((long long)hdr[0] <= cap) break;
            size_t ncap = (size_t)hdr[0] + 1024;
            GL_CHECK(gl_buf_reserve(ctx, ctx->run_start, ncap * 4));
            GL_CHECK(gl_buf_reserve(ctx, ctx->run_class, ncap));
            GL_CHECK(gl_buf_reserve(ctx, ctx->run_tmp_start, ncap * 4));
            GL_CHECK(gl_buf_reserve(ctx, ctx->run_tmp_class, ncap));
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
    if (ctx->n_runs < 0) ctx->n_runs = (int64_t)hdr[0];
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
    // difference array (padded to whole tiles) followed by the per-tile net sums: one memset clears both
    const size_t entries = diff_entries_for(len) + tile_entries_for(len) + super_entries_for(len);
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
