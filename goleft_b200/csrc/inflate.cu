// inflate.cu — BGZF (RFC 1951 DEFLATE in RFC 1952 members, SAM spec 4.1) inflated ON THE GPU.
//
// The reference's `samtools depth` children spend their time in zlib's inflate (depth/depth.go:45); so does this repo's
// host feeder (host/bam_feed.cpp: ~120 MB/s per hyper-thread with a whole socket busy, the CLI's wall clock).  A BGZF
// file is a sequence of independent members of at most 64 KB, which is exactly the parallelism a GPU wants: one WARP per
// member, thousands in flight.
//
//   K_inflate   one warp per BGZF block.  The warp's lanes build the Huffman decoding tables of a deflate block together
//               (code-length counting, canonical first codes, a 10-bit primary lookup table in shared memory); lane 0 then
//               decodes symbols — the bit stream is inherently sequential — keeping a 64-bit bit buffer refilled with
//               32-bit loads; literals are stored by lane 0, matches (length, distance) are copied by ALL lanes
//               (out[p+i] = out[p-dist + i mod dist] is correct for overlapping matches, every source byte is already
//               written).  Codes longer than 10 bits take a canonical bit-by-bit walk (rare).  Stored, fixed and dynamic
//               blocks are handled; anything malformed sets the block's status and the host falls back to zlib.
//
// Written from RFC 1951; checked against zlib on every test input (tests/test_inflate_gpu.py).
#include "gl_common.cuh"
#include <string.h>

namespace {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kInfWarps = 4;                      // warps (= BGZF blocks) per CTA
constexpr int kPrimBits = 10;                     // primary literal/length table
constexpr int kDistBits = 8;                      // primary distance table

struct InfTables {
    unsigned short lit_prim[1 << kPrimBits];      // (symbol << 4) | code length, 0 = longer than kPrimBits / invalid
    unsigned short dist_prim[1 << kDistBits];
    unsigned short lit_sym[288];                  // symbols in canonical order (for the slow walk)
    unsigned short dist_sym[32];
    int lit_count[16], dist_count[16];            // codes per length
    unsigned char lens[320];                      // code lengths being read (litlen then dist)
};

__constant__ unsigned short c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ unsigned char c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ unsigned short c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ unsigned char c_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ unsigned char c_clen_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// ---- bit reader (lane 0 only): LSB-first, 64-bit buffer, refilled 4 bytes at a time from global memory
struct BitReader {
    const unsigned char* p;       // next byte to load
    const unsigned char* end;     // end of the deflate payload
    unsigned long long buf;
    int cnt;                      // bits in buf
    int phantom;                  // of which zeros fed past the end of the payload (consuming one of them is an error)
    __device__ __forceinline__ void init(const unsigned char* b, const unsigned char* e) { p = b; end = e; buf = 0; cnt = 0; phantom = 0; }
    __device__ __forceinline__ void refill() {
        while (cnt <= 32) {
            unsigned v;
            if (p + 4 <= end && ((reinterpret_cast<uintptr_t>(p) & 3) == 0)) { v = *reinterpret_cast<const unsigned*>(p); p += 4; buf |= (unsigned long long)v << cnt; cnt += 32; }
            else if (p < end) { v = *p++; buf |= (unsigned long long)v << cnt; cnt += 8; }
            else { cnt += 8; phantom += 8; }       // past the end: zeros (a well-formed stream never consumes them)
        }
    }
    __device__ __forceinline__ unsigned peek(int n) const { return (unsigned)(buf & ((1ull << n) - 1ull)); }
    __device__ __forceinline__ void drop(int n) { buf >>= n; cnt -= n; }
    __device__ __forceinline__ unsigned bits(int n) { if (cnt < n) refill(); const unsigned v = peek(n); drop(n); return v; }
    __device__ __forceinline__ bool overrun() const { return cnt < phantom; }
};

__device__ __forceinline__ unsigned rev_bits(unsigned v, int n) { return __brev(v) >> (32 - n); }

// Build the decoding tables for one alphabet from code lengths lens[0..n): count per length, canonical symbol order, and the
// primary table (reversed codes, replicated).  All 32 lanes take part.  Returns false for an over-subscribed code.
__device__ bool build_table(const unsigned char* lens, int n, int* count, unsigned short* sym, unsigned short* prim, int prim_bits, int lane) {
    if (lane < 16) count[lane] = 0;
    __syncwarp();
    for (int i = lane; i < n; i += 32) if (lens[i]) atomicAdd(&count[lens[i]], 1);
    __syncwarp();
    for (int i = lane; i < (1 << prim_bits); i += 32) prim[i] = 0;
    __syncwarp();
    // offsets (first index in sym[] of each length) and first codes, computed redundantly by every lane
    int offs[16], first[16];
    int code = 0, left = 1, o = 0;
    bool ok = true;
#pragma unroll
    for (int l = 1; l < 16; l++) {
        left <<= 1;
        left -= count[l];
        if (left < 0) ok = false;
        code <<= 1;
        first[l] = code;
        code += count[l];
        offs[l] = o;
        o += count[l];
    }
    if (!ok) return false;
    // place the symbols: symbol i of length l gets rank = number of symbols j < i with the same length (warp-cooperative
    // would need a sort; the alphabets are tiny, so lane l handles length l sequentially)
    if (lane >= 1 && lane < 16) {
        const int l = lane;
        int k = 0;
        for (int i = 0; i < n; i++) {
            if (lens[i] == l) {
                sym[offs[l] + k] = (unsigned short)i;
                if (l <= prim_bits) {
                    const unsigned c = rev_bits((unsigned)(first[l] + k), l);
                    for (unsigned r = c; r < (1u << prim_bits); r += (1u << l)) prim[r] = (unsigned short)((i << 4) | l);
                }
                k++;
            }
        }
    }
    __syncwarp();
    return true;
}

// slow canonical walk for codes longer than the primary table (RFC 1951 3.2.2): returns the symbol or -1
__device__ int decode_slow(BitReader& br, const int* count, const unsigned short* sym) {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len < 16; len++) {
        code |= (int)br.bits(1);
        const int c = count[len];
        if (code - c < first) return sym[index + (code - first)];
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

__device__ __forceinline__ int decode_sym(BitReader& br, const unsigned short* prim, int prim_bits, const int* count, const unsigned short* sym) {
    if (br.cnt < 15) br.refill();
    const unsigned e = prim[br.peek(prim_bits)];
    if (e) { br.drop((int)(e & 15)); return (int)(e >> 4); }
    return decode_slow(br, count, sym);
}

// status: 0 ok, 1 bad header, 2 bad deflate data, 3 output size mismatch
__global__ void __launch_bounds__(kInfWarps * 32) bgzf_inflate_kernel(const unsigned char* __restrict__ comp, const long long* __restrict__ comp_off,
                                                                     const long long* __restrict__ out_off, long long n_blocks,
                                                                     unsigned char* __restrict__ out, int* __restrict__ status) {
    __shared__ InfTables s_tab[kInfWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long blk = (long long)blockIdx.x * kInfWarps + warp;
    if (blk >= n_blocks) return;
    InfTables& T = s_tab[warp];
    const unsigned char* src = comp + comp_off[blk];
    const long long csize = comp_off[blk + 1] - comp_off[blk];
    unsigned char* dst = out + out_off[blk];
    const int isize = (int)(out_off[blk + 1] - out_off[blk]);
    // ---- gzip member header (RFC 1952) with the BGZF extra field; payload = [12 + xlen, csize - 8)
    int st = 0;
    int hdr = 0;
    if (csize < 26 || src[0] != 0x1f || src[1] != 0x8b || src[2] != 8 || !(src[3] & 4)) st = 1;
    else hdr = 12 + (int)(src[10] | (src[11] << 8));
    if (!st && hdr + 8 > csize) st = 1;
    if (st) { if (lane == 0) status[blk] = st; return; }
    BitReader br;
    br.init(src + hdr, src + csize - 8);
    int pos = 0;                                   // bytes written (uniform across the warp)
    bool last = false;
    while (!last && st == 0) {
        // ---- block header, read by lane 0 and broadcast
        unsigned h = 0;
        if (lane == 0) h = br.bits(3);
        h = __shfl_sync(kFullMask, h, 0);
        last = h & 1;
        const int type = (int)(h >> 1);
        if (type == 0) {                                                       // stored
            int len = 0;
            unsigned long long srcp = 0;
            if (lane == 0) {
                br.drop(br.cnt & 7);                                           // to a byte boundary
                // bytes still in the bit buffer belong to the stream: rewind the pointer over them
                const unsigned char* q = br.p - (br.cnt >> 3);
                br.buf = 0; br.cnt = 0;
                if (q + 4 > br.end) len = -1;
                else {
                    len = q[0] | (q[1] << 8);
                    const int nlen = q[2] | (q[3] << 8);
                    if ((len ^ 0xffff) != nlen || q + 4 + len > br.end) len = -1;
                    srcp = (unsigned long long)(q + 4);
                    br.p = q + 4 + (len > 0 ? len : 0);
                }
            }
            len = __shfl_sync(kFullMask, len, 0);
            srcp = __shfl_sync(kFullMask, srcp, 0);
            if (len < 0 || pos + len > isize) { st = 2; break; }
            const unsigned char* q = reinterpret_cast<const unsigned char*>(srcp);
            for (int i = lane; i < len; i += 32) dst[pos + i] = q[i];
            pos += len;
            __syncwarp();
            continue;
        }
        if (type == 3) { st = 2; break; }
        int nlit = 288, ndist = 30;
        if (type == 1) {                                                       // fixed codes (RFC 1951 3.2.6)
            for (int i = lane; i < 288; i += 32) T.lens[i] = (unsigned char)(i < 144 ? 8 : (i < 256 ? 9 : (i < 280 ? 7 : 8)));
            for (int i = lane; i < 30; i += 32) T.lens[288 + i] = 5;
            __syncwarp();
        } else {                                                               // dynamic: read the code lengths (lane 0), tables by all
            int ok = 1;
            if (lane == 0) {
                nlit = (int)br.bits(5) + 257;
                ndist = (int)br.bits(5) + 1;
                const int ncode = (int)br.bits(4) + 4;
                if (nlit > 286 || ndist > 30) ok = 0;
                for (int i = 0; i < 19; i++) T.lens[i] = 0;
                for (int i = 0; i < ncode; i++) T.lens[c_clen_order[i]] = (unsigned char)br.bits(3);
            }
            ok = __shfl_sync(kFullMask, ok, 0);
            nlit = __shfl_sync(kFullMask, nlit, 0);
            ndist = __shfl_sync(kFullMask, ndist, 0);
            if (!ok) { st = 2; break; }
            __syncwarp();
            // the code-length alphabet uses the distance slots of the table set (7-bit primary is plenty: lengths <= 7)
            if (!build_table(T.lens, 19, T.dist_count, T.dist_sym, T.dist_prim, 7, lane)) { st = 2; break; }
            if (lane == 0) {
                int i = 0;
                unsigned char tmp_prev = 0;
                // decoded lengths are staged in lit_sym's bytes?  No: they go to a separate area above the 19 code-length entries
                unsigned char* L = T.lens + 0;                                  // lens[] is re-used: decode into a scratch first
                unsigned char* scratch = reinterpret_cast<unsigned char*>(T.lit_prim);   // 2 KB, free until the literal table is built
                while (i < nlit + ndist) {
                    const int sy = decode_sym(br, T.dist_prim, 7, T.dist_count, T.dist_sym);
                    if (sy < 0) { ok = 0; break; }
                    if (sy < 16) { scratch[i++] = (unsigned char)sy; tmp_prev = (unsigned char)sy; }
                    else {
                        int rep, val = 0;
                        if (sy == 16) { if (i == 0) { ok = 0; break; } val = tmp_prev; rep = 3 + (int)br.bits(2); }
                        else if (sy == 17) rep = 3 + (int)br.bits(3);
                        else rep = 11 + (int)br.bits(7);
                        if (i + rep > nlit + ndist) { ok = 0; break; }
                        while (rep--) scratch[i++] = (unsigned char)val;
                        tmp_prev = (unsigned char)val;
                    }
                }
                if (ok) {
                    for (int k = 0; k < nlit; k++) L[k] = scratch[k];
                    for (int k = nlit; k < 288; k++) L[k] = 0;
                    for (int k = 0; k < ndist; k++) L[288 + k] = scratch[nlit + k];
                    for (int k = ndist; k < 32; k++) L[288 + k] = 0;
                    if (L[256] == 0) ok = 0;                                    // no end-of-block code
                }
            }
            ok = __shfl_sync(kFullMask, ok, 0);
            if (!ok) { st = 2; break; }
            __syncwarp();
            nlit = 288; ndist = 30;
        }
        if (!build_table(T.lens, nlit, T.lit_count, T.lit_sym, T.lit_prim, kPrimBits, lane)) { st = 2; break; }
        if (!build_table(T.lens + 288, ndist, T.dist_count, T.dist_sym, T.dist_prim, kDistBits, lane)) { st = 2; break; }
        // ---- symbols: lane 0 decodes; literals stored by lane 0, matches copied by the warp
        for (;;) {
            int kind = 0, mlen = 0, mdist = 0;                                  // kind 1: match, 2: end of block, 3: error
            if (lane == 0) {
                for (;;) {
                    const int sy = decode_sym(br, T.lit_prim, kPrimBits, T.lit_count, T.lit_sym);
                    if (sy < 0) { kind = 3; break; }
                    if (sy < 256) {
                        if (pos >= isize) { kind = 3; break; }
                        dst[pos++] = (unsigned char)sy;
                        continue;
                    }
                    if (sy == 256) { kind = 2; break; }
                    const int li = sy - 257;
                    if (li >= 29) { kind = 3; break; }
                    mlen = c_len_base[li] + (int)br.bits(c_len_extra[li]);
                    const int ds = decode_sym(br, T.dist_prim, kDistBits, T.dist_count, T.dist_sym);
                    if (ds < 0 || ds >= 30) { kind = 3; break; }
                    mdist = c_dist_base[ds] + (int)br.bits(c_dist_extra[ds]);
                    if (mdist > pos || pos + mlen > isize) { kind = 3; break; }
                    if (mlen <= 8) {                                            // short match: not worth a warp round trip
                        for (int i = 0; i < mlen; i++) dst[pos + i] = dst[pos - mdist + i];
                        pos += mlen;
                        continue;
                    }
                    kind = 1;
                    break;
                }
                if (br.overrun()) kind = 3;
            }
            kind = __shfl_sync(kFullMask, kind, 0);
            pos = __shfl_sync(kFullMask, pos, 0);
            if (kind == 1) {
                mlen = __shfl_sync(kFullMask, mlen, 0);
                mdist = __shfl_sync(kFullMask, mdist, 0);
                __syncwarp();                                                   // lane 0's earlier stores are visible to the warp
                for (int i = lane; i < mlen; i += 32) dst[pos + i] = dst[pos - mdist + (mdist >= mlen ? i : i % mdist)];
                pos += mlen;
                __syncwarp();
                continue;
            }
            if (kind == 3) st = 2;
            break;
        }
        __syncwarp();
    }
    if (st == 0 && pos != isize) st = 3;
    if (lane == 0) status[blk] = st;
}

}  // namespace

extern "C" {

// Inflate n_blocks BGZF members that are resident on the device: block b = d_comp[comp_off[b], comp_off[b+1]) (the whole
// member, header and trailer included) -> d_out[out_off[b], out_off[b+1]) (out_off from the members' ISIZE fields).
// d_status[b]: 0 ok, != 0 malformed / unsupported (the caller falls back to zlib for the file).  Asynchronous on the ctx
// stream.  CRC32 is not checked (the host feeder does not check it either; ISIZE is).
int gl_bgzf_inflate_device(gl_ctx* ctx, const uint8_t* d_comp, const int64_t* d_comp_off, const int64_t* d_out_off, int64_t n_blocks,
                           uint8_t* d_out, int32_t* d_status) {
    GL_CHECK(gl_use(ctx));
    if (n_blocks < 0 || (n_blocks > 0 && (!d_comp || !d_comp_off || !d_out_off || !d_out || !d_status)))
        return gl_fail(ctx, GL_EINVAL, "gl_bgzf_inflate_device: bad argument");
    if (n_blocks == 0) return GL_OK;
    {
        gl_prof_scope prof(ctx, "bgzf_inflate_kernel");
        bgzf_inflate_kernel<<<(unsigned)((n_blocks + kInfWarps - 1) / kInfWarps), kInfWarps * 32, 0, ctx->stream>>>(
            d_comp, reinterpret_cast<const long long*>(d_comp_off), reinterpret_cast<const long long*>(d_out_off), n_blocks, d_out, d_status);
    }
    GL_LAUNCHED(ctx, 1);
    return GL_OK;
}

}  // extern "C"
