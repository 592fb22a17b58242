// bam_feed.h — the index-guided, parallel BAM feeder of `goleft depth`.
// Replaces what the reference gets from its `samtools depth -Q q -r chr:b-e` children (depth/depth.go:45,116,152):
// only the BGZF blocks of the requested reference range are read (BAI linear index -> record-aligned virtual offsets),
// the range is cut into units at those offsets, and every pool thread inflates AND parses its own units — no serial
// record walk.  Each unit yields its M/=/X blocks already in start order (the parser knows the read position, so the
// second block of a deletion read waits in a tiny pending list until the stream has passed it); the units are then
// assembled into the engine's packed8 words (short reads) or sorted int32 arrays (long reads) by seg_pack.h.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include "hts_io.h"

namespace glhts {

struct DecodeStats {
    int64_t n_records = 0, n_pass = 0;          // records seen / passing the samtools-depth filter
    int64_t bytes_in = 0, bytes_out = 0;        // compressed bytes read, bytes inflated
    double inflate_s = 0, parse_s = 0;          // summed over the workers
    double wall_s = 0;                          // of the whole call
    int units = 0;
};

struct ContigSegs {
    int format = 0;                             // 8: packed8 words, 32: (start,end) sorted by start, 0: nothing decoded
    std::vector<int32_t> anchors; std::vector<uint8_t> ds, len; int64_t n_blocks = 0;
    std::vector<int32_t> start, end; int64_t n = 0;
    int64_t n_pieces = 0; int32_t max_len = 0;
};

class BamFile {
public:
    BamFile() {}
    ~BamFile();
    BamFile(const BamFile&) = delete;
    // opens path and path + ".bai" (or path with .bam replaced by .bai); has_index tells whether one was found
    std::string open(const std::string& path);
    bool has_index = false;
    BamHeader header;
    BaiIndex bai;
    std::string path;
    // M/=/X blocks of the records of reference `tid` that pass (flag & 0x704) == 0 && MAPQ >= min_mapq and can overlap
    // [beg, end).  want: 0 = choose (packed8 unless the blocks are long), 8, 32.  Needs the index.
    std::string decode(int tid, int64_t beg, int64_t end, int min_mapq, int threads, int want, ContigSegs& out, DecodeStats* st);
    int fd() const { return fd_; }
    int64_t file_size() const { return size_; }
private:
    int fd_ = -1;
    int64_t size_ = 0;
};

}  // namespace glhts

// the C ABI's opaque BAM handle (feeder_api.cpp, bamgpu.cu)
struct gl_bam { glhts::BamFile f; glhts::ContigSegs segs; };
