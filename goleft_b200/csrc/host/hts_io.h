// hts_io.h — host-side readers the feeder needs: BGZF (multi-threaded inflate), BAM records -> filtered
// M/=/X segments, BAI linear index + stats bin, .fai, BGZF writer.  Written from the SAM/BAM spec
// (SURVEY.md Appendix D); replaces what the reference gets from the `samtools depth` child
// (depth/depth.go:45) and from github.com/biogo/hts (indexcov/indexcov.go:514, covstats/covstats.go:122-173).
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

namespace glhts {

struct RefInfo {
    std::string name;
    int64_t length = 0;
    int64_t offset = 0, line_bases = 0, line_width = 0;   // .fai columns 3-5 (0 when the source has none, e.g. a BAM header)
};

// byte offset of 0-based base p of a FASTA record, relative to the record's first base (faidx `position`)
inline int64_t fasta_position(const RefInfo& r, int64_t p) { return r.line_bases > 0 ? p / r.line_bases * r.line_width + p % r.line_bases : p; }

// ---------------------------------------------------------------------------------------------- BGZF
// Inflates a whole BGZF stream with `threads` workers, in batches of blocks, handing every batch's
// concatenated payload to `sink(data, n)` in file order.  Returns "" or an error message.
typedef bool (*bgzf_sink_fn)(void* user, const uint8_t* data, size_t n);
std::string bgzf_inflate_stream(const std::string& path, int threads, bgzf_sink_fn sink, void* user);

// BGZF writer (indexcov's *.bed.gz, indexcov.go:248-257): 64 KB blocks, ModTime 0, OS 0xff, EOF block.
class BgzfWriter {
public:
    explicit BgzfWriter(const std::string& path);
    ~BgzfWriter();
    bool ok() const { return f_ != nullptr; }
    void write(const char* p, size_t n);
    void close();
private:
    void flush_batch();                              // deflates the pending 64 KB blocks on all cores, writes them in order
    void* f_ = nullptr;
    std::vector<uint8_t> buf_;                       // pending uncompressed bytes (up to kBatchBlocks blocks)
};

// ---------------------------------------------------------------------------------------------- BAM
struct BamHeader {
    std::string text;
    std::vector<RefInfo> refs;
    std::vector<std::string> sample_names() const;      // distinct @RG SM values, in order of appearance
};

// What `samtools depth -Q q` counts: records with (flag & 0x704) == 0 and MAPQ >= q;
// their M/=/X blocks (adjacent blocks separated only by I/S/H/P are merged); D and N advance without counting.
struct SegmentSet {
    BamHeader header;
    std::vector<std::vector<int32_t>> start, end;        // per reference id, in record order
    int64_t n_records = 0, n_pass = 0;
};
std::string bam_decode_segments(const std::string& path, int min_mapq, int threads, int only_tid, SegmentSet& out);
std::string bam_read_header(const std::string& path, BamHeader& out);

// covstats (covstats/covstats.go:122-173): the sampled per-record values
struct CovstatsSample {
    std::vector<int32_t> read_len, insert, tmpl;
    int64_t n_unmapped = 0, k = 0, n_bad = 0, n_dup = 0, n_proper = 0;
};
std::string bam_covstats_sample(const std::string& path, int n, int skip, BamHeader& hdr, CovstatsSample& out);

// ---------------------------------------------------------------------------------------------- BAI
struct BaiIndex {
    std::vector<std::vector<uint64_t>> ioffsets;         // per reference: linear index (16 KB tiles)
    std::vector<uint64_t> mapped, unmapped;              // stats pseudo-bin 0x924a (indexcov/types.go:19,33-43)
    std::vector<uint8_t> has_stats;
    uint64_t n_no_coor = 0;
    std::vector<uint64_t> ref_beg, ref_end;              // virtual-offset range of the reference's records (min chunk begin /
                                                         // max chunk end over its bins; 0 = no records)
};
std::string bai_read(const std::string& path, BaiIndex& out);

// ---------------------------------------------------------------------------------------------- CRAI
// .crai (gzip'd text, 6 tab columns) -> per reference slices; sizes = the 16 KB pseudo-tiles indexcov uses for
// CRAM (indexcov/crai/crai.go:56-192).  crai_make_sizes returns false where the reference panics.
struct CraiSlices { std::vector<int64_t> start, span; std::vector<int32_t> bytes; };
std::string crai_read(const std::string& path, std::vector<CraiSlices>& out);
bool crai_make_sizes(const int64_t* start, const int64_t* span, const int32_t* bytes, int64_t n, std::vector<int64_t>& sizes);

// ---------------------------------------------------------------------------------------------- FAI
std::string fai_read(const std::string& path, std::vector<RefInfo>& out);

}  // namespace glhts
