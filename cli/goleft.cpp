// goleft — C++ host for the B200 engine, keeping the `goleft depth|indexcov|covstats|depthwed` command-line
// and BED/TSV output surface of brentp/goleft v0.2.6 (cmd/goleft/goleft.go:24-69).  Every number it prints is
// computed by libgoleft_b200.so (CUDA, sm_100a) through the C ABI in include/goleft_b200.h; there is no CPU
// fallback: without a GPU the subcommands fail with the library's error.
//
// Out of scope here: HTML/PNG plots, PCA columns of the .ped, --stats GC columns.  .crai input needs --fai.
#include <errno.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <array>
#include <chrono>
#include <functional>
#include <condition_variable>
#include <map>
#include <mutex>
#include <regex>
#include <string>
#include <thread>
#include <vector>
#include <zlib.h>

#include "../goleft_b200/csrc/host/hts_io.h"
#include "../goleft_b200/csrc/host/bam_feed.h"
#include "../goleft_b200/csrc/host/thread_pool.h"
#include "goleft_b200.h"

static const char* kVersion = "0.2.6";          // goleft.go:3 (the surface this build mirrors)

// ------------------------------------------------------------------------------------------------ util
[[noreturn]] static void fatal(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fputc('\n', stderr);
    exit(code);
}

static void glck(gl_ctx* ctx, int rc, const char* what) {
    if (rc != GL_OK) fatal(1, "%s: %s", what, gl_last_error(ctx));
}

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
extern "C" int glhost_pool_size(void);              // libgoleft_b200.so: size of the host thread pool (creates it)
static void glhost_pool_warm() { (void)glhost_pool_size(); }

// Go's fmt spells non-finite floats "NaN", "+Inf", "-Inf" (C prints nan / inf): a sample with no tiles on a chromosome
// gets 0/0 in its ROC and bins.out / bins.in can be x/0, and parsers of .roc / .ped must see the reference's tokens.
static std::string go_float(const char* cfmt, double v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "+Inf" : "-Inf";
    char b[64];
    snprintf(b, sizeof b, cfmt, v);
    return b;
}

static bool ends_with(const std::string& s, const std::string& suf) {
    return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
}

static bool file_exists(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0; }

// lines of a (possibly gzipped) text file; like Go's ReadString('\n') loops that break on io.EOF, a final
// line WITHOUT a trailing newline is dropped when drop_unterminated is set (depth.go:106-110,137-141)
static std::vector<std::string> read_lines(const std::string& path, bool drop_unterminated) {
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) fatal(1, "open %s: %s", path.c_str(), strerror(errno));
    std::vector<std::string> out;
    std::string cur;
    char buf[1 << 16];
    int n;
    while ((n = gzread(f, buf, sizeof buf)) > 0) {
        for (int i = 0; i < n; i++) {
            if (buf[i] == '\n') { out.push_back(cur); cur.clear(); }
            else cur.push_back(buf[i]);
        }
    }
    gzclose(f);
    if (!cur.empty() && !drop_unterminated) out.push_back(cur);
    return out;
}

// go-arg style parser (alexflint/go-arg v1.4.3): -x v, --long v, --long=v, -long v; bool flags take no value
struct Spec { std::string lng; char sht; bool is_bool; std::string* sval; bool* bval; bool required; bool seen; };
struct ArgParser {
    std::string prog;
    std::vector<Spec> specs;
    std::vector<std::string> positional;
    void add(const std::string& lng, char sht, std::string* v, bool required = false) { specs.push_back({lng, sht, false, v, nullptr, required, false}); }
    void addb(const std::string& lng, char sht, bool* v) { specs.push_back({lng, sht, true, nullptr, v, false, false}); }
    [[noreturn]] void fail(const std::string& msg) const {
        fprintf(stderr, "Usage: %s [options] ...\nerror: %s\n", prog.c_str(), msg.c_str());
        exit(255);                                         // functional-tests.sh:40-42
    }
    void parse(int argc, char** argv) {
        for (int i = 1; i < argc; i++) {
            std::string a = argv[i];
            if (a == "-h" || a == "--help") {
                printf("Usage: %s", prog.c_str());
                for (auto& s : specs) printf(" [--%s%s]", s.lng.c_str(), s.is_bool ? "" : " VALUE");
                printf(" ...\n");
                exit(0);
            }
            if (a.size() > 1 && a[0] == '-' && !(a.size() > 1 && isdigit((unsigned char)a[1]))) {
                std::string name = a.substr(a[1] == '-' ? 2 : 1), val;
                bool has_val = false;
                size_t eq = name.find('=');
                if (eq != std::string::npos) { val = name.substr(eq + 1); name = name.substr(0, eq); has_val = true; }
                Spec* sp = nullptr;
                for (auto& s : specs) if (s.lng == name || (name.size() == 1 && s.sht && s.sht == name[0])) sp = &s;
                if (!sp) fail("unknown argument " + a);
                sp->seen = true;
                if (sp->is_bool) { *sp->bval = has_val ? (val == "true" || val == "1") : true; continue; }
                if (!has_val) {
                    if (i + 1 >= argc) fail("missing value for " + a);
                    val = argv[++i];
                }
                *sp->sval = val;
            } else {
                positional.push_back(a);
            }
        }
        for (auto& s : specs) if (s.required && !s.seen) fail(s.lng + " is required");
    }
};

// depth.go:73-94: regexp "(.+?)[:\t](\d+)([\-\t])(\d+).*?" — leftmost match, lazy first group
static bool chrom_start_end(const std::string& line, std::string& chrom, long long& start, long long& end) {
    const size_t n = line.size();
    for (size_t m0 = 0; m0 < n; m0++) {
        for (size_t k = m0 + 1; k < n; k++) {
            if (line[k] != ':' && line[k] != '\t') continue;
            size_t b = k + 1;
            while (b < n && isdigit((unsigned char)line[b])) b++;
            if (b == k + 1 || b >= n || (line[b] != '-' && line[b] != '\t')) continue;
            size_t d = b + 1;
            while (d < n && isdigit((unsigned char)line[d])) d++;
            if (d == b + 1) continue;
            chrom = line.substr(m0, k - m0);
            long long s = atoll(line.c_str() + k + 1);
            if (line[b] == '-') s--;
            start = s < 0 ? 0 : s;
            end = atoll(line.c_str() + b + 1);
            return true;
        }
    }
    return false;
}

// ------------------------------------------------------------------------------------------------ depth
// `goleft depth`: depth/depth.go:103-159 builds the work list (10 Mb chunks of every .fai contig, or the BED lines), runs one
// `samtools depth -r` child per entry and concatenates the per-chunk BED files (-o: in order).  Here the unit of work is a
// contig (all its chunks in one GPU pass, run_break = step keeps the per-chunk row structure) or, in BED mode, a contig's
// BED lines; contigs are dealt to the visible GPUs longest first (gl_lpt_assign), one host thread + one gl_ctx + one BAM
// handle per GPU; the host pool (inflate + parse) is shared.  Output is always in work-list order (a valid -o output).
namespace {

struct DepthOpts {
    int W = 250, maxmean = 0, Q = 1, mincov = 4, threads = 0;
    long long step = 10000000;
    bool stats = false;
    std::string reference;
};

struct DepthRegion { std::string chrom; long long s, e; };

struct DepthJob {                                   // one contig (.fai mode) or the BED lines of one contig (BED mode)
    std::string chrom;
    long long len = 0;                              // .fai mode: contig length
    std::vector<size_t> lines;                      // BED mode: indices into the region list
    std::string hd, ca;                             // .fai mode output
    double decode_s = 0, gpu_s = 0, inflate_s = 0, parse_s = 0;
    long long records = 0, bytes_in = 0, bytes_out = 0;
    int gpu_fed = 0;                                // passes whose segments came from the GPU feeder
    bool done = false;
};

struct DepthShared {
    const DepthOpts* opt;
    std::string bam_path;
    const glhts::SegmentSet* preload = nullptr;     // BAM without an index: every contig decoded once up front
    std::vector<DepthRegion>* regions = nullptr;    // BED mode
    std::vector<std::string>* out_hd = nullptr;     // BED mode: per line
    std::vector<std::string>* out_ca = nullptr;
    std::vector<glhts::RefInfo> fa_index;
    const uint8_t* fa_map = nullptr; size_t fa_len = 0;
    std::mutex mu; std::condition_variable cv;
};

// One GPU's worker: ctx + BAM handle + buffers
struct DepthEngine {
    DepthShared& sh;
    const DepthOpts& o;
    gl_ctx* ctx = nullptr;
    gl_bam gbam;                                            // the library's BAM handle: its BamFile is the host feeder,
    glhts::BamFile& bam = gbam.f;                           // gl_bam_decode_device runs the feeder on the GPU
    std::map<std::string, int> tid_of;
    glhts::ContigSegs segs;
    const int32_t* dev_s = nullptr; const int32_t* dev_e = nullptr; int64_t dev_n = -1;   // segments left on the device by the GPU feeder (-1: none)
    std::vector<int64_t> sums; std::vector<int32_t> rstart; std::vector<uint8_t> rclass;
    std::vector<char> tbuf_hd, tbuf_ca;
    // --stats state (depth.go:191-200,246-252): chunks of a contig are parked, one gl_fasta_stats launch for all their rows
    struct Pending { std::string chrom; long long rs, re; std::vector<int64_t> sums; std::vector<int32_t> rs_; std::vector<uint8_t> rc_; std::string *hd, *ca; int64_t row0, nrows; };
    std::vector<Pending> pending;
    std::vector<int64_t> row_s, row_e;

    DepthEngine(DepthShared& s, int device) : sh(s), o(*s.opt) {
        if (gl_ctx_create(device, &ctx) != GL_OK) fatal(1, "goleft depth: %s", gl_last_error(nullptr));
        if (!sh.preload) {
            std::string e = bam.open(sh.bam_path);
            if (!e.empty()) fatal(1, "%s", e.c_str());
            for (size_t i = 0; i < bam.header.refs.size(); i++) tid_of[bam.header.refs[i].name] = (int)i;
        } else {
            for (size_t i = 0; i < sh.preload->header.refs.size(); i++) tid_of[sh.preload->header.refs[i].name] = (int)i;
        }
        rstart.resize(1 << 16); rclass.resize(1 << 16);
    }
    ~DepthEngine() { gl_ctx_destroy(ctx); }

    // segments of [beg,end) of a contig into `segs` (packed8 or sorted int32); want: 0 auto, 32 int32.
    // whole = the complete reference is wanted: the GPU feeder (BGZF inflate + record parse on the device, only compressed
    // bytes over PCIe) takes it when it can; GL_GPU_FEED=0 keeps everything on the host feeder.
    void fetch(const std::string& c, long long beg, long long end, int want, DepthJob* job, bool whole = false) {
        segs = glhts::ContigSegs();
        dev_n = -1;
        auto it = tid_of.find(c);
        if (it == tid_of.end()) return;
        static const bool gpu_feed = [] { const char* e = getenv("GL_GPU_FEED"); return !e || atoi(e) != 0; }();
        if (whole && gpu_feed && !sh.preload) {
            gl_bam_segments st;
            const int rc = gl_bam_decode_device(ctx, &gbam, it->second, o.Q, &dev_s, &dev_e, &dev_n, &st);
            if (rc == GL_OK) {
                if (job) { job->decode_s += st.wall_s; job->inflate_s += st.inflate_s; job->parse_s += st.parse_s; job->records += st.n_records;
                           job->bytes_in += st.bytes_in; job->bytes_out += st.bytes_out; job->gpu_fed++; }
                return;
            }
            dev_n = -1;
            if (rc != GL_ESTATE) glck(ctx, rc, "gl_bam_decode_device");      // GL_ESTATE: not possible on the device for this reference
        }
        if (sh.preload) {                                              // unindexed BAM: slices of the preloaded arrays (record order)
            const std::vector<int32_t>& S = sh.preload->start[(size_t)it->second];
            const std::vector<int32_t>& E = sh.preload->end[(size_t)it->second];
            segs.format = S.empty() ? 0 : 32;
            segs.start = S; segs.end = E; segs.n = (int64_t)S.size();
            return;
        }
        glhts::DecodeStats st;
        std::string e = bam.decode(it->second, beg, end, o.Q, o.threads, want, segs, &st);
        if (!e.empty()) fatal(1, "%s", e.c_str());
        if (job) { job->decode_s += st.wall_s; job->inflate_s += st.inflate_s; job->parse_s += st.parse_s; job->records += st.n_records; job->bytes_in += st.bytes_in; job->bytes_out += st.bytes_out; }
    }

    void add_segments() {
        if (dev_n >= 0) { if (dev_n > 0) glck(ctx, gl_depth_add_segments_device(ctx, dev_s, dev_e, dev_n), "gl_depth_add_segments_device"); return; }
        if (segs.format == 8 && segs.n_blocks > 0)
            glck(ctx, gl_depth_add_segments_packed8(ctx, segs.anchors.data(), segs.ds.data(), segs.len.data(), segs.n_blocks), "gl_depth_add_segments_packed8");
        else if (segs.format == 32 && segs.n > 0)
            glck(ctx, gl_depth_add_segments(ctx, segs.start.data(), segs.end.data(), segs.n), "gl_depth_add_segments");
    }

    // window sums + class runs of [rs,re) on the host (the --stats and BED paths)
    void reduce_to_host(long long rs, long long re, long long run_break, int64_t& nw, int64_t& nr) {
        glck(ctx, gl_depth_begin(ctx, rs, re), "gl_depth_begin");
        add_segments();
        glck(ctx, gl_depth_reduce(ctx, o.W, o.mincov, o.maxmean, run_break), "gl_depth_reduce");
        int32_t md = 0;
        glck(ctx, gl_depth_result_sizes(ctx, &nw, &nr, &md), "gl_depth_result_sizes");
        sums.resize((size_t)nw);
        if ((int64_t)rstart.size() < nr) { rstart.resize((size_t)nr + 1024); rclass.resize((size_t)nr + 1024); }
        glck(ctx, gl_depth_get_windows(ctx, sums.data(), nw), "gl_depth_get_windows");
        glck(ctx, gl_depth_get_runs(ctx, rstart.data(), nullptr, rclass.data(), (int64_t)rstart.size()), "gl_depth_get_runs");
    }

    void emit(const std::string& c, long long rs, long long re, const int64_t* ws, int64_t nw, const int32_t* rs_, const uint8_t* rc_, int64_t nr,
              std::string* hd_to, std::string* ca_to) {
        if (!o.stats) {
            char *hd = nullptr, *ca = nullptr;
            int64_t hl = 0, cl = 0;
            if (gl_depth_format_chunk(c.c_str(), rs, re, o.W, ws, nw, rs_, rc_, nr, &hd, &hl, &ca, &cl) != GL_OK) fatal(1, "gl_depth_format_chunk failed");
            hd_to->append(hd, (size_t)hl); ca_to->append(ca, (size_t)cl);
            gl_free_text(hd); gl_free_text(ca);
            return;
        }
        Pending pc{c, rs, re, std::vector<int64_t>(ws, ws + nw), std::vector<int32_t>(rs_, rs_ + nr), std::vector<uint8_t>(rc_, rc_ + nr), hd_to, ca_to,
                   (int64_t)row_s.size(), 0};
        int64_t n = 0;
        gl_depth_chunk_rows(rs, re, o.W, rs_, rc_, nr, nullptr, nullptr, 0, &n);
        row_s.resize(row_s.size() + (size_t)n); row_e.resize(row_s.size());
        if (gl_depth_chunk_rows(rs, re, o.W, rs_, rc_, nr, row_s.data() + pc.row0, row_e.data() + pc.row0, n, &n) != GL_OK) fatal(1, "gl_depth_chunk_rows failed");
        pc.nrows = n;
        pending.push_back(std::move(pc));
    }

    void flush_stats() {                                    // all parked chunks belong to one contig
        if (pending.empty()) return;
        const std::string& c = pending[0].chrom;
        std::vector<double> st3(row_s.size() * 3, 0.0);
        const glhts::RefInfo* ri = nullptr;
        for (const glhts::RefInfo& r : sh.fa_index) if (r.name == c) { ri = &r; break; }
        if (!ri) fprintf(stderr, "GC: unknown sequence %s\n", c.c_str());                 // faidx error text; zeros are printed (depth.go:195-199)
        else if (!row_s.empty()) {
            const int64_t rec_bytes = std::min<int64_t>((int64_t)sh.fa_len - ri->offset, glhts::fasta_position(*ri, ri->length) + 1);
            if (ri->offset < 0 || rec_bytes < 0) fatal(1, "bad .fai entry for %s", c.c_str());
            glck(ctx, gl_fasta_load(ctx, sh.fa_map + ri->offset, rec_bytes), "gl_fasta_load");
            std::vector<int64_t> ba(row_s.size()), bb(row_s.size());
            for (size_t i = 0; i < row_s.size(); i++) {
                // faidx panics on coordinates past the contig; rows are clipped to it here instead
                const int64_t s = std::min<int64_t>(std::max<int64_t>(row_s[i], 0), ri->length), e = std::min<int64_t>(std::max<int64_t>(row_e[i], s), ri->length);
                const int64_t ps = glhts::fasta_position(*ri, s), pe = glhts::fasta_position(*ri, e);
                ba[i] = std::min(ps, rec_bytes);
                bb[i] = std::max(ba[i], std::min<int64_t>(pe + ((ri->offset + pe < (int64_t)sh.fa_len) ? 1 : 0), rec_bytes));
            }
            glck(ctx, gl_fasta_stats(ctx, ba.data(), bb.data(), (int64_t)ba.size(), nullptr, st3.data()), "gl_fasta_stats");
        }
        for (const Pending& pc : pending) {
            char *hd = nullptr, *ca = nullptr;
            int64_t hl = 0, cl = 0;
            if (gl_depth_format_chunk_stats(pc.chrom.c_str(), pc.rs, pc.re, o.W, pc.sums.data(), (int64_t)pc.sums.size(), pc.rs_.data(), pc.rc_.data(),
                                            (int64_t)pc.rs_.size(), st3.data() + pc.row0 * 3, pc.nrows, &hd, &hl, &ca, &cl) != GL_OK)
                fatal(1, "gl_depth_format_chunk_stats failed");
            pc.hd->append(hd, (size_t)hl); pc.ca->append(ca, (size_t)cl);
            gl_free_text(hd); gl_free_text(ca);
        }
        pending.clear(); row_s.clear(); row_e.clear();
    }

    // ---- .fai mode: one contig.  The GPU pass covers at most 2^30-1 bases; longer contigs go in whole-chunk pieces
    //      whose texts concatenate (chunk edges are run breaks anyway).
    void run_contig(DepthJob& job) {
        const long long kMaxPass = ((1LL << 30) - 1) / o.step * o.step;
        const double t_all = now_s();
        for (long long ps = 0; ps < job.len; ps += kMaxPass) {
            const long long pe = std::min(job.len, ps + kMaxPass);
            fetch(job.chrom, ps, pe, 0, &job, ps == 0 && pe == job.len);
            const double tg = now_s();
            const bool device_text = !o.stats && job.chrom.size() <= 64;
            if (device_text && dev_n >= 0) {                                   // segments already on the device (GPU feeder)
                const int64_t nwin = (pe - 1) / o.W - ps / o.W + 1;
                const size_t need_hd = (size_t)gl_depth_text_bound(job.chrom.c_str(), nwin);
                if (tbuf_hd.size() < need_hd) tbuf_hd.resize(need_hd);
                if (tbuf_ca.size() < (size_t)(1 << 20)) tbuf_ca.resize(1 << 20);
                glck(ctx, gl_depth_begin(ctx, ps, pe), "gl_depth_begin");
                add_segments();
                glck(ctx, gl_depth_reduce(ctx, o.W, o.mincov, o.maxmean, o.step), "gl_depth_reduce");
                int64_t hl = 0, cl = 0;
                int rc = gl_depth_text(ctx, job.chrom.c_str(), tbuf_hd.data(), (int64_t)tbuf_hd.size(), &hl, tbuf_ca.data(), (int64_t)tbuf_ca.size(), &cl);
                if (rc == GL_ERANGE) {
                    tbuf_hd.resize((size_t)hl + 16); tbuf_ca.resize((size_t)cl + 16);
                    rc = gl_depth_text(ctx, job.chrom.c_str(), tbuf_hd.data(), (int64_t)tbuf_hd.size(), &hl, tbuf_ca.data(), (int64_t)tbuf_ca.size(), &cl);
                }
                glck(ctx, rc, "gl_depth_text");
                job.hd.append(tbuf_hd.data(), (size_t)hl);
                job.ca.append(tbuf_ca.data(), (size_t)cl);
            } else if (device_text) {
                const int64_t nwin = (pe - 1) / o.W - ps / o.W + 1;
                const size_t need_hd = (size_t)gl_depth_text_bound(job.chrom.c_str(), nwin);
                if (tbuf_hd.size() < need_hd) tbuf_hd.resize(need_hd);
                if (tbuf_ca.size() < (size_t)(1 << 20)) tbuf_ca.resize(1 << 20);
                int64_t hl = 0, cl = 0;
                for (int attempt = 0; attempt < 2; attempt++) {
                    int rc;
                    if (segs.format == 8)
                        rc = gl_depth_bed_region_packed8(ctx, job.chrom.c_str(), ps, pe, segs.anchors.data(), segs.ds.data(), segs.len.data(), segs.n_blocks, o.W,
                                                         o.mincov, o.maxmean, o.step, tbuf_hd.data(), (int64_t)tbuf_hd.size(), &hl, tbuf_ca.data(), (int64_t)tbuf_ca.size(), &cl);
                    else
                        rc = gl_depth_bed_region(ctx, job.chrom.c_str(), ps, pe, segs.start.data(), segs.end.data(), segs.n, o.W, o.mincov, o.maxmean, o.step,
                                                 o.threads, tbuf_hd.data(), (int64_t)tbuf_hd.size(), &hl, tbuf_ca.data(), (int64_t)tbuf_ca.size(), &cl);
                    if (rc == GL_ERANGE && attempt == 0) { tbuf_hd.resize((size_t)hl + 16); tbuf_ca.resize((size_t)cl + 16); continue; }
                    glck(ctx, rc, "gl_depth_bed_region");
                    break;
                }
                job.hd.append(tbuf_hd.data(), (size_t)hl);
                job.ca.append(tbuf_ca.data(), (size_t)cl);
            } else {
                int64_t nw = 0, nr = 0;
                reduce_to_host(ps, pe, o.step, nw, nr);
                for (long long cs = ps; cs < pe; cs += o.step) {
                    const long long ce = std::min(cs + o.step, pe);
                    const int32_t* lo = std::lower_bound(rstart.data(), rstart.data() + nr, (int32_t)cs);
                    const int32_t* hi = std::lower_bound(rstart.data(), rstart.data() + nr, (int32_t)ce);
                    emit(job.chrom, cs, ce, sums.data() + (cs / o.W - ps / o.W), (ce - 1) / o.W - cs / o.W + 1, lo, rclass.data() + (lo - rstart.data()), hi - lo,
                         &job.hd, &job.ca);
                }
                flush_stats();
            }
            job.gpu_s += now_s() - tg;
        }
        (void)t_all;
    }

    // ---- BED mode: the lines of one contig.  One pass over the span they cover (window sums + class runs, run_break 0),
    //      the clipped first/last window of every line from gl_depth_interval_sums, then the reference's rows line by line
    //      (depth.go:293-358 sees each line as its own chunk).
    void run_bed_chrom(DepthJob& job) {
        std::vector<DepthRegion>& regions = *sh.regions;
        const std::vector<size_t>& ks = job.lines;
        const int W = o.W;
        long long lo = regions[ks[0]].s, hi = regions[ks[0]].e;
        for (size_t k : ks) { lo = std::min(lo, regions[k].s); hi = std::max(hi, regions[k].e); }
        const long long span_s = lo / W * W;
        fetch(job.chrom, span_s, hi, 32, &job);                               // sorted int32: the interval sums and the per-line slices want arrays
        const double tg = now_s();
        int64_t nw = 0, nr = 0;
        auto format_one = [&](size_t k, const int64_t* ws, int64_t nw_, const int32_t* rs_, const uint8_t* rc_, int64_t nr_) {
            emit(regions[k].chrom, regions[k].s, regions[k].e, ws, nw_, rs_, rc_, nr_, &(*sh.out_hd)[k], &(*sh.out_ca)[k]);
        };
        bool batched = ks.size() > 1 && hi - span_s < (1LL << 30) - 2 * (long long)W;
        if (batched) {
            reduce_to_host(span_s, hi, 0, nw, nr);
            int32_t path = 0;
            batched = gl_depth_last_path(ctx, &path) == GL_OK && (path == 1 || segs.n == 0);   // the interval sums want the fused path's cell index
        }
        if (!batched) {
            // per-line passes over the slice of segments that can reach the line (the arrays are sorted by start)
            glhts::ContigSegs all;
            all.start.swap(segs.start); all.end.swap(segs.end); all.n = segs.n; all.format = segs.format; all.max_len = segs.max_len;
            const int32_t maxlen = std::max<int32_t>(all.max_len, 1);
            const bool sorted = !sh.preload;                                   // the index-guided decoder sorts; preloaded arrays are in record order
            for (size_t k : ks) {
                if (regions[k].e - regions[k].s >= (1LL << 30)) fatal(1, "BED region longer than 2^30-1 bases: %s", regions[k].chrom.c_str());
                size_t a = 0, b = (size_t)all.n;
                if (sorted && all.n > 0) {
                    a = (size_t)(std::lower_bound(all.start.begin(), all.start.end(), (int32_t)std::max<long long>(regions[k].s - maxlen, INT32_MIN)) - all.start.begin());
                    b = (size_t)(std::lower_bound(all.start.begin(), all.start.end(), (int32_t)std::min<long long>(regions[k].e, INT32_MAX)) - all.start.begin());
                }
                segs.format = b > a ? 32 : 0;
                segs.start.assign(all.start.begin() + (long)a, all.start.begin() + (long)b);
                segs.end.assign(all.end.begin() + (long)a, all.end.begin() + (long)b);
                segs.n = (int64_t)(b - a);
                reduce_to_host(regions[k].s, regions[k].e, 0, nw, nr);
                format_one(k, sums.data(), nw, rstart.data(), rclass.data(), nr);
            }
            flush_stats();
            job.gpu_s += now_s() - tg;
            return;
        }
        std::vector<int64_t> rsum, edge;
        std::vector<int32_t> ia, ib, rrs;
        std::vector<uint8_t> rrc;
        for (size_t k : ks) {                                            // clipped edge windows
            const long long s = regions[k].s, e = regions[k].e, w0 = s / W, w1 = (e - 1) / W;
            if (s % W != 0 || (w0 == w1 && e != (w0 + 1) * W)) { ia.push_back((int32_t)s); ib.push_back((int32_t)std::min(e, (w0 + 1) * W)); }
            if (w1 > w0 && e != (w1 + 1) * W) { ia.push_back((int32_t)(w1 * W)); ib.push_back((int32_t)e); }
        }
        edge.resize(ia.size());
        if (!ia.empty()) glck(ctx, gl_depth_interval_sums(ctx, ia.data(), ib.data(), (int64_t)ia.size(), edge.data()), "gl_depth_interval_sums");
        size_t ei = 0;
        const long long sw0 = span_s / W;
        for (size_t k : ks) {
            const long long s = regions[k].s, e = regions[k].e, w0 = s / W, w1 = (e - 1) / W;
            rsum.assign(sums.begin() + (w0 - sw0), sums.begin() + (w1 - sw0) + 1);
            if (s % W != 0 || (w0 == w1 && e != (w0 + 1) * W)) rsum[0] = edge[ei++];
            if (w1 > w0 && e != (w1 + 1) * W) rsum.back() = edge[ei++];
            const int32_t* rb = rstart.data();
            const int32_t* first = std::upper_bound(rb, rb + nr, (int32_t)s);          // first run starting after s
            const int32_t* last = std::lower_bound(rb, rb + nr, (int32_t)e);
            rrs.assign(1, (int32_t)s);
            rrc.assign(1, rclass[(size_t)(first - rb) - 1]);                           // run_start[0] == span_s <= s
            rrs.insert(rrs.end(), first, last);
            rrc.insert(rrc.end(), rclass.data() + (first - rb), rclass.data() + (last - rb));
            format_one(k, rsum.data(), (int64_t)rsum.size(), rrs.data(), rrc.data(), (int64_t)rrs.size());
        }
        flush_stats();
        job.gpu_s += now_s() - tg;
    }
};

}  // namespace

static int cmd_depth(int argc, char** argv) {
    std::string w = "250", m = "0", q = "1", chrom, mincov = "4", reference, procs = "0", bed, prefix, gpus = "0";
    bool ordered = false, stats = false, timing = false;
    ArgParser ap;
    ap.prog = "goleft depth";
    ap.add("windowsize", 'w', &w); ap.add("maxmeandepth", 'm', &m); ap.addb("ordered", 'o', &ordered);
    ap.add("q", 'Q', &q); ap.add("chrom", 'c', &chrom); ap.add("mincov", 0, &mincov); ap.addb("stats", 's', &stats);
    ap.add("reference", 'r', &reference); ap.add("processes", 'p', &procs); ap.add("bed", 'b', &bed);
    ap.add("prefix", 0, &prefix, true);
    ap.add("gpus", 0, &gpus);                       // extension: GPUs to use (0 = every visible one)
    ap.addb("timing", 0, &timing);                  // extension: per-phase wall clock on stderr
    ap.parse(argc, argv);
    if (ap.positional.empty()) ap.fail("bam is required");
    const std::string bam = ap.positional[0];
    DepthOpts o;
    o.W = atoi(w.c_str()); o.maxmean = atoi(m.c_str()); o.Q = atoi(q.c_str()); o.mincov = atoi(mincov.c_str());
    o.stats = stats; o.reference = reference;
    if (o.W <= 0) ap.fail("windowsize must be > 0");
    o.threads = atoi(procs.c_str());
    if (o.threads < 0) o.threads = 0;               // 0 = every thread of the host pool
    o.step = std::max(1LL, 10000000LL / o.W) * o.W;                          // depth.go:132
    const double t_start = now_s();

    // work list (depth.go:103-159)
    std::vector<DepthRegion> regions;
    std::vector<DepthJob> jobs;
    if (!bed.empty()) {
        for (const std::string& ln : read_lines(bed, true)) {
            if (ln.empty()) continue;
            DepthRegion r;
            if (!chrom_start_end(ln, r.chrom, r.s, r.e)) fatal(1, "couldn't get region from line%s", ln.c_str());
            regions.push_back(r);
        }
        std::map<std::string, size_t> job_of;
        for (size_t k = 0; k < regions.size(); k++) {
            if (regions[k].e <= regions[k].s) continue;
            auto ins = job_of.emplace(regions[k].chrom, jobs.size());
            if (ins.second) { jobs.emplace_back(); jobs.back().chrom = regions[k].chrom; }
            DepthJob& j = jobs[ins.first->second];
            j.lines.push_back(k);
            j.len = std::max(j.len, regions[k].e);
        }
    } else {
        for (const std::string& ln : read_lines(reference + ".fai", true)) {
            size_t t1 = ln.find('\t');
            if (t1 == std::string::npos) continue;
            std::string c = ln.substr(0, t1);
            if (!chrom.empty() && c != chrom) continue;
            const long long len = atoll(ln.c_str() + t1 + 1);
            if (len <= 0) continue;
            jobs.emplace_back();
            jobs.back().chrom = c; jobs.back().len = len;
        }
    }

    DepthShared sh;
    sh.opt = &o;
    sh.bam_path = bam;
    std::vector<std::string> out_hd(regions.size()), out_ca(regions.size());
    sh.regions = &regions; sh.out_hd = &out_hd; sh.out_ca = &out_ca;
    // an unindexed BAM cannot be seeked: decode it once, whole (the reference's `samtools depth -r` would refuse it)
    glhts::SegmentSet preload;
    {
        glhts::BamFile probe;
        std::string e = probe.open(bam);
        if (!e.empty()) fatal(1, "%s", e.c_str());
        bool usable = probe.has_index;
        // a stub index (stats bin but an empty / all-zero linear index for a reference that has reads) cannot be seeked with
        for (size_t r = 0; usable && r < probe.bai.ioffsets.size(); r++) {
            if (!probe.bai.has_stats[r] || probe.bai.mapped[r] + probe.bai.unmapped[r] == 0) continue;
            bool any = false;
            for (uint64_t v : probe.bai.ioffsets[r]) if (v) { any = true; break; }
            usable = any;
        }
        if (!usable) {
            fprintf(stderr, "goleft depth: no usable .bai next to %s: decoding the whole file (index it for -c / --bed to read only what they need)\n", bam.c_str());
            int thr = o.threads > 0 ? o.threads : (int)std::max(1u, std::thread::hardware_concurrency());
            int only = -1;
            if (!chrom.empty()) for (size_t i = 0; i < probe.header.refs.size(); i++) if (probe.header.refs[i].name == chrom) only = (int)i;
            e = glhts::bam_decode_segments(bam, o.Q, thr, only, preload);
            if (!e.empty()) fatal(1, "%s", e.c_str());
            sh.preload = &preload;
        }
    }
    if (stats) {
        if (reference.empty()) fatal(1, "goleft depth: --stats needs --reference");
        std::string e2 = glhts::fai_read(reference + ".fai", sh.fa_index);
        if (!e2.empty()) fatal(1, "%s", e2.c_str());
        int fd = open(reference.c_str(), O_RDONLY);
        struct stat st;
        if (fd < 0 || fstat(fd, &st) != 0) fatal(1, "cannot open %s", reference.c_str());
        sh.fa_len = (size_t)st.st_size;
        if (sh.fa_len) {
            void* mm = mmap(nullptr, sh.fa_len, PROT_READ, MAP_PRIVATE, fd, 0);
            if (mm == MAP_FAILED) fatal(1, "cannot mmap %s", reference.c_str());
            sh.fa_map = static_cast<const uint8_t*>(mm);
        }
        close(fd);
    }

    const std::string sfx = chrom.empty() ? "" : "." + chrom;               // depth.go:378-389
    FILE* fca = fopen((prefix + sfx + ".callable.bed").c_str(), "w");
    FILE* fhd = fopen((prefix + sfx + ".depth.bed").c_str(), "w");
    if (!fca || !fhd) fatal(1, "cannot create output files with prefix %s", prefix.c_str());

    // ---- GPUs and placement
    int n_dev = 0;
    if (gl_device_count(&n_dev) != GL_OK || n_dev < 1) fatal(1, "goleft depth: %s", gl_last_error(nullptr));
    int G = atoi(gpus.c_str());
    const bool oversub = getenv("GL_OVERSUBSCRIBE") != nullptr;               // testing aid: --gpus N workers on fewer devices (worker g -> device g % n_dev)
    if (G <= 0 || (G > n_dev && !oversub)) G = n_dev;
    G = (int)std::min<size_t>((size_t)G, std::max<size_t>(jobs.size(), 1));
    std::vector<int32_t> gpu_of(jobs.size(), 0);
    {
        std::vector<int64_t> wts(jobs.size());
        for (size_t i = 0; i < jobs.size(); i++) wts[i] = jobs[i].len;
        if (!jobs.empty() && gl_lpt_assign(wts.data(), (int32_t)jobs.size(), G, gpu_of.data(), nullptr) != GL_OK) fatal(1, "gl_lpt_assign failed");
    }
    auto worker = [&](int g) {
        int node = -1, ncpu = 0;
        if (G > 1) gl_bind_numa_for_device(g % n_dev, 0, 1, &node, &ncpu);    // this feeder thread next to its GPU (the pool keeps all cores)
        DepthEngine eng(sh, g % n_dev);
        for (size_t i = 0; i < jobs.size(); i++) {
            if (gpu_of[i] != g) continue;
            if (bed.empty()) eng.run_contig(jobs[i]); else eng.run_bed_chrom(jobs[i]);
            { std::lock_guard<std::mutex> lk(sh.mu); jobs[i].done = true; }
            sh.cv.notify_all();
        }
    };
    glhost_pool_warm();                                                       // create the host pool before the feeder threads narrow their affinity
    std::vector<std::thread> th;
    for (int g = 0; g < G; g++) th.emplace_back(worker, g);
    // ---- ordered output as the jobs finish
    if (bed.empty()) {
        for (size_t i = 0; i < jobs.size(); i++) {
            { std::unique_lock<std::mutex> lk(sh.mu); sh.cv.wait(lk, [&] { return jobs[i].done; }); }
            fwrite(jobs[i].hd.data(), 1, jobs[i].hd.size(), fhd);
            fwrite(jobs[i].ca.data(), 1, jobs[i].ca.size(), fca);
            std::string().swap(jobs[i].hd); std::string().swap(jobs[i].ca);
        }
    }
    for (auto& t : th) t.join();
    if (!bed.empty()) {
        for (size_t k = 0; k < regions.size(); k++) {
            fwrite(out_hd[k].data(), 1, out_hd[k].size(), fhd);
            fwrite(out_ca[k].data(), 1, out_ca[k].size(), fca);
        }
    }
    fclose(fca); fclose(fhd);
    if (sh.fa_map) munmap(const_cast<uint8_t*>(sh.fa_map), sh.fa_len);
    if (timing) {
        double dec = 0, gpu = 0, inf = 0, par = 0; long long rec = 0, bin = 0, bout = 0; int fed = 0;
        for (const DepthJob& j : jobs) { fed += j.gpu_fed; dec += j.decode_s; gpu += j.gpu_s; inf += j.inflate_s; par += j.parse_s; rec += j.records; bin += j.bytes_in; bout += j.bytes_out; }
        fprintf(stderr, "{\"goleft_depth_timing\": {\"wall_s\": %.4f, \"gpus\": %d, \"jobs\": %zu, \"decode_wall_s_sum\": %.4f, \"gpu_call_s_sum\": %.4f, "
                        "\"inflate_thread_s_sum\": %.4f, \"parse_thread_s_sum\": %.4f, "
                        "\"records\": %lld, \"bgzf_bytes_in\": %lld, \"bgzf_bytes_out\": %lld, \"host_threads\": %d, \"gpu_fed_passes\": %d}}\n",
                now_s() - t_start, G, jobs.size(), dec, gpu, inf, par, rec, bin, bout, glhost_pool_size(), fed);
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ indexcov
// indexcov.go:213-246
static std::string short_name(const std::string& b, bool from_filename, const glhts::BamHeader* hdr) {
    if (!from_filename && hdr) {
        std::vector<std::string> sm = hdr->sample_names();
        if (sm.size() > 1) fatal(1, "bam reagroup: more than one RG for %s", b.c_str());
        if (sm.size() == 1) return sm[0];
    }
    std::string v = b.substr(b.find_last_of('/') == std::string::npos ? 0 : b.find_last_of('/') + 1);
    std::vector<std::string> parts;
    size_t p = 0;
    for (;;) { size_t d = v.find('.', p); parts.push_back(v.substr(p, d == std::string::npos ? std::string::npos : d - p)); if (d == std::string::npos) break; p = d + 1; }
    if (parts.size() <= 2) return parts[0];
    std::string out = parts[0];
    for (size_t i = 1; i + 1 < parts.size(); i++) out += "-" + parts[i];
    return out;
}

// indexcov.go:530-547
static bool same_chrom(const std::vector<std::string>& as, const std::string& b) {
    for (const std::string& a : as) {
        if (a == b) return true;
        std::string na = a;
        if (a.compare(0, 3, "chr") == 0) na = a.substr(3);
        else if (b.compare(0, 3, "chr") == 0) na = "chr" + a;
        if (na == b) return true;
    }
    return false;
}

// indexcov.go:957-991 for one sample
static double get_cn(const float* d, size_t n) {
    std::vector<float> tmp;
    size_t lows = 0;
    for (size_t i = 0; i < n; i++) if (d[i] != 0) { tmp.push_back(d[i]); if (d[i] < 0.02f) lows++; }
    if (tmp.empty()) return -0.1;
    std::sort(tmp.begin(), tmp.end());
    const double pLo = (double)lows / (double)n;
    const float* t = tmp.data();
    size_t tk = tmp.size();
    if (pLo > 0.3) { t += lows; tk -= lows; }
    if (tk == 0) return 0;
    return (double)(2.0f * t[(size_t)((double)tk * 0.4)]);
}

struct IcRef { std::string name; long long len; int id; };

template <class T> struct RawBuf {                          // malloc'ed, uninitialised, sized once
    T* p = nullptr; size_t n = 0;
    RawBuf() = default;
    RawBuf(const RawBuf&) = delete;
    RawBuf& operator=(const RawBuf&) = delete;
    ~RawBuf() { free(p); }
    void resize(size_t m) { free(p); p = static_cast<T*>(malloc(std::max<size_t>(m, 1) * sizeof(T))); n = m; if (!p) fatal(1, "out of memory (%zu bytes)", m * sizeof(T)); }
    T* data() { return p; }
    T& operator[](size_t i) { return p[i]; }
};

static int cmd_indexcov(int argc, char** argv) {
    std::string dir, exclude = "^chrEBV$|^NC|_random$|Un_|^HLA\\-|_alt$|hap\\d$", sex = "X,Y", chrom, fai, gpus = "0";
    bool includegl = false, extranorm = false;
    ArgParser ap;
    ap.prog = "goleft indexcov";
    ap.add("directory", 'd', &dir, true); ap.addb("includegl", 'e', &includegl); ap.add("excludepatt", 0, &exclude);
    ap.add("sex", 'X', &sex); ap.add("chrom", 'c', &chrom); ap.add("fai", 'f', &fai); ap.addb("extranormalize", 'n', &extranorm);
    ap.add("gpus", 0, &gpus);                       // extension: GPUs to use (0 = every visible one); samples are sharded over them
    ap.parse(argc, argv);
    std::vector<std::string> bams = ap.positional;
    if (bams.empty()) ap.fail("bam is required");
    std::vector<std::string> sexes_wanted;
    { size_t p = 0; for (;;) { size_t c = sex.find(',', p); std::string t = sex.substr(p, c == std::string::npos ? std::string::npos : c - p); if (!t.empty()) sexes_wanted.push_back(t); if (c == std::string::npos) break; p = c + 1; } }
    std::regex excl(exclude);
    mkdir(dir.c_str(), 0755);

    // references (indexcov.go:344-374): BAM header of the first input, or the .fai
    std::vector<IcRef> refs;
    if (ends_with(bams[0], ".bam")) {
        glhts::BamHeader h;
        std::string e = glhts::bam_read_header(bams[0], h);
        if (!e.empty()) fatal(1, "%s", e.c_str());
        for (size_t i = 0; i < h.refs.size(); i++) refs.push_back({h.refs[i].name, h.refs[i].length, (int)i});
        if (!chrom.empty()) {                                               // RefsFromBam appends the named chrom (:336-338)
            std::string c = chrom.compare(0, 3, "chr") == 0 ? chrom.substr(3) : chrom;
            bool found = false;
            for (size_t i = 0; i < h.refs.size() && !found; i++) {
                const std::string& n = h.refs[i].name;
                if (c == n || (n.compare(0, 3, "chr") == 0 && c == n.substr(3))) { refs.push_back({n, h.refs[i].length, (int)i}); found = true; }
            }
            if (!found) fatal(1, "indexcov: chromosome: %s not found", chrom.c_str());
        }
    } else if (!fai.empty()) {
        std::vector<glhts::RefInfo> r;
        std::string e = glhts::fai_read(fai, r);
        if (!e.empty()) fatal(1, "error opening fai: %s. Is is present?", fai.c_str());
        for (auto& x : r) if (chrom.empty() || x.name == chrom) refs.push_back({x.name, x.length, (int)refs.size()});
        if (refs.empty()) fatal(1, "ERROR: didn't find any usable chromosomes in %s", fai.c_str());
    } else {
        fatal(1, "indexcov: since no .fai was specified, expected input to be a list of bams");
    }

    // indexes (indexcov.go:471-525)
    const bool ic_timing = getenv("GL_TIMING") != nullptr;               // phase wall clocks on stderr
    double ic_t0 = now_s(), ic_tp = ic_t0;
    auto ic_mark = [&](const char* what) { if (ic_timing) { const double t = now_s(); fprintf(stderr, "[indexcov timing] %-28s %.3f s\n", what, t - ic_tp); ic_tp = t; } };
    const size_t S = bams.size();
    std::vector<std::string> names(S);
    std::vector<glhts::BaiIndex> idx(S);
    std::vector<std::vector<glhts::CraiSlices>> crai(S);
    // the reference reads the indexes on 8 goroutines (indexcov.go:417-434); here every pool thread takes files.  The first
    // error in INPUT order is the one reported, whatever thread met it.
    std::vector<std::string> read_err(S);
    glhost::ThreadPool::global().run((int64_t)S, [&](int64_t ii, int) {
        const size_t i = (size_t)ii;
        const std::string& b = bams[i];
        if (ends_with(b, ".crai")) {                                       // indexcov.go:474-496
            read_err[i] = glhts::crai_read(b, crai[i]);
            names[i] = short_name(b, true, nullptr);
            return;
        }
        std::string p = ends_with(b, ".bai") ? b : b + ".bai";
        if (!file_exists(p)) p = b.substr(0, b.size() - 4) + (ends_with(b, ".bai") ? "" : ".bai");
        read_err[i] = glhts::bai_read(p, idx[i]);
        if (!read_err[i].empty()) return;
        if (ends_with(b, ".bai")) names[i] = short_name(b, true, nullptr);
        else { glhts::BamHeader h; read_err[i] = glhts::bam_read_header(b, h); if (read_err[i].empty()) names[i] = short_name(b, false, &h); }
    });
    for (size_t i = 0; i < S; i++) if (!read_err[i].empty()) fatal(1, "%s", read_err[i].c_str());

    ic_mark("read indexes");
    // ---- GPUs (samples are independent through I5 -> sample-sharded over the GPUs; the rows need every sample
    //      of a tile, and they are assembled on the host, so the "gather" is each GPU's D2H into its slice of the cohort arrays)
    int n_dev = 0;
    if (gl_device_count(&n_dev) != GL_OK || n_dev < 1) fatal(1, "goleft indexcov: %s", gl_last_error(nullptr));
    int G = atoi(gpus.c_str());
    const bool oversub = getenv("GL_OVERSUBSCRIBE") != nullptr;               // testing aid: --gpus N shards on fewer devices (shard g -> device g % n_dev)
    if (G <= 0 || (G > n_dev && !oversub)) G = n_dev;
    fprintf(stderr, "indexcov: running on %zu indexes\n", bams.size());

    // ---- I1 for the whole cohort (indexcov/types.go:45-82): every sample's linear-index offsets go up once, one descriptor per
    //      (sample, reference with >= 2 entries); CRAM samples bring host-made pseudo-tile sizes
    std::vector<int64_t> sample_ptr(1, 0);
    std::vector<std::vector<int64_t>> size_ptr(S);                          // per sample: CSR over its refs
    std::vector<uint64_t> mapped(S, 0), unmapped(S, 0);
    std::vector<uint64_t> voff_all;
    std::vector<int64_t> desc_voff, desc_size;
    std::vector<int32_t> desc_n;
    std::vector<size_t> desc_begin(S + 1, 0), voff_begin(S + 1, 0);         // per sample: its first descriptor / first offset
    std::vector<std::vector<int64_t>> crai_sizes(S);                        // CRAM samples: the values themselves
    for (size_t i = 0; i < S; i++) {
        const int64_t base = sample_ptr.back();
        desc_begin[i] = desc_n.size(); voff_begin[i] = voff_all.size();
        if (ends_with(bams[i], ".crai")) {                                  // CRAM: interpolated pseudo-tiles (crai.go:56-127)
            const size_t nr = crai[i].size();
            std::vector<int64_t> sp(nr + 1, 0), all;
            for (size_t r = 0; r < nr; r++) {
                std::vector<int64_t> one;
                if (!glhts::crai_make_sizes(crai[i][r].start.data(), crai[i][r].span.data(), crai[i][r].bytes.data(),
                                            (int64_t)crai[i][r].start.size(), one)) fatal(2, "panic: tilewidth logic error");
                all.insert(all.end(), one.begin(), one.end());
                sp[r + 1] = sp[r] + (int64_t)one.size();
            }
            if (sp[nr] < 1) fatal(1, "indexcov: no usable chromsomes in bam: %s", bams[i].c_str());
            sample_ptr.push_back(base + sp[nr]);
            crai_sizes[i] = std::move(all);
            size_ptr[i] = sp;
            continue;
        }
        const size_t nr = idx[i].ioffsets.size();
        std::vector<int64_t> sp(nr + 1, 0);
        for (size_t r = 0; r < nr; r++) {
            const std::vector<uint64_t>& io = idx[i].ioffsets[r];
            mapped[i] += idx[i].mapped[r]; unmapped[i] += idx[i].unmapped[r];
            sp[r + 1] = sp[r] + (io.size() >= 2 ? (int64_t)io.size() - 1 : 0);          // types.go:68-72
            if (io.size() >= 2) {
                desc_voff.push_back((int64_t)voff_all.size()); desc_n.push_back((int32_t)io.size()); desc_size.push_back(base + sp[r]);
                voff_all.insert(voff_all.end(), io.begin(), io.end());
            }
        }
        if (sp[nr] < 1) fatal(1, "indexcov: no usable chromsomes in bam: %s", bams[i].c_str());   // indexcov.go:100-102
        sample_ptr.push_back(base + sp[nr]);
        size_ptr[i] = sp;
    }
    desc_begin[S] = desc_n.size(); voff_begin[S] = voff_all.size();
    const int64_t total = sample_ptr.back();
    G = (int)std::min<size_t>((size_t)G, S);
    // contiguous sample ranges with about total/G tiles each
    std::vector<size_t> cut((size_t)G + 1, S);
    cut[0] = 0;
    for (int g = 1; g < G; g++) {
        const int64_t want = total / G * g;
        cut[(size_t)g] = (size_t)(std::lower_bound(sample_ptr.begin(), sample_ptr.end(), want) - sample_ptr.begin());
        cut[(size_t)g] = std::min(std::max(cut[(size_t)g], cut[(size_t)g - 1]), S);
    }

    const std::string base = dir + "/" + dir.substr(dir.find_last_of('/') == std::string::npos ? 0 : dir.find_last_of('/') + 1) + "-indexcov";
    // the references that are reported (exclude pattern, indexcov.go:637-646)
    std::vector<const IcRef*> kept;
    {
        int n_reported = 0;
        for (const IcRef& ref : refs) {
            if (!exclude.empty() && std::regex_search(ref.name, excl)) {
                if (n_reported < 10) { fprintf(stderr, "indexcv: excluding chromosome: %s because of exclude-pattern: %s\n", ref.name.c_str(), exclude.c_str()); if (++n_reported == 10) fprintf(stderr, "not reporting further skipped chromosomes\n"); }
                continue;
            }
            kept.push_back(&ref);
        }
    }
    // cohort-sized arrays that the GPUs overwrite completely: not zero-filled (a std::vector's constructor page-faulted through
    // 0.8 GB on one thread: 0.6 s of the 313-sample run)
    std::vector<double> med(S);
    RawBuf<float> dep; dep.resize((size_t)total);
    RawBuf<uint8_t> tok_all;
    RawBuf<int32_t> counts_all;
    RawBuf<int64_t> b4_all;
    if (!extranorm) { tok_all.resize((size_t)total * 10 + 16); counts_all.resize(kept.size() * S * GL_INDEXCOV_SLOTS); b4_all.resize(kept.size() * S * 4); }

    std::vector<gl_ctx*> ctxs((size_t)G, nullptr);
    // phase A (per GPU, its samples): I1 + I2 + I3 on the device, medians + depths home; the depths stay resident for phase B
    struct Shard { int64_t* d_sizes = nullptr; float* d_dep = nullptr; int64_t tiles = 0; };
    std::vector<Shard> shard((size_t)G);
    auto run_on_gpus = [&](const std::function<void(int)>& fn) {
        if (G == 1) { fn(0); return; }
        std::vector<std::thread> th;
        for (int g = 0; g < G; g++) th.emplace_back([&, g] { int node = -1, ncpu = 0; gl_bind_numa_for_device(g % n_dev, 0, 1, &node, &ncpu); fn(g); });
        for (auto& t : th) t.join();
    };
    glhost_pool_warm();
    ic_mark("host arrays");
    std::vector<std::array<double, 4>> gpu_t((size_t)G);                     // per GPU: ctx, upload + I1, I2 + I3, D2H
    run_on_gpus([&](int g) {
        gl_ctx* c = nullptr;
        double tg0 = now_s();
        if (gl_ctx_create(g % n_dev, &c) != GL_OK) fatal(1, "goleft indexcov: %s", gl_last_error(nullptr));
        ctxs[(size_t)g] = c;
        gpu_t[(size_t)g] = {now_s() - tg0, 0, 0, 0};
        tg0 = now_s();
        const size_t a = cut[(size_t)g], b = cut[(size_t)g + 1];
        const int64_t t0 = sample_ptr[a], tiles = sample_ptr[b] - t0;
        Shard& sd = shard[(size_t)g];
        sd.tiles = tiles;
        if (b == a) return;
        auto dalloc = [&](size_t bytes) { void* p = nullptr; glck(c, gl_dev_alloc(c, (int64_t)std::max<size_t>(bytes, 16), &p), "gl_dev_alloc"); return p; };
        auto up = [&](const void* h, size_t bytes) { void* p = dalloc(bytes); if (bytes) glck(c, gl_memcpy_h2d(c, p, h, (int64_t)bytes), "gl_memcpy_h2d"); return p; };
        sd.d_sizes = static_cast<int64_t*>(dalloc((size_t)tiles * 8));
        const size_t d0 = desc_begin[a], d1 = desc_begin[b], v0 = voff_begin[a], v1 = voff_begin[b];
        if (d1 > d0) {
            std::vector<int64_t> dv(desc_voff.begin() + (long)d0, desc_voff.begin() + (long)d1), ds(desc_size.begin() + (long)d0, desc_size.begin() + (long)d1);
            for (auto& x : dv) x -= (int64_t)v0;
            for (auto& x : ds) x -= t0;
            void* d_voff = up(voff_all.data() + v0, (v1 - v0) * 8);
            void* d_dv = up(dv.data(), dv.size() * 8);
            void* d_dn = up(desc_n.data() + d0, (d1 - d0) * 4);
            void* d_ds = up(ds.data(), ds.size() * 8);
            const int rc = gl_indexcov_sizes_batch_device(c, static_cast<const uint64_t*>(d_voff), static_cast<const int64_t*>(d_dv),
                                                          static_cast<const int32_t*>(d_dn), static_cast<const int64_t*>(d_ds), (int64_t)(d1 - d0), sd.d_sizes);
            if (rc == GL_ERANGE) fatal(2, "panic: expected positive change in vOffset");                 // types.go:75-77
            glck(c, rc, "gl_indexcov_sizes_batch_device");
            gl_dev_free(c, d_voff); gl_dev_free(c, d_dv); gl_dev_free(c, d_dn); gl_dev_free(c, d_ds);
        }
        for (size_t i = a; i < b; i++)
            if (!crai_sizes[i].empty())
                glck(c, gl_memcpy_h2d(c, sd.d_sizes + (sample_ptr[i] - t0), crai_sizes[i].data(), (int64_t)crai_sizes[i].size() * 8), "gl_memcpy_h2d");
        glck(c, gl_sync(c), "gl_sync");
        gpu_t[(size_t)g][1] = now_s() - tg0; tg0 = now_s();
        // I2+I3 for the shard in one kernel, on the resident sizes (indexcov.go:83-151)
        std::vector<int64_t> sp(sample_ptr.begin() + (long)a, sample_ptr.begin() + (long)b + 1);
        for (auto& x : sp) x -= t0;
        void* d_sp = up(sp.data(), sp.size() * 8);
        double* d_med = static_cast<double*>(dalloc((b - a) * 8));
        sd.d_dep = static_cast<float*>(dalloc((size_t)tiles * 4));
        glck(c, gl_indexcov_cohort_device(c, sd.d_sizes, static_cast<const int64_t*>(d_sp), (int32_t)(b - a), d_med, sd.d_dep), "gl_indexcov_cohort_device");
        glck(c, gl_sync(c), "gl_sync");
        gpu_t[(size_t)g][2] = now_s() - tg0; tg0 = now_s();
        glck(c, gl_memcpy_d2h(c, med.data() + a, d_med, (int64_t)(b - a) * 8), "gl_memcpy_d2h");
        glck(c, gl_memcpy_d2h(c, dep.data() + t0, sd.d_dep, tiles * 4), "gl_memcpy_d2h");
        gpu_t[(size_t)g][3] = now_s() - tg0;
        gl_dev_free(c, sd.d_sizes); gl_dev_free(c, d_sp); gl_dev_free(c, d_med);
        sd.d_sizes = nullptr;
    });
    std::vector<uint64_t>().swap(voff_all);
    if (ic_timing)
        for (int g = 0; g < G; g++)
            fprintf(stderr, "[indexcov timing]   gpu %d: gl_ctx_create %.3f s, upload + I1 %.3f s, I2 + I3 %.3f s, D2H %.3f s\n", g, gpu_t[(size_t)g][0],
                    gpu_t[(size_t)g][1], gpu_t[(size_t)g][2], gpu_t[(size_t)g][3]);
    ic_mark("I1 + I2 + I3 (GPU, + D2H)");

    glhts::BgzfWriter bgz(base + ".bed.gz");
    FILE* roc = fopen((base + ".roc").c_str(), "w");
    if (!bgz.ok() || !roc) fatal(1, "cannot create outputs in %s", dir.c_str());
    { std::string h = "#chrom\tstart\tend"; for (auto& n : names) h += "\t" + n; h += "\n"; bgz.write(h.data(), h.size()); }

    // ---- phase B.  Without -n the values never change after I3: the "%.3g" token of EVERY value (I6) and the slot histograms
    //      + counters of EVERY (reference, sample) pair (I4+I5) come from two launches per GPU over its resident depths.
    //      `longest` of a reference is over all samples (and needs every median), hence the barrier between the phases.
    if (!extranorm) {
        const size_t nq = kept.size();
        std::vector<int64_t> longest_q(nq, 0);
        for (size_t q = 0; q < nq; q++)
            for (size_t k = 0; k < S; k++)
                if (kept[q]->id + 1 < (int)size_ptr[k].size() && med[k] != 0)
                    longest_q[q] = std::max(longest_q[q], size_ptr[k][kept[q]->id + 1] - size_ptr[k][kept[q]->id]);
        run_on_gpus([&](int g) {
            gl_ctx* c = ctxs[(size_t)g];
            const size_t a = cut[(size_t)g], b = cut[(size_t)g + 1], Sg = b - a;
            Shard& sd = shard[(size_t)g];
            if (Sg == 0) return;
            const int64_t t0 = sample_ptr[a], tiles = sd.tiles;
            auto dalloc = [&](size_t bytes) { void* p = nullptr; glck(c, gl_dev_alloc(c, (int64_t)std::max<size_t>(bytes, 16), &p), "gl_dev_alloc"); return p; };
            auto up = [&](const void* h, size_t bytes) { void* p = dalloc(bytes); if (bytes) glck(c, gl_memcpy_h2d(c, p, h, (int64_t)bytes), "gl_memcpy_h2d"); return p; };
            uint8_t* d_tok = static_cast<uint8_t*>(dalloc((size_t)tiles * 10));
            if (tiles) glck(c, gl_format_g3_device(c, sd.d_dep, tiles, d_tok), "gl_format_g3_device");
            if (tiles) glck(c, gl_memcpy_d2h(c, tok_all.data() + (size_t)t0 * 10, d_tok, tiles * 10), "gl_memcpy_d2h");
            gl_dev_free(c, d_tok);
            const size_t nseg = nq * Sg;
            if (nseg) {
                std::vector<int64_t> seg_start(nseg, 0), seg_len(nseg, 0), seg_longest(nseg, 0);
                for (size_t q = 0; q < nq; q++)
                    for (size_t k = a; k < b; k++) {
                        const size_t j = q * Sg + (k - a);
                        if (kept[q]->id + 1 < (int)size_ptr[k].size() && med[k] != 0) {
                            seg_start[j] = sample_ptr[k] - t0 + size_ptr[k][kept[q]->id];
                            seg_len[j] = size_ptr[k][kept[q]->id + 1] - size_ptr[k][kept[q]->id];
                        }
                        seg_longest[j] = longest_q[q];
                    }
                void* d_ss = up(seg_start.data(), nseg * 8);
                void* d_sl = up(seg_len.data(), nseg * 8);
                void* d_lg = up(seg_longest.data(), nseg * 8);
                int32_t* d_c = static_cast<int32_t*>(dalloc(nseg * GL_INDEXCOV_SLOTS * 4));
                int64_t* d_b = static_cast<int64_t*>(dalloc(nseg * 32));
                glck(c, gl_indexcov_counts_segs_device(c, sd.d_dep, static_cast<const int64_t*>(d_ss), static_cast<const int64_t*>(d_sl),
                                                       static_cast<const int64_t*>(d_lg), (int32_t)nseg, d_c, d_b), "gl_indexcov_counts_segs_device");
                if (G == 1) {
                    glck(c, gl_memcpy_d2h(c, counts_all.data(), d_c, (int64_t)nseg * GL_INDEXCOV_SLOTS * 4), "gl_memcpy_d2h");
                    glck(c, gl_memcpy_d2h(c, b4_all.data(), d_b, (int64_t)nseg * 32), "gl_memcpy_d2h");
                } else {                                                         // this GPU's columns of every reference's block
                    std::vector<int32_t> cc(nseg * GL_INDEXCOV_SLOTS);
                    std::vector<int64_t> bb(nseg * 4);
                    glck(c, gl_memcpy_d2h(c, cc.data(), d_c, (int64_t)nseg * GL_INDEXCOV_SLOTS * 4), "gl_memcpy_d2h");
                    glck(c, gl_memcpy_d2h(c, bb.data(), d_b, (int64_t)nseg * 32), "gl_memcpy_d2h");
                    for (size_t q = 0; q < nq; q++) {
                        memcpy(&counts_all[(q * S + a) * GL_INDEXCOV_SLOTS], &cc[q * Sg * GL_INDEXCOV_SLOTS], Sg * GL_INDEXCOV_SLOTS * 4);
                        memcpy(&b4_all[(q * S + a) * 4], &bb[q * Sg * 4], Sg * 32);
                    }
                }
                gl_dev_free(c, d_ss); gl_dev_free(c, d_sl); gl_dev_free(c, d_lg); gl_dev_free(c, d_c); gl_dev_free(c, d_b);
            }
        });
    }
    for (int g = 0; g < G; g++) if (shard[(size_t)g].d_dep) { gl_dev_free(ctxs[(size_t)g], shard[(size_t)g].d_dep); shard[(size_t)g].d_dep = nullptr; }
    for (int g = 1; g < G; g++) if (ctxs[(size_t)g]) gl_ctx_destroy(ctxs[(size_t)g]);
    gl_ctx* ctx = ctxs[0];                                                  // -n (I7) runs per reference on GPU 0
    ic_mark("tokens + counts (GPU, + D2H)");

    std::map<std::string, std::vector<double>> sexes;
    std::vector<int64_t> bins(S * 4, 0);
    std::vector<float> slopes(S, 0);
    int n_slopes = 0;
    for (size_t q = 0; q < kept.size(); q++) {
        const IcRef& ref = *kept[q];
        // per-sample slices of this reference (ref.ID() indexes the index's reference array, indexcov.go:656)
        std::vector<const float*> dptr(S, nullptr);
        std::vector<int64_t> doff(S, 0);
        std::vector<int32_t> lens(S, 0);
        size_t longest = 0;
        for (size_t k = 0; k < S; k++) {
            if (ref.id + 1 < (int)size_ptr[k].size() && med[k] != 0) {
                const int64_t a = size_ptr[k][ref.id], b = size_ptr[k][ref.id + 1];
                doff[k] = sample_ptr[k] + a;
                dptr[k] = dep.data() + doff[k];
                lens[k] = (int32_t)(b - a);
            }
            longest = std::max(longest, (size_t)lens[k]);
        }
        const bool is_sex = same_chrom(sexes_wanted, ref.name);
        std::vector<float> mat;                                             // S x longest, used when values are rewritten
        std::vector<int32_t> counts_ref;
        std::vector<int64_t> b4_ref;
        std::vector<uint8_t> tok_ref;
        std::vector<int64_t> seg_ptr(1, 0);
        const int32_t* counts = nullptr;
        const int64_t* b4 = nullptr;
        if (extranorm) {
            if (!is_sex && longest > 0) {
                mat.assign(S * longest, 0.f);
                for (size_t k = 0; k < S; k++) if (lens[k]) memcpy(&mat[k * longest], dptr[k], (size_t)lens[k] * 4);
                glck(ctx, gl_indexcov_xnorm(ctx, mat.data(), lens.data(), (int32_t)S, (int32_t)longest), "gl_indexcov_xnorm");
                for (size_t k = 0; k < S; k++) dptr[k] = &mat[k * longest];
            }
            // -n rewrites the values per reference: slot histograms, counters and tokens from this reference's values
            std::vector<float> flat;
            std::vector<int64_t> lg(S, (int64_t)longest);
            for (size_t k = 0; k < S; k++) { flat.insert(flat.end(), dptr[k], dptr[k] + lens[k]); seg_ptr.push_back((int64_t)flat.size()); }
            counts_ref.resize(S * GL_INDEXCOV_SLOTS); b4_ref.resize(S * 4);
            glck(ctx, gl_indexcov_counts_batch(ctx, flat.data(), seg_ptr.data(), lg.data(), (int32_t)S, counts_ref.data(), b4_ref.data()), "gl_indexcov_counts_batch");
            tok_ref.resize(flat.size() * 10 + 10);
            glck(ctx, gl_format_g3(ctx, flat.data(), (int64_t)flat.size(), tok_ref.data()), "gl_format_g3");
            counts = counts_ref.data(); b4 = b4_ref.data();
        } else {
            counts = &counts_all[q * S * GL_INDEXCOV_SLOTS];
            b4 = &b4_all[q * S * 4];
        }
        // bed.gz rows (indexcov.go:678-680,1038-1048): every "%.3g" comes from the GPU formatter as a 10-byte token; the rows
        // are put together by all host threads, 256 rows per task, and handed to the BGZF writer in order
        {
            const size_t kRows = 256, n_tasks = (longest + kRows - 1) / kRows;
            std::vector<std::string> parts(n_tasks);
            glhost::ThreadPool::global().run((int64_t)n_tasks, [&](int64_t t, int) {
                std::string& out = parts[(size_t)t];
                out.reserve(kRows * (ref.name.size() + 24 + S * 7));
                char num[48];
                const size_t i0 = (size_t)t * kRows, i1 = std::min(longest, i0 + kRows);
                for (size_t i = i0; i < i1; i++) {
                    out += ref.name;
                    out.append(num, (size_t)snprintf(num, sizeof num, "\t%zu\t%zu", i * 16384, (i + 1) * 16384));
                    for (size_t k = 0; k < S; k++) {
                        out.push_back('\t');
                        if (i >= (size_t)lens[k]) { out.push_back('0'); continue; }
                        const uint8_t* tk = extranorm ? &tok_ref[((size_t)seg_ptr[k] + i) * 10] : &tok_all[((size_t)doff[k] + i) * 10];
                        if (tk[9]) out.append(reinterpret_cast<const char*>(tk), tk[9]);
                        else out += go_float("%.3g", (double)dptr[k][i]);                           // magnitude outside the kernel's range
                    }
                    out.push_back('\n');
                }
            });
            for (const std::string& part : parts) bgz.write(part.data(), part.size());
        }
        if (is_sex) {
            if (longest > 0) { std::vector<double> cn(S); for (size_t k = 0; k < S; k++) cn[k] = get_cn(dptr[k], (size_t)lens[k]); sexes[ref.name] = cn; }
        } else {
            for (size_t k = 0; k < S * 4; k++) bins[k] += b4[k];
        }
        if (longest > 0) {                                                  // writeROCs (indexcov.go:1018-1036)
            std::string h = "#chrom\tcov"; for (auto& n : names) h += "\t" + n; h += "\n";
            fputs(h.c_str(), roc);
            std::vector<float> rocs(S * GL_INDEXCOV_SLOTS);
            for (size_t k = 0; k < S; k++) {
                int tot[GL_INDEXCOV_SLOTS];
                const int32_t* c = &counts[k * GL_INDEXCOV_SLOTS];
                tot[GL_INDEXCOV_SLOTS - 1] = c[GL_INDEXCOV_SLOTS - 1];
                for (int i = GL_INDEXCOV_SLOTS - 2; i >= 0; i--) tot[i] = tot[i + 1] + c[i];
                const float mx = (float)tot[0];
                for (int i = 0; i < GL_INDEXCOV_SLOTS; i++) rocs[k * GL_INDEXCOV_SLOTS + i] = (float)tot[i] / mx;
            }
            for (int i = 0; i < GL_INDEXCOV_SLOTS; i++) {
                fprintf(roc, "%s\t%.2f", ref.name.c_str(), (double)i / (70.0 * (2.0 / 3.0)));
                for (size_t k = 0; k < S; k++) fprintf(roc, "\t%s", go_float("%.2f", (double)rocs[k * GL_INDEXCOV_SLOTS + i]).c_str());
                fputc('\n', roc);
            }
            if ((includegl || ref.name.compare(0, 2, "GL") != 0) && longest > 2 && !is_sex && longest > 100) {
                const double sm = 2.0 / 3.0;                                // updateSlopes (indexcov.go:739-750)
                const int ilo = (int)(0.5 + (sm - 0.1) * 70), ihi = (int)(0.5 + (sm + 0.1) * 70);
                const float scalar = (float)ref.len / 1e6f;
                for (size_t k = 0; k < S; k++) slopes[k] += (rocs[k * 70 + ilo] - rocs[k * 70 + ihi]) * scalar;
                n_slopes++;
            }
        }
    }
    ic_mark("rows + roc per reference");
    bgz.close();
    fclose(roc);
    ic_mark("bgzf close");
    for (size_t k = 0; k < S; k++) slopes[k] = slopes[k] / (float)n_slopes;
    if (sexes.size() != sexes_wanted.size()) {                              // checkSexes (indexcov.go:760-770)
        std::string keys; for (auto& kv : sexes) keys += (keys.empty() ? "" : ",") + kv.first;
        if (sexes.empty() && sex != "X,Y") fatal(1, "(FATAL) indexcov: expected %zu sex chromosomes, found: %zu.", sexes_wanted.size(), sexes.size());
        fprintf(stderr, "(WARNING) indexcov: expected %zu sex chromosomes, found: %zu.\nyou can set the expected with --sex '%s'\n", sexes_wanted.size(), sexes.size(), keys.c_str());
    }
    // .ped (indexcov.go:815-894), without the PCA columns
    if (sexes.empty()) fprintf(stderr, "sex chromosomes not found.\n");
    FILE* ped = fopen((base + ".ped").c_str(), "w");
    if (!ped) fatal(1, "cannot create %s.ped", base.c_str());
    bool anygt = false;
    for (size_t k = 0; k < S; k++) if (mapped[k] > 0 || unmapped[k] > 0) anygt = true;
    fprintf(ped, "#family_id\tsample_id\tpaternal_id\tmaternal_id\tsex\tphenotype");
    for (auto& kv : sexes) fprintf(ped, "\tCN%s", kv.first.c_str());
    fprintf(ped, "\tbins.out\tbins.lo\tbins.hi\tbins.in\tslope\tp.out");
    if (anygt) fprintf(ped, "\tmapped\tunmapped");
    fputc('\n', ped);
    for (size_t k = 0; k < S; k++) {
        const int inferred = sexes.empty() ? -9 : (int)(0.5 + sexes.begin()->second[k]);
        fprintf(ped, "unknown\t%s\t-9\t-9\t%d\t-9", names[k].c_str(), inferred);
        for (auto& kv : sexes) fprintf(ped, "\t%s", go_float("%.2f", kv.second[k]).c_str());
        fprintf(ped, "\t%lld\t%lld\t%lld\t%lld\t%s\t%s", (long long)bins[k * 4 + 0], (long long)bins[k * 4 + 1], (long long)bins[k * 4 + 2],
                (long long)bins[k * 4 + 3], go_float("%.3f", (double)slopes[k]).c_str(),
                go_float("%.2f", (double)bins[k * 4 + 0] / (double)bins[k * 4 + 3]).c_str());
        if (anygt) fprintf(ped, "\t%llu\t%llu", (unsigned long long)mapped[k], (unsigned long long)unmapped[k]);
        fputc('\n', ped);
    }
    fclose(ped);
    // The reference's PCA (indexcov.go:772-810) and its HTML/PNG plots are out of scope; the log lines its
    // functional tests look for (functional-tests.sh:49,82) are kept, and index.html is a plain list of the outputs.
    if (S < 3) {
        fprintf(stderr, "got: %zu principal components\n", S);
        fprintf(stderr, "indexcov: %zu principal components, not plotting\n", S);
    }
    {
        const std::string index_path = dir + "/index.html";
        FILE* ih = fopen(index_path.c_str(), "w");
        if (ih) {
            const std::string stem = base.substr(base.find_last_of('/') + 1);
            fprintf(ih, "<html><head><title>indexcov</title></head><body><p>goleft_b200 indexcov: %zu samples; plots are not built by this engine.</p><ul>"
                        "<li><a href=\"%s.bed.gz\">%s.bed.gz</a></li><li><a href=\"%s.roc\">%s.roc</a></li><li><a href=\"%s.ped\">%s.ped</a></li></ul></body></html>\n",
                    S, stem.c_str(), stem.c_str(), stem.c_str(), stem.c_str(), stem.c_str(), stem.c_str());
            fclose(ih);
            fprintf(stderr, "indexcov finished: see %s for overview of output\n", index_path.c_str());   // indexcov.go:454
        }
    }
    gl_ctx_destroy(ctx);
    if (ic_timing) fprintf(stderr, "[indexcov timing] %-28s %.3f s\n", "total", now_s() - ic_t0);
    return 0;
}

// ------------------------------------------------------------------------------------------------ covstats
// V2 (covstats.go:57-89,175-218).  The reference sorts the sampled insert sizes / template lengths / read lengths and reads
// order statistics off the sorted slices.  Here the multiset is a histogram made on the GPU (gl_bincount_i32): the k-th
// smallest value is a search in its running count, madFilter's cut is a count, and meanStd's float64 accumulation visits
// the values in ascending order with their multiplicities — the same sequence of additions as over the sorted slice, so
// the results are bit-identical without a sort.
struct SortedCounts {
    int32_t lo = 0;
    std::vector<uint64_t> cum;                               // cum[i] = how many values are <= lo + i
    std::vector<int32_t> sorted;                             // fallback when the value range is too wide for a histogram
    size_t n = 0;
    void build(gl_ctx* ctx, const std::vector<int32_t>& v) {
        n = v.size();
        cum.clear(); sorted.clear();
        if (n == 0) return;
        int32_t mn = v[0], mx = v[0];
        for (int32_t x : v) { mn = std::min(mn, x); mx = std::max(mx, x); }
        const long long range = (long long)mx - mn + 1;
        if (range > (1LL << 22)) { sorted = v; std::sort(sorted.begin(), sorted.end()); return; }
        lo = mn;
        cum.assign((size_t)range, 0);
        glck(ctx, gl_bincount_i32(ctx, v.data(), (int64_t)n, mn, (int32_t)((long long)mx + 1 > INT32_MAX ? INT32_MAX : mx + 1), cum.data()), "gl_bincount_i32");
        if ((long long)mx + 1 > INT32_MAX) { sorted = v; std::sort(sorted.begin(), sorted.end()); cum.clear(); return; }
        for (size_t i = 1; i < cum.size(); i++) cum[i] += cum[i - 1];
    }
    int32_t kth(size_t k) const {                            // k-th smallest, 0-based
        if (!sorted.empty()) return sorted[k];
        return lo + (int32_t)(std::upper_bound(cum.begin(), cum.end(), (uint64_t)k) - cum.begin());
    }
    size_t count_le(long long v) const {
        if (!sorted.empty()) return (size_t)(std::upper_bound(sorted.begin(), sorted.end(), v, [](long long a, int32_t b) { return a < (long long)b; }) - sorted.begin());
        if (v < lo) return 0;
        const long long i = v - lo;
        return (size_t)(i >= (long long)cum.size() ? cum.back() : cum[(size_t)i]);
    }
    template <class F> void for_first(size_t m, F f) const { // the m smallest values, ascending, with multiplicity
        if (!sorted.empty()) { for (size_t i = 0; i < m; i++) f(sorted[i]); return; }
        size_t done = 0;
        for (size_t i = 0; i < cum.size() && done < m; i++) {
            const size_t c = std::min<size_t>((size_t)(cum[i] - (i ? cum[i - 1] : 0)), m - done);
            for (size_t k = 0; k < c; k++) f(lo + (int32_t)i);
            done += c;
        }
    }
};
static void mean_std(const SortedCounts& a, size_t n, double& mean, double& sd) {     // covstats.go:78-89 over the n smallest
    const double l = (double)n;
    mean = 0; sd = 0;
    a.for_first(n, [&](int32_t x) { mean += (double)x / l; });
    const double mu = mean;
    a.for_first(n, [&](int32_t x) { sd += pow((double)x - mu, 2) / l; });
    sd = sqrt(sd);
}
static size_t mad_filter(const SortedCounts& a, int nmads) {                        // covstats.go:57-76: how many values are kept
    const size_t n = a.n;
    const int32_t med = a.kth(n / 2);
    // upper_mads = arr[n/2+1:] - med is already sorted: its middle element is arr[n/2+1 + len/2] - med
    const size_t len = n - (n / 2 + 1);
    const long long umad = (long long)a.kth(n / 2 + 1 + len / 2) - med;
    const long long upper = (long long)med + (long long)nmads * umad;
    size_t i = a.count_le(upper);                                                   // first index with arr[i] > upper
    if (i == n) i = n - 1;                                                          // Go's range loop leaves i at the last index
    return i;
}

static int cmd_covstats(int argc, char** argv) {
    // the header is printed before the arguments are parsed (covstats.go:224-226)
    printf("coverage\tinsert_mean\tinsert_sd\tinsert_5th\tinsert_95th\ttemplate_mean\ttemplate_sd\tpct_unmapped\tpct_bad_reads\tpct_duplicate\tpct_proper_pair\tread_length\tbam\tsample\n");
    std::string nstr = "1000000", regions, fasta;
    ArgParser ap;
    ap.prog = "goleft covstats";
    ap.add("n", 'n', &nstr); ap.add("regions", 'r', &regions); ap.add("fasta", 'f', &fasta);
    ap.parse(argc, argv);
    if (ap.positional.empty()) ap.fail("bams is required");
    const int N = atoi(nstr.c_str());
    gl_ctx* ctx = nullptr;
    if (gl_ctx_create(0, &ctx) != GL_OK) fatal(1, "goleft covstats: %s", gl_last_error(nullptr));
    for (const std::string& bam : ap.positional) {
        glhts::BamHeader hdr;
        glhts::CovstatsSample cs;
        std::string e = glhts::bam_covstats_sample(bam, N, 100000, hdr, cs);
        if (!e.empty()) fatal(1, "%s", e.c_str());
        std::string names;
        for (auto& s : hdr.sample_names()) names += (names.empty() ? "" : ",") + s;
        if (names.empty()) names = "<no-read-groups>";
        glhts::BaiIndex bai;
        bool have_idx = false;
        if (ends_with(bam, ".bam")) {
            std::string p = bam + ".bai";
            if (!file_exists(p)) p = bam.substr(0, bam.size() - 4) + ".bai";
            std::string be = glhts::bai_read(p, bai);
            if (!be.empty()) fatal(1, "%s", be.c_str());
            have_idx = true;
        }
        double pBad = 0, pDup = 0, pProper = 0, pUnmapped = 0, rlMean = 0, insMean = 0, insSd = 0, tMean = 0, tSd = 0;
        int pct5 = 0, pct95 = 0, maxRead = 0;
        SortedCounts sc;
        if (!cs.read_len.empty()) {                                                 // covstats.go:177-186
            const double den = (double)(cs.k + cs.n_unmapped);
            pBad = cs.n_bad / den; pDup = cs.n_dup / den; pProper = cs.n_proper / den; pUnmapped = cs.n_unmapped / den;
            sc.build(ctx, cs.read_len);
            double sd;
            mean_std(sc, sc.n, rlMean, sd);
            maxRead = sc.kth(sc.n - 1);
        }
        if (!cs.insert.empty()) {                                                   // covstats.go:188-199
            sc.build(ctx, cs.insert);
            const double l = (double)(cs.insert.size() - 1);
            pct5 = sc.kth((size_t)(0.05 * l + 0.5));
            pct95 = sc.kth((size_t)(0.95 * l + 0.5));
            if (sc.n < 3) fatal(2, "panic: runtime error: index out of range (covstats madFilter needs at least 3 insert sizes)");
            const size_t ki = mad_filter(sc, 10);
            mean_std(sc, ki, insMean, insSd);
            sc.build(ctx, cs.tmpl);
            const size_t kt = mad_filter(sc, 10);
            mean_std(sc, kt, tMean, tSd);
            // (Stats.H, the template-length histogram of covstats.go:202-217, is consumed by smoove through the Go API and is
            //  not part of the TSV row this command prints; it is not computed here)
        }
        long long genome = 0;
        unsigned long long mapped = 0;
        std::string not_found;
        for (size_t r = 0; r < hdr.refs.size(); r++) {                              // covstats.go:256-267
            genome += hdr.refs[r].length;
            if (have_idx) {
                if (r >= bai.has_stats.size() || !bai.has_stats[r]) {
                    if (hdr.refs[r].name.find("random") == std::string::npos && hdr.refs[r].length > 10000) not_found += (not_found.empty() ? "" : ",") + hdr.refs[r].name;
                    continue;
                }
                mapped += bai.mapped[r];
            }
        }
        if (!not_found.empty()) fprintf(stderr, "chromosomes: %s not found in %s\n", not_found.c_str(), bam.c_str());
        if (!regions.empty()) {                                                     // covstats.go:36-55
            genome = 0;
            for (const std::string& ln : read_lines(regions, true)) {
                size_t t1 = ln.find('\t');
                if (t1 == std::string::npos) continue;
                size_t t2 = ln.find('\t', t1 + 1);
                genome += atoll(ln.c_str() + t2 + 1) - atoll(ln.c_str() + t1 + 1);
            }
        }
        const double coverage = (1 - pBad) * (double)mapped * rlMean / (double)genome;     // covstats.go:277
        printf("%.2f\t%.2f\t%.2f\t%d\t%d\t%.2f\t%.2f\t%.2f\t%.1f\t%.1f\t%.1f\t%d\t%s\t%s\n", coverage, insMean, insSd, pct5, pct95, tMean, tSd,
               100 * pUnmapped, 100 * pBad, 100 * pDup, 100 * pProper, maxRead, bam.c_str(), names.c_str());
    }
    gl_ctx_destroy(ctx);
    return 0;
}

// ------------------------------------------------------------------------------------------------ depthwed
static int cmd_depthwed(int argc, char** argv) {
    std::string size;
    ArgParser ap;
    ap.prog = "goleft depthwed";
    ap.add("size", 's', &size, true);
    ap.parse(argc, argv);
    if (ap.positional.empty()) ap.fail("beds is required");
    const long long sz = atoll(size.c_str());
    const size_t S = ap.positional.size();
    std::vector<std::string> chroms;
    std::map<std::string, int> chrom_id;
    std::vector<int32_t> starts, ends, cid;
    std::vector<double> means;                       // kept only for the rare case a depth does not fit int32
    std::vector<int32_t> depth32;                    // int(0.5 + mean), rounded where the reference rounds: at parse time (depthwed.go:103)
    bool fits32 = true;
    size_t R = 0;
    std::string header = "#chrom\tstart\tend";
    for (size_t f = 0; f < S; f++) {
        std::string nm = ap.positional[f];                                          // getNameFromFile (depthwed.go:37-46)
        nm = nm.substr(nm.find_last_of('/') == std::string::npos ? 0 : nm.find_last_of('/') + 1);
        for (const char* suf : {".gz", ".bed", ".depth"}) if (ends_with(nm, suf)) nm = nm.substr(0, nm.size() - strlen(suf));
        header += "\t" + nm;
        std::vector<std::string> lines = read_lines(ap.positional[f], true);
        if (f == 0) { R = lines.size(); means.resize(S * R); depth32.resize(S * R); }
        else if (lines.size() != R) fatal(2, "panic: not all files have same number of records");
        for (size_t r = 0; r < R; r++) {
            const std::string& ln = lines[r];
            size_t t1 = ln.find('\t'), t2 = ln.find('\t', t1 + 1), t3 = ln.find('\t', t2 + 1);
            if (t1 == std::string::npos || t2 == std::string::npos || t3 == std::string::npos) fatal(1, "bad line in %s: %s", ap.positional[f].c_str(), ln.c_str());
            const double dep = strtod(ln.c_str() + t3 + 1, nullptr);
            means[f * R + r] = dep;
            const double rd = 0.5 + dep;                                            // Go: int(0.5 + dep) truncates toward zero
            if (!(rd > -2147483648.0 && rd < 2147483648.0)) fits32 = false;
            else depth32[f * R + r] = (int32_t)rd;
            if (f == 0) {
                std::string c = ln.substr(0, t1);
                auto it = chrom_id.find(c);
                if (it == chrom_id.end()) { it = chrom_id.emplace(c, (int)chroms.size()).first; chroms.push_back(c); }
                cid.push_back(it->second);
                starts.push_back(atoi(ln.c_str() + t1 + 1));
                ends.push_back(atoi(ln.c_str() + t2 + 1));
            }
        }
    }
    printf("%s\n", header.c_str());
    if (R == 0) return 0;
    gl_ctx* ctx = nullptr;
    if (gl_ctx_create(0, &ctx) != GL_OK) fatal(1, "goleft depthwed: %s", gl_last_error(nullptr));
    std::vector<int32_t> os(R), oe(R), oc(R);
    std::vector<int64_t> out;
    std::vector<int32_t> out32;
    int64_t n_out = 0;
    bool have32 = false;
    if (fits32) {                                    // 4 B in + 4 B out per cell; a group sum beyond int32 falls through to the 64-bit path
        out32.resize(R * S);
        const int rc = gl_depthwed_aggregate_i32(ctx, depth32.data(), (int32_t)S, (int64_t)R, starts.data(), ends.data(), cid.data(), sz, os.data(),
                                                 oe.data(), oc.data(), out32.data(), (int64_t)R, &n_out);
        if (rc == GL_OK) have32 = true;
        else if (!(rc == GL_ERANGE && n_out == -1)) glck(ctx, rc, "gl_depthwed_aggregate_i32");
    }
    if (!have32) {
        out.resize(R * S);
        glck(ctx, gl_depthwed_aggregate(ctx, means.data(), (int32_t)S, (int64_t)R, starts.data(), ends.data(), cid.data(), sz, os.data(), oe.data(),
                                        oc.data(), out.data(), (int64_t)R, &n_out), "gl_depthwed_aggregate");
    }
    std::string row;
    char num[32];
    for (int64_t g = 0; g < n_out; g++) {
        row = chroms[oc[g]];
        snprintf(num, sizeof num, "\t%d\t%d", os[g], oe[g]); row += num;
        for (size_t s = 0; s < S; s++) { snprintf(num, sizeof num, "\t%lld", have32 ? (long long)out32[(size_t)g * S + s] : (long long)out[(size_t)g * S + s]); row += num; }
        row += "\n";
        fwrite(row.data(), 1, row.size(), stdout);
    }
    gl_ctx_destroy(ctx);
    return 0;
}

// ------------------------------------------------------------------------------------------------ dispatcher
// ------------------------------------------------------------------------------------------------ indexsplit
// indexsplit/indexsplit.go:24-30,197-216: N regions of about equal amounts of data across the cohort, from the indexes alone
static int cmd_indexsplit(int argc, char** argv) {
    std::string n, fai, problematic;
    ArgParser ap;
    ap.prog = "goleft indexsplit";
    ap.add("n", 'n', &n, true); ap.add("fai", 0, &fai); ap.add("problematic", 'p', &problematic);
    ap.parse(argc, argv);
    if (ap.positional.empty()) ap.fail("indexes is required");
    const std::vector<std::string>& paths = ap.positional;
    const int N = atoi(n.c_str());

    // references: from the first BAM's header, else from the .fai (the reference's ReadFai references carry ID -1 and
    // Split indexes a slice with it; here a .fai reference's id is its position in the file)
    struct Ref { std::string name; int64_t len; int id; };
    std::vector<Ref> refs;
    if (ends_with(paths[0], ".bam")) {
        glhts::BamHeader h;
        std::string e = glhts::bam_read_header(paths[0], h);
        if (!e.empty()) fatal(1, "%s", e.c_str());
        for (size_t i = 0; i < h.refs.size(); i++) refs.push_back({h.refs[i].name, h.refs[i].length, (int)i});
    } else {
        std::vector<glhts::RefInfo> r;
        std::string e = glhts::fai_read(fai, r);
        if (!e.empty()) fatal(1, "error opening fai: %s. Is is present?", fai.c_str());
        std::stable_sort(r.begin(), r.end(), [](const glhts::RefInfo& a, const glhts::RefInfo& b) { return a.offset < b.offset; });   // indexcov.go:293
        for (auto& x : r) refs.push_back({x.name, x.length, (int)refs.size()});
    }
    const int32_t R = (int32_t)refs.size();

    gl_ctx* ctx = nullptr;
    if (gl_ctx_create(0, &ctx) != GL_OK) fatal(1, "goleft indexsplit: %s", gl_last_error(nullptr));
    const size_t S = paths.size();
    std::vector<int64_t> all_sizes, ptr(S * (size_t)(R + 1), 0), max_len((size_t)R, 0);
    for (size_t i = 0; i < S; i++) {                                        // indexcov.ReadIndex(path).Sizes(), indexsplit.go:92-93
        const std::string& b = paths[i];
        std::vector<int64_t> sz, sp;
        if (ends_with(b, ".crai")) {
            std::vector<glhts::CraiSlices> cr;
            std::string e = glhts::crai_read(b, cr);
            if (!e.empty()) fatal(1, "%s", e.c_str());
            sp.assign(cr.size() + 1, 0);
            for (size_t r = 0; r < cr.size(); r++) {
                std::vector<int64_t> one;
                if (!glhts::crai_make_sizes(cr[r].start.data(), cr[r].span.data(), cr[r].bytes.data(), (int64_t)cr[r].start.size(), one)) fatal(2, "panic: tilewidth logic error");
                sz.insert(sz.end(), one.begin(), one.end());
                sp[r + 1] = (int64_t)sz.size();
            }
        } else {
            std::string p = ends_with(b, ".bai") ? b : b + ".bai";
            if (!file_exists(p)) p = b.substr(0, b.size() - 4) + (ends_with(b, ".bai") ? "" : ".bai");
            glhts::BaiIndex bai;
            std::string e = glhts::bai_read(p, bai);
            if (!e.empty()) fatal(1, "%s", e.c_str());
            const size_t nr = bai.ioffsets.size();
            std::vector<uint64_t> voff;
            std::vector<int64_t> rp(nr + 1, 0);
            for (size_t r = 0; r < nr; r++) { voff.insert(voff.end(), bai.ioffsets[r].begin(), bai.ioffsets[r].end()); rp[r + 1] = (int64_t)voff.size(); }
            sz.resize(voff.size() + 1); sp.assign(nr + 1, 0);
            glck(ctx, gl_indexcov_sizes(ctx, voff.data(), rp.data(), (int32_t)nr, sz.data(), sp.data()), "gl_indexcov_sizes");
            sz.resize((size_t)sp[nr]);
        }
        if (sp.back() < 1) fatal(1, "indexcov: no usable chromsomes in bam: %s", b.c_str());     // Index.init, indexcov.go:100-102
        const int64_t base = (int64_t)all_sizes.size();
        const int32_t nr = (int32_t)sp.size() - 1;
        for (int32_t r = 0; r <= R; r++) ptr[i * (size_t)(R + 1) + r] = base + sp[(size_t)std::min(r, nr)];
        for (int32_t r = 0; r < R && r < nr; r++) max_len[(size_t)r] = std::max(max_len[(size_t)r], sp[(size_t)r + 1] - sp[(size_t)r]);
        all_sizes.insert(all_sizes.end(), sz.begin(), sz.begin() + sp[(size_t)std::min(R, nr)]);
    }
    std::vector<int64_t> out_ptr((size_t)R + 1, 0);
    for (int32_t r = 0; r < R; r++) out_ptr[(size_t)r + 1] = out_ptr[(size_t)r] + max_len[(size_t)r];
    std::vector<double> tile_sum((size_t)out_ptr[(size_t)R] + 1, 0.0);
    glck(ctx, gl_indexsplit_accumulate(ctx, all_sizes.data(), ptr.data(), (int32_t)S, R, out_ptr.data(), tile_sum.data()), "gl_indexsplit_accumulate");

    // -p: BED of problematic regions (depth.ReadTree, depth/intervals.go:42-82)
    std::vector<int32_t> pr; std::vector<int64_t> ps, pe;
    if (!problematic.empty()) {
        std::map<std::string, int> num;
        for (int32_t q = 0; q < R; q++) num.emplace(refs[(size_t)q].name, q);
        for (const std::string& ln : read_lines(problematic, true)) {
            std::string c; long long s0, e0;
            if (!chrom_start_end(ln, c, s0, e0)) fatal(2, "panic: couldn't get region from line %s", ln.c_str());
            if (s0 >= e0) continue;
            auto it = num.find(c);
            if (it != num.end()) { pr.push_back(it->second); ps.push_back(s0); pe.push_back(e0); }
        }
    }
    std::vector<const char*> nm; std::vector<int64_t> ln_; std::vector<int32_t> ids;
    for (const Ref& r : refs) { nm.push_back(r.name.c_str()); ln_.push_back(r.len); ids.push_back(r.id); }
    char* text = nullptr; int64_t tl = 0;
    if (gl_indexsplit_chunks(tile_sum.data(), out_ptr.data(), R, nm.data(), ln_.data(), ids.data(), R, N, pr.data(), ps.data(), pe.data(),
                             (int64_t)pr.size(), &text, &tl) != GL_OK) fatal(1, "gl_indexsplit_chunks failed");
    fwrite(text, 1, (size_t)tl, stdout);
    gl_free_text(text);
    gl_ctx_destroy(ctx);
    return 0;
}

static void print_progs() {                                                          // cmd/goleft/goleft.go:33-53
    fprintf(stderr, "goleft Version: %s (B200 engine: %s)\n\n", kVersion, gl_version());
    fprintf(stderr, "covstats    : coverage and insert-size statistics on bams by sampling\n");
    fprintf(stderr, "depth       : parallelize calls to samtools in user-defined windows\n");
    fprintf(stderr, "depthwed    : matricize output from depth to n-sites * n-samples\n");
    fprintf(stderr, "indexcov    : quick coverage estimate using only the bam index\n");
    fprintf(stderr, "indexsplit  : create regions of even coverage across bams/crams\n");
}

int main(int argc, char** argv) {
    if (argc < 2) { print_progs(); return 0; }
    const std::string prog = argv[1];
    if (prog == "depth") return cmd_depth(argc - 1, argv + 1);
    if (prog == "indexcov") return cmd_indexcov(argc - 1, argv + 1);
    if (prog == "covstats") return cmd_covstats(argc - 1, argv + 1);
    if (prog == "depthwed") return cmd_depthwed(argc - 1, argv + 1);
    if (prog == "indexsplit") return cmd_indexsplit(argc - 1, argv + 1);
    print_progs();
    return 1;
}
