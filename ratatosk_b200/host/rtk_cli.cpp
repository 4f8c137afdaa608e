// rtk_cli.cpp — host driver: FASTQ -> librtk_b200.so -> FASTQ.  The ticket loop of the reference's search()
// (src/Ratatosk.cpp:727-906), its ordered block writer (:869-886, :919-999) and writeCorrectedOutput with -t trimming
// (:510-616) around the C ABI of include/rtk.h.  Same command line as `Ratatosk correct -1 | -2` from an index
// (src/Ratatosk.cpp:151-296): the files it writes are byte-identical to the reference's.
//
//   rtk_correct correct -1 -g G.k31.fasta.gz -d G.k31.rtsk -l reads.fastq[.gz] -o out            -> out.2.fastq
//   rtk_correct correct -2 -g G.k63.fasta.gz -d G.k63.rtsk -l out.2.fastq -L reads.fastq -o out  -> out.fastq[.gz with -G]
//
// What differs from the reference, by design:
//   * a ticket is --ticket-bases (default 32 Mi) of reads, not buffer_sz = 1 MiB: one ticket = one rtk_correct_batch call, and
//     the GPU wants thousands of regions per call.  Ticket size is not observable in the output (blocks are written in ticket
//     order, which the reference restores with its re-ordering pass for pass 1 and for pass 2 under -O).
//   * tickets are dealt to --gpus devices (one calling thread + one context per device, graph uploaded once per device);
//     blocks are formatted (and gzip-compressed, one member per block, when -G) by the device threads in parallel and written by
//     one ordered writer: no temporary file, no second pass over the output.
//   * pass 2 always runs phasing() (the reference's multi-thread branch, `-c` >= 2) and always writes in input order (`-O`).
//   * there is no CPU path: without a CUDA device rtk_ctx_create fails and the driver exits with its message.
#include <getopt.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rtk.h"

namespace {

// ---------------------------------------------------------------------------------------------- FASTA / FASTQ reader
// kseq semantics (Bifrost/src/File_Parser.hpp -> kseq.h): name = header up to the first white space, multi-line records,
// '>' records have no quality string; plain or gzip input.
class SeqReader {
  public:
    explicit SeqReader(const std::vector<std::string>& files) : files_(files) {}
    ~SeqReader() { if (gz_) gzclose(gz_); }
    bool next(std::string& name, std::string& seq, std::string& qual, bool& has_qual) {
        for (;;) {
            if (!gz_ && !open_next()) return false;
            if (read_record(name, seq, qual, has_qual)) return true;
            gzclose(gz_); gz_ = nullptr;
        }
    }
    const std::string& error() const { return err_; }

  private:
    bool open_next() {
        if (fi_ >= files_.size()) return false;
        gz_ = gzopen(files_[fi_].c_str(), "rb");
        if (!gz_) { err_ = "cannot open " + files_[fi_]; fi_ = files_.size(); return false; }
        gzbuffer(gz_, 1u << 20);
        ++fi_; pos_ = len_ = 0; eof_ = false; pending_ = 0;
        return true;
    }
    int getc_() {
        if (pos_ == len_) {
            if (eof_) return -1;
            const int n = gzread(gz_, buf_, sizeof(buf_));
            if (n <= 0) { eof_ = true; return -1; }
            len_ = (size_t)n; pos_ = 0;
        }
        return (unsigned char)buf_[pos_++];
    }
    // appends the rest of the current line to s (without the line end); false at end of file with nothing read
    bool getline_(std::string& s) {
        bool any = false;
        for (;;) {
            if (pos_ == len_) { if (getc_() < 0) return any; --pos_; }
            const char* b = buf_ + pos_;
            const char* e = (const char*)memchr(b, '\n', len_ - pos_);
            any = true;
            if (e) { s.append(b, e); pos_ += (size_t)(e - b) + 1; break; }
            s.append(b, (size_t)(len_ - pos_)); pos_ = len_;
        }
        if (!s.empty() && s.back() == '\r') s.pop_back();
        return true;
    }
    bool read_record(std::string& name, std::string& seq, std::string& qual, bool& has_qual) {
        int c = pending_;
        pending_ = 0;
        while (c != '>' && c != '@') { c = getc_(); if (c < 0) return false; }   // skip to the next header
        line_.clear();
        if (!getline_(line_)) return false;
        size_t e = 0;
        while (e < line_.size() && !isspace((unsigned char)line_[e])) ++e;
        name.assign(line_, 0, e);
        seq.clear(); qual.clear(); has_qual = false;
        for (;;) {   // sequence lines up to '+', the next header or the end of the file
            c = getc_();
            if (c < 0) return true;
            if (c == '+' || c == '>' || c == '@') break;
            if (c == '\n' || c == '\r') continue;
            seq.push_back((char)c);
            getline_(seq);
        }
        if (c != '+') { pending_ = c; return true; }
        line_.clear(); getline_(line_);   // rest of the '+' line
        has_qual = true;
        while (qual.size() < seq.size()) { const size_t before = qual.size(); if (!getline_(qual) && qual.size() == before) break; }
        return true;
    }
    std::vector<std::string> files_;
    size_t fi_ = 0;
    gzFile gz_ = nullptr;
    char buf_[1 << 16];
    size_t pos_ = 0, len_ = 0;
    bool eof_ = false;
    int pending_ = 0;
    std::string line_, err_;
};

// ---------------------------------------------------------------------------------------------- options
struct Options {   // the `correct` fields of Correct_Opt the path reads (src/Common.hpp:16-158), reference defaults
    std::vector<std::string> long_in, long_raw;
    std::string out, graph, data, graph2, data2;   // graph2 / data2: the k2 index of the two-pass mode
    bool pass1 = false, pass2 = false, gzip_out = false, verbose = false, force_snp = false, force_order = false;
    int threads = 1, trim_qual = 0, max_qual = 40, insert_sz = 500, k1 = 31, k2 = 63, rounds = 1, w1 = 1000, w2 = 5000;
    double min_conf_snp = 0.9;
    int gpus = 1, first_gpu = 0;
    uint64_t ticket_bases = 32ull << 20;
    bool has_phase_files = false;
    bool no_cache = false;       // parse the index files even when a flat cache exists, and write none
    std::string cache;           // flat cache file (default: <rtsk>.k<k>.rtkflat next to the index)
};

void usage() {
    fprintf(stderr,
            "rtk_correct correct (-1 | -2) -g <graph.fasta[.gz]> -d <graph.rtsk> -l <long reads> [-L <raw long reads>] -o <prefix>\n"
            "  same options as `Ratatosk correct` from an index: -c -t -m -i -k -K -w -W -r -Q -O -G -v\n"
            "  -c 1 (default) = the reference's single-thread branch: the 2nd pass runs without phasing; -c N > 1 = with phasing\n"
            "  --gpus N          devices to deal tickets to (default 1)\n"
            "  --first-gpu D     first CUDA device ordinal (default 0)\n"
            "  --ticket-bases B  read bases per library call (default 33554432)\n"
            "  --in-graph2 G2 --in-unitig-data2 D2   (without -1 / -2) both passes as one pipeline: -g -d = k1 index, G2 D2 = k2 index;\n"
            "                    writes <prefix>.fastq[.gz] like `correct -1` followed by `correct -2 -O`\n"
            "  --cache FILE      flat graph cache (default <rtsk>.k<k>.rtkflat; written on first use, mapped in place afterwards)\n"
            "  --no-cache        always parse the index files, write no cache\n"
            "rtk_correct annotate -g <graph.fasta[.gz]> -d <graph.rtsk> -o <out.rtsk> [-k K] [--min-cov N] [--no-snp] [--first-gpu D] [-v]\n"
            "  detectSNPs + detectShortCycles of the reference's index build, recomputed on the GPU from the colours of the index\n"
            "rtk_correct index2 -g <graph.k2.fasta[.gz]> -l <pass-1 corrected long reads> -o <prefix> [-K k2] [-M f] [-C n] [-Q n] [--min-cov N] [--no-snp] [--keep-all-reads] [-v]\n"
            "  `Ratatosk index -2`: colours the k2 graph with the long reads, adds SNP and short-cycle annotations, writes <prefix>.index.k<k2>.rtsk\n");
}

int parse(int argc, char** argv, Options& o) {
    if (argc < 2 || strcmp(argv[1], "correct") != 0) { usage(); return 1; }
    static struct option lo[] = {{"in-long", required_argument, 0, 'l'}, {"out-long", required_argument, 0, 'o'}, {"cores", required_argument, 0, 'c'},
                                 {"trim-split", required_argument, 0, 't'}, {"in-graph", required_argument, 0, 'g'}, {"in-unitig-data", required_argument, 0, 'd'},
                                 {"min-conf-snp-corr", required_argument, 0, 'm'}, {"insert-sz", required_argument, 0, 'i'}, {"k1", required_argument, 0, 'k'},
                                 {"k2", required_argument, 0, 'K'}, {"max-len-weak1", required_argument, 0, 'w'}, {"max-len-weak2", required_argument, 0, 'W'},
                                 {"correction-rounds", required_argument, 0, 'r'}, {"in-long-raw", required_argument, 0, 'L'}, {"in-long-phase", required_argument, 0, 'P'},
                                 {"in-short-phase", required_argument, 0, 'p'}, {"max-base-qual", required_argument, 0, 'Q'}, {"1st-pass-only", no_argument, 0, '1'},
                                 {"2nd-pass-only", no_argument, 0, '2'}, {"force-correct-snp", no_argument, 0, 'f'}, {"force-io-order", no_argument, 0, 'O'},
                                 {"gzip-out", no_argument, 0, 'G'}, {"verbose", no_argument, 0, 'v'}, {"gpus", required_argument, 0, 1000},
                                 {"first-gpu", required_argument, 0, 1001}, {"ticket-bases", required_argument, 0, 1002}, {"no-cache", no_argument, 0, 1003}, {"in-graph2", required_argument, 0, 1005}, {"in-unitig-data2", required_argument, 0, 1006},
                                 {"cache", required_argument, 0, 1004}, {0, 0, 0, 0}};
    int c;
    while ((c = getopt_long(argc - 1, argv + 1, "l:o:c:t:g:d:m:i:k:K:w:W:r:L:P:p:Q:12fOGv", lo, nullptr)) != -1) {
        switch (c) {
            case 'l': o.long_in.push_back(optarg); break;
            case 'L': o.long_raw.push_back(optarg); break;
            case 'o': o.out = optarg; break;
            case 'g': o.graph = optarg; break;
            case 'd': o.data = optarg; break;
            case 'c': o.threads = atoi(optarg); break;
            case 't': o.trim_qual = atoi(optarg); break;
            case 'm': o.min_conf_snp = atof(optarg); break;
            case 'i': o.insert_sz = atoi(optarg); break;
            case 'k': o.k1 = atoi(optarg); break;
            case 'K': o.k2 = atoi(optarg); break;
            case 'w': o.w1 = atoi(optarg); break;
            case 'W': o.w2 = atoi(optarg); break;
            case 'r': o.rounds = atoi(optarg); break;
            case 'Q': o.max_qual = atoi(optarg); break;
            case 'P': case 'p': o.has_phase_files = true; break;   // accepted and unused, like the reference when it corrects from an index
            case '1': o.pass1 = true; break;
            case '2': o.pass2 = true; break;
            case 'f': o.force_snp = true; break;
            case 'O': o.force_order = true; break;
            case 'G': o.gzip_out = true; break;
            case 'v': o.verbose = true; break;
            case 1000: o.gpus = atoi(optarg); break;
            case 1001: o.first_gpu = atoi(optarg); break;
            case 1002: o.ticket_bases = strtoull(optarg, nullptr, 10); break;
            case 1003: o.no_cache = true; break;
            case 1004: o.cache = optarg; break;
            case 1005: o.graph2 = optarg; break;
            case 1006: o.data2 = optarg; break;
            default: usage(); return 1;
        }
    }
    // check_ProgramOptions (src/Ratatosk.cpp:303-420), the checks that concern this path, same wording
    bool ok = true;
    auto bad = [&](const std::string& m) { fprintf(stderr, "Ratatosk::Ratatosk(): %s\n", m.c_str()); ok = false; };
    if (o.trim_qual < 0 || o.trim_qual > o.max_qual) bad("Quality score trimming threshold cannot be less than 0 or more than " + std::to_string(o.max_qual) + " (" + std::to_string(o.trim_qual) + " given).");
    if (o.k2 <= o.k1) bad("Length of long k-mers for 2nd correction pass cannot be less than or equal to length of short k-mers for 1st correction pass (" + std::to_string(o.k2) + " and " + std::to_string(o.k1) + " given).");
    if (o.insert_sz <= 0) bad("Insert size of short reads cannot be less than or equal to 0");
    if (o.max_qual < 0) bad("Maximum base quality cannot be less than 0");
    if (o.rounds < 1) bad("At least one short read correction round is required.");
    if (o.min_conf_snp < 0.0) bad("Minimum confidence threshold to correct a SNP must be greater or equal to 0.0.");
    if (o.min_conf_snp > 1.0) bad("Minimum confidence threshold to correct a SNP must be lower or equal to 1.0.");
    if (o.w1 <= 0 || o.w2 <= 0) bad("Maximum length of a weak region to correct cannot be less than or equal to 0");
    if (o.pass1 && o.pass2) bad("-1 and -2 are mutually exclusive (perform *only* one of the two correction passes). To perform both, remove -1 and -2 from your command line.");
    const bool two_pass = !o.pass1 && !o.pass2;
    if (two_pass && (o.graph2.empty() || o.data2.empty()))
        bad("rtk_correct corrects from indexes: give -1 or -2 with the index of that pass (-g -d), or both indexes (-g -d for k1, --in-graph2 --in-unitig-data2 for k2) to run the two passes as one pipeline (building / colouring the graphs is the reference's `index` step).");
    if (!two_pass && (!o.graph2.empty() || !o.data2.empty())) bad("--in-graph2 / --in-unitig-data2 belong to the two-pass mode (no -1 / -2).");
    if (two_pass && !o.cache.empty()) bad("--cache names one file: not usable with two indexes.");
    if (o.graph.empty() != o.data.empty()) bad("One of the input index files is missing (either the graph or the data).");
    if (o.graph.empty()) bad("rtk_correct needs the index of the pass (-g and -d).");
    if (o.long_in.empty()) bad("Missing input long reads (-l).");
    if (o.out.empty()) bad("Missing output prefix (-o).");
    if (o.pass2 && o.long_raw.empty()) bad("Missing input raw long reads (-L) for the 2nd correction pass.");
    // -p / -P: with an index (-g -d) the reference never fills its HapReads (hapPass1 / hapPass2 stay empty on the hasIndex
    // branches, src/Ratatosk.cpp:1062-1090, :1213-1217), so hap_id is ~0 for every read: the files are ignored here as well
    if (o.has_phase_files && o.verbose) fprintf(stderr, "rtk_correct: phasing files are not used when correcting from an index (as in the reference)\n");
    if (o.gpus < 1) bad("--gpus must be at least 1.");
    return ok ? 0 : 1;
}

// ---------------------------------------------------------------------------------------------- tickets
struct Ticket {
    uint64_t id = 0;
    std::vector<std::string> names;
    std::string seq, qual, raw;
    std::vector<uint64_t> off{0}, raw_off{0};
    bool has_qual = true;
};
struct Block { std::string bytes; uint64_t reads = 0, bases = 0; };

template <class T>
class Channel {   // bounded queue between the reader and the device threads
  public:
    explicit Channel(size_t cap) : cap_(cap) {}
    void push(T&& v) {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&] { return q_.size() < cap_ || closed_; });
        q_.push_back(std::move(v));
        cv_.notify_all();
    }
    bool pop(T& v) {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&] { return !q_.empty() || closed_; });
        if (q_.empty()) return false;
        v = std::move(q_.front());
        q_.pop_front();
        cv_.notify_all();
        return true;
    }
    void close() { std::lock_guard<std::mutex> l(m_); closed_ = true; cv_.notify_all(); }

  private:
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<T> q_;
    size_t cap_;
    bool closed_ = false;
};

// writeCorrectedOutput (src/Ratatosk.cpp:510-558): a record, or with trim > 0 its stretches of >= k bases whose quality
// is at least trim, named name/1, name/2, ...
void format_record(std::string& out, const std::string& name, const char* seq, const char* qual, const uint64_t len, const int k, const int trim) {
    if (trim == 0) {
        out.push_back('@'); out += name; out.push_back('\n'); out.append(seq, len); out += "\n+\n"; out.append(qual, len); out.push_back('\n');
        return;
    }
    const char c_min = (char)(trim + 33);
    int64_t start = -1, run = -1, id = 1;
    auto flush = [&] {
        if (run >= k) {
            out.push_back('@'); out += name; out.push_back('/'); out += std::to_string(id++); out.push_back('\n');
            out.append(seq + start, (size_t)run); out += "\n+\n"; out.append(qual + start, (size_t)run); out.push_back('\n');
        }
    };
    for (int64_t p = 0; p < (int64_t)len; ++p) {
        if (qual[p] >= c_min) { if (start == -1) { start = p; run = 0; } ++run; }
        else { flush(); start = -1; run = -1; }
    }
    flush();
}

bool gzip_member(const std::string& in, std::string& out) {   // one gzip member per block: members concatenate into a valid .gz
    z_stream z; memset(&z, 0, sizeof(z));
    if (deflateInit2(&z, Z_DEFAULT_COMPRESSION, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    out.resize(deflateBound(&z, in.size()) + 64);
    z.next_in = (Bytef*)in.data(); z.avail_in = (uInt)in.size();
    z.next_out = (Bytef*)&out[0]; z.avail_out = (uInt)out.size();
    const int rc = deflate(&z, Z_FINISH);
    out.resize(z.total_out);
    deflateEnd(&z);
    return rc == Z_STREAM_END;
}

struct Shared {
    std::mutex m;
    std::condition_variable cv;
    std::map<uint64_t, Block> done;   // finished blocks waiting for their turn
    std::string error;
    std::atomic<bool> failed{false};
    void fail(const std::string& e) { std::lock_guard<std::mutex> l(m); if (error.empty()) error = e; failed = true; cv.notify_all(); }
};



// `rtk_correct annotate`: the last two steps of the reference's index build (src/Ratatosk.cpp:1124-1134) on an existing index -
// detectSNPs + detectShortCycles recomputed on the GPU from the colours and edge flags the index holds, written as a new .rtsk
int run_annotate(int argc, char** argv) {
    std::string graph, data, out;
    int k = 31, device = 0, min_cov = 2;
    bool verbose = false, no_snp = false;
    static struct option lo[] = {{"in-graph", required_argument, 0, 'g'}, {"in-unitig-data", required_argument, 0, 'd'}, {"out-unitig-data", required_argument, 0, 'o'},
                                 {"k", required_argument, 0, 'k'}, {"min-cov", required_argument, 0, 1000}, {"no-snp", no_argument, 0, 1001},
                                 {"first-gpu", required_argument, 0, 1002}, {"verbose", no_argument, 0, 'v'}, {0, 0, 0, 0}};
    int c;
    while ((c = getopt_long(argc - 1, argv + 1, "g:d:o:k:v", lo, nullptr)) != -1) {
        switch (c) {
            case 'g': graph = optarg; break;
            case 'd': data = optarg; break;
            case 'o': out = optarg; break;
            case 'k': k = atoi(optarg); break;
            case 'v': verbose = true; break;
            case 1000: min_cov = atoi(optarg); break;
            case 1001: no_snp = true; break;
            case 1002: device = atoi(optarg); break;
            default: usage(); return 1;
        }
    }
    if (graph.empty() || data.empty() || out.empty() || min_cov < 1) { usage(); return 1; }
    const auto t0 = std::chrono::steady_clock::now();
    auto secs = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    auto die = [&]() { fprintf(stderr, "Ratatosk::Ratatosk(): %s\n", rtk_last_error()); return 1; };
    rtk_host_graph* hg = nullptr;
    rtk_ctx* ctx = nullptr;
    if (rtk_graph_load(graph.c_str(), data.c_str(), k, &hg) != RTK_OK) return die();
    if (rtk_ctx_create(device, &ctx) != RTK_OK || rtk_graph_upload(ctx, hg) != RTK_OK) return die();
    rtk_graph_info gi;
    rtk_graph_get_info(hg, &gi);
    rtk_opt ropt;
    rtk_opt_default(&ropt, 1);
    ropt.k = (uint32_t)k; ropt.min_cov_vertices = (uint32_t)min_cov;
    uint64_t *amb_off = nullptr, *cyc_off = nullptr, st[10] = {0};
    uint32_t* amb_ids = nullptr;
    uint8_t* is_cycle = nullptr;
    char* cyc_pool = nullptr;
    if (no_snp) {   // force_no_snp_corr: no candidates are added (src/Ratatosk.cpp:1124)
        amb_off = (uint64_t*)calloc(gi.n_unitigs + 1, 8);
        amb_ids = (uint32_t*)calloc(1, 4);
        if (verbose) printf("Ratatosk::Ratatosk(): SNPs candidate detection is disabled.\n");
    } else {
        if (verbose) printf("Ratatosk::Ratatosk(): Adding SNPs candidates to graph.\n");
        if (rtk_detect_snps(ctx, &ropt, &amb_off, &amb_ids, st) != RTK_OK) return die();
    }
    if (verbose) printf("Ratatosk::Ratatosk(): Adding micro/mini-satellites motif candidates to graph.\n");
    if (rtk_detect_short_cycles(ctx, &ropt, &is_cycle, &cyc_off, &cyc_pool, st) != RTK_OK) return die();
    if (verbose) printf("Ratatosk::Ratatosk(): Writing index to disk.\n");
    if (rtk_rtsk_write_annotations(hg, data.c_str(), out.c_str(), amb_off, amb_ids, is_cycle, cyc_off, cyc_pool) != RTK_OK) return die();
    if (verbose) {
        uint64_t n_cyc = 0;
        for (uint64_t u = 0; u < gi.n_unitigs; ++u) n_cyc += is_cycle[u];
        printf("rtk_correct: %llu unitigs, %llu SNP marks (%llu candidates, %llu traversals), %llu unitigs in short cycles, %.2f s\n", (unsigned long long)gi.n_unitigs,
               (unsigned long long)amb_off[gi.n_unitigs], (unsigned long long)st[4], (unsigned long long)st[6], (unsigned long long)n_cyc, secs());
    }
    rtk_free(amb_off); rtk_free(amb_ids); rtk_free(is_cycle); rtk_free(cyc_off); rtk_free(cyc_pool);
    rtk_ctx_destroy(ctx);
    rtk_graph_free(hg);
    return 0;
}

// `rtk_correct index2`: `Ratatosk index -2` (src/Ratatosk.cpp:1205-1260) - the k2 graph coloured with the pass-1 corrected long reads
// (addCoverage, long-read branch), then detectSNPs and detectShortCycles, then <prefix>.index.k<K>.rtsk
int run_index2(int argc, char** argv) {
    std::string graph, out;
    std::vector<std::string> reads;
    int k = 63, device = 0, min_cov = 2, min_len = 3000, max_qual = 40;
    double min_conf = 0.0;
    bool verbose = false, no_snp = false;
    static struct option lo[] = {{"in-graph", required_argument, 0, 'g'}, {"in-long", required_argument, 0, 'l'}, {"out-long", required_argument, 0, 'o'},
                                 {"k2", required_argument, 0, 'K'}, {"min-conf-color2", required_argument, 0, 'M'}, {"min-len-color2", required_argument, 0, 'C'},
                                 {"max-base-qual", required_argument, 0, 'Q'}, {"min-cov", required_argument, 0, 1000}, {"no-snp", no_argument, 0, 1001},
                                 {"first-gpu", required_argument, 0, 1002}, {"keep-all-reads", no_argument, 0, 1003}, {"verbose", no_argument, 0, 'v'}, {0, 0, 0, 0}};
    bool keep_all = false;
    int c;
    while ((c = getopt_long(argc - 1, argv + 1, "g:l:o:K:M:C:Q:v", lo, nullptr)) != -1) {
        switch (c) {
            case 'g': graph = optarg; break;
            case 'l': reads.push_back(optarg); break;
            case 'o': out = optarg; break;
            case 'K': k = atoi(optarg); break;
            case 'M': min_conf = atof(optarg); break;
            case 'C': min_len = atoi(optarg); break;
            case 'Q': max_qual = atoi(optarg); break;
            case 'v': verbose = true; break;
            case 1000: min_cov = atoi(optarg); break;
            case 1001: no_snp = true; break;
            case 1002: device = atoi(optarg); break;
            case 1003: keep_all = true; break;
            default: usage(); return 1;
        }
    }
    if (graph.empty() || reads.empty() || out.empty() || min_cov < 1 || min_len < 0 || min_conf < 0.0 || min_conf > 1.0) { usage(); return 1; }
    const auto t0 = std::chrono::steady_clock::now();
    auto secs = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    auto die = [&]() { fprintf(stderr, "Ratatosk::Ratatosk(): %s\n", rtk_last_error()); return 1; };
    if (verbose) printf("Ratatosk::Ratatosk(): Loading graph (2/2).\n");
    rtk_host_graph *hg = nullptr, *hg2 = nullptr;
    rtk_ctx *ctx = nullptr, *ctx2 = nullptr;
    if (rtk_graph_load(graph.c_str(), nullptr, k, &hg) != RTK_OK) return die();
    if (rtk_ctx_create(device, &ctx) != RTK_OK || rtk_graph_upload(ctx, hg) != RTK_OK) return die();
    rtk_graph_info gi;
    rtk_graph_get_info(hg, &gi);
    // the reads: sequences, qualities and names as three pools
    std::string seq_pool, qual_pool, name_pool, name, seq, qual;
    std::vector<uint64_t> seq_off(1, 0), name_off(1, 0);
    {
        SeqReader rd(reads);
        bool has_qual = false, all_qual = true;
        while (rd.next(name, seq, qual, has_qual)) {
            if (!has_qual || qual.size() != seq.size()) { all_qual = false; qual.assign(seq.size(), 'I'); }
            seq_pool += seq; qual_pool += qual; name_pool += name;
            seq_off.push_back(seq_pool.size()); name_off.push_back(name_pool.size());
        }
        if (!rd.error().empty()) { fprintf(stderr, "Ratatosk::Ratatosk(): %s\n", rd.error().c_str()); return 1; }
        if (!all_qual && verbose) printf("rtk_correct: records without quality string are not masked\n");
    }
    const uint32_t n_reads = (uint32_t)(seq_off.size() - 1);
    rtk_opt ropt;
    rtk_opt_default(&ropt, 2);
    ropt.k = (uint32_t)k; ropt.min_cov_vertices = (uint32_t)min_cov; ropt.max_qual = max_qual; ropt.reserved = keep_all ? 1u : 0u;
    if (verbose) printf("Ratatosk::Ratatosk(): Adding colors and coverage to graph (2/2).\n");
    uint64_t *kmcov = nullptr, *shared = nullptr, *col_off = nullptr, st[10] = {0};
    uint32_t* col_ids = nullptr;
    if (rtk_color_long_reads(ctx, &ropt, n_reads, seq_pool.data(), seq_off.data(), qual_pool.data(), seq_off.data(), name_pool.data(), name_off.data(),
                             (uint32_t)min_len, min_conf, &kmcov, &shared, &col_off, &col_ids, nullptr, st) != RTK_OK) return die();
    if (rtk_graph_recolor(hg, kmcov, shared, col_off, col_ids, &hg2) != RTK_OK) return die();
    rtk_ctx_destroy(ctx);
    rtk_graph_free(hg);
    if (rtk_ctx_create(device, &ctx2) != RTK_OK || rtk_graph_upload(ctx2, hg2) != RTK_OK) return die();
    uint64_t *amb_off = nullptr, *cyc_off = nullptr, st2[10] = {0};
    uint32_t* amb_ids = nullptr;
    uint8_t* is_cycle = nullptr;
    char* cyc_pool = nullptr;
    if (no_snp) {
        amb_off = (uint64_t*)calloc(gi.n_unitigs + 1, 8);
        amb_ids = (uint32_t*)calloc(1, 4);
        if (verbose) printf("Ratatosk::Ratatosk(): SNPs candidate detection is disabled (2/2).\n");
    } else {
        if (verbose) printf("Ratatosk::Ratatosk(): Adding SNPs candidates to graph (2/2).\n");
        if (rtk_detect_snps(ctx2, &ropt, &amb_off, &amb_ids, st2) != RTK_OK) return die();
    }
    if (verbose) printf("Ratatosk::Ratatosk(): Adding micro/mini-satellites motif candidates to graph (2/2).\n");
    if (rtk_detect_short_cycles(ctx2, &ropt, &is_cycle, &cyc_off, &cyc_pool, st2) != RTK_OK) return die();
    const std::string path = out + ".index.k" + std::to_string(k) + ".rtsk";
    if (verbose) printf("Ratatosk::Ratatosk(): Writing index to disk (2/2).\n");
    if (rtk_rtsk_write(hg2, path.c_str(), amb_off, amb_ids, is_cycle, cyc_off, cyc_pool) != RTK_OK) return die();
    if (verbose) {
        uint64_t n_cyc = 0;
        for (uint64_t u = 0; u < gi.n_unitigs; ++u) n_cyc += is_cycle[u];
        printf("rtk_correct: %llu unitigs coloured by %llu of %u reads (%llu unitig-read pairs), %llu SNP marks, %llu unitigs in short cycles, %.2f s\n",
               (unsigned long long)gi.n_unitigs, (unsigned long long)st[5], n_reads, (unsigned long long)st[4], (unsigned long long)amb_off[gi.n_unitigs],
               (unsigned long long)n_cyc, secs());
    }
    rtk_free(kmcov); rtk_free(shared); rtk_free(col_off); rtk_free(col_ids);
    rtk_free(amb_off); rtk_free(amb_ids); rtk_free(is_cycle); rtk_free(cyc_off); rtk_free(cyc_pool);
    rtk_ctx_destroy(ctx2);
    rtk_graph_free(hg2);
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc >= 2 && strcmp(argv[1], "annotate") == 0) return run_annotate(argc, argv);
    if (argc >= 2 && strcmp(argv[1], "index2") == 0) return run_index2(argc, argv);
    Options o;
    if (parse(argc, argv, o)) return 1;
    const int pass = o.pass2 ? 2 : o.pass1 ? 1 : 0;   // 0: both passes as one pipeline (rtk_correct_two_pass_batch)
    const int k = pass == 1 ? o.k1 : (pass == 2 ? o.k2 : o.k1);   // k of the graph given with -g / -d
    const int k_out = pass == 1 ? o.k1 : o.k2;   // k of the pass that writes the file (minimum length of a trimmed sub-read)
    const auto t_start = std::chrono::steady_clock::now();
    auto secs = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count(); };

    // ---- graph: loaded + flattened once, uploaded to every device (src/Ratatosk.cpp:1087-1089)
    if (o.verbose) printf("Ratatosk::Ratatosk(): Loading graph (%d/2).\n", pass ? pass : 1);
    rtk_host_graph* hg = nullptr;
    int from_cache = 0;
    const int grc = o.no_cache ? rtk_graph_load(o.graph.c_str(), o.data.c_str(), k, &hg)
                               : rtk_graph_load_cached(o.graph.c_str(), o.data.c_str(), k, o.cache.empty() ? nullptr : o.cache.c_str(), &hg, &from_cache);
    if (grc != RTK_OK) { fprintf(stderr, "Ratatosk::Ratatosk(): %s\n", rtk_last_error()); return 1; }
    std::vector<rtk_ctx*> ctx((size_t)o.gpus, nullptr);
    for (int d = 0; d < o.gpus; ++d) {
        if (rtk_ctx_create(o.first_gpu + d, &ctx[d]) != RTK_OK || rtk_graph_upload(ctx[d], hg) != RTK_OK) {
            fprintf(stderr, "Ratatosk::Ratatosk(): %s\n", rtk_last_error());
            return 1;
        }
    }
    if (o.verbose) { rtk_graph_info gi; rtk_graph_get_info(hg, &gi); printf("rtk_correct: graph%s resident on %d GPU(s) after %.1f s (%llu unitigs, %llu k-mers, %.1f MB slab)\n", from_cache ? " (flat cache, mapped in place)" : "", o.gpus, secs(), (unsigned long long)gi.n_unitigs, (unsigned long long)gi.n_kmers, gi.slab_bytes / 1e6); }

    // two-pass mode: the k2 graph next to the k1 graph on every device
    rtk_host_graph* hg2 = nullptr;
    std::vector<rtk_ctx*> ctx2((size_t)o.gpus, nullptr);
    if (pass == 0) {
        if (o.verbose) printf("Ratatosk::Ratatosk(): Loading graph (2/2).\n");
        const int rc2 = o.no_cache ? rtk_graph_load(o.graph2.c_str(), o.data2.c_str(), o.k2, &hg2)
                                   : rtk_graph_load_cached(o.graph2.c_str(), o.data2.c_str(), o.k2, nullptr, &hg2, nullptr);
        if (rc2 != RTK_OK) { fprintf(stderr, "Ratatosk::Ratatosk(): %s\n", rtk_last_error()); return 1; }
        for (int d = 0; d < o.gpus; ++d)
            if (rtk_ctx_create(o.first_gpu + d, &ctx2[d]) != RTK_OK || rtk_graph_upload(ctx2[d], hg2) != RTK_OK) { fprintf(stderr, "Ratatosk::Ratatosk(): %s\n", rtk_last_error()); return 1; }
    }
    rtk_opt ropt2;   // second pass of the two-pass mode
    rtk_opt_default(&ropt2, 2);
    ropt2.k = (uint32_t)o.k2; ropt2.insert_sz = (uint32_t)o.insert_sz; ropt2.max_len_weak_region1 = (uint32_t)o.w1; ropt2.max_len_weak_region2 = (uint32_t)o.w2;
    ropt2.nb_correction_rounds = (uint32_t)o.rounds; ropt2.max_qual = o.max_qual; ropt2.trim_qual = o.trim_qual;
    ropt2.min_confidence_snp_corr = o.min_conf_snp; ropt2.force_unres_snp_corr = o.force_snp ? 1u : 0u;

    rtk_opt ropt;
    rtk_opt_default(&ropt, pass == 2 ? 2 : 1);
    ropt.k = (uint32_t)k; ropt.insert_sz = (uint32_t)o.insert_sz; ropt.max_len_weak_region1 = (uint32_t)o.w1; ropt.max_len_weak_region2 = (uint32_t)o.w2;
    ropt.nb_correction_rounds = (uint32_t)o.rounds; ropt.max_qual = o.max_qual; ropt.trim_qual = o.trim_qual;
    ropt.min_confidence_snp_corr = o.min_conf_snp; ropt.force_unres_snp_corr = o.force_snp ? 1u : 0u;
    if (pass != 1 && o.verbose && o.force_snp) fprintf(stderr, "Ratatosk::search(): Force unresolved SNP correction is activated.\n");

    // ---- output file: <out>.2.fastq after pass 1 (src/Ratatosk.cpp:1079), <out>.fastq[.gz] after pass 2 (:621, :909-911)
    const bool gz_out = o.gzip_out && pass != 1;
    const std::string fn_out = o.out + (pass == 1 ? ".2" : "") + ".fastq" + (gz_out ? ".gz" : "");
    for (const auto* v : {&o.long_in, &o.long_raw})
        for (const auto& f : *v)
            if (f == fn_out) { fprintf(stderr, "Ratatosk::search(): output file %s is also an input file\n", fn_out.c_str()); return 1; }
    FILE* fout = fopen(fn_out.c_str(), "wb");
    if (!fout) { fprintf(stderr, "Ratatosk::search(): cannot open %s for writing\n", fn_out.c_str()); return 1; }
    setvbuf(fout, nullptr, _IOFBF, 8 << 20);
    if (o.verbose) printf(pass ? "Ratatosk::Ratatosk(): Correcting long reads (%d/2).\n" : "Ratatosk::Ratatosk(): Correcting long reads (both passes).\n", pass);

    Shared sh;
    Channel<Ticket> tickets((size_t)o.gpus * 2);
    std::atomic<uint64_t> n_tickets{0};
    std::atomic<bool> reader_done{false};

    // ---- reader: the ticket dispenser (src/Ratatosk.cpp:746-800)
    std::thread reader([&] {
        SeqReader in(o.long_in), raw(o.long_raw);
        Ticket t;
        std::string name, seq, qual, rname, rseq, rqual;
        bool hq = false, rhq = false;
        uint64_t id = 0, n_reads = 0;
        auto flush = [&] {
            if (t.names.empty()) return;
            t.id = id++;
            n_tickets = id;
            tickets.push(std::move(t));
            t = Ticket();
        };
        while (!sh.failed && in.next(name, seq, qual, hq)) {
            if (pass == 2) {   // the raw read of the same rank must carry the same name (:788-799, first character skipped)
                if (!raw.next(rname, rseq, rqual, rhq) || name.compare(1, std::string::npos, rname, 1, std::string::npos) != 0) {
                    sh.fail("Ratatosk::correct(): Corrected read file is not in the same order as input long read file. Abort.");
                    break;
                }
                t.raw += rseq;
                t.raw_off.push_back(t.raw.size());
            }
            if (hq && qual.size() != seq.size()) qual.resize(seq.size(), '!');
            if (!hq) t.has_qual = false;
            t.names.push_back(name);
            t.seq += seq;
            if (hq) t.qual += qual; else t.qual.append(seq.size(), '!');
            t.off.push_back(t.seq.size());
            if (o.verbose && (++n_reads % 1000 == 0)) printf("Ratatosk::correct(): Processed %llu reads \n", (unsigned long long)n_reads);
            if (t.seq.size() >= o.ticket_bases) flush();
        }
        if (!in.error().empty()) sh.fail(in.error());
        flush();
        reader_done = true;
        tickets.close();
        std::lock_guard<std::mutex> l(sh.m);
        sh.cv.notify_all();
    });

    // ---- device threads: one per GPU, each owning a context; a ticket = one library call per stage
    std::vector<std::thread> workers;
    for (int d = 0; d < o.gpus; ++d) {
        workers.emplace_back([&, d] {
            Ticket t;
            while (!sh.failed && tickets.pop(t)) {
                const uint32_t n = (uint32_t)t.names.size();
                char *cs = nullptr, *cq = nullptr;
                uint64_t* co = nullptr;
                const char* qual_in = t.has_qual ? t.qual.data() : nullptr;
                int rc;
                if (pass == 0) {
                    if (!t.has_qual) { sh.fail("Ratatosk::search(): the two-pass mode needs FASTQ input (quality strings)"); break; }
                    rc = rtk_correct_two_pass_batch(ctx[d], ctx2[d], &ropt, &ropt2, n, t.seq.data(), t.off.data(), t.qual.data(), t.off.data(), &cs, &cq, &co,
                                                    nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
                } else if (pass == 1) {
                    rc = rtk_correct_batch(ctx[d], &ropt, 1, n, t.seq.data(), t.off.data(), qual_in, t.off.data(), &cs, &cq, &co, nullptr);
                } else {
                    // upper-case the pass-1 reads (:814); phasing (:832) then getSeeds + correctSequence (:838/:840)
                    for (char& ch : t.seq) ch = (char)toupper((unsigned char)ch);
                    char *ps = nullptr, *pq = nullptr;
                    uint64_t* po = nullptr;
                    if (o.threads == 1) {
                        // -c 1 (the reference's default): the single-thread branch of search() has no phasing step - fixSNPs when
                        // forced (:672), then getSeeds + correctSequence (:674/:676)
                        rc = RTK_OK;
                        if (o.force_snp) rc = rtk_fix_snps_batch(ctx[d], &ropt, n, t.seq.data(), t.off.data(), &ps, nullptr);
                        if (rc == RTK_OK) rc = rtk_correct_batch(ctx[d], &ropt, 2, n, ps ? ps : t.seq.data(), t.off.data(), t.qual.data(), t.off.data(), &cs, &cq, &co, nullptr);
                        rtk_free(ps);
                    } else {
                    rc = rtk_phasing_batch(ctx[d], &ropt, n, t.raw.data(), t.raw_off.data(), t.seq.data(), t.off.data(), t.qual.data(), t.off.data(), &ps, &pq, &po);
                    if (rc == RTK_OK) rc = rtk_correct_batch(ctx[d], &ropt, 2, n, ps, po, pq, po, &cs, &cq, &co, nullptr);
                    rtk_free(ps); rtk_free(pq); rtk_free(po);
                    }
                }
                if (rc != RTK_OK) { sh.fail(std::string("Ratatosk::search(): ") + rtk_last_error()); break; }
                Block b;
                b.reads = n; b.bases = t.seq.size();
                std::string text;
                text.reserve((size_t)(co[n] * 2 + (uint64_t)n * 64));
                for (uint32_t i = 0; i < n; ++i) format_record(text, t.names[i], cs + co[i], cq + co[i], co[i + 1] - co[i], k_out, pass != 1 ? o.trim_qual : 0);
                rtk_free(cs); rtk_free(cq); rtk_free(co);
                if (gz_out) { if (!gzip_member(text, b.bytes)) { sh.fail("Ratatosk::search(): gzip compression failed"); break; } }
                else b.bytes.swap(text);
                std::lock_guard<std::mutex> l(sh.m);
                sh.done.emplace(t.id, std::move(b));
                sh.cv.notify_all();
            }
        });
    }

    // ---- ordered writer (replaces the (ticket_id, size, offset) list + re-ordering pass, :869-886, :919-999)
    uint64_t next = 0, reads = 0, bases = 0;
    {
        std::unique_lock<std::mutex> l(sh.m);
        for (;;) {
            sh.cv.wait(l, [&] { return sh.failed || sh.done.count(next) || (reader_done && next >= n_tickets); });
            if (sh.failed) break;
            auto it = sh.done.find(next);
            if (it == sh.done.end()) break;   // every ticket written
            Block b = std::move(it->second);
            sh.done.erase(it);
            l.unlock();
            if (fwrite(b.bytes.data(), 1, b.bytes.size(), fout) != b.bytes.size()) sh.fail("Ratatosk::search(): write error on " + fn_out);
            reads += b.reads; bases += b.bases;
            ++next;
            l.lock();
        }
    }
    tickets.close();
    reader.join();
    for (auto& w : workers) w.join();
    fclose(fout);
    for (auto c : ctx) rtk_ctx_destroy(c);
    for (auto c : ctx2) rtk_ctx_destroy(c);
    rtk_graph_free(hg);
    rtk_graph_free(hg2);
    if (sh.failed) { fprintf(stderr, "%s\n", sh.error.c_str()); remove(fn_out.c_str()); return 1; }
    if (o.verbose) printf("rtk_correct: %llu reads, %llu bases, %llu tickets in %.1f s -> %s\n", (unsigned long long)reads, (unsigned long long)bases, (unsigned long long)next, secs(), fn_out.c_str());
    return 0;
}
