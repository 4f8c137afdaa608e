// oracle/ref_seams.cpp — TEST INFRASTRUCTURE ONLY.
//
// extern "C" probes into the UNMODIFIED reference objects (compiled from
// /root/reference by oracle/Makefile into oracle/_ref/libref_seams.so).  They expose
// the inner seams of the correction hot path so that tests/ can compare the CUDA
// path with the reference itself, and so that tests/golden/make_golden.py can
// freeze reference outputs into committed fixtures:
//
//   ref_graph_load / ref_graph_dump   CompactedDBG<UnitigData>::read (Bifrost IO.tcc:124)
//                                     + readGraphData (src/Graph.cpp:722)
//   ref_search_sequence               CompactedDBG::searchSequence (Bifrost/src/Search.tcc:526)
//   ref_get_seeds                     getSeeds (src/Graph.cpp:3)
//   ref_correct_read                  the per-read body of search() (src/Ratatosk.cpp:808-867)
//   ref_phasing                       phasing (src/Graph.cpp:869), second pass, multi-thread branch
//   ref_fix_snps                      fixSNPs (src/Alignment.cpp:846), second pass with -f
//   ref_edlib                         edlibAlign (src/edlib.cpp:141)
//
// Nothing in the product library links, includes or dlopens this file.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>
#include <algorithm>
#include <iostream>
#include <sstream>
#include <map>
#include <set>
#include <queue>
#include <stack>
#include <random>
#include <thread>
#include <atomic>
#include <mutex>
#include <unordered_map>
#include <unordered_set>
#include <functional>
#include <memory>
#include <bitset>
#include <iterator>
#include <utility>
#include <tuple>
#include <array>
#include <list>
#include <deque>
#include <numeric>
#include <limits>
#include <type_traits>
#include <condition_variable>
#include <future>
#include <chrono>
#include <iomanip>
#include <cmath>
#include <cassert>
#include <climits>
#include <zlib.h>

// UnitigMap::{pos_unitig,isShort,isAbundant} identify the reference unitig but are
// private (Bifrost/src/UnitigMap.hpp:281-284).  Access control does not change object
// layout, so this translation unit (only) reads them by lifting the keyword.  All
// standard headers are included above so that only reference headers see it.
#define private public
#define protected public
#include "CompactedDBG.hpp"
#include "Common.hpp"
#include "Correction.hpp"
#include "Graph.hpp"
#include "UnitigData.hpp"
#include "edlib.h"
#undef private
#undef protected

using namespace std;

struct RefGraph {
    std::unordered_map<uint64_t, const_UnitigMap<UnitigData>> by_key;  // filled lazily (ref_explore_subgraph)
    CompactedDBG<UnitigData>* dbg;
    Correct_Opt opt;
    size_t max_km_cov;
    pair<HapReads, HapReads> hap;
};

struct ref_hit {
    uint64_t pos;      // position in the query
    uint64_t unitig;   // pos_unitig | isShort<<62 | isAbundant<<63
    uint32_t dist;
    uint32_t len;
    uint32_t size;
    uint32_t strand;
};

static inline uint64_t um_key(const const_UnitigMap<UnitigData>& um) {
    return (uint64_t)um.pos_unitig | ((uint64_t)um.isShort << 62) | ((uint64_t)um.isAbundant << 63);
}

static ref_hit* to_hits(const vector<pair<size_t, const_UnitigMap<UnitigData>>>& v) {
    ref_hit* h = (ref_hit*)malloc(sizeof(ref_hit) * (v.size() + 1));
    for (size_t i = 0; i < v.size(); ++i) {
        h[i].pos = v[i].first;
        h[i].unitig = um_key(v[i].second);
        h[i].dist = (uint32_t)v[i].second.dist;
        h[i].len = (uint32_t)v[i].second.len;
        h[i].size = (uint32_t)v[i].second.size;
        h[i].strand = v[i].second.strand ? 1u : 0u;
    }
    return h;
}

extern "C" {

void* ref_graph_load(const char* fasta, const char* rtsk, int k, int pass2, int threads) {
    RefGraph* g = new RefGraph();
    g->opt.k = k;
    g->opt.nb_threads = threads;
    g->dbg = new CompactedDBG<UnitigData>(k);
    if (!g->dbg->read(string(fasta), (size_t)threads, false)) { delete g->dbg; delete g; return nullptr; }
    if (rtsk != nullptr && rtsk[0] != 0) {
        if (!readGraphData(string(rtsk), *g->dbg, false, false)) { delete g->dbg; delete g; return nullptr; }
    }
    // src/Ratatosk.cpp:625
    g->max_km_cov = max(getMaxKmerCoverage(*g->dbg, g->opt.top_km_cov_ratio), g->opt.max_km_cov);
    (void)pass2;
    return g;
}

void ref_graph_free(void* h) {
    RefGraph* g = (RefGraph*)h;
    if (!g) return;
    delete g->dbg;
    delete g;
}

uint64_t ref_graph_num_unitigs(void* h) { return ((RefGraph*)h)->dbg->size(); }
uint64_t ref_graph_max_km_cov(void* h) { return ((RefGraph*)h)->max_km_cov; }

// Option setters (Correct_Opt fields on the hot path, src/Common.hpp:16-158)
void ref_set_opt(void* h, const char* name, double v) {
    Correct_Opt& o = ((RefGraph*)h)->opt;
    const string n(name);
    if (n == "insert_sz") o.insert_sz = (size_t)v;
    else if (n == "min_cov_vertices") o.min_cov_vertices = (size_t)v;
    else if (n == "max_qual") o.max_qual = (int)v;
    else if (n == "out_qual") o.out_qual = (int)v;
    else if (n == "max_len_weak_region1") o.max_len_weak_region1 = (size_t)v;
    else if (n == "max_len_weak_region2") o.max_len_weak_region2 = (size_t)v;
    else if (n == "weak_region_len_factor") o.weak_region_len_factor = v;
    else if (n == "min_score") o.min_score = v;
    else if (n == "nb_correction_rounds") o.nb_correction_rounds = (size_t)v;
    else if (n == "force_unres_snp_corr") o.force_unres_snp_corr = (v != 0);
}

// One text line per unitig:
//   key \t seq \t kmCov_cardBranches \t shared_pids \t global ids \t local ids \t ambiguity ids \t hap ids \t cycles(hex)
int ref_graph_dump(void* h, const char* path) {
    RefGraph* g = (RefGraph*)h;
    FILE* f = fopen(path, "w");
    if (!f) return -1;
    for (const auto& um : *g->dbg) {
        const UnitigData* ud = um.getData();
        fprintf(f, "%llu\t%s\t", (unsigned long long)um_key(um), um.referenceUnitigToString().c_str());
        // kmCov_cardBranches / shared_pids are private: recover them through write()
        std::stringstream ss;
        ud->write(ss);
        const string blob = ss.str();
        uint64_t w0, w1;
        memcpy(&w0, blob.data(), 8);
        memcpy(&w1, blob.data() + 8, 8);
        fprintf(f, "%llu\t%llu\t", (unsigned long long)w0, (unsigned long long)w1);
        const pair<const PairID*, const PairID*> pp = ud->getPairID().getPairIDs();
        if (pp.first) for (const uint32_t id : *pp.first) fprintf(f, "%u,", id);
        fputc('\t', f);
        if (pp.second) for (const uint32_t id : *pp.second) fprintf(f, "%u,", id);
        fputc('\t', f);
        {
            std::stringstream s2;  // ambiguity ids are private too: re-parse via write()
        }
        // ambiguity chars through the public accessor on the full fw mapping
        {
            const_UnitigMap<UnitigData> full(um);
            full.dist = 0; full.len = full.size - g->dbg->getK() + 1; full.strand = true;
            const vector<pair<size_t, char>> v = ud->get_ambiguity_char(full);
            for (const auto& p : v) fprintf(f, "%llu:%c,", (unsigned long long)p.first, p.second);
        }
        fputc('\t', f);
        for (const uint32_t id : ud->get_hapID()) fprintf(f, "%u,", id);
        fputc('\t', f);
        {
            const vector<const char*> cyc = ud->getCompactCycles();
            for (const char* c : cyc) fprintf(f, "%s;", c);
        }
        fprintf(f, "\t%d\t%d\n", (int)ud->isBranching(), (int)ud->isShortCycle());
    }
    fclose(f);
    return 0;
}

int64_t ref_search_sequence(void* h, const char* s, int exact, int ins, int del, int subst, int or_excl, ref_hit** out) {
    RefGraph* g = (RefGraph*)h;
    const vector<pair<size_t, const_UnitigMap<UnitigData>>> v =
        static_cast<const CompactedDBG<UnitigData>*>(g->dbg)->searchSequence(string(s), exact != 0, ins != 0, del != 0, subst != 0, or_excl != 0);
    *out = to_hits(v);
    return (int64_t)v.size();
}

int ref_get_seeds(void* h, const char* s, const char* q, int pass2, ref_hit** solid, int64_t* n_solid, ref_hit** weak, int64_t* n_weak) {
    RefGraph* g = (RefGraph*)h;
    unordered_map<Kmer, vector<const_UnitigMap<UnitigData>>, KmerHash> m_km_um;
    const auto p = getSeeds(g->opt, *g->dbg, string(s), string(q ? q : ""), pass2 != 0, 0xffffffffffffffffULL, m_km_um, false);
    *solid = to_hits(p.first);  *n_solid = (int64_t)p.first.size();
    *weak = to_hits(p.second);  *n_weak = (int64_t)p.second.size();
    return 0;
}

// Per-read body of search() for pass 1 (src/Ratatosk.cpp:808-867, multi-thread branch) and,
// for pass 2 without phasing, the single-thread branch (:670-683).
int ref_correct_read(void* h, const char* s_in, const char* q_in, int pass2, char** s_out, char** q_out) {
    RefGraph* g = (RefGraph*)h;
    const Correct_Opt& opt = g->opt;
    string in_read(s_in), in_qual(q_in ? q_in : "");
    std::transform(in_read.begin(), in_read.end(), in_read.begin(), ::toupper);
    const uint64_t hap_id = 0xffffffffffffffffULL;
    unordered_map<Kmer, vector<const_UnitigMap<UnitigData>>, KmerHash> m_km_um;
    if (!pass2) {
        if (!in_qual.empty()) getStdQual(in_qual, opt.max_qual);
        const double step_min_score = 1.00 / static_cast<double>(opt.nb_correction_rounds);
        const double step_wrlf = (opt.nb_correction_rounds == 1) ? 0.0 : ((opt.weak_region_len_factor - 0.10) / static_cast<double>(opt.nb_correction_rounds - 1));
        const size_t step_mlwr1 = opt.max_len_weak_region1 / opt.nb_correction_rounds;
        for (size_t j = 0; j < opt.nb_correction_rounds; ++j) {
            Correct_Opt l_opt = opt;
            l_opt.min_score = 1.00 - (j + 1) * step_min_score;
            l_opt.weak_region_len_factor = opt.weak_region_len_factor - (opt.nb_correction_rounds - j - 1) * step_wrlf;
            l_opt.max_len_weak_region1 = (j + 1) * step_mlwr1;
            const auto p = getSeeds(l_opt, *g->dbg, in_read, in_qual, false, hap_id, m_km_um, (j + 1) != opt.nb_correction_rounds);
            pair<string, string> c = correctSequence(*g->dbg, l_opt, in_read, in_qual, p.first, p.second, false, nullptr, hap_id, g->hap, g->max_km_cov);
            in_read = move(c.first);
            in_qual = move(c.second);
        }
    } else {
        const auto p = getSeeds(opt, *g->dbg, in_read, in_qual, true, hap_id, m_km_um, false);
        pair<string, string> c = correctSequence(*g->dbg, opt, in_read, in_qual, p.first, p.second, true, nullptr, hap_id, g->hap, g->max_km_cov);
        in_read = move(c.first);
        in_qual = move(c.second);
    }
    *s_out = strdup(in_read.c_str());
    *q_out = strdup(in_qual.c_str());
    return 0;
}

// phasing() (src/Graph.cpp:869): the step the multi-thread branch of search() runs before getSeeds in the second pass
int ref_phasing(void* h, const char* s_raw, const char* s_corr, const char* q_corr, char** s_out, char** q_out) {
    RefGraph* g = (RefGraph*)h;
    string raw(s_raw), corr(s_corr), qual(q_corr);
    std::transform(raw.begin(), raw.end(), raw.begin(), ::toupper);
    std::transform(corr.begin(), corr.end(), corr.begin(), ::toupper);
    pair<string, string> r = phasing(*g->dbg, g->opt, raw, corr, qual);
    *s_out = strdup(r.first.c_str());
    *q_out = strdup(r.second.c_str());
    return 0;
}

// fixSNPs() (src/Alignment.cpp:846): `-f`, applied to the pass-1 read before phasing / getSeeds of the second pass
int ref_fix_snps(void* h, const char* s_corr, char** s_out) {
    RefGraph* g = (RefGraph*)h;
    string corr(s_corr);
    std::transform(corr.begin(), corr.end(), corr.begin(), ::toupper);
    const string r = fixSNPs(g->opt, *g->dbg, corr);
    *s_out = strdup(r.c_str());
    return 0;
}

// detectSNPs (src/Graph.cpp:484) + detectShortCycles (src/Graph.cpp:4660) re-run on the loaded graph with the given
// min_cov_vertices: the stored annotations are cleared first, then one text line per unitig with a non-empty result goes to
// `path`:  sequence \t ambiguity ids (decimal, comma separated) \t isShortCycle \t compacted cycles (';' after each string)
int ref_annotate(void* h, int min_cov, int threads, const char* path) {
    RefGraph* g = (RefGraph*)h;
    Correct_Opt opt = g->opt;
    opt.min_cov_vertices = (size_t)min_cov;
    opt.nb_threads = (size_t)threads;
    opt.verbose = false;
    for (auto& um : *g->dbg) {
        UnitigData* ud = um.getData();
        ud->ambiguity_ids = PairID();
        if (ud->compactedCycles.second != nullptr) delete[] ud->compactedCycles.second;
        ud->compactedCycles = pair<size_t, char*>(0, nullptr);
        ud->setIsCycle(false);
    }
    detectSNPs(*g->dbg, opt);
    detectShortCycles(*g->dbg, opt);
    FILE* f = fopen(path, "w");
    if (!f) return -1;
    for (const auto& um : *g->dbg) {
        const UnitigData* ud = um.getData();
        const vector<const char*> cyc = ud->getCompactCycles();
        if (ud->ambiguity_ids.isEmpty() && cyc.empty() && !ud->isShortCycle()) continue;
        fprintf(f, "%s\t", um.referenceUnitigToString().c_str());
        for (const uint32_t id : ud->ambiguity_ids) fprintf(f, "%u,", id);
        fprintf(f, "\t%d\t", (int)ud->isShortCycle());
        for (const char* c : cyc) fprintf(f, "%s;", c);
        fputc('\n', f);
    }
    fclose(f);
    return 0;
}

// edlibAlign with the reference's IUPAC equality table (src/Common.hpp:262-276) when iupac != 0.
// mode: 0 NW, 1 SHW, 2 HW (EdlibAlignMode); task: 0 DISTANCE, 1 LOC, 2 PATH.
int ref_edlib(const char* q, int ql, const char* t, int tl, int mode, int task, int k, int iupac,
              int* dist, int* nloc, int** endloc, int** startloc, unsigned char** aln, int* alnlen) {
    EdlibAlignConfig cfg = edlibNewAlignConfig(k, (EdlibAlignMode)mode, (EdlibAlignTask)task,
                                               iupac ? edlib_iupac_alpha : NULL, iupac ? (int)sz_edlib_iupac_alpha : 0);
    EdlibAlignResult r = edlibAlign(q, ql, t, tl, cfg);
    *dist = r.editDistance;
    *nloc = r.numLocations;
    *endloc = NULL; *startloc = NULL; *aln = NULL; *alnlen = 0;
    if (r.numLocations > 0 && r.endLocations) {
        *endloc = (int*)malloc(sizeof(int) * r.numLocations);
        memcpy(*endloc, r.endLocations, sizeof(int) * r.numLocations);
    }
    if (r.numLocations > 0 && r.startLocations) {
        *startloc = (int*)malloc(sizeof(int) * r.numLocations);
        memcpy(*startloc, r.startLocations, sizeof(int) * r.numLocations);
    }
    if (r.alignment && r.alignmentLength > 0) {
        *aln = (unsigned char*)malloc(r.alignmentLength);
        memcpy(*aln, r.alignment, r.alignmentLength);
        *alnlen = r.alignmentLength;
    }
    const int status = r.status;
    edlibFreeAlignResult(r);
    return status;
}

// exploreSubGraph (src/GraphTraversal.cpp:456-587) on explicit arguments.
// start / end unitigs are given by key (see um_key) + strand (+ end dist); end_key = ~0 -> no target.
// pids = w_pid.all_pids.  Output (malloc'd u32 stream): for terminal then non-terminal paths:
//   n_paths, then per path: n_um, (key_lo, key_hi, strand, dist, len) * n_um, qual_len, qual bytes padded to 4
static void serialize_paths(const vector<Path<UnitigData>>& v, vector<uint32_t>& out) {
    out.push_back((uint32_t)v.size());
    for (const auto& p : v) {
        const Path<UnitigData>::PathOut po = p.toStringVector();
        const vector<const_UnitigMap<UnitigData>>& vu = po.toVector();
        out.push_back((uint32_t)vu.size());
        for (const auto& um : vu) {
            const uint64_t key = um_key(um);
            out.push_back((uint32_t)key); out.push_back((uint32_t)(key >> 32));
            out.push_back(um.strand ? 1u : 0u); out.push_back((uint32_t)um.dist); out.push_back((uint32_t)um.len);
        }
        const string& q = po.toQualityString();
        out.push_back((uint32_t)q.size());
        for (size_t i = 0; i < q.size(); i += 4) {
            uint32_t w = 0;
            for (size_t j = 0; j < 4 && i + j < q.size(); ++j) w |= ((uint32_t)(unsigned char)q[i + j]) << (8 * j);
            out.push_back(w);
        }
    }
}

int ref_explore_subgraph(void* h, uint64_t start_key, int start_strand, uint64_t end_key, int end_strand, uint32_t end_dist,
                         const char* ref, uint32_t level, uint32_t max_len_path, const uint32_t* pids, uint32_t n_pids,
                         double* scores, uint32_t** out, uint64_t* out_words) {
    RefGraph* g = (RefGraph*)h;
    if (g->by_key.empty()) for (const auto& um : *g->dbg) g->by_key[um_key(um)] = um;
    const size_t k = g->dbg->getK();
    auto it = g->by_key.find(start_key);
    if (it == g->by_key.end()) return -1;
    const_UnitigMap<UnitigData> um = it->second;
    um.dist = 0; um.len = um.size - k + 1; um.strand = (start_strand != 0);
    const_UnitigMap<UnitigData> um_e;
    if (end_key != 0xffffffffffffffffULL) {
        auto ie = g->by_key.find(end_key);
        if (ie == g->by_key.end()) return -1;
        um_e = ie->second;
        um_e.dist = end_dist; um_e.len = 1; um_e.strand = (end_strand != 0);
    }
    WeightsPairID w_pid;
    for (uint32_t i = 0; i < n_pids; ++i) w_pid.all_pids.add(pids[i]);
    vector<Path<UnitigData>> term, nonterm;
    unordered_map<const SharedPairID*, pair<double, bool>, HashSharedPairIDptr> m_pid;
    const pair<double, double> sc = exploreSubGraph(g->opt, w_pid, ref, strlen(ref), max_len_path, um, um_e, level, term, nonterm,
                                                    0xffffffffffffffffULL, m_pid);
    scores[0] = sc.first; scores[1] = sc.second;
    vector<uint32_t> ser;
    serialize_paths(term, ser);
    serialize_paths(nonterm, ser);
    *out = (uint32_t*)malloc(sizeof(uint32_t) * (ser.size() + 1));
    memcpy(*out, ser.data(), sizeof(uint32_t) * ser.size());
    *out_words = ser.size();
    return 0;
}

// explorePathsBFS2 (src/GraphTraversal.cpp:212) when end_key != ~0, else explorePathsBFS (:3).
// Anchors are k-mer hits: (key, strand, dist), len = 1.  Output stream as serialize_paths().
int ref_explore_paths(void* h, uint64_t start_key, int start_strand, uint32_t start_dist, uint64_t end_key, int end_strand,
                      uint32_t end_dist, const char* ref, const uint32_t* pids, uint32_t n_pids, int pass2,
                      uint32_t** out, uint64_t* out_words) {
    RefGraph* g = (RefGraph*)h;
    if (g->by_key.empty()) for (const auto& um : *g->dbg) g->by_key[um_key(um)] = um;
    auto it = g->by_key.find(start_key);
    if (it == g->by_key.end()) return -1;
    const_UnitigMap<UnitigData> um_s = it->second;
    um_s.dist = start_dist; um_s.len = 1; um_s.strand = (start_strand != 0);
    WeightsPairID w_pid;
    for (uint32_t i = 0; i < n_pids; ++i) w_pid.all_pids.add(pids[i]);
    pair<vector<Path<UnitigData>>, bool> r;
    if (end_key != 0xffffffffffffffffULL) {
        auto ie = g->by_key.find(end_key);
        if (ie == g->by_key.end()) return -1;
        const_UnitigMap<UnitigData> um_e = ie->second;
        um_e.dist = end_dist; um_e.len = 1; um_e.strand = (end_strand != 0);
        r = explorePathsBFS2(g->opt, ref, strlen(ref), w_pid, um_s, um_e, pass2 != 0, 0xffffffffffffffffULL);
    } else {
        r = explorePathsBFS(g->opt, ref, strlen(ref), w_pid, um_s, pass2 != 0, 0xffffffffffffffffULL);
    }
    vector<uint32_t> ser;
    serialize_paths(r.first, ser);
    *out = (uint32_t*)malloc(sizeof(uint32_t) * (ser.size() + 1));
    memcpy(*out, ser.data(), sizeof(uint32_t) * ser.size());
    *out_words = ser.size();
    return 0;
}

void ref_free(void* p) { free(p); }

}  // extern "C"
