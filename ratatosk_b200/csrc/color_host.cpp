// color_host.cpp — colouring a graph with long reads: the `long_read_correct` branch of addCoverage (src/Graph.cpp:1561-3366), i.e.
// what `Ratatosk index -2` runs on the k2 graph with the pass-1 corrected long reads (src/Ratatosk.cpp:1213-1240) before detectSNPs /
// detectShortCycles (annotate_host.cpp).  Shared by the CUDA build and tests/hostsim.
//
// Steps, with the reference lines they restate:
//   reading     reads shorter than k or than min_len_2nd_pass are skipped; bases whose quality is below
//               getQual(min_confidence_2nd_pass) become 'N' (:1796-1815; with the defaults that is every base pass 1 left
//               uncorrected, quality '!'); reads with the same name share one id (name_hmap, :1800-1804)
//   mapping     findUnitig over the k-mers of the read, jumping over each mapped run (:1636-1668 / :1720-1744): the K1 exact sweep
//               of the whole batch on the device, runs = maximal stretches of consecutive k-mers of one unitig
//   anchoring   a first pass picks one unitig per read (the longest mapped unitig that already anchors a read, else the longest
//               one, :1655-1663) and the final read ids are dealt in unitig order, ascending first-pass id within a unitig
//               (:2085-2118; sampling_rate = 1).  In the reference this pass is multi-threaded and which unitig "already anchors a
//               read" depends on thread timing, so its ids differ from run to run (and its single-thread branch drops the last
//               read buffer: `index -2 -c 1` colours nothing); here the pass is the sequential reading of that code.  The
//               colouring is therefore defined up to a relabelling of the read ids - that is what the tests compare.
//   colouring   every unitig a read maps to receives the read's id, its unphased coverage grows by the k-mers mapped (:1720-1744)
//   flags       isBranching and the per-edge "shared by >= min_cov_vertices reads" bits (postProcessUnitigs, :1986-2023): one warp
//               per unitig on the device (rtk_edge_flags_kernel)
// Not restated: the subsampling branch taken when estimateHaplotypeCoverage() >= 10 (:2312-3083; it draws from
// std::random_device) - the call fails with RTK_EINVAL and says so when the estimate reaches 10, unless the caller asks to keep
// every read (rtk_opt::reserved bit 0); reads longer than the reference's 1 MB reading buffer are mapped whole (the reference
// would cut them).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "kmer.cuh"
#include "rtk_host_common.hpp"

namespace rtk {

namespace {

struct Run { uint32_t unitig, len; };

// runs of a read from its exact hit list in v_um order (Search.tcc:685-705): a run is emitted contiguously, ascending unitig offset
void runs_of(const rtk_graph_view& g, const std::vector<rtk_hit>& h, std::vector<Run>& out) {
    out.clear();
    size_t i = 0;
    while (i < h.size()) {
        size_t j = i + 1;
        if (g.unitig_off[h[i].unitig + 1] - g.unitig_off[h[i].unitig] != g.k) {
            while (j < h.size() && h[j].unitig == h[i].unitig && h[j].strand == h[i].strand && h[j].dist == h[j - 1].dist + 1 &&
                   (h[i].strand ? h[j].pos == h[j - 1].pos + 1 : h[j].pos + 1 == h[j - 1].pos)) ++j;
        }
        out.push_back(Run{h[i].unitig, (uint32_t)(j - i)});
        i = j;
    }
}

// estimateHaplotypeCoverage (src/Graph.cpp:4185-4233): mean coverage of the arms of the simple bubbles of the graph
uint64_t estimate_hap_cov(const rtk_graph_view& g, const std::vector<uint64_t>& cov) {
    uint64_t tot_cov = 0, nb_km = 0;
    auto succ = [&](uint32_t v, uint32_t s, uint32_t b) -> uint32_t {   // (unitig | strand << 31) or RTK_NONE32
        const uint32_t slot = s ? g.adj[8 * (uint64_t)v + b] : g.adj[8 * (uint64_t)v + 4 + (3 - b)];
        if (slot == RTK_NONE32) return RTK_NONE32;
        const uint32_t vs = s ? (slot >> 31) : (1u - (slot >> 31));
        return (slot & 0x7fffffffu) | (vs << 31);
    };
    auto n_succ = [&](uint32_t v, uint32_t s) { uint32_t c = 0; for (uint32_t b = 0; b < 4; ++b) c += succ(v, s, b) != RTK_NONE32; return c; };
    for (uint64_t u = 0; u < g.n_unitigs; ++u) {
        if (n_succ((uint32_t)u, 1) <= 1) continue;
        uint32_t arms[4], na = 0;
        bool branches = false;
        for (uint32_t b = 0; b < 4; ++b) {
            const uint32_t x = succ((uint32_t)u, 1, b);
            if (x == RTK_NONE32) continue;
            arms[na++] = x;
            branches = branches || n_succ(x & 0x7fffffffu, x >> 31) > 1 || n_succ(x & 0x7fffffffu, 1u - (x >> 31)) > 1;
        }
        if (na < 2 || branches) continue;
        uint32_t end = RTK_NONE32;
        bool simple = true;
        for (uint32_t i = 0; i < na; ++i)
            for (uint32_t b = 0; b < 4; ++b) {
                const uint32_t y = succ(arms[i] & 0x7fffffffu, arms[i] >> 31, b);
                if (y == RTK_NONE32) continue;
                if (end == RTK_NONE32) end = y; else simple = simple && y == end;
            }
        if (!simple) continue;
        for (uint32_t i = 0; i < na; ++i) {
            const uint32_t v = arms[i] & 0x7fffffffu;
            nb_km += g.unitig_off[v + 1] - g.unitig_off[v] - g.k + 1;
            tot_cov += cov[v];
        }
    }
    return nb_km ? tot_cov / nb_km : 0;
}

}  // namespace

struct ColorOut {
    std::vector<uint64_t> kmcov, shared, col_off;
    std::vector<uint32_t> col_ids;
    uint64_t n_ids = 0;           // read ids dealt (ids are 0 .. n_ids - 1)
    std::vector<uint32_t> read_id;   // per input read: its id, ~0u when it received none
};

void color_long_reads_host(rtk_ctx* ctx, const rtk_opt* opt, uint32_t n_reads, const char* seq_pool, const uint64_t* seq_off,
                           const char* qual_pool, const uint64_t* qual_off, const char* name_pool, const uint64_t* name_off,
                           uint32_t min_len, double min_conf, ColorOut& out, uint64_t* stats) {
    if (!ctx->has_graph || !ctx->host_graph) throw std::invalid_argument("no graph uploaded to this context");
    const rtk_graph_view& g = ctx->host_graph->view;
    const uint32_t k = g.k;
    const uint64_t n = g.n_unitigs;
    const uint32_t min_cov = opt ? opt->min_cov_vertices : 2u;
    const int out_qual = opt ? opt->out_qual : 1, max_qual = opt ? opt->max_qual : 40;
    // getQual (src/Common.hpp:410-418)
    const char c_min = (char)(std::min(min_conf, 1.0) * (double)((char)max_qual - out_qual) + 33 + out_qual);

    // ---- reading: first-pass ids by name, masked upper-case copies of the reads that take part
    std::vector<uint32_t> first_id(n_reads, ~0u);
    uint32_t next_first = 0;
    {
        std::unordered_map<std::string, uint32_t> by_name;
        for (uint32_t r = 0; r < n_reads; ++r) {
            const uint64_t len = seq_off[r + 1] - seq_off[r];
            if (len < k || len < min_len) continue;
            if (name_pool) {
                const std::string nm(name_pool + name_off[r], name_off[r + 1] - name_off[r]);
                const auto it = by_name.emplace(nm, next_first);
                first_id[r] = it.first->second;
                if (it.second) ++next_first;
            } else first_id[r] = next_first++;
        }
    }
    std::vector<char> pool(seq_off[n_reads] - seq_off[0] + 1);
    std::vector<uint64_t> off(n_reads + 1);
    for (uint32_t r = 0; r <= n_reads; ++r) off[r] = seq_off[r] - seq_off[0];
    parallel_for(n_reads, [&](size_t b, size_t e) {
        for (size_t r = b; r < e; ++r) {
            const char* s = seq_pool + seq_off[r];
            const char* q = qual_pool ? qual_pool + qual_off[r] : nullptr;
            char* d = pool.data() + off[r];
            const uint64_t len = off[r + 1] - off[r];
            const bool live = first_id[r] != ~0u;
            for (uint64_t i = 0; i < len; ++i) {
                char c = s[i];
                if (c >= 'a' && c <= 'z') c = (char)(c - 32);
                d[i] = (!live || (q && q[i] < c_min)) ? 'N' : c;
            }
        }
    });

    // ---- mapping: K1 exact sweep, batches bounded by the labelled-hit layout
    std::vector<std::vector<Run>> runs(n_reads);
    {
        const uint64_t max_bases = 512ull << 20;
        uint32_t r0 = 0;
        while (r0 < n_reads) {
            uint32_t r1 = r0;
            while (r1 < n_reads && r1 - r0 < (1u << RTK_HIT_READ_BITS) - 1 && (r1 == r0 || off[r1 + 1] - off[r0] <= max_bases)) ++r1;
            std::vector<std::vector<rtk_hit>> per_read;
            search_sequence_host(ctx, r1 - r0, pool.data(), off.data() + r0, RTK_SEARCH_EXACT, per_read, stats);
            parallel_for(r1 - r0, [&](size_t b, size_t e) {
                for (size_t r = b; r < e; ++r) runs_of(g, per_read[r], runs[r0 + r]);
            });
            r0 = r1;
        }
    }

    // ---- anchoring (first pass) and the final ids
    std::vector<std::vector<uint32_t>> anchored(n);   // first-pass ids per unitig (a set: ascending, unique)
    {
        std::vector<uint32_t> centroid(next_first, ~0u);
        for (uint32_t r = 0; r < n_reads; ++r) {
            if (first_id[r] == ~0u) continue;
            uint32_t canon = ~0u, canon_size = 0, len_centroid = 0;
            for (const Run& x : runs[r]) {
                const uint32_t size = (uint32_t)(g.unitig_off[x.unitig + 1] - g.unitig_off[x.unitig]);
                const bool holds = !anchored[x.unitig].empty();
                if (canon == ~0u || (holds && size > len_centroid) || (len_centroid == 0 && size > canon_size)) {
                    canon = x.unitig; canon_size = size;
                    if (holds && size > len_centroid) len_centroid = size;
                }
            }
            if (canon != ~0u) {
                std::vector<uint32_t>& a = anchored[canon];
                if (a.empty() || a.back() != first_id[r]) { if (std::find(a.begin(), a.end(), first_id[r]) == a.end()) a.push_back(first_id[r]); }
            }
        }
    }
    std::vector<uint32_t> final_id(next_first, ~0u);
    uint32_t next_final = 0;
    for (uint64_t u = 0; u < n; ++u) {
        std::vector<uint32_t>& a = anchored[u];
        std::sort(a.begin(), a.end());
        for (const uint32_t id : a) if (final_id[id] == ~0u) final_id[id] = next_final++;
    }
    anchored.clear(); anchored.shrink_to_fit();

    // ---- colouring: (unitig, id) pairs -> sorted unique lists, unphased coverage
    out.kmcov.assign(n, 0); out.shared.assign(n, 0);
    out.read_id.assign(n_reads, ~0u);
    std::vector<uint64_t> pairs;
    std::vector<uint64_t> cov(n, 0);
    for (uint32_t r = 0; r < n_reads; ++r) {
        if (first_id[r] == ~0u || final_id[first_id[r]] == ~0u) continue;
        const uint32_t id = final_id[first_id[r]];
        out.read_id[r] = id;
        for (const Run& x : runs[r]) {
            pairs.push_back(((uint64_t)x.unitig << 32) | id);
            cov[x.unitig] += x.len;
        }
    }
    std::sort(pairs.begin(), pairs.end());
    pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
    out.col_off.assign(n + 1, 0);
    out.col_ids.resize(pairs.size());
    for (size_t i = 0; i < pairs.size(); ++i) { ++out.col_off[(pairs[i] >> 32) + 1]; out.col_ids[i] = (uint32_t)pairs[i]; }
    for (uint64_t u = 0; u < n; ++u) out.col_off[u + 1] += out.col_off[u];
    for (uint64_t u = 0; u < n; ++u) out.kmcov[u] = std::min<uint64_t>(cov[u], 0x7fffffffULL) << 31;   // increaseUnphasedCoverage saturates
    out.n_ids = next_final;

    // ---- the subsampling branch (estimated haplotype coverage >= 10, src/Graph.cpp:2312-2314) is not restated
    for (uint64_t u = 0; u < n; ++u) cov[u] = std::min<uint64_t>(cov[u], 0x7fffffffULL);
    const uint64_t hap_cov = estimate_hap_cov(g, cov);
    if (stats) stats[8] = hap_cov;
    const bool keep_all = opt && (opt->reserved & 1u);   // caller accepts the un-subsampled colouring (a superset of any subsample)
    if (hap_cov >= 10 && !keep_all) throw std::invalid_argument("colouring: estimated haplotype coverage " + std::to_string(hap_cov) + " >= 10: the reference subsamples the reads here (src/Graph.cpp:2312-3083, random), which this library does not do (rtk_opt::reserved bit 0 keeps all reads instead)");

    // ---- flags on the device
    float ms = 0.f;
    edge_flags_run(ctx, min_cov, out.col_off.data(), out.col_ids.data(), out.kmcov.data(), out.shared.data(), &ms);
    if (stats) { stats[4] += pairs.size(); stats[5] += next_final; stats[7] += (uint64_t)(ms * 1e6); }
}

}  // namespace rtk

using namespace rtk;

extern "C" {

int rtk_color_long_reads(rtk_ctx* c, const rtk_opt* opt, uint32_t n_reads, const char* seq_pool, const uint64_t* seq_off, const char* qual_pool,
                         const uint64_t* qual_off, const char* name_pool, const uint64_t* name_off, uint32_t min_len, double min_conf,
                         uint64_t** kmcov, uint64_t** shared, uint64_t** col_off, uint32_t** col_ids, uint32_t** read_id, uint64_t* stats) {
    return guarded([&] {
        if (!c || !seq_pool || !seq_off || !kmcov || !shared || !col_off || !col_ids) throw std::invalid_argument("null argument");
        if ((qual_pool == nullptr) != (qual_off == nullptr) || (name_pool == nullptr) != (name_off == nullptr)) throw std::invalid_argument("pool without offsets");
        DeviceBind bind(c);
        ColorOut o;
        color_long_reads_host(c, opt, n_reads, seq_pool, seq_off, qual_pool, qual_off, name_pool, name_off, min_len, min_conf, o, stats);
        const size_t n = o.kmcov.size();
        *kmcov = (uint64_t*)malloc((n + 1) * 8);
        *shared = (uint64_t*)malloc((n + 1) * 8);
        *col_off = (uint64_t*)malloc((n + 1) * 8);
        *col_ids = (uint32_t*)malloc((o.col_ids.size() + 1) * 4);
        if (read_id) *read_id = (uint32_t*)malloc(((size_t)n_reads + 1) * 4);
        if (!*kmcov || !*shared || !*col_off || !*col_ids || (read_id && !*read_id)) throw std::bad_alloc();
        if (n) { memcpy(*kmcov, o.kmcov.data(), n * 8); memcpy(*shared, o.shared.data(), n * 8); }
        memcpy(*col_off, o.col_off.data(), (n + 1) * 8);
        if (!o.col_ids.empty()) memcpy(*col_ids, o.col_ids.data(), o.col_ids.size() * 4);
        if (read_id && n_reads) memcpy(*read_id, o.read_id.data(), (size_t)n_reads * 4);
    });
}

int rtk_graph_recolor(const rtk_host_graph* g, const uint64_t* kmcov, const uint64_t* shared, const uint64_t* col_off, const uint32_t* col_ids,
                      rtk_host_graph** out) {
    return guarded([&] {
        if (!g || !kmcov || !shared || !col_off || !col_ids || !out) throw std::invalid_argument("null argument");
        const HostGraph hg = recolored_graph(g->view, kmcov, shared, col_off, col_ids);
        rtk_host_graph* r = new rtk_host_graph();
        r->slab = build_slab(hg);
        finish_host_graph(r);
        *out = r;
    });
}

int rtk_rtsk_write(const rtk_host_graph* g, const char* path, const uint64_t* amb_off, const uint32_t* amb_ids, const uint8_t* is_cycle,
                   const uint64_t* cyc_off, const char* cyc_pool) {
    return guarded([&] {
        if (!g || !path || !amb_off || !amb_ids || !is_cycle || !cyc_off || !cyc_pool) throw std::invalid_argument("null argument");
        write_rtsk(g->view, path, amb_off, amb_ids, is_cycle, cyc_off, cyc_pool);
    });
}

}  // extern "C"
