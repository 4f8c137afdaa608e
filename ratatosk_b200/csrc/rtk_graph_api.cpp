// rtk_graph_api.cpp — host-only part of the C ABI (include/rtk.h): errors, options, graph
// loading / flattening / accessors.  No CUDA here; shared by librtk_b200.so and tests/hostsim.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <thread>

#include "kmer.cuh"
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "rtk_host_common.hpp"

namespace rtk {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }

unsigned host_threads() {
    static unsigned n = [] {
        const char* e = getenv("RTK_HOST_THREADS");
        unsigned v = e ? (unsigned)atoi(e) : std::thread::hardware_concurrency();
        if (!e) {   // one process per GPU (torchrun): the ranks of a node share its cores
            const char* lws = getenv("LOCAL_WORLD_SIZE");
            const unsigned ranks = lws ? (unsigned)atoi(lws) : 1u;
            if (ranks > 1) v = std::max(2u, v / ranks);
        }
        if (v == 0) v = 1;
        return std::min(v, 64u);
    }();
    return n;
}

// per-thread override of the host thread count: a correction gang (correct.cpp) owns a share of the cores, and everything
// it starts (parallel_for, the broker's workers) inherits that share through this thread-local
static thread_local unsigned tl_thread_budget = 0;
void set_thread_budget(unsigned n) { tl_thread_budget = n; }
unsigned thread_budget() { return tl_thread_budget ? tl_thread_budget : host_threads(); }

void parallel_for(size_t n, const std::function<void(size_t, size_t)>& body) {
    const unsigned nt = (unsigned)std::min<size_t>(thread_budget(), n);
    if (nt <= 1) { if (n) body(0, n); return; }
    // dynamic chunks: reads differ in length by orders of magnitude
    const size_t chunk = std::max<size_t>(1, n / (nt * 8));
    std::atomic<size_t> next(0);
    std::exception_ptr err;
    std::mutex err_mu;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) {
        th.emplace_back([&] {
            try {
                for (;;) {
                    const size_t b = next.fetch_add(chunk);
                    if (b >= n) break;
                    body(b, std::min(n, b + chunk));
                }
            } catch (...) {
                std::lock_guard<std::mutex> g(err_mu);
                if (!err) err = std::current_exception();
            }
        });
    }
    for (auto& x : th) x.join();
    if (err) std::rethrow_exception(err);
}

void flatten_hits(const std::vector<std::vector<rtk_hit>>& per_read, rtk_hit** hits, uint64_t** off) {
    const size_t n = per_read.size();
    uint64_t total = 0;
    for (const auto& v : per_read) total += v.size();
    *off = (uint64_t*)malloc((n + 1) * sizeof(uint64_t));
    *hits = (rtk_hit*)malloc((total + 1) * sizeof(rtk_hit));
    if (!*off || !*hits) throw std::bad_alloc();
    uint64_t t = 0;
    for (size_t i = 0; i < n; ++i) {
        (*off)[i] = t;
        if (!per_read[i].empty()) memcpy(*hits + t, per_read[i].data(), per_read[i].size() * sizeof(rtk_hit));
        t += per_read[i].size();
    }
    (*off)[n] = t;
}
}  // namespace rtk

using namespace rtk;

extern "C" {

int rtk_version(void) { return 100; }
const char* rtk_last_error(void) { return rtk::g_err.c_str(); }
void rtk_free(void* p) { free(p); }

void rtk_opt_default(rtk_opt* o, int pass) {
    memset(o, 0, sizeof(*o));
    o->k = (pass == 2) ? 63 : 31;
    o->insert_sz = 500; o->min_cov_vertices = 2; o->max_km_cov = 128;
    o->max_len_weak_region1 = 1000; o->max_len_weak_region2 = 5000; o->nb_correction_rounds = 1;
    o->out_qual = 1; o->max_qual = 40; o->trim_qual = 0;
    o->weak_region_len_factor = 0.25; o->large_k_factor = 1.5; o->min_score = 0.0; o->min_confidence_snp_corr = 0.9;
}

int rtk_graph_load(const char* fasta_path, const char* rtsk_path, int k, rtk_host_graph** out) {
    return guarded([&] {
        if (!fasta_path || !out) throw std::invalid_argument("null argument");
        HostGraph hg = load_index(fasta_path, rtsk_path ? rtsk_path : "", k);
        rtk_host_graph* g = new rtk_host_graph();
        g->slab = build_slab(hg);
        finish_host_graph(g);
        *out = g;
    });
}

int rtk_graph_from_unitigs(int k, uint64_t n, const char* const* seqs, rtk_host_graph** out) {
    return guarded([&] {
        if (!seqs || !out) throw std::invalid_argument("null argument");
        HostGraph hg;
        hg.k = k;
        hg.unitigs.reserve(n);
        for (uint64_t i = 0; i < n; ++i) hg.unitigs.emplace_back(seqs[i]);
        rtk_host_graph* g = new rtk_host_graph();
        g->slab = build_slab(hg);
        finish_host_graph(g);
        *out = g;
    });
}

void rtk_graph_free(rtk_host_graph* g) {
    if (!g) return;
    if (g->slab.mapped) munmap(g->slab.data, (size_t)g->slab.bytes);
    else free(g->slab.data);
    delete g;
}

int rtk_graph_get_info(const rtk_host_graph* g, rtk_graph_info* info) {
    if (!g || !info) { set_error("null argument"); return RTK_EINVAL; }
    info->k = g->hdr.k; info->n_unitigs = g->hdr.n_unitigs; info->n_kmers = g->hdr.n_kmers;
    info->pool_bases = g->hdr.pool_bases; info->n_buckets = g->hdr.n_buckets; info->n_gsets = g->hdr.n_gsets;
    info->slab_bytes = g->hdr.total_bytes; info->max_km_cov_graph = g->hdr.max_km_cov_graph;
    return RTK_OK;
}

const void* rtk_graph_slab(const rtk_host_graph* g, uint64_t* bytes) {
    if (!g) return nullptr;
    if (bytes) *bytes = g->slab.bytes;
    return g->slab.data;
}

// The flat cache file IS the slab (flat_graph.h): a versioned header followed by 256-byte aligned SoA / CSR sections whose
// offsets are relative to the file start, so it can be used in place from a read-only mapping - by the host-side logic, by
// the one H2D copy of rtk_graph_upload, and by several processes of one node sharing the page cache.  Replaces re-running
// CompactedDBG::read + readGraphData + the flattening on every `correct` (src/Graph.cpp:722-801, Bifrost/src/IO.tcc:936-1165).
int rtk_graph_save(const rtk_host_graph* g, const char* path) {
    return guarded([&] {
        if (!g || !path) throw std::invalid_argument("null argument");
        const std::string tmp = std::string(path) + ".tmp." + std::to_string((long)getpid());
        FILE* f = fopen(tmp.c_str(), "wb");
        if (!f) throw std::runtime_error(std::string("cannot write ") + tmp);
        const size_t w = fwrite(g->slab.data, 1, g->slab.bytes, f);
        const bool ok = (fclose(f) == 0) && w == g->slab.bytes;
        if (!ok || rename(tmp.c_str(), path) != 0) { remove(tmp.c_str()); throw std::runtime_error(std::string("cannot write ") + path); }
    });
}

int rtk_graph_open(const char* path, rtk_host_graph** out) {
    return guarded([&] {
        if (!path || !out) throw std::invalid_argument("null argument");
        const int fd = open(path, O_RDONLY);
        if (fd < 0) throw std::runtime_error(std::string("cannot open ") + path);
        struct stat st;
        if (fstat(fd, &st) != 0 || st.st_size < (off_t)sizeof(rtk_slab_header)) { close(fd); throw std::runtime_error(std::string("not a flat graph file: ") + path); }
        void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
        close(fd);
        if (m == MAP_FAILED) throw std::runtime_error(std::string("cannot map ") + path);
        madvise(m, (size_t)st.st_size, MADV_WILLNEED);
        rtk_host_graph* g = new rtk_host_graph();
        g->slab.data = (unsigned char*)m;
        g->slab.bytes = (uint64_t)st.st_size;
        g->slab.mapped = true;
        try { finish_host_graph(g); } catch (...) { munmap(m, (size_t)st.st_size); delete g; throw; }
        *out = g;
    });
}

// rtk_graph_load through the cache: `cache_path` (NULL: <rtsk_path>.k<k>.rtkflat, or <fasta_path>.k<k>.rtkflat without colours) is
// opened in place when it exists, is at least as recent as both index files and holds a slab of this k; otherwise the index
// is parsed and flattened and the cache is (re)written - a failed write is not an error, the in-memory graph is returned.
int rtk_graph_load_cached(const char* fasta_path, const char* rtsk_path, int k, const char* cache_path, rtk_host_graph** out, int* from_cache) {
    if (from_cache) *from_cache = 0;
    if (!fasta_path || !out) { set_error("null argument"); return RTK_EINVAL; }
    const bool has_rtsk = rtsk_path && *rtsk_path;
    const std::string cache = (cache_path && *cache_path) ? std::string(cache_path)
                                                          : std::string(has_rtsk ? rtsk_path : fasta_path) + ".k" + std::to_string(k) + ".rtkflat";
    struct stat sc, sf, sr;
    if (stat(cache.c_str(), &sc) == 0 && stat(fasta_path, &sf) == 0 && sc.st_mtime >= sf.st_mtime &&
        (!has_rtsk || (stat(rtsk_path, &sr) == 0 && sc.st_mtime >= sr.st_mtime))) {
        rtk_host_graph* g = nullptr;
        if (rtk_graph_open(cache.c_str(), &g) == RTK_OK) {
            if ((int)g->hdr.k == k) { *out = g; if (from_cache) *from_cache = 1; return RTK_OK; }
            rtk_graph_free(g);
        }
    }
    const int rc = rtk_graph_load(fasta_path, rtsk_path, k, out);
    if (rc != RTK_OK) return rc;
    rtk_graph_save(*out, cache.c_str());   // best effort
    return RTK_OK;
}

int rtk_graph_unitig_seq(const rtk_host_graph* g, uint32_t u, char* buf, uint64_t cap, uint64_t* len) {
    if (!g || u >= g->hdr.n_unitigs) { set_error("bad unitig id"); return RTK_EINVAL; }
    const uint64_t b = g->view.unitig_off[u], e = g->view.unitig_off[u + 1];
    if (len) *len = e - b;
    if (buf) {
        if (cap < e - b) { set_error("buffer too small"); return RTK_EINVAL; }
        for (uint64_t p = b; p < e; ++p) buf[p - b] = "ACGT"[rtk_pool_base(g->view.pool, p)];
    }
    return RTK_OK;
}

int rtk_graph_unitig_words(const rtk_host_graph* g, uint32_t u, uint64_t* kmcov, uint64_t* shared, uint32_t adj[8]) {
    if (!g || u >= g->hdr.n_unitigs) { set_error("bad unitig id"); return RTK_EINVAL; }
    if (kmcov) *kmcov = g->view.kmcov[u];
    if (shared) *shared = g->view.shared[u];
    if (adj) memcpy(adj, g->view.adj + 8 * (uint64_t)u, 32);
    return RTK_OK;
}

int rtk_graph_unitig_colors(const rtk_host_graph* g, uint32_t u, const uint32_t** gids, uint64_t* n_g,
                            const uint32_t** lids, uint64_t* n_l) {
    if (!g || u >= g->hdr.n_unitigs) { set_error("bad unitig id"); return RTK_EINVAL; }
    const uint32_t gs = g->view.gset_of[u];
    if (gs == RTK_NONE32) { *gids = nullptr; *n_g = 0; }
    else { *gids = g->view.gset_ids + g->view.gset_off[gs]; *n_g = g->view.gset_off[gs + 1] - g->view.gset_off[gs]; }
    *lids = g->view.loc_ids + g->view.loc_off[u];
    *n_l = g->view.loc_off[u + 1] - g->view.loc_off[u];
    return RTK_OK;
}


void rtk_seeds_free(rtk_seeds* s) {
    if (!s) return;
    free(s->solid); free(s->solid_off); free(s->weak); free(s->weak_off);
    memset(s, 0, sizeof(*s));
}

}  // extern "C"
