// sim_api.cpp — TEST INFRASTRUCTURE: the device-facing half of the C ABI implemented by
// running the repo's kernel sources on the CPU simulator (cuda_sim.h).  Builds into
// tests/hostsim/_build/librtk_hostsim.so, which the CPU test-suite loads through the same
// ctypes wrapper as the real library.  Never shipped, never loaded by the product.
#include "cuda_sim.h"

#include <cstdlib>
#include <cstring>

#include "../../ratatosk_b200/csrc/k1_lookup.cuh"
#include "../../ratatosk_b200/csrc/rtk_host_common.hpp"

namespace rtk {

static uint64_t sim_k1(rtk_ctx* ctx, uint32_t n_reads, const char* seq, const uint64_t* seq_off, uint32_t flags,
                       RawHitVec& raw, uint64_t* n_probes, std::vector<uint64_t>* dense = nullptr) {
    const rtk_graph_view& g = ctx->host_graph->view;
    const uint32_t k = g.k;
    const bool exact = flags & RTK_SEARCH_EXACT;
    const bool inexact = flags & (RTK_SEARCH_INS | RTK_SEARCH_DEL | RTK_SEARCH_SUBST);
    if (exact && inexact) throw std::invalid_argument("exact and inexact search in one call is not a combination the reference path uses");
    const uint32_t tile = k1_tile_size(k, exact);
    std::vector<uint32_t> tiles;
    build_tiles(n_reads, seq_off, k, tile, tiles, seq, exact ? k : k - 1, exact ? k : k + 1);
    const uint32_t n_tiles = (uint32_t)(tiles.size() / 2);
    const uint64_t total = seq_off[n_reads] - seq_off[0];
    if (dense) dense->assign(total + 1, ~0ULL);
    if (!n_tiles) return 0;
    uint64_t cap = 4 * total + 1024;
    std::vector<rtk_raw_hit> hits(cap);
    unsigned long long counters[2] = {0, 0};
    rtk_k1_params p;
    p.table = g.table; p.n_buckets = g.n_buckets; p.pool = g.pool; p.k = (int)k;
    p.seq = seq; p.seq_off = seq_off; p.tiles = tiles.data(); p.n_tiles = n_tiles; p.tile = tile;
    p.do_subst = (flags & RTK_SEARCH_SUBST) ? 1 : 0;
    p.do_ins = (flags & RTK_SEARCH_INS) ? 1 : 0;
    p.do_del = (flags & RTK_SEARCH_DEL) ? 1 : 0;
    p.hits = hits.data(); p.n_hits = &counters[0]; p.hit_cap = cap; p.n_probes = &counters[1];
    p.dense = nullptr;
    if (dense) p.dense = dense->data();
    const unsigned grid = std::min<uint32_t>(n_tiles, 7);  // fewer blocks than tiles: exercises the grid-stride loop
    if (exact) {
        if (k <= 32) sim_launch(grid, RTK_K1_THREADS, [&] { rtk_k1_exact_kernel<uint64_t>(p); });
        else sim_launch(grid, RTK_K1_THREADS, [&] { rtk_k1_exact_kernel<rtk_u128>(p); });
    } else {
        if (k <= 32) sim_launch(grid, RTK_K1_THREADS, [&] { rtk_k1_inexact_kernel<uint64_t>(p); });
        else sim_launch(grid, RTK_K1_THREADS, [&] { rtk_k1_inexact_kernel<rtk_u128>(p); });
    }
    if (n_probes) *n_probes = counters[1];
    if (dense) return counters[0];
    if (counters[0] > cap) throw std::runtime_error("hostsim: hit buffer overflow");
    raw.resize(counters[0]);
    for (size_t i = 0; i < raw.size(); ++i) { raw[i].a = hits[i].a; raw[i].b = hits[i].b; }
    if (n_probes) *n_probes = counters[1];
    return counters[0];
}

void search_sequence_host(rtk_ctx* ctx, uint32_t n_reads, const char* seq_pool, const uint64_t* seq_off, uint32_t flags,
                          std::vector<std::vector<rtk_hit>>& per_read, uint64_t* stats) {
    if (!ctx->has_graph || !ctx->host_graph) throw std::invalid_argument("no graph uploaded to this context");
    const bool sparse_hint = (flags & RTK_SEARCH_SPARSE_HINT) != 0;
    flags &= ~RTK_SEARCH_SPARSE_HINT;
    RawHitVec raw;
    uint64_t probes = 0, n_raw;
    static const bool list_only = getenv("RTK_K1_LIST") != nullptr;
    if (flags == RTK_SEARCH_EXACT && !list_only && !sparse_hint && seq_off[n_reads] != seq_off[0]) {   // dense exact sweep, like the product
        std::vector<uint64_t> dense;
        std::vector<uint64_t> rel(n_reads + 1);
        for (uint32_t i = 0; i <= n_reads; ++i) rel[i] = seq_off[i] - seq_off[0];
        n_raw = sim_k1(ctx, n_reads, seq_pool + seq_off[0], rel.data(), flags, raw, &probes, &dense);
        resolve_exact_dense(ctx->host_graph->view, n_reads, rel.data(), dense.data(), per_read);
    } else {
        n_raw = sim_k1(ctx, n_reads, seq_pool, seq_off, flags, raw, &probes);
        resolve_batch(ctx->host_graph->view, n_reads, seq_pool, seq_off, flags, raw, per_read);
    }
    if (stats) { stats[0] += probes; stats[1] += n_raw; }
}

}  // namespace rtk

using namespace rtk;

extern "C" {

int rtk_ctx_create(int, rtk_ctx** out) { *out = new rtk_ctx(); return RTK_OK; }
void rtk_ctx_destroy(rtk_ctx* c) { delete c; }
int rtk_ctx_fork(const rtk_ctx* parent, rtk_ctx** out) { *out = new rtk_ctx(*parent); return RTK_OK; }
}
int rtk::ctx_fork_priority(const rtk_ctx* parent, bool, rtk_ctx** out) { *out = new rtk_ctx(*parent); return RTK_OK; }
rtk_ctx* rtk::fork_acquire(rtk_ctx* parent, bool) { return new rtk_ctx(*parent); }
void rtk::fork_release(rtk_ctx*, rtk_ctx* child, bool) { delete child; }
extern "C" {
int rtk_graph_upload(rtk_ctx* c, const rtk_host_graph* g) {
    c->host_graph = g; c->hdr = g->hdr; c->has_graph = true;
    return RTK_OK;
}
int rtk_ctx_sync(rtk_ctx*) { return RTK_OK; }
int rtk_ctx_resident_reads(rtk_ctx*, const char*, const uint64_t*, uint32_t, uint64_t) { return RTK_OK; }
int rtk_is_hostsim(void) { return 1; }

int rtk_search_sequence(rtk_ctx* c, uint32_t n_reads, const char* seq_pool, const uint64_t* seq_off, uint32_t flags,
                        rtk_hit** hits, uint64_t** hit_off, uint64_t* stats) {
    return guarded([&] {
        std::vector<std::vector<rtk_hit>> per_read;
        search_sequence_host(c, n_reads, seq_pool, seq_off, flags, per_read, stats);
        flatten_hits(per_read, hits, hit_off);
    });
}

int rtk_get_seeds(rtk_ctx* c, const rtk_opt* opt, int pass, uint32_t n_reads, const char* seq_pool,
                  const uint64_t* seq_off, rtk_seeds* out, uint64_t* stats) {
    return guarded([&] {
        std::vector<std::vector<rtk_hit>> solid, weak;
        get_seeds_host(c, *opt, pass, n_reads, seq_pool, seq_off, solid, weak, stats);
        flatten_hits(solid, &out->solid, &out->solid_off);
        flatten_hits(weak, &out->weak, &out->weak_off);
    });
}

}  // extern "C"
