// rtk_api.cu — C ABI (include/rtk.h): contexts, graph residency, K1 driver.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "k1_lookup.cuh"
#include "rtk_internal.hpp"
#include "rtk_host_common.hpp"

namespace rtk {

std::atomic<uint64_t> g_launches{0}, g_h2d_bytes{0}, g_d2h_bytes{0};


static inline double now_ns() {
    return (double)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------- K1 driver
uint64_t k1_launch(rtk_ctx* ctx, uint32_t n_reads, const char* d_seq, const uint64_t* d_seq_off,
                   const uint64_t* h_seq_off, uint32_t flags, uint64_t* n_probes, float* kernel_ms, const char* h_seq, bool dense, bool sparse_tiles) {
    if (!ctx->has_graph) throw std::invalid_argument("no graph uploaded to this context");
    if (n_reads >= (1u << RTK_HIT_READ_BITS)) throw std::invalid_argument("more than 2^24 reads in one batch");
    const uint32_t k = ctx->hdr.k;
    const bool exact = flags & RTK_SEARCH_EXACT;
    const bool inexact = flags & (RTK_SEARCH_INS | RTK_SEARCH_DEL | RTK_SEARCH_SUBST);
    if (exact && inexact) throw std::invalid_argument("exact and inexact search in one call is not a combination the reference path uses");
    const uint64_t total = h_seq_off[n_reads] - h_seq_off[0];

    const uint32_t tile = k1_tile_size(k, exact);
    std::vector<uint32_t> tiles;
    // the exact sweep normally runs on whole reads (nothing to skip) unless the caller says they are mostly masked (list form + h_seq)
    if (exact && !dense && h_seq && sparse_tiles) build_tiles(n_reads, h_seq_off, k, tile, tiles, h_seq, k, k);
    else build_tiles(n_reads, h_seq_off, k, tile, tiles, exact ? nullptr : h_seq, k - 1, k + 1);
    const uint32_t n_tiles = (uint32_t)(tiles.size() / 2);
    ctx->d_counters.reserve(64);
    RTK_CUDA(cudaMemsetAsync(ctx->d_counters.p, 0, 64, ctx->stream));
    if (n_tiles == 0) {
        if (dense) {   // nothing to probe: every position answers "not in the graph"
            ctx->d_hits.reserve(total * 8 + 64);
            RTK_CUDA(cudaMemsetAsync(ctx->d_hits.p, 0xFF, total * 8, ctx->stream));
        }
        if (n_probes) *n_probes = 0;
        if (kernel_ms) *kernel_ms = 0.f;
        return 0;
    }
    ctx->d_tiles.reserve(tiles.size() * 4);
    RTK_CUDA(counted_memcpy_async(ctx->d_tiles.p, tiles.data(), tiles.size() * 4, cudaMemcpyHostToDevice, ctx->stream));

    if (dense && !exact) throw std::invalid_argument("dense output is an exact-sweep mode");
    uint64_t cap = std::max<uint64_t>(1u << 20, exact ? total + 1024 : 2 * total + 1024);
    for (int attempt = 0; attempt < 3; ++attempt) {
        ctx->d_hits.reserve(dense ? total * 8 + 64 : cap * sizeof(rtk_raw_hit));
        cap = dense ? ~0ull : ctx->d_hits.cap / sizeof(rtk_raw_hit);
        if (dense) RTK_CUDA(cudaMemsetAsync(ctx->d_hits.p, 0xFF, total * 8, ctx->stream));
        rtk_k1_params p;
        p.table = ctx->dview.table; p.n_buckets = ctx->dview.n_buckets; p.pool = ctx->dview.pool; p.k = (int)k;
        p.seq = d_seq; p.seq_off = d_seq_off; p.tiles = ctx->d_tiles.as<uint32_t>(); p.n_tiles = n_tiles; p.tile = tile;
        p.do_subst = (flags & RTK_SEARCH_SUBST) ? 1 : 0;
        p.do_ins = (flags & RTK_SEARCH_INS) ? 1 : 0;
        p.do_del = (flags & RTK_SEARCH_DEL) ? 1 : 0;
        p.hits = ctx->d_hits.as<rtk_raw_hit>();
        p.n_hits = ctx->d_counters.as<unsigned long long>();
        p.hit_cap = cap;
        p.n_probes = n_probes ? ctx->d_counters.as<unsigned long long>() + 1 : nullptr;
        p.dense = dense ? ctx->d_hits.as<uint64_t>() : nullptr;
        RTK_CUDA(cudaMemsetAsync(ctx->d_counters.p, 0, 64, ctx->stream));
        // grid: whole waves of resident CTAs (148 SMs x 8 CTAs of 256 threads), grid-stride over tiles
        const uint32_t grid = std::min<uint32_t>(n_tiles, (uint32_t)ctx->sm_count * 8u);
        RTK_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
        ++g_launches;
        if (exact) {
            if (k <= 32) rtk_k1_exact_kernel<uint64_t><<<grid, RTK_K1_THREADS, 0, ctx->stream>>>(p);
            else rtk_k1_exact_kernel<rtk_u128><<<grid, RTK_K1_THREADS, 0, ctx->stream>>>(p);
        } else {
            if (k <= 32) rtk_k1_inexact_kernel<uint64_t><<<grid, RTK_K1_THREADS, 0, ctx->stream>>>(p);
            else rtk_k1_inexact_kernel<rtk_u128><<<grid, RTK_K1_THREADS, 0, ctx->stream>>>(p);
        }
        RTK_CUDA(cudaGetLastError());
        RTK_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
        unsigned long long cnt[2];
        RTK_CUDA(counted_memcpy_async(cnt, ctx->d_counters.p, sizeof(cnt), cudaMemcpyDeviceToHost, ctx->stream));
        RTK_CUDA(cudaStreamSynchronize(ctx->stream));
        if (kernel_ms) RTK_CUDA(cudaEventElapsedTime(kernel_ms, ctx->ev0, ctx->ev1));
        if (n_probes) *n_probes = cnt[1];
        if (cnt[0] <= cap) return cnt[0];
        cap = cnt[0] + 1024;  // overflow: rerun with the exact size
    }
    throw std::runtime_error("K1 hit buffer overflow persisted");
}

void search_sequence_host(rtk_ctx* ctx, uint32_t n_reads, const char* seq_pool, const uint64_t* seq_off, uint32_t flags,
                          std::vector<std::vector<rtk_hit>>& per_read, uint64_t* stats) {
    const double t_start = now_ns();
    DeviceBind bind(ctx);
    const bool sparse_hint = (flags & RTK_SEARCH_SPARSE_HINT) != 0;
    flags &= ~RTK_SEARCH_SPARSE_HINT;
    if (!ctx->has_graph || !ctx->host_graph) throw std::invalid_argument("no graph uploaded to this context");
    const uint64_t total = seq_off[n_reads] - seq_off[0];
    // offsets relative to the start of this batch's pool
    std::vector<uint64_t> rel(n_reads + 1);
    for (uint32_t i = 0; i <= n_reads; ++i) rel[i] = seq_off[i] - seq_off[0];
    const char* d_seq;
    const uint64_t* d_off;
    if (flags == RTK_SEARCH_EXACT && ctx->resident_seq && ctx->resident_n == n_reads && ctx->resident_total == total) {
        d_seq = ctx->resident_seq; d_off = ctx->resident_off;   // the caller's copy in HBM (same bytes as seq_pool)
        ctx->resident_seq = nullptr;
    } else {
        ctx->d_seq.reserve(total + 16);
        ctx->d_seq_off.reserve((n_reads + 1) * 8);
        RTK_CUDA(counted_memcpy_async(ctx->d_seq.p, seq_pool + seq_off[0], total, cudaMemcpyHostToDevice, ctx->stream));
        RTK_CUDA(counted_memcpy_async(ctx->d_seq_off.p, rel.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        d_seq = ctx->d_seq.as<char>(); d_off = ctx->d_seq_off.as<uint64_t>();
    }
    uint64_t probes = 0;
    float kms = 0.f;
    // exact sweeps come back DENSE (8 bytes per read position, read order) through pinned memory: no labels, no sort, no
    // bucketing; the list form (16 bytes per hit, any order) is kept for the sparse one-edit sweeps.  RTK_K1_LIST=1: list form.
    static const bool list_only = getenv("RTK_K1_LIST") != nullptr;
    const bool dense = (flags == RTK_SEARCH_EXACT) && !list_only && !sparse_hint && total != 0;
    const uint64_t n_raw = k1_launch(ctx, n_reads, d_seq, d_off, rel.data(), flags, &probes, &kms, seq_pool + seq_off[0], dense, sparse_hint);
    if (dense) {
        ctx->h_dense.reserve(total * 8 + 64);
        // chunked so that decoding could overlap; one stream, pinned landing zone
        RTK_CUDA(counted_memcpy_async(ctx->h_dense.p, ctx->d_hits.p, total * 8, cudaMemcpyDeviceToHost, ctx->stream));
        stream_wait(ctx->stream);
        resolve_exact_dense(ctx->host_graph->view, n_reads, rel.data(), ctx->h_dense.as<uint64_t>(), per_read);
    } else {
        RawHitVec raw(n_raw);
        if (n_raw) {
            RTK_CUDA(counted_memcpy_async(raw.data(), ctx->d_hits.p, n_raw * sizeof(RawHit), cudaMemcpyDeviceToHost, ctx->stream));
            RTK_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        resolve_batch(ctx->host_graph->view, n_reads, seq_pool, seq_off, flags, raw, per_read);
    }
    if (stats) {
        stats[0] += probes;
        stats[1] += n_raw;
        stats[2] += (uint64_t)(kms * 1e6);
        stats[3] += (uint64_t)(now_ns() - t_start);
    }
}

}  // namespace rtk

using namespace rtk;

static void create_side_streams(rtk_ctx* c, int prio = 0) {
    RTK_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    for (int k = 0; k < 6; ++k) {
        RTK_CUDA(cudaStreamCreateWithPriority(&c->side[k], cudaStreamNonBlocking, prio));
        RTK_CUDA(cudaEventCreateWithFlags(&c->ev_join[k], cudaEventDisableTiming));
    }
}

rtk_ctx* rtk::fork_acquire(rtk_ctx* parent, bool high) {
    {
        std::lock_guard<std::mutex> lk(parent->fork_mu);
        auto& v = parent->fork_cache[high ? 1 : 0];
        if (!v.empty()) { rtk_ctx* c = v.back(); v.pop_back(); return c; }
    }
    rtk_ctx* c = nullptr;
    if (ctx_fork_priority(parent, high, &c) != RTK_OK) throw std::runtime_error(std::string("rtk_ctx_fork: ") + rtk_last_error());
    return c;
}
void rtk::fork_release(rtk_ctx* parent, rtk_ctx* child, bool high) {
    if (!child) return;
    std::lock_guard<std::mutex> lk(parent->fork_mu);
    parent->fork_cache[high ? 1 : 0].push_back(child);
}
static void drop_fork_cache(rtk_ctx* c) {
    std::vector<rtk_ctx*> all;
    {
        std::lock_guard<std::mutex> lk(c->fork_mu);
        for (auto& v : c->fork_cache) { all.insert(all.end(), v.begin(), v.end()); v.clear(); }
    }
    for (rtk_ctx* f : all) rtk_ctx_destroy(f);
}

extern "C" {

int rtk_ctx_create(int device, rtk_ctx** out) {
    return guarded([&] {
        // 32 hardware work queues instead of the default 8: the broker's service contexts own ~20 streams, and streams that
        // alias onto one queue serialise on each other's (latency-bound) kernels.  Only effective before CUDA initialises in
        // this process; a host program that initialises CUDA first (bench.py: torch) sets the variable itself.
        setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0) throw CudaError(std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU path)");
        if (device < 0 || device >= n) throw std::invalid_argument("device index out of range");
        RTK_CUDA(cudaSetDevice(device));
        // The service threads of the correction broker spin on their streams by default: measured on B200 with 16 host
        // cores, blocking in the driver (RTK_BLOCKING_SYNC=1) frees cores for the workers but adds wake-up latency to
        // every batch and lowers throughput by ~30 % (the region chains are latency-bound).
        if (getenv("RTK_BLOCKING_SYNC")) { if (cudaSetDeviceFlags(cudaDeviceScheduleBlockingSync) != cudaSuccess) cudaGetLastError(); }
        rtk_ctx* c = new rtk_ctx();
        c->device = device;
        RTK_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        RTK_CUDA(cudaEventCreate(&c->ev0));
        RTK_CUDA(cudaEventCreate(&c->ev1));
        create_side_streams(c);
        RTK_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
        *out = c;
    });
}

int rtk_ctx_fork(const rtk_ctx* parent, rtk_ctx** out) { return rtk::ctx_fork_priority(parent, false, out); }

}  // extern "C"

// fork whose streams run at the device's highest (high = true) or lowest stream priority: the short K4 / K5 batches of the
// correction broker must not queue behind the long region-engine launches
int rtk::ctx_fork_priority(const rtk_ctx* parent, bool high, rtk_ctx** out) {
    return guarded([&] {
        if (!parent || !out) throw std::invalid_argument("null argument");
        RTK_CUDA(cudaSetDevice(parent->device));
        int lo = 0, hi = 0;   // numerically lowest = highest priority
        RTK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        const int prio = high ? hi : lo;
        rtk_ctx* c = new rtk_ctx();
        c->device = parent->device;
        RTK_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio));
        RTK_CUDA(cudaEventCreate(&c->ev0));
        RTK_CUDA(cudaEventCreate(&c->ev1));
        create_side_streams(c, prio);
        c->sm_count = parent->sm_count;
        // the graph is shared, never owned: the parent must outlive the fork
        c->d_slab = parent->d_slab; c->owns_slab = false; c->has_graph = parent->has_graph;
        c->hdr = parent->hdr; c->dview = parent->dview; c->host_graph = parent->host_graph;
        *out = c;
    });
}

extern "C" {

void rtk_ctx_destroy(rtk_ctx* c) {
    if (!c) return;
    drop_fork_cache(c);
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->owns_slab && c->d_slab) cudaFree((void*)c->d_slab);
    c->d_seq.release(); c->d_seq_off.release(); c->d_tiles.release(); c->d_hits.release(); c->d_counters.release();
    for (auto& b : c->d_aux) b.release();
    for (auto& b : c->d_sub) b.release();
    for (auto& b : c->h_pin) b.release();
    for (auto& b : c->d_rg) b.release();
    for (auto& b : c->h_rg) b.release();
    c->d_fs.release(); c->h_fs.release(); c->h_dense.release();
    if (c->host_copy.data) free(c->host_copy.data);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    for (int k = 0; k < 6; ++k) {
        if (c->ev_join[k]) cudaEventDestroy(c->ev_join[k]);
        if (c->side[k]) cudaStreamDestroy(c->side[k]);
    }
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int rtk_graph_upload(rtk_ctx* c, const rtk_host_graph* g) {
    return guarded([&] {
        if (!c || !g) throw std::invalid_argument("null argument");
        RTK_CUDA(cudaSetDevice(c->device));
        drop_fork_cache(c);   // cached forks share the previous graph
        if (c->owns_slab && c->d_slab) RTK_CUDA(cudaFree((void*)c->d_slab));
        void* d = nullptr;
        RTK_CUDA(cudaMalloc(&d, g->slab.bytes));
        RTK_CUDA(counted_memcpy_async(d, g->slab.data, g->slab.bytes, cudaMemcpyHostToDevice, c->stream));
        RTK_CUDA(cudaStreamSynchronize(c->stream));
        c->d_slab = (const unsigned char*)d;
        c->owns_slab = true;
        c->hdr = g->hdr;
        c->dview = rtk_make_view(d, c->hdr);
        c->host_graph = g;
        c->has_graph = true;
    });
}

int rtk_graph_adopt_device(rtk_ctx* c, const void* dev_slab, uint64_t bytes) {
    return guarded([&] {
        if (!c || !dev_slab) throw std::invalid_argument("null argument");
        RTK_CUDA(cudaSetDevice(c->device));
        if (bytes < sizeof(rtk_slab_header)) throw std::invalid_argument("slab too small");
        drop_fork_cache(c);   // cached forks share the previous graph
        // the host-side anchor logic needs a host mirror: copy the slab back once
        if (c->host_copy.data) { free(c->host_copy.data); c->host_copy.data = nullptr; }
        c->host_copy.data = (unsigned char*)aligned_alloc(256, (bytes + 255) & ~(uint64_t)255);
        c->host_copy.bytes = bytes;
        RTK_CUDA(counted_memcpy_async(c->host_copy.data, dev_slab, bytes, cudaMemcpyDeviceToHost, c->stream));
        RTK_CUDA(cudaStreamSynchronize(c->stream));
        c->host_graph_owned.slab = c->host_copy;
        finish_host_graph(&c->host_graph_owned);
        if (c->owns_slab && c->d_slab) RTK_CUDA(cudaFree((void*)c->d_slab));
        c->d_slab = (const unsigned char*)dev_slab;
        c->owns_slab = false;
        c->hdr = c->host_graph_owned.hdr;
        c->dview = rtk_make_view(dev_slab, c->hdr);
        c->host_graph = &c->host_graph_owned;
        c->has_graph = true;
    });
}

int rtk_ctx_sync(rtk_ctx* c) {
    return guarded([&] { RTK_CUDA(cudaStreamSynchronize(c->stream)); });
}

int rtk_search_sequence(rtk_ctx* c, uint32_t n_reads, const char* seq_pool, const uint64_t* seq_off, uint32_t flags,
                        rtk_hit** hits, uint64_t** hit_off, uint64_t* stats) {
    return guarded([&] {
        if (!c || !seq_pool || !seq_off || !hits || !hit_off) throw std::invalid_argument("null argument");
        RTK_CUDA(cudaSetDevice(c->device));
        std::vector<std::vector<rtk_hit>> per_read;
        search_sequence_host(c, n_reads, seq_pool, seq_off, flags, per_read, stats);
        flatten_hits(per_read, hits, hit_off);
    });
}

int rtk_k1_sweep_device(rtk_ctx* c, uint32_t n_reads, const char* dev_seq, const uint64_t* dev_seq_off,
                        const uint64_t* host_seq_off, uint32_t flags, uint64_t* n_probes, uint64_t* n_raw_hits,
                        float* kernel_ms) {
    return guarded([&] {
        if (!c || !dev_seq || !dev_seq_off || !host_seq_off) throw std::invalid_argument("null argument");
        RTK_CUDA(cudaSetDevice(c->device));
        uint64_t probes = 0;
        const uint64_t n = k1_launch(c, n_reads, dev_seq, dev_seq_off, host_seq_off, flags, &probes, kernel_ms);
        if (n_probes) *n_probes = probes;
        if (n_raw_hits) *n_raw_hits = n;
    });
}

int rtk_get_seeds(rtk_ctx* c, const rtk_opt* opt, int pass, uint32_t n_reads, const char* seq_pool,
                  const uint64_t* seq_off, rtk_seeds* out, uint64_t* stats) {
    return guarded([&] {
        if (!c || !opt || !seq_pool || !seq_off || !out) throw std::invalid_argument("null argument");
        RTK_CUDA(cudaSetDevice(c->device));
        std::vector<std::vector<rtk_hit>> solid, weak;
        get_seeds_host(c, *opt, pass, n_reads, seq_pool, seq_off, solid, weak, stats);
        flatten_hits(solid, &out->solid, &out->solid_off);
        flatten_hits(weak, &out->weak, &out->weak_off);
    });
}

int rtk_ctx_resident_reads(rtk_ctx* c, const char* dev_seq_pool, const uint64_t* dev_seq_off, uint32_t n_reads, uint64_t total_bases) {
    if (!c) { set_error("null argument"); return RTK_EINVAL; }
    c->resident_seq = dev_seq_pool; c->resident_off = dev_seq_off; c->resident_n = n_reads; c->resident_total = total_bases;
    return RTK_OK;
}

int rtk_correct_batch_resident(rtk_ctx* c, const rtk_opt* opt, int pass, uint32_t n_reads, const char* seq_pool,
                               const uint64_t* seq_off, const char* dev_seq_pool, const uint64_t* dev_seq_off,
                               const char* qual_pool, const uint64_t* qual_off, char** out_seq_pool, char** out_qual_pool,
                               uint64_t** out_off, uint64_t* stats) {
    if (!c || !seq_off || !dev_seq_pool || !dev_seq_off) { set_error("null argument"); return RTK_EINVAL; }
    c->resident_seq = dev_seq_pool; c->resident_off = dev_seq_off; c->resident_n = n_reads;
    c->resident_total = seq_off[n_reads] - seq_off[0];
    const int rc = rtk_correct_batch(c, opt, pass, n_reads, seq_pool, seq_off, qual_pool, qual_off, out_seq_pool, out_qual_pool, out_off, stats);
    c->resident_seq = nullptr;
    return rc;
}

}  // extern "C"
