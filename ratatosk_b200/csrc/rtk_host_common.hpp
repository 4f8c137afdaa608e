// rtk_host_common.hpp — host helpers shared by the CUDA build and the tests/hostsim build.
#pragma once
#include <algorithm>
#include <cstring>
#include <functional>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "k1_lookup_layout.h"
#include "rtk_internal.hpp"

#ifndef RTK_K1_THREADS
#define RTK_K1_THREADS 256
#define RTK_K1_TILE 224 /* read positions per CTA: leaves room for the 1/(k-1) stretch of insertion strings */
#endif

namespace rtk {

#ifdef RTK_HOSTSIM
struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
#endif

template <typename F> inline int guarded(F&& f) {
    try {
        f();
        return RTK_OK;
    } catch (const CudaError& e) {
        set_error(e.what());
        return RTK_ECUDA;
    } catch (const std::bad_alloc&) {
        set_error("out of host memory");
        return RTK_ENOMEM;
    } catch (const std::invalid_argument& e) {
        set_error(e.what());
        return RTK_EINVAL;
    } catch (const std::exception& e) {
        set_error(e.what());
        return RTK_EIO;
    }
}

inline void finish_host_graph(rtk_host_graph* g) {
    memcpy(&g->hdr, g->slab.data, sizeof(rtk_slab_header));
    if (g->hdr.magic != RTK_SLAB_MAGIC || g->hdr.version != RTK_SLAB_VERSION) throw std::runtime_error("not a flat graph slab");
    if (g->hdr.total_bytes != g->slab.bytes) throw std::runtime_error("flat graph slab size mismatch");
    g->view = rtk_make_view(g->slab.data, g->hdr);
}


// read positions per K1 tile: one per thread
inline uint32_t k1_tile_size(uint32_t, bool) { return RTK_K1_THREADS; }

// h_seq (optional): the reads on the host, h_seq + h_seq_off[r] = read r.  When given, a tile is emitted only if some window
// STARTING inside it can produce a k-mer: `need` A/C/G/T bases among the `span` read bases it draws from (exact sweep: k of
// k; one-edit sweeps: at least k-1 of k+1 - a substitution / insertion window uses k or k-1 consecutive read bases, a
// deletion window k of k+1 with the dropped base free to be anything, e.g. an 'N').  The masked copies getSeeds sweeps
// inexactly are mostly 'N' (src/Graph.cpp:106-191); a tile without such a window can neither probe nor hit.  Conservative: a
// live tile may still hold no valid window, the kernel decides per window.
unsigned host_threads();
void set_thread_budget(unsigned n);   // thread-local share of the host threads (0 = all); inherited by parallel_for / the broker
unsigned thread_budget();
void parallel_for(size_t n, const std::function<void(size_t, size_t)>& body);  // body(begin, end) on chunks

inline void build_tiles(uint32_t n_reads, const uint64_t* h_seq_off, uint32_t k, uint32_t tile, std::vector<uint32_t>& tiles,
                        const char* h_seq = nullptr, uint32_t need = 0, uint32_t span = 0) {
    tiles.clear();
    if (!(h_seq && need && span >= need)) {
        for (uint32_t r = 0; r < n_reads; ++r) {
            const uint64_t len = h_seq_off[r + 1] - h_seq_off[r];
            if (len >= (1ULL << RTK_HIT_POS_BITS)) throw std::invalid_argument("read longer than 2^30 bases");
            if (len < k) continue;
            // positions l with l + k - 1 <= len (one past the last full k-mer: insertion windows use k-1 read bases)
            const uint32_t npos = (uint32_t)(len - k + 2);
            for (uint32_t t0 = 0; t0 < npos; t0 += tile) { tiles.push_back(r); tiles.push_back(t0); }
        }
        return;
    }
    for (uint32_t r = 0; r < n_reads; ++r)
        if (h_seq_off[r + 1] - h_seq_off[r] >= (1ULL << RTK_HIT_POS_BITS)) throw std::invalid_argument("read longer than 2^30 bases");
    // liveness of every tile, reads in parallel: a tile is live if a window of `span` bases starting in it holds >= need A/C/G/T
    std::vector<uint64_t> first_tile(n_reads + 1, 0);
    for (uint32_t r = 0; r < n_reads; ++r) {
        const uint64_t len = h_seq_off[r + 1] - h_seq_off[r];
        first_tile[r + 1] = first_tile[r] + (len < k ? 0 : ((len - k + 2) + tile - 1) / tile);
    }
    std::vector<uint8_t> live(first_tile[n_reads], 0);
    parallel_for(n_reads, [&](size_t rb, size_t re) {
        for (size_t r = rb; r < re; ++r) {
            const uint64_t len = h_seq_off[r + 1] - h_seq_off[r];
            if (len < k) continue;
            const uint64_t npos = len - k + 2;
            const unsigned char* s = (const unsigned char*)h_seq + h_seq_off[r];
            uint8_t* lv = live.data() + first_tile[r];
            auto ok = [](unsigned char c) -> uint32_t { return (uint32_t)(c == 'A') | (uint32_t)(c == 'C') | (uint32_t)(c == 'G') | (uint32_t)(c == 'T'); };
            uint32_t cnt = 0;   // valid bases in [l, min(len, l + span))
            for (uint64_t i = 0; i < std::min<uint64_t>(len, span); ++i) cnt += ok(s[i]);
            for (uint64_t l = 0; l < npos; ++l) {
                if (cnt >= need) { lv[l / tile] = 1; const uint64_t nl = (l / tile + 1) * tile; l = nl - 1; cnt = 0; for (uint64_t i = nl; i < std::min<uint64_t>(len, nl + span); ++i) cnt += ok(s[i]); continue; }
                cnt -= ok(s[l]);
                if (l + span < len) cnt += ok(s[l + span]);
            }
        }
    });
    for (uint32_t r = 0; r < n_reads; ++r) {
        const uint64_t nt = first_tile[r + 1] - first_tile[r];
        for (uint64_t t = 0; t < nt; ++t) if (live[first_tile[r] + t]) { tiles.push_back(r); tiles.push_back((uint32_t)(t * tile)); }
    }
}

// Host worker threads for the per-read anchor logic (reads are independent, like the reference's worker
// loop, src/Ratatosk.cpp:727-906).  RTK_HOST_THREADS overrides the default of hardware_concurrency (<= 64).  (declared above)

void flatten_hits(const std::vector<std::vector<rtk_hit>>& per_read, rtk_hit** hits, uint64_t** off);

}  // namespace rtk
