// region_host.hpp — host half of the device-resident region engine shared by the CUDA launcher (region.cu) and the CPU
// simulator (tests/hostsim/sim_region.cpp): capacities, argument checks, result unpacking.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <stdexcept>
#include <vector>

#include "../../include/rtk.h"
#include "region.cuh"
#include "traceback_host.hpp"

namespace rtk {

struct RegionCaps {
    uint32_t str_cap, mat_cells, tmp_cap, arena_cap, chain_nodes_cap, chain_len_cap;
};

// capacities of the per-warp scratch; RTK_RG_* environment variables override (debugging / tests of the bail path)
inline RegionCaps region_caps() {
    auto env = [](const char* n, uint32_t d) { const char* e = getenv(n); return e ? (uint32_t)strtoul(e, nullptr, 10) : d; };
    RegionCaps c;
    c.str_cap = env("RTK_RG_STR_CAP", 16384);
    c.mat_cells = env("RTK_RG_MAT_CELLS", (1u << 20) / 20 + 64);   // edlib's direct traceback holds < 1 MiB of state (20 B per cell)
    c.tmp_cap = env("RTK_RG_TMP_CAP", 192u << 10);
    c.arena_cap = env("RTK_RG_ARENA_CAP", 1536u << 10);
    c.chain_nodes_cap = env("RTK_RG_CHAIN_NODES", 4096);
    c.chain_len_cap = env("RTK_RG_CHAIN_LEN", 32768);
    return c;
}

inline void region_fill_params(rtk_rg_params& p, const rtk_opt& opt, int pass, const RegionCaps& caps) {
    p.str_cap = caps.str_cap; p.mat_cells = caps.mat_cells; p.tmp_cap = caps.tmp_cap; p.arena_cap = caps.arena_cap;
    p.chain_nodes_cap = caps.chain_nodes_cap; p.chain_len_cap = caps.chain_len_cap;
    p.scratch_per_warp = rtk_rg_make_layout(caps.str_cap, caps.mat_cells, caps.tmp_cap, caps.arena_cap, caps.chain_nodes_cap, caps.chain_len_cap).total;
    p.min_cov = opt.min_cov_vertices;
    p.pass2 = (pass == 2) ? 1u : 0u;
    p.max_len_weak_region = (pass == 2) ? opt.max_len_weak_region2 : opt.max_len_weak_region1;
    p.max_len_subpath = (uint32_t)static_cast<size_t>(opt.k * opt.large_k_factor);   // src/GraphTraversal.cpp:594
    p.out_qual = opt.out_qual; p.max_qual = opt.max_qual;
    p.wrlf = opt.weak_region_len_factor; p.min_score = opt.min_score;
    p.tb_limit = tb_limit();
}

inline void region_check_calls(uint32_t n_calls, const rtk_region_call_t* calls, uint64_t win_bytes, uint64_t n_weak, uint64_t n_pids, uint64_t n_unitigs,
                               uint32_t k) {
    for (uint32_t i = 0; i < n_calls; ++i) {
        const rtk_region_call_t& c = calls[i];
        if (c.win_off + c.win_len > win_bytes || c.weak_off + c.n_weak > n_weak || c.pid_off + c.pid_len > n_pids)
            throw std::invalid_argument("region call range outside its pool");
        if (c.start_unitig >= n_unitigs || (c.has_end && c.end_unitig >= n_unitigs)) throw std::invalid_argument("bad unitig id");
        const uint64_t pos2 = c.has_end ? (uint64_t)c.end_pos : (uint64_t)c.s_len - k;
        if ((!c.has_end && c.s_len < k) || pos2 < c.start_pos || pos2 - c.start_pos + k != c.win_len)
            throw std::invalid_argument("region window does not span [start_pos, pos_um_solid2 + k)");
    }
}

// longest windows first (a region is a chain of dependent steps roughly proportional to its span)
inline std::vector<uint32_t> region_order(uint32_t n_calls, const rtk_region_call_t* calls) {
    std::vector<uint32_t> order(n_calls);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return calls[a].win_len > calls[b].win_len; });
    return order;
}

// output pools: vertices and characters (spelled path + quality, 8-byte padded each) all calls may publish
inline void region_out_caps(uint32_t n_calls, const rtk_region_call_t* calls, uint64_t& nodes_cap, uint64_t& chars_cap, uint64_t& segs_cap) {
    uint64_t w = 0;
    for (uint32_t i = 0; i < n_calls; ++i) w += calls[i].win_len;
    segs_cap = 4ull * n_calls + 1024;
    chars_cap = 2 * (2 * w + 80ull * n_calls) + 1024;   // a path is at most ~1.6 x its window (1.25 x + 10 per hop) + k
    nodes_cap = w / 2 + 64ull * n_calls + 1024;               // vertices: far fewer than bases in practice; overflow bails
}

}  // namespace rtk
