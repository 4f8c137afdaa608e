// sim_region.cpp — TEST INFRASTRUCTURE: the device-resident region engine (region.cuh, the same kernel source) on the CPU
// simulator, behind the same internal / C-ABI entries as region.cu.
#include "cuda_sim.h"

#include <cstring>

#include "../../ratatosk_b200/csrc/region.cuh"
#include "../../ratatosk_b200/csrc/region_host.hpp"
#include "../../ratatosk_b200/csrc/rtk_host_common.hpp"

namespace rtk {

void region_batch_run(rtk_ctx* c, const rtk_opt& opt, int pass, uint32_t n_calls, const rtk_region_call_t* calls, const char* win_pool,
                      uint64_t win_bytes, const rtk_hit* weak_pool, uint64_t n_weak, const uint32_t* pid_pool, uint64_t n_pids, RegionBatchOut& out) {
    out.results.assign(n_calls, rtk_region_result_t());
    out.nodes.clear(); out.chars.clear(); out.segs.clear(); out.kernel_ms = 0.f;
    if (!n_calls) return;
    if (!c->has_graph) throw std::invalid_argument("no graph uploaded to this context");
    const rtk_graph_view& g = c->host_graph->view;
    if (opt.k != g.k) throw std::invalid_argument("rtk_opt.k does not match the graph's k");
    region_check_calls(n_calls, calls, win_bytes, n_weak, n_pids, g.n_unitigs, g.k);
    const RegionCaps caps = region_caps();
    rtk_rg_params p;
    memset(&p, 0, sizeof(p));
    region_fill_params(p, opt, pass, caps);
    const std::vector<uint32_t> order = region_order(n_calls, calls);
    uint64_t nodes_cap, chars_cap, segs_cap;
    region_out_caps(n_calls, calls, nodes_cap, chars_cap, segs_cap);
    std::vector<rtk_region_seg_t> segs(segs_cap);
    std::vector<rtk_path_node> nodes(nodes_cap);
    std::vector<char> chars(chars_cap);
    unsigned long long counters[4] = {0, 0, 0, 0};
    const unsigned n_slots = 16;           // scratch slots (the simulator runs at most 16 blocks at a time)
    std::vector<unsigned char> scratch((size_t)n_slots * RTK_RG_WARPS * p.scratch_per_warp);
    std::vector<uint32_t> slot_flags(n_slots, 0);
    const rtk_hit no_weak = {0, 0, 0, 0};
    const uint32_t no_pid = 0;
    p.unitig_off = g.unitig_off; p.pool = g.pool; p.shared = g.shared; p.adj = g.adj; p.gset_of = g.gset_of;
    p.gset_off = g.gset_off; p.gset_ids = g.gset_ids; p.loc_off = g.loc_off; p.loc_ids = g.loc_ids; p.cyc_off = g.cyc_off; p.cyc_pool = g.cyc_pool; p.k = g.k;
    p.tasks = calls; p.order = order.data(); p.n_tasks = n_calls;
    p.win_pool = win_pool; p.weak_pool = n_weak ? weak_pool : &no_weak; p.pid_pool = n_pids ? pid_pool : &no_pid;
    p.results = out.results.data();
    p.out_nodes = nodes.data(); p.out_chars = chars.data(); p.out_top = counters; p.slot_flags = slot_flags.data(); p.n_slots = n_slots;
    p.out_nodes_cap = nodes_cap; p.out_chars_cap = chars_cap; p.out_segs = segs.data(); p.out_segs_cap = segs_cap;
    p.scratch = scratch.data();
    sim_launch((n_calls + RTK_RG_WARPS - 1) / RTK_RG_WARPS, RTK_RG_WARPS * 32, [&] { rtk_region_kernel(p); });
    const uint64_t un = std::min<uint64_t>(counters[0], nodes_cap), uc = std::min<uint64_t>(counters[1], chars_cap);
    out.nodes.assign(nodes.begin(), nodes.begin() + un);
    out.chars.assign(chars.begin(), chars.begin() + uc);
    out.segs.assign(segs.begin(), segs.begin() + std::min<uint64_t>(counters[2], segs_cap));
}

}  // namespace rtk

using namespace rtk;

extern "C" int rtk_region_paths_batch(rtk_ctx* c, const rtk_opt* opt, int pass, uint32_t n_calls, const rtk_region_call_t* calls, const char* win_pool,
                                      uint64_t win_bytes, const rtk_hit* weak_pool, uint64_t n_weak, const uint32_t* pid_pool, uint64_t n_pids,
                                      rtk_region_out* out, uint64_t* stats) {
    return guarded([&] {
        if (!c || !opt || !out || (n_calls && (!calls || !win_pool))) throw std::invalid_argument("null argument");
        if (pass != 1 && pass != 2) throw std::invalid_argument("pass must be 1 or 2");
        memset(out, 0, sizeof(*out));
        RegionBatchOut r;
        region_batch_run(c, *opt, pass, n_calls, calls, win_pool, win_bytes, weak_pool, n_weak, pid_pool, n_pids, r);
        out->results = (rtk_region_result_t*)malloc(sizeof(rtk_region_result_t) * ((size_t)n_calls + 1));
        out->nodes = (rtk_path_node*)malloc(sizeof(rtk_path_node) * (r.nodes.size() + 1));
        out->chars = (char*)malloc(r.chars.size() + 1);
        out->segs = (rtk_region_seg_t*)malloc(sizeof(rtk_region_seg_t) * (r.segs.size() + 1));
        if (!out->results || !out->nodes || !out->chars || !out->segs) throw std::bad_alloc();
        if (n_calls) memcpy(out->results, r.results.data(), sizeof(rtk_region_result_t) * (size_t)n_calls);
        if (!r.nodes.empty()) memcpy(out->nodes, r.nodes.data(), sizeof(rtk_path_node) * r.nodes.size());
        if (!r.chars.empty()) memcpy(out->chars, r.chars.data(), r.chars.size());
        if (!r.segs.empty()) memcpy(out->segs, r.segs.data(), sizeof(rtk_region_seg_t) * r.segs.size());
        out->n_nodes = r.nodes.size(); out->n_chars = r.chars.size(); out->n_segs = r.segs.size();
        if (stats) stats[0] += n_calls;
    });
}

extern "C" void rtk_region_out_free(rtk_region_out* o) {
    if (!o) return;
    free(o->results); free(o->nodes); free(o->chars); free(o->segs);
    memset(o, 0, sizeof(*o));
}
