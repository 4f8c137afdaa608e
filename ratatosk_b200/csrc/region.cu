// region.cu — launcher and C ABI entry of the device-resident region engine (region.cuh): one packed upload, ONE kernel
// launch for all calls of the batch (one warp per region, longest first), results + the used part of
// the output pools back.  Compiled with -fmad=false: the engine's few double-precision score / quality expressions must
// round like the host's (no FMA contraction), see rg_get_qual / rg_min_max.
#include <cstring>
#include <vector>

#include "region.cuh"
#include "region_host.hpp"
#include "rtk_host_common.hpp"

namespace rtk {

// scratch slots = CTAs (= warps = regions) that can be resident at once (RTK_RG_CTAS_PER_SM overrides the occupancy query);
// a slot is ~2 MB
static uint32_t region_slots(const rtk_ctx* c) {
    static int per_sm = [] {
        const char* e = getenv("RTK_RG_CTAS_PER_SM");
        if (e) return std::max(1, atoi(e));
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, rtk_region_kernel, RTK_RG_WARPS * 32, 0) != cudaSuccess || n < 1) n = 12;
        return n;
    }();
    return (uint32_t)c->sm_count * (uint32_t)per_sm;
}

void region_batch_run(rtk_ctx* c, const rtk_opt& opt, int pass, uint32_t n_calls, const rtk_region_call_t* calls, const char* win_pool,
                      uint64_t win_bytes, const rtk_hit* weak_pool, uint64_t n_weak, const uint32_t* pid_pool, uint64_t n_pids, RegionBatchOut& out) {
    out.results.assign(n_calls, rtk_region_result_t());
    out.nodes.clear(); out.chars.clear(); out.segs.clear(); out.kernel_ms = 0.f;
    if (!n_calls) return;
    if (!c->has_graph) throw std::invalid_argument("no graph uploaded to this context");
    if (opt.k != c->hdr.k) throw std::invalid_argument("rtk_opt.k does not match the graph's k");
    DeviceBind bind(c);   // called from the broker's service threads as well as from the C entry
    region_check_calls(n_calls, calls, win_bytes, n_weak, n_pids, c->hdr.n_unitigs, c->hdr.k);
    cudaStream_t st = c->stream;
    const RegionCaps caps = region_caps();
    rtk_rg_params p;
    memset(&p, 0, sizeof(p));
    region_fill_params(p, opt, pass, caps);
    const std::vector<uint32_t> order = region_order(n_calls, calls);
    // ---- one packed upload: tasks | order | weak anchors | colour ids | windows
    const uint64_t b_tasks = (uint64_t)n_calls * sizeof(rtk_region_call_t), b_order = (uint64_t)n_calls * 4, b_weak = n_weak * sizeof(rtk_hit),
                   b_pids = n_pids * 4;
    const uint64_t o_tasks = 0, o_order = rtk_rg_align16(o_tasks + b_tasks), o_weak = rtk_rg_align16(o_order + b_order),
                   o_pids = rtk_rg_align16(o_weak + b_weak), o_win = rtk_rg_align16(o_pids + b_pids), total = rtk_rg_align16(o_win + win_bytes + 16);
    PinBuf& H = c->h_rg[0];
    H.reserve(total);
    char* h = H.as<char>();
    memcpy(h + o_tasks, calls, b_tasks);
    memcpy(h + o_order, order.data(), b_order);
    if (b_weak) memcpy(h + o_weak, weak_pool, b_weak);
    if (b_pids) memcpy(h + o_pids, pid_pool, b_pids);
    if (win_bytes) memcpy(h + o_win, win_pool, win_bytes);
    c->d_rg[0].reserve(total);
    char* d = c->d_rg[0].as<char>();
    RTK_CUDA(counted_memcpy_async(d, h, total, cudaMemcpyHostToDevice, st));
    // ---- outputs and scratch
    uint64_t nodes_cap, chars_cap, segs_cap;
    region_out_caps(n_calls, calls, nodes_cap, chars_cap, segs_cap);
    c->d_rg[6].reserve(segs_cap * sizeof(rtk_region_seg_t));
    c->d_rg[1].reserve((uint64_t)n_calls * sizeof(rtk_region_result_t));
    c->d_rg[2].reserve(nodes_cap * sizeof(rtk_path_node));
    c->d_rg[3].reserve(chars_cap);
    const uint32_t n_slots = region_slots(c);
    c->d_rg[4].reserve(64 + (uint64_t)n_slots * 4);
    c->d_rg[5].reserve((uint64_t)n_slots * RTK_RG_WARPS * p.scratch_per_warp);
    RTK_CUDA(cudaMemsetAsync(c->d_rg[4].p, 0, 64 + (uint64_t)n_slots * 4, st));
    const rtk_graph_view& g = c->dview;
    p.unitig_off = g.unitig_off; p.pool = g.pool; p.shared = g.shared; p.adj = g.adj; p.gset_of = g.gset_of;
    p.gset_off = g.gset_off; p.gset_ids = g.gset_ids; p.loc_off = g.loc_off; p.loc_ids = g.loc_ids; p.cyc_off = g.cyc_off; p.cyc_pool = g.cyc_pool; p.k = c->hdr.k;
    p.tasks = (const rtk_rg_task*)(d + o_tasks); p.order = (const uint32_t*)(d + o_order); p.n_tasks = n_calls;
    p.win_pool = d + o_win; p.weak_pool = (const rtk_hit*)(d + o_weak); p.pid_pool = (const uint32_t*)(d + o_pids);
    p.results = c->d_rg[1].as<rtk_rg_result>();
    p.out_nodes = c->d_rg[2].as<rtk_rg_node>(); p.out_chars = c->d_rg[3].as<char>();
    p.out_top = c->d_rg[4].as<unsigned long long>(); p.slot_flags = (uint32_t*)(c->d_rg[4].as<char>() + 64); p.n_slots = n_slots;
    p.out_nodes_cap = nodes_cap; p.out_chars_cap = chars_cap; p.out_segs = c->d_rg[6].as<rtk_region_seg_t>(); p.out_segs_cap = segs_cap;
    p.scratch = c->d_rg[5].as<unsigned char>();
    RTK_CUDA(cudaEventRecord(c->ev0, st));
    ++g_launches;
    rtk_region_kernel<<<(n_calls + RTK_RG_WARPS - 1) / RTK_RG_WARPS, RTK_RG_WARPS * 32, 0, st>>>(p);
    RTK_CUDA(cudaGetLastError());
    RTK_CUDA(cudaEventRecord(c->ev1, st));
    // ---- results + counters, then the used part of the pools
    const uint64_t b_res = (uint64_t)n_calls * sizeof(rtk_region_result_t);
    PinBuf& HR = c->h_rg[1];
    HR.reserve(b_res + 64);
    RTK_CUDA(counted_memcpy_async(HR.p, c->d_rg[1].p, b_res, cudaMemcpyDeviceToHost, st));
    RTK_CUDA(counted_memcpy_async(HR.as<char>() + b_res, c->d_rg[4].p, 24, cudaMemcpyDeviceToHost, st));
    stream_wait(st);
    memcpy(out.results.data(), HR.p, b_res);
    unsigned long long tops[3];
    memcpy(tops, HR.as<char>() + b_res, 24);
    const uint64_t used_nodes = std::min<uint64_t>(tops[0], nodes_cap), used_chars = std::min<uint64_t>(tops[1], chars_cap),
                   used_segs = std::min<uint64_t>(tops[2], segs_cap);
    const uint64_t b_nodes = used_nodes * sizeof(rtk_path_node), b_segs = used_segs * sizeof(rtk_region_seg_t);
    PinBuf& HO = c->h_rg[2];
    HO.reserve(b_nodes + b_segs + used_chars + 64);
    if (used_nodes) RTK_CUDA(counted_memcpy_async(HO.p, c->d_rg[2].p, b_nodes, cudaMemcpyDeviceToHost, st));
    if (used_segs) RTK_CUDA(counted_memcpy_async(HO.as<char>() + b_nodes, c->d_rg[6].p, b_segs, cudaMemcpyDeviceToHost, st));
    if (used_chars) RTK_CUDA(counted_memcpy_async(HO.as<char>() + b_nodes + b_segs, c->d_rg[3].p, used_chars, cudaMemcpyDeviceToHost, st));
    stream_wait(st);
    out.nodes.resize(used_nodes);
    out.segs.resize(used_segs);
    out.chars.resize(used_chars);
    if (used_nodes) memcpy(out.nodes.data(), HO.p, b_nodes);
    if (used_segs) memcpy(out.segs.data(), HO.as<char>() + b_nodes, b_segs);
    if (used_chars) memcpy(out.chars.data(), HO.as<char>() + b_nodes + b_segs, used_chars);
    RTK_CUDA(cudaEventElapsedTime(&out.kernel_ms, c->ev0, c->ev1));
}

}  // namespace rtk

using namespace rtk;

extern "C" int rtk_region_paths_batch(rtk_ctx* c, const rtk_opt* opt, int pass, uint32_t n_calls, const rtk_region_call_t* calls, const char* win_pool,
                                      uint64_t win_bytes, const rtk_hit* weak_pool, uint64_t n_weak, const uint32_t* pid_pool, uint64_t n_pids,
                                      rtk_region_out* out, uint64_t* stats) {
    return guarded([&] {
        if (!c || !opt || !out || (n_calls && (!calls || !win_pool))) throw std::invalid_argument("null argument");
        if (pass != 1 && pass != 2) throw std::invalid_argument("pass must be 1 or 2");
        if ((n_weak && !weak_pool) || (n_pids && !pid_pool)) throw std::invalid_argument("null pool");
        DeviceBind bind(c);
        memset(out, 0, sizeof(*out));
        RegionBatchOut r;
        region_batch_run(c, *opt, pass, n_calls, calls, win_pool, win_bytes, weak_pool, n_weak, pid_pool, n_pids, r);
        out->results = (rtk_region_result_t*)malloc(sizeof(rtk_region_result_t) * ((size_t)n_calls + 1));
        out->nodes = (rtk_path_node*)malloc(sizeof(rtk_path_node) * (r.nodes.size() + 1));
        out->chars = (char*)malloc(r.chars.size() + 1);
        out->segs = (rtk_region_seg_t*)malloc(sizeof(rtk_region_seg_t) * (r.segs.size() + 1));
        if (!out->results || !out->nodes || !out->chars || !out->segs) { free(out->results); free(out->nodes); free(out->chars); free(out->segs); memset(out, 0, sizeof(*out)); throw std::bad_alloc(); }
        if (n_calls) memcpy(out->results, r.results.data(), sizeof(rtk_region_result_t) * (size_t)n_calls);
        if (!r.nodes.empty()) memcpy(out->nodes, r.nodes.data(), sizeof(rtk_path_node) * r.nodes.size());
        if (!r.chars.empty()) memcpy(out->chars, r.chars.data(), r.chars.size());
        if (!r.segs.empty()) memcpy(out->segs, r.segs.data(), sizeof(rtk_region_seg_t) * r.segs.size());
        out->n_nodes = r.nodes.size(); out->n_chars = r.chars.size(); out->n_segs = r.segs.size();
        if (stats) { stats[0] += n_calls; stats[2] += (uint64_t)(r.kernel_ms * 1e6); }
    });
}

extern "C" void rtk_region_out_free(rtk_region_out* o) {
    if (!o) return;
    free(o->results); free(o->nodes); free(o->chars); free(o->segs);
    memset(o, 0, sizeof(*o));
}
