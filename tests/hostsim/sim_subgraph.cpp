// sim_subgraph.cpp — TEST INFRASTRUCTURE: rtk_explore_subgraph_batch on the CPU simulator (same kernel sources:
// subgraph.cuh for enumeration, myers.cuh for scoring; same host planner / selection).
#include "cuda_sim.h"

#include <cstring>

#include "../../ratatosk_b200/csrc/myers.cuh"
#include "../../ratatosk_b200/csrc/myers_host.hpp"
#include "../../ratatosk_b200/csrc/rtk_host_common.hpp"
#include "../../ratatosk_b200/csrc/subgraph.cuh"
#include "../../ratatosk_b200/csrc/subgraph_host.hpp"

using namespace rtk;

template <int G> static void sim_myers_class(rtk_myers_params p, const std::vector<uint32_t>& order) {
    if (order.empty()) return;
    p.order = order.data();
    p.n = (uint32_t)order.size();
    const uint64_t threads = (uint64_t)p.n * G;
    sim_launch((unsigned)((threads + RTK_MYERS_THREADS - 1) / RTK_MYERS_THREADS), RTK_MYERS_THREADS, [&] { rtk_myers_kernel<G>(p); });
}

extern "C" int rtk_explore_subgraph_batch(rtk_ctx* c, uint32_t n_calls, const rtk_subgraph_call_t* calls, const char* ref_pool,
                                          uint64_t ref_bytes, const uint32_t* pid_pool, uint64_t n_pids, double wrlf,
                                          rtk_subgraph_out* out, uint64_t*) {
    return guarded([&] {
        if (!c->has_graph) throw std::invalid_argument("no graph uploaded to this context");
        memset(out, 0, sizeof(*out));
        const rtk_graph_view& g = c->host_graph->view;
        for (uint32_t i = 0; i < n_calls; ++i) if (calls[i].level + 1 > RTK_DFS_MAX_NODES) throw std::invalid_argument("level too large");
        std::vector<uint32_t> ncand(n_calls + 1, 0);
        std::vector<uint64_t> nchars(n_calls + 1, 0);
        std::vector<uint32_t> no_pids(1, 0);
        rtk_dfs_params p;
        p.unitig_off = g.unitig_off; p.pool = g.pool; p.shared = g.shared; p.adj = g.adj; p.gset_of = g.gset_of;
        p.gset_off = g.gset_off; p.gset_ids = g.gset_ids; p.loc_off = g.loc_off; p.loc_ids = g.loc_ids; p.k = g.k;
        p.calls = calls; p.pid_pool = n_pids ? pid_pool : no_pids.data(); p.n_calls = n_calls;
        p.n_cand = ncand.data(); p.n_chars = nchars.data(); p.cand_off = nullptr; p.char_off = nullptr; p.cands = nullptr; p.chars = nullptr;
        uint32_t overflow = 0;
        p.overflow = &overflow;
        const unsigned grid = (n_calls + RTK_DFS_WARPS - 1) / RTK_DFS_WARPS;
        if (n_calls) sim_launch(grid, RTK_DFS_WARPS * 32, [&] { rtk_dfs_kernel<false>(p); });
        if (overflow) throw std::runtime_error("exploreSubGraph: a burst exceeded the DFS capacity (RTK_DFS_MAX_NODES / RTK_DFS_STACK)");
        std::vector<uint64_t> cand_off(n_calls + 1, 0), char_off(n_calls + 1, 0);
        for (uint32_t i = 0; i < n_calls; ++i) { cand_off[i + 1] = cand_off[i] + ncand[i]; char_off[i + 1] = char_off[i] + nchars[i]; }
        const uint64_t n_cands = cand_off[n_calls], n_chars = char_off[n_calls];
        std::vector<rtk_cand> cands(n_cands + 1);
        std::vector<char> pool(ref_bytes + n_chars + 16, 0);
        memcpy(pool.data(), ref_pool, ref_bytes);
        p.cand_off = cand_off.data(); p.char_off = char_off.data(); p.cands = cands.data(); p.chars = pool.data() + ref_bytes;
        if (n_calls) sim_launch(grid, RTK_DFS_WARPS * 32, [&] { rtk_dfs_kernel<true>(p); });
        cands.resize(n_cands);
        std::vector<CandAlign> plan(n_cands);
        std::vector<uint64_t> qb(n_cands + 1), tb(n_cands + 1);
        std::vector<uint32_t> ql(n_cands + 1), tl(n_cands + 1);
        std::vector<uint8_t> md(n_cands + 1);
        for (uint64_t i = 0; i < n_cands; ++i) {
            plan[i] = plan_candidate(cands[i], calls[cands[i].call], ref_bytes, wrlf);
            qb[i] = plan[i].q_beg; tb[i] = plan[i].t_beg; ql[i] = plan[i].q_len; tl[i] = plan[i].t_len; md[i] = plan[i].mode;
        }
        std::vector<int32_t> ed(n_cands + 1, -1), ne(n_cands + 1, 0), kmax(n_cands + 1, -1);
        if (n_cands) {
            const MyersPlan pl = plan_myers((uint32_t)n_cands, ql.data(), tl.data());
            std::vector<int32_t> ends(pl.ends_off[n_cands] + 1, 0);
            std::vector<int8_t> hb(pl.hb_off[n_cands] + 1, 0);
            rtk_myers_params mp;
            mp.q_pool = pool.data(); mp.q_beg = qb.data(); mp.q_len = ql.data(); mp.t_pool = pool.data(); mp.t_beg = tb.data(); mp.t_len = tl.data();
            mp.mode = md.data(); mp.kmax = kmax.data(); mp.order = nullptr; mp.n = 0; mp.dist = ed.data(); mp.n_ends = ne.data();
            mp.ends = ends.data(); mp.ends_off = pl.ends_off.data(); mp.hbound = hb.data(); mp.hb_off = pl.hb_off.data();
            sim_myers_class<1>(mp, pl.order[0]); sim_myers_class<2>(mp, pl.order[1]); sim_myers_class<4>(mp, pl.order[2]);
            sim_myers_class<8>(mp, pl.order[3]); sim_myers_class<16>(mp, pl.order[4]); sim_myers_class<32>(mp, pl.order[5]);
            for (uint32_t a : pl.trivial) { int32_t d, e; myers_trivial(ql[a], tl[a], md[a], d, e); ed[a] = d; }
        }
        fill_subgraph_out(n_calls, cands, cand_off, ed, plan, g.unitig_off, g.k, out);
    });
}
