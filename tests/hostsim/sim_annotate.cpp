// sim_annotate.cpp — TEST INFRASTRUCTURE: the graph annotation kernels (annotate.cuh) on the CPU simulator.
#include "cuda_sim.h"

#include <cstring>
#include <vector>

#include "../../ratatosk_b200/csrc/annotate.cuh"
#include "../../ratatosk_b200/csrc/rtk_host_common.hpp"

namespace rtk {

static rtk_an_graph sim_an_graph(const rtk_graph_view& g, uint32_t min_cov) {
    rtk_an_graph G;
    G.unitig_off = g.unitig_off; G.pool = g.pool; G.shared = g.shared; G.adj = g.adj; G.gset_of = g.gset_of; G.gset_off = g.gset_off;
    G.gset_ids = g.gset_ids; G.loc_off = g.loc_off; G.loc_ids = g.loc_ids; G.k = g.k; G.min_cov = min_cov;
    return G;
}

void cycles_run(rtk_ctx* c, uint32_t min_cov, const uint32_t* list, uint32_t first, uint32_t n, uint32_t arena_cap,
                std::vector<uint8_t>& status, std::vector<uint32_t>& records, float* kernel_ms) {
    status.assign(n, 0);
    records.clear();
    if (kernel_ms) *kernel_ms = 0.f;
    if (!n) return;
    const unsigned grid = std::min<uint32_t>((n + RTK_AN_WARPS - 1) / RTK_AN_WARPS, 5);   // fewer warps than jobs: the job loop runs
    std::vector<uint32_t> arena((size_t)grid * RTK_AN_WARPS * 3 * arena_cap);
    uint64_t cap = 64;                                                                     // small on purpose: the re-run path runs
    for (;;) {
        records.assign(cap, 0);
        unsigned long long used = 0;
        rtk_cyc_params p;
        p.g = sim_an_graph(c->host_graph->view, min_cov);
        p.list = list; p.first = first; p.n = n; p.arena = arena.data(); p.arena_cap = arena_cap; p.status = status.data();
        p.out = records.data(); p.out_used = &used; p.out_cap = cap;
        sim_launch(grid, RTK_AN_WARPS * 32, [&] { rtk_cycles_kernel(p); });
        if (used <= cap) { records.resize(used); return; }
        cap = used;
    }
}

void snp_run(rtk_ctx* c, uint32_t min_cov, const std::vector<rtk_snp_job>& jobs, const std::vector<rtk_snp_cand>& cands,
             std::vector<uint8_t>& fin, uint32_t n_bslots, uint32_t arena_cap, std::vector<uint8_t>& status, uint64_t* n_walks,
             float* kernel_ms) {
    status.assign(jobs.size(), 0);
    if (kernel_ms) *kernel_ms = 0.f;
    if (jobs.empty()) return;
    const uint32_t n = (uint32_t)jobs.size();
    const unsigned grid = std::min<uint32_t>((n + RTK_AN_WARPS - 1) / RTK_AN_WARPS, 5);
    std::vector<uint32_t> arena((size_t)grid * RTK_AN_WARPS * 4 * arena_cap);
    std::vector<uint8_t> tried = fin, verdict(n_bslots + 1, 0);
    unsigned long long walks = 0;
    rtk_snp_params p;
    p.g = sim_an_graph(c->host_graph->view, min_cov);
    p.jobs = jobs.data(); p.n_jobs = n; p.cands = cands.data(); p.fin = fin.data(); p.tried = tried.data(); p.verdict = verdict.data();
    p.arena = arena.data(); p.arena_cap = arena_cap; p.status = status.data(); p.n_walks = &walks;
    sim_launch(grid, RTK_AN_WARPS * 32, [&] { rtk_snp_kernel(p); });
    if (n_walks) *n_walks += walks;
}

void edge_flags_run(rtk_ctx* c, uint32_t min_cov, const uint64_t* col_off, const uint32_t* col_ids, uint64_t* kmcov, uint64_t* shared,
                    float* kernel_ms) {
    if (kernel_ms) *kernel_ms = 0.f;
    const rtk_graph_view& g = c->host_graph->view;
    const uint32_t n = (uint32_t)g.n_unitigs;
    if (!n) return;
    rtk_edge_params p;
    p.adj = g.adj; p.col_off = col_off; p.col_ids = col_ids; p.kmcov = kmcov; p.shared = shared; p.n = n; p.min_cov = min_cov;
    const unsigned grid = std::min<uint32_t>((n + RTK_AN_WARPS - 1) / RTK_AN_WARPS, 5);
    sim_launch(grid, RTK_AN_WARPS * 32, [&] { rtk_edge_flags_kernel(p); });
}

}  // namespace rtk
