// traverse_api.cpp — C ABI entry for the anchor-to-anchor path search (host orchestration in traverse.cpp).
#include <cstdlib>
#include <cstring>

#include "rtk_host_common.hpp"
#include "traverse.hpp"

using namespace rtk;

extern "C" int rtk_explore_paths(rtk_ctx* ctx, const rtk_opt* opt, const rtk_hit* um_s, const rtk_hit* um_e, const char* ref,
                                 uint32_t ref_len, const uint32_t* pids, uint32_t n_pids, rtk_path_node** nodes, uint32_t* n_nodes,
                                 char** qual, uint32_t* path_len) {
    return guarded([&] {
        if (!ctx || !opt || !um_s || !ref || !nodes || !n_nodes || !qual || !path_len) throw std::invalid_argument("null argument");
        if (!ctx->has_graph || !ctx->host_graph) throw std::invalid_argument("no graph uploaded to this context");
        const rtk_graph_view& g = ctx->host_graph->view;
        if (opt->k != g.k) throw std::invalid_argument("rtk_opt.k does not match the graph's k");
        if (ref_len <= g.k) throw std::invalid_argument("read window not longer than k");
        TraverseOpt to;
        to.k = g.k; to.min_cov_vertices = opt->min_cov_vertices; to.out_qual = opt->out_qual; to.max_qual = opt->max_qual;
        to.weak_region_len_factor = opt->weak_region_len_factor; to.large_k_factor = opt->large_k_factor; to.min_score = opt->min_score;
        auto to_node = [&](const rtk_hit* h) {
            PNode n;
            if (!h) return n;
            if (h->unitig >= g.n_unitigs) throw std::invalid_argument("bad unitig id");
            n.unitig = h->unitig; n.strand = h->strand; n.dist = h->dist; n.len = 1;
            return n;
        };
        const std::vector<uint32_t> all_pids(pids, pids + n_pids);
        const std::string r(ref, ref_len);
        const std::vector<GPath> v = um_e ? explore_paths_bfs2(ctx, g, to, r, all_pids, to_node(um_s), to_node(um_e))
                                          : explore_paths_bfs(ctx, g, to, r, all_pids, to_node(um_s));
        *nodes = nullptr; *n_nodes = 0; *qual = nullptr; *path_len = 0;
        if (v.empty()) return;
        const GPath& p = v[0];
        *n_nodes = (uint32_t)p.v.size();
        *nodes = (rtk_path_node*)malloc(sizeof(rtk_path_node) * (p.v.size() + 1));
        *qual = (char*)malloc(p.qual.size() + 1);
        if (!*nodes || !*qual) throw std::bad_alloc();
        for (size_t i = 0; i < p.v.size(); ++i) { (*nodes)[i].unitig = p.v[i].unitig; (*nodes)[i].strand = p.v[i].strand; (*nodes)[i].dist = p.v[i].dist; (*nodes)[i].len = p.v[i].len; }
        memcpy(*qual, p.qual.c_str(), p.qual.size() + 1);
        *path_len = (uint32_t)p.length();
    });
}
