// subgraph_host.cpp — see subgraph_host.hpp.
#include "subgraph_host.hpp"

#include <cstdlib>
#include <cstring>
#include <new>

namespace rtk {

void fill_subgraph_out(uint32_t n_calls, const std::vector<rtk_cand>& cands, const std::vector<uint64_t>& cand_off,
                       const std::vector<int32_t>& ed, const std::vector<CandAlign>& plan, const uint64_t* unitig_off, uint32_t k,
                       rtk_subgraph_out* out) {
    std::vector<double> score(cands.size());
    for (size_t i = 0; i < cands.size(); ++i) score[i] = cand_score(ed[i], plan[i].norm);
    std::vector<SubgraphSelection> sel(n_calls);
    uint64_t n_paths = 0, n_nodes = 0;
    for (uint32_t c = 0; c < n_calls; ++c) {
        sel[c] = select_candidates(cands.data(), score.data(), (uint32_t)cand_off[c], (uint32_t)cand_off[c + 1]);
        for (uint32_t i : sel[c].terminal) { ++n_paths; n_nodes += cands[i].n_nodes; }
        for (uint32_t i : sel[c].nonterminal) { ++n_paths; n_nodes += cands[i].n_nodes; }
    }
    out->scores = (double*)malloc(sizeof(double) * 4 * (n_calls + 1));
    out->path_off = (uint64_t*)malloc(sizeof(uint64_t) * (n_calls + 1));
    out->n_terminal = (uint32_t*)malloc(sizeof(uint32_t) * (n_calls + 1));
    out->node_off = (uint64_t*)malloc(sizeof(uint64_t) * (n_paths + 1));
    out->nodes = (rtk_path_node*)malloc(sizeof(rtk_path_node) * (n_nodes + 1));
    out->path_len = (uint32_t*)malloc(sizeof(uint32_t) * (n_paths + 1));
    out->path_ed = (int32_t*)malloc(sizeof(int32_t) * (n_paths + 1));
    if (!out->scores || !out->path_off || !out->n_terminal || !out->node_off || !out->nodes || !out->path_len || !out->path_ed) throw std::bad_alloc();
    uint64_t pi = 0, ni = 0;
    for (uint32_t c = 0; c < n_calls; ++c) {
        out->scores[4 * c] = sel[c].t1; out->scores[4 * c + 1] = sel[c].t2;
        out->scores[4 * c + 2] = sel[c].nt1; out->scores[4 * c + 3] = sel[c].nt2;
        out->path_off[c] = pi;
        out->n_terminal[c] = (uint32_t)sel[c].terminal.size();
        for (int pass = 0; pass < 2; ++pass) {
            for (uint32_t i : (pass == 0 ? sel[c].terminal : sel[c].nonterminal)) {
                const rtk_cand& cd = cands[i];
                out->node_off[pi] = ni;
                for (uint32_t j = 0; j < cd.n_nodes; ++j) {
                    const uint32_t u = cd.nodes[j] & 0x7fffffffu;
                    rtk_path_node nd;
                    nd.unitig = u; nd.strand = cd.nodes[j] >> 31;
                    if (j + 1 == cd.n_nodes) { nd.dist = cd.last_dist; nd.len = cd.last_len; }
                    else { nd.dist = 0; nd.len = (uint32_t)(unitig_off[u + 1] - unitig_off[u]) - k + 1; }
                    out->nodes[ni++] = nd;
                }
                out->path_len[pi] = cd.path_len;
                out->path_ed[pi] = ed[i];
                ++pi;
            }
        }
    }
    out->path_off[n_calls] = pi;
    out->node_off[pi] = ni;
}

}  // namespace rtk

extern "C" void rtk_subgraph_out_free(rtk_subgraph_out* o) {
    if (!o) return;
    free(o->scores); free(o->path_off); free(o->n_terminal); free(o->node_off); free(o->nodes); free(o->path_len); free(o->path_ed);
    memset(o, 0, sizeof(*o));
}
