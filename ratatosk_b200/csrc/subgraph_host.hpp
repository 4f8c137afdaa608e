// subgraph_host.hpp — host half of exploreSubGraph shared by the CUDA driver and tests/hostsim:
// alignment planning for the candidates K2 emitted, and the reference's selection of equal-best paths
// in discovery order (src/GraphTraversal.cpp:511-523, :536-548).
#pragma once
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "../../include/rtk.h"
#include "subgraph.cuh"

namespace rtk {

struct CandAlign {          // one K4 job per candidate
    uint64_t q_beg, t_beg;  // into the combined pool [refs | spelled paths]
    uint32_t q_len, t_len;
    uint8_t mode;           // 0 NW, 2 HW
    uint32_t norm;          // score = 1 - ed / norm (getScorePath, src/GraphTraversal.cpp:867-909)
};

// refs occupy [0, ref_bytes) of the combined pool, spelled paths follow
inline CandAlign plan_candidate(const rtk_cand& c, const rtk_subgraph_call& call, uint64_t ref_bytes, double wrlf) {
    CandAlign a;
    const uint64_t path_beg = ref_bytes + c.str_off;
    if (c.terminal) {
        a.q_beg = path_beg; a.q_len = c.path_len; a.t_beg = call.ref_off; a.t_len = call.ref_len; a.mode = 0; a.norm = c.path_len;
    } else if (c.path_len >= call.ref_len) {
        a.q_beg = call.ref_off; a.q_len = call.ref_len; a.t_beg = path_beg; a.t_len = c.path_len; a.mode = 2; a.norm = call.ref_len;
    } else {
        const size_t lref = std::min<size_t>(call.ref_len, static_cast<size_t>(c.path_len * (1.0 + wrlf)));
        a.q_beg = path_beg; a.q_len = c.path_len; a.t_beg = call.ref_off; a.t_len = (uint32_t)lref; a.mode = 2; a.norm = c.path_len;
    }
    return a;
}

inline double cand_score(int32_t ed, uint32_t norm) {
    const double s = 1.0 - (static_cast<double>(ed) / norm);
    return std::min(std::max(s, 0.0), 1.0);
}

struct SubgraphSelection {
    double t1 = 0.0, t2 = 0.0, nt1 = 0.0, nt2 = 0.0;
    std::vector<uint32_t> terminal, nonterminal;  // candidate indices kept, discovery order
};

// cands[first, last) are one call's candidates in discovery order
inline SubgraphSelection select_candidates(const rtk_cand* cands, const double* score, uint32_t first, uint32_t last) {
    SubgraphSelection s;
    for (uint32_t i = first; i < last; ++i) {
        const double sc = score[i];
        if (cands[i].terminal) {
            if (sc >= s.t1) { if (sc > s.t1) s.terminal.clear(); s.terminal.push_back(i); s.t2 = s.t1; s.t1 = sc; }
            else if (sc > s.t2) s.t2 = sc;
        } else {
            if (sc >= s.nt1) { if (sc > s.nt1) s.nonterminal.clear(); s.nonterminal.push_back(i); s.nt2 = s.nt1; s.nt1 = sc; }
            else if (sc > s.nt2) s.nt2 = sc;
        }
    }
    return s;
}

// flatten the per-call selections into the C ABI output
void fill_subgraph_out(uint32_t n_calls, const std::vector<rtk_cand>& cands, const std::vector<uint64_t>& cand_off,
                       const std::vector<int32_t>& ed, const std::vector<CandAlign>& plan, const uint64_t* unitig_off, uint32_t k,
                       rtk_subgraph_out* out);

}  // namespace rtk
