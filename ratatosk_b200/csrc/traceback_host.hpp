// traceback_host.hpp — host side of K5 shared by the CUDA driver and tests/hostsim: batch planning and edlib's
// divide-and-conquer for large problems (obtainAlignment / obtainAlignmentHirschberg, src/edlib.cpp:1164-1399).
//
// edlib stores the whole DP state and walks it back when that state is below 1 MiB (:1191-1193); above, it
// splits the target in two halves, finds the FIRST query row r (scanning rows upwards from 0, then the two
// boundary rows) where  D_left[r] + D_right_reversed[r+1] == best, and recurses on the upper-left and lower-
// right sub-problems (:1330-1356).  The split row and the size switch decide which of the equally optimal
// paths comes out, so both are reproduced: the device supplies last DP columns (forward and reversed
// problems) and direct tracebacks, this file drives the recursion level by level so that every level is
// one batch.
#pragma once
#include <cstdlib>
#include <functional>
#include <stdint.h>

#include <algorithm>
#include <stdexcept>
#include <vector>

namespace rtk {

// chunked parallel loop supplied by the including library (rtk_graph_api.cpp: parallel_for)
void parallel_for(size_t n, const std::function<void(size_t, size_t)>& body);
inline void tb_parallel_for(size_t n, const std::function<void(size_t, size_t)>& body) { parallel_for(n, body); }

// edlib's switch between the direct traceback and its divide-and-conquer: 1 MiB of alignment state (src/edlib.cpp:1191-1193).
// RTK_TB_LIMIT lowers it for tests that compare the two implementations of the divide-and-conquer (device engine vs K5 host
// driver) on small inputs; never set in production (the reference's switch is part of its output).
inline uint64_t tb_limit() {
    static const uint64_t v = [] { const char* e = getenv("RTK_TB_LIMIT"); return e ? (uint64_t)strtoull(e, nullptr, 10) : 1024ull * 1024ull; }();
    return v;
}
inline bool tb_needs_hirschberg(uint64_t qlen, uint64_t tlen) {
    const uint64_t nb = (qlen + 63) / 64;
    return (2ull * 8 + 4) * nb * tlen + 2ull * 4 * tlen >= tb_limit();
}

struct TbItem {
    uint64_t q_beg, t_beg;  // into the forward pools, or into the reversed pools when rev
    uint32_t q_len, t_len;
    uint8_t rev;
};

struct TbBackend {
    virtual ~TbBackend() {}
    // NW fill + walk-back of every item (all below the size switch): ops[i] (0 match, 1 query base unaligned,
    // 2 target base unaligned, 3 mismatch) and the NW distance
    virtual void direct(const std::vector<TbItem>& items, std::vector<std::vector<uint8_t>>& ops, std::vector<int32_t>& dist) = 0;
    // rows[i][r] = D[r+1][t_len] (NW) for r in [0, q_len)
    virtual void last_column(const std::vector<TbItem>& items, std::vector<std::vector<int32_t>>& rows) = 0;
};

// lane-group class for a query of nb blocks: smallest power of two covering it, at most 32 lanes (more blocks: rounds)
inline int tb_class_of(uint64_t nb) {
    int c = 0;
    while (c < 5 && (1u << c) < nb) ++c;
    return c;
}

// rows of one column from its stored block words: score(r) = anchor - popc(P & m) + popc(M & m), m = rows r+1..anchor row
inline void tb_rows_from_column(uint32_t qlen, const uint64_t* P, const uint64_t* M, const int32_t* anchor, std::vector<int32_t>& rows) {
    rows.resize(qlen);
    const uint32_t nb = (qlen + 63) / 64;
    for (uint32_t b = 0; b < nb; ++b) {
        const int arow = (b == nb - 1) ? (int)((qlen - 1) & 63) : 63;
        int s = anchor[b];
        rows[b * 64 + arow] = s;
        for (int r = arow - 1; r >= 0; --r) {  // D(r) = D(r+1) - delta(r+1)
            s -= (int)((P[b] >> (r + 1)) & 1) - (int)((M[b] >> (r + 1)) & 1);
            rows[b * 64 + r] = s;
        }
    }
}

// Which target columns of a problem matter to the caller (phasing() only reads the path where the corrected read is marked
// for reversal, src/Graph.cpp:1001-1052): prefix[off[a] + p] = number of marked target positions < p, p in [0, t_len + 1].
// A sub-problem of the divide-and-conquer whose closed column range [tu, tv] holds no marked position is not solved: its
// path is replaced by "all its query bases unaligned, then all its target bases unaligned", which consumes the same bases.
struct TbNeed {
    const uint32_t* prefix;
    const uint64_t* off;
    bool any(uint32_t a, uint32_t tu, uint32_t tv) const { return prefix[off[a] + tv + 1] != prefix[off[a] + tu]; }
};
struct TbRun { uint8_t kind; uint32_t len; };   // kind 0: aligned pair (match or mismatch), 1: query base unaligned, 2: target base unaligned

struct TbNode {
    uint32_t a;                 // problem
    uint32_t qx, qy, tu, tv;    // sub-ranges [qx,qy) x [tu,tv) of the problem's strings
    int child[2];
    std::vector<uint8_t> ops;
    int32_t dist;
    uint8_t pruned;
};

// The recursion tree of n NW problems: problem a = q_pool[q_beg[a], +q_len[a]) vs t_pool[t_beg[a], +t_len[a]); the reversed
// pools hold each string reversed at the same offsets.  nodes[a] is the root of problem a; leaves carry ops.
inline void solve_nw_tree(TbBackend& be, uint32_t n, const uint64_t* q_beg, const uint32_t* q_len, const uint64_t* t_beg,
                          const uint32_t* t_len, const TbNeed* need, std::vector<TbNode>& nodes) {
    typedef TbNode Node;
    nodes.clear();
    std::vector<int> work;
    for (uint32_t a = 0; a < n; ++a) { nodes.push_back({a, 0, q_len[a], 0, t_len[a], {-1, -1}, {}, -1, 0}); work.push_back((int)a); }
    auto fw_item = [&](const Node& nd, uint32_t tu, uint32_t tv) {
        return TbItem{q_beg[nd.a] + nd.qx, t_beg[nd.a] + tu, nd.qy - nd.qx, tv - tu, 0};
    };
    auto rev_item = [&](const Node& nd, uint32_t tu, uint32_t tv) {  // reversed strings of the same sub-ranges
        return TbItem{q_beg[nd.a] + (q_len[nd.a] - nd.qy), t_beg[nd.a] + (t_len[nd.a] - tv), nd.qy - nd.qx, tv - tu, 1};
    };
    while (!work.empty()) {
        std::vector<int> direct_ids, big_ids;
        for (int id : work) {
            Node& nd = nodes[id];
            const uint32_t ql = nd.qy - nd.qx, tl = nd.tv - nd.tu;
            if (need && !need->any(nd.a, nd.tu, nd.tv)) { nd.pruned = 1; nd.dist = -1; }
            else if (ql == 0 || tl == 0) { nd.ops.assign((size_t)ql + tl, ql == 0 ? 2 : 1); nd.dist = (int32_t)(ql + tl); }  // :1171-1178
            else if (!tb_needs_hirschberg(ql, tl)) direct_ids.push_back(id);
            else big_ids.push_back(id);
        }
        if (!direct_ids.empty()) {
            std::vector<TbItem> items;
            for (int id : direct_ids) items.push_back(fw_item(nodes[id], nodes[id].tu, nodes[id].tv));
            std::vector<std::vector<uint8_t>> o;
            std::vector<int32_t> d;
            be.direct(items, o, d);
            for (size_t i = 0; i < direct_ids.size(); ++i) { nodes[direct_ids[i]].ops = std::move(o[i]); nodes[direct_ids[i]].dist = d[i]; }
        }
        std::vector<int> next;
        if (!big_ids.empty()) {
            std::vector<TbItem> items;
            for (int id : big_ids) {
                const Node& nd = nodes[id];
                const uint32_t left = (nd.tv - nd.tu) / 2;
                items.push_back(fw_item(nd, nd.tu, nd.tu + left));            // left half, forward
                items.push_back(rev_item(nd, nd.tu + left, nd.tv));           // right half, reversed
            }
            std::vector<std::vector<int32_t>> rows;
            be.last_column(items, rows);
            // The split row of every big problem (independent of one another): the first query row at which the two halves add
            // up to the problem's distance (src/edlib.cpp:1330-1337).  The distance is the minimum of those sums over all the
            // ways to cut, so it needs no sweep of its own.
            std::vector<int> splits(big_ids.size(), -2);
            std::vector<int32_t> bests(big_ids.size(), -1);
            tb_parallel_for(big_ids.size(), [&](size_t ib, size_t ie) {
                for (size_t i = ib; i < ie; ++i) {
                    const int id = big_ids[i];
                    const uint32_t ql = nodes[id].qy - nodes[id].qx, tl = nodes[id].tv - nodes[id].tu;
                    const uint32_t left = tl / 2, right = tl - left;
                    const int32_t* L = rows[2 * i].data();
                    const int32_t* Rr = rows[2 * i + 1].data();   // Rr[i'] = dist(rev q prefix i'+1, rev right half)
                    auto R = [&](uint32_t j) { return Rr[ql - 1 - j]; };  // dist(q[j:], right half)
                    int32_t best = std::min((int32_t)left + R(0), L[ql - 1] + (int32_t)right);
                    for (uint32_t r = 0; r + 1 < ql; ++r) best = std::min(best, L[r] + Rr[ql - 2 - r]);
                    int split = -2;
                    for (uint32_t r = 0; r + 1 < ql; ++r)
                        if (L[r] + Rr[ql - 2 - r] == best) { split = (int)r; break; }
                    if (split == -2 && (int32_t)left + R(0) == best) split = -1;
                    if (split == -2 && L[ql - 1] + (int32_t)right == best) split = (int)ql - 1;
                    splits[i] = split; bests[i] = best;
                }
            });
            for (size_t i = 0; i < big_ids.size(); ++i) {
                const int id = big_ids[i];
                const uint32_t tl = nodes[id].tv - nodes[id].tu;
                const uint32_t left = tl / 2;
                const int split = splits[i];
                if (split == -2) throw std::runtime_error("alignment split not found");
                const uint32_t ul_h = (uint32_t)(split + 1);
                const uint32_t pa = nodes[id].a, pqx = nodes[id].qx, pqy = nodes[id].qy, ptu = nodes[id].tu, ptv = nodes[id].tv;
                nodes[id].dist = bests[i];
                const int c0 = (int)nodes.size();
                nodes.push_back({pa, pqx, pqx + ul_h, ptu, ptu + left, {-1, -1}, {}, -1, 0});
                nodes.push_back({pa, pqx + ul_h, pqy, ptu + left, ptv, {-1, -1}, {}, -1, 0});
                nodes[id].child[0] = c0; nodes[id].child[1] = c0 + 1;
                next.push_back(c0); next.push_back(c0 + 1);
            }
        }
        work.swap(next);
    }
}

// in-order concatenation of the leaves: ops[a] / dist[a]
inline void solve_nw_paths(TbBackend& be, uint32_t n, const uint64_t* q_beg, const uint32_t* q_len, const uint64_t* t_beg,
                           const uint32_t* t_len, std::vector<std::vector<uint8_t>>& ops, std::vector<int32_t>& dist) {
    std::vector<TbNode> nodes;
    solve_nw_tree(be, n, q_beg, q_len, t_beg, t_len, nullptr, nodes);
    ops.assign(n, {});
    dist.assign(n, -1);
    tb_parallel_for(n, [&](size_t ab, size_t ae) {
        for (size_t a = ab; a < ae; ++a) {
            std::vector<int> st(1, (int)a);
            size_t tot = 0;
            while (!st.empty()) {
                const int id = st.back();
                st.pop_back();
                if (nodes[id].child[0] < 0) tot += nodes[id].ops.size();
                else { st.push_back(nodes[id].child[1]); st.push_back(nodes[id].child[0]); }
            }
            ops[a].reserve(tot);
            st.assign(1, (int)a);
            while (!st.empty()) {
                const int id = st.back();
                st.pop_back();
                if (nodes[id].child[0] < 0) ops[a].insert(ops[a].end(), nodes[id].ops.begin(), nodes[id].ops.end());
                else { st.push_back(nodes[id].child[1]); st.push_back(nodes[id].child[0]); }
            }
            dist[a] = nodes[a].dist;
        }
    });
}

// the same paths as runs of equal kind (match and mismatch are one kind), sub-problems outside `need` replaced as described there
inline void solve_nw_runs(TbBackend& be, uint32_t n, const uint64_t* q_beg, const uint32_t* q_len, const uint64_t* t_beg,
                          const uint32_t* t_len, const TbNeed* need, std::vector<std::vector<TbRun>>& runs) {
    std::vector<TbNode> nodes;
    solve_nw_tree(be, n, q_beg, q_len, t_beg, t_len, need, nodes);
    runs.assign(n, {});
    tb_parallel_for(n, [&](size_t ab, size_t ae) {
        for (size_t a = ab; a < ae; ++a) {
            std::vector<TbRun>& out = runs[a];
            auto push = [&](uint8_t kind, uint32_t len) {
                if (!len) return;
                if (!out.empty() && out.back().kind == kind) out.back().len += len; else out.push_back({kind, len});
            };
            std::vector<int> st(1, (int)a);
            while (!st.empty()) {
                const int id = st.back();
                st.pop_back();
                const TbNode& nd = nodes[id];
                if (nd.child[0] >= 0) { st.push_back(nd.child[1]); st.push_back(nd.child[0]); continue; }
                if (nd.pruned) { push(1, nd.qy - nd.qx); push(2, nd.tv - nd.tu); continue; }
                for (const uint8_t o : nd.ops) push((o == 1) ? 1 : (o == 2) ? 2 : 0, 1);
            }
        }
    });
}

// placement of a set of items in the matrix / ops scratch, per lane-group class
struct TbPlan {
    std::vector<uint32_t> order[6];
    std::vector<uint32_t> ids;
    std::vector<uint64_t> mat_off, ops_off, hb_off;
    uint64_t cells = 0;
};

inline TbPlan plan_items(const std::vector<TbItem>& items, bool lastcol) {
    TbPlan pl;
    const uint32_t n = (uint32_t)items.size();
    pl.mat_off.assign(n + 1, 0); pl.ops_off.assign(n + 1, 0); pl.hb_off.assign(n + 1, 0);
    for (uint32_t a = 0; a < n; ++a) {
        const uint64_t nb = ((uint64_t)items[a].q_len + 63) / 64;
        pl.order[tb_class_of(nb)].push_back(a);
        pl.mat_off[a] = pl.cells;
        pl.cells += lastcol ? nb : nb * items[a].t_len;
        pl.ops_off[a + 1] = pl.ops_off[a] + (lastcol ? 0 : (uint64_t)items[a].q_len + items[a].t_len);
        pl.hb_off[a + 1] = pl.hb_off[a] + items[a].t_len;
    }
    pl.mat_off[n] = pl.cells;
    for (int c = 0; c < 6; ++c) {
        std::stable_sort(pl.order[c].begin(), pl.order[c].end(), [&](uint32_t x, uint32_t y) { return items[x].t_len > items[y].t_len; });
        pl.ids.insert(pl.ids.end(), pl.order[c].begin(), pl.order[c].end());
    }
    return pl;
}

}  // namespace rtk
