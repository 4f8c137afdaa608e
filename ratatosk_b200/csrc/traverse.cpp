// traverse.cpp — see traverse.hpp.
#include "traverse.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <stdexcept>
#include <unordered_map>

#include "broker.hpp"
#include "kmer.cuh"

namespace rtk {

namespace {

inline uint32_t usize(const rtk_graph_view& g, uint32_t u) { return (uint32_t)(g.unitig_off[u + 1] - g.unitig_off[u]); }
inline uint32_t ufull(const rtk_graph_view& g, uint32_t u) { return usize(g, u) - g.k + 1; }
inline bool has_shared_pids(const rtk_graph_view& g, uint32_t u) { return (g.shared[u] & 0xffULL) != 0; }
inline bool is_short_cycle(const rtk_graph_view& g, uint32_t u) { return (g.shared[u] & 0x100ULL) != 0; }

// oriented spelling of a mapping: unitig[dist, dist+len+k-1) forward, or its reverse complement
void append_mapped(const rtk_graph_view& g, const PNode& n, size_t skip, std::string& out) {
    const uint64_t ub = g.unitig_off[n.unitig];
    const size_t mlen = (size_t)n.len + g.k - 1;
    for (size_t t = skip; t < mlen; ++t) {
        const uint64_t pos = n.strand ? (n.dist + t) : (n.dist + (mlen - 1 - t));
        uint32_t b = rtk_pool_base(g.pool, ub + pos);
        if (!n.strand) b = 3 - b;
        out.push_back("ACGT"[b]);
    }
}

// find(um.getMappedTail().forwardBase(c), true) for a vertex whose mapping ends at the unitig's end in
// traversal orientation (every vertex the traversal extends from)
PNode successor(const rtk_graph_view& g, const PNode& n, uint32_t base) {
    const uint32_t slot = n.strand ? g.adj[8 * (uint64_t)n.unitig + base] : g.adj[8 * (uint64_t)n.unitig + 4 + (3 - base)];
    PNode r;
    if (slot == RTK_NONE32) return r;
    r.unitig = slot & 0x7fffffffu;
    r.strand = n.strand ? (slot >> 31) : (1u - (slot >> 31));
    r.dist = 0;
    r.len = ufull(g, r.unitig);
    return r;
}

inline uint32_t base_code(char c) { return c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 4u; }

}  // namespace

// ---------------------------------------------------------------------------------------------- GPath
bool GPath::extend(const rtk_graph_view& g, const PNode& um) {
    if (um.empty()) return false;
    if (v.empty()) { v.push_back(um); l = (size_t)um.len + g.k - 1; }
    else {
        if (v.size() >= 2) { v.back().dist = 0; v.back().len = ufull(g, v.back().unitig); }  // old end becomes an interior (whole) vertex
        v.push_back(um);
        l += um.len;
    }
    return true;
}

bool GPath::extend(const rtk_graph_view& g, const PNode& um, const std::string& qual_s) {
    if (um.empty()) return false;
    const size_t k = g.k, sub = (size_t)um.len + k - 1;
    if (v.empty()) {
        v.push_back(um);
        l = sub;
        if (qual_s.length() == sub) qual = qual_s;
        else return false;
    } else {
        if (v.size() >= 2) { v.back().dist = 0; v.back().len = ufull(g, v.back().unitig); }
        v.push_back(um);
        l += um.len;
        if (qual_s.length() == sub) qual += qual_s.substr(k - 1);
        else return false;
    }
    return true;
}

bool GPath::merge(const rtk_graph_view& g, const GPath& o) {
    if (o.length() == 0) return true;
    if (length() == 0) { *this = o; return true; }
    if (qual.empty() != o.qual.empty()) return false;
    const PNode& last = v.back();
    if (last.unitig != o.v.front().unitig || last.strand != o.v.front().strand) return false;
    const size_t k = g.k;
    PNode& e = v.back();
    if (!e.strand) e.dist = o.v.front().dist;
    e.len += o.v.front().len - 1;
    if (v.size() == 1) {
        v.insert(v.end(), o.v.begin() + 1, o.v.end());
    } else if (o.v.size() >= 2) {
        e.dist = 0; e.len = ufull(g, e.unitig);  // becomes interior
        v.insert(v.end(), o.v.begin() + 1, o.v.end());
    }
    l += o.l - k;
    if (o.qual.length() != 0) qual.append(o.qual.substr(k));
    return true;
}

void GPath::prune_prefix(const rtk_graph_view& g, size_t len) {
    if (v.empty() || l == 0 || len >= l) return;
    const size_t k = g.k;
    size_t cum = 0;
    for (size_t i = 0; i < v.size(); ++i) {
        cum += (i == 0) ? ((size_t)v[i].len + k - 1) : (size_t)v[i].len;
        if (cum >= len || i + 1 == v.size()) {
            const size_t excess = cum - len;
            if (!v[i].strand) v[i].dist += (uint32_t)excess;
            v[i].len -= (uint32_t)excess;
            v.resize(i + 1);
            break;
        }
    }
    l = len;
    if (qual.length() != 0) qual = qual.substr(0, l);
}

std::string GPath::to_string(const rtk_graph_view& g) const {
    std::string s;
    if (v.empty()) return s;
    s.reserve(l + 8);
    for (size_t i = 0; i < v.size(); ++i) append_mapped(g, v[i], i == 0 ? 0 : g.k - 1, s);
    return s;
}

GPath GPath::rev_comp() const {
    GPath out;
    if (v.empty()) return out;
    out.l = l;
    out.qual = qual;
    std::reverse(out.qual.begin(), out.qual.end());
    out.v.assign(v.rbegin(), v.rend());
    for (auto& n : out.v) n.strand = 1u - n.strand;
    return out;
}

GPath GPath::from_compact(const rtk_graph_view& g, const PNode& um_start, const std::string& ext, const PNode& um_end) {
    GPath p;
    if (um_start.empty()) return p;
    const size_t k = g.k;
    if (um_end.empty() && ext.empty()) { p.v.push_back(um_start); p.l = (size_t)um_start.len + k - 1; return p; }
    PNode curr = um_start;
    size_t len = (size_t)um_start.len + k - 1;
    std::vector<PNode> mid;
    for (const char c : ext) {
        const uint32_t b = base_code(c);
        curr = (b < 4) ? successor(g, curr, b) : PNode();
        if (curr.empty()) return p;
        len += curr.len;
        mid.push_back(curr);
    }
    if (um_end.empty()) return p;
    len += um_end.len;
    p.v.push_back(um_start);
    p.v.insert(p.v.end(), mid.begin(), mid.end());
    p.v.push_back(um_end);
    p.l = len;
    return p;
}

// ---------------------------------------------------------------------------------------------- GPU services
// A request is handed to the wave broker when this thread runs under one (broker.hpp), else executed at once.
void gpu_distances_fl(rtk_ctx* ctx, const std::vector<AlignJob>& jobs, std::vector<int32_t>& dist, std::vector<int32_t>& first_end,
                      std::vector<int32_t>& last_end) {
    dist.assign(jobs.size(), -1);
    first_end.assign(jobs.size(), -1);
    last_end.assign(jobs.size(), -1);
    if (jobs.empty()) return;
    DistReq r{&jobs, &dist, &first_end, &last_end};
    if (GpuBroker* b = current_broker()) b->submit(&r);
    else run_dist_batch(ctx, std::vector<DistReq*>(1, &r));
}

void gpu_distances(rtk_ctx* ctx, const std::vector<AlignJob>& jobs, std::vector<int32_t>& dist, std::vector<int32_t>& first_end) {
    std::vector<int32_t> last_end;
    gpu_distances_fl(ctx, jobs, dist, first_end, last_end);
}

void gpu_paths(rtk_ctx* ctx, const std::vector<AlignJob>& jobs, std::vector<int32_t>& dist, std::vector<std::vector<uint8_t>>& ops) {
    dist.assign(jobs.size(), -1);
    ops.assign(jobs.size(), {});
    if (jobs.empty()) return;
    PathReq r{&jobs, &dist, &ops};
    if (GpuBroker* b = current_broker()) b->submit(&r);
    else run_path_batch(ctx, std::vector<PathReq*>(1, &r));
}

// ---------------------------------------------------------------------------------------------- selectors
// The reference aligns candidate i > 0 with a band derived from the best so far and requires a strictly
// better normalised distance; a candidate that can win is always inside that band, so aligning all of them
// unbounded in one batch and replaying the scan gives the same winner.
template <typename GetPath>
static std::pair<int, int> select_generic(rtk_ctx* ctx, const rtk_graph_view& g, size_t n, GetPath get, const std::string& ref,
                                          uint8_t mode, bool norm_by_max) {
    std::vector<AlignJob> jobs(n);
    for (size_t i = 0; i < n; ++i) { jobs[i].q = get(i).to_string(g); jobs[i].tref = &ref; jobs[i].mode = mode; }
    std::vector<int32_t> dist, fe;
    gpu_distances(ctx, jobs, dist, fe);
    double best = 0.0;
    int best_id = -1, best_end = -1;
    for (size_t i = 0; i < n; ++i) {
        const size_t norm = norm_by_max ? std::max(jobs[i].q.length(), ref.length()) : jobs[i].q.length();
        const double d = static_cast<double>(dist[i]) / norm;
        if (i == 0 || (dist[i] >= 0 && d < best)) { best = d; best_id = (int)i; best_end = fe[i]; }
    }
    return {best_id, best_end};
}

std::pair<int, int> select_best_alignment(rtk_ctx* ctx, const rtk_graph_view& g, const std::vector<GPath>& c, const std::string& ref) {
    return select_generic(ctx, g, c.size(), [&](size_t i) -> const GPath& { return c[i]; }, ref, 0, true);
}
std::pair<int, int> select_best_prefix_alignment(rtk_ctx* ctx, const rtk_graph_view& g, const std::vector<const GPath*>& c, const std::string& ref) {
    return select_generic(ctx, g, c.size(), [&](size_t i) -> const GPath& { return *c[i]; }, ref, 1, false);
}
std::pair<int, int> select_best_substring_alignment(rtk_ctx* ctx, const rtk_graph_view& g, const std::vector<GPath>& c, const std::string& ref) {
    return select_generic(ctx, g, c.size(), [&](size_t i) -> const GPath& { return c[i]; }, ref, 2, false);
}

// ---------------------------------------------------------------------------------------------- exploreSubGraph + qualities
namespace {

struct Burst {
    std::vector<GPath> terminal, nonterminal;
    double t1 = 0.0, nt1 = 0.0;
};

// getScorePath(opt, path, ref, ref_len, score_best, score_second_best) (src/GraphTraversal.cpp:722-772) for the kept
// terminal and non-terminal paths of one burst
void set_qualities2(rtk_ctx* ctx, const rtk_graph_view& g, const TraverseOpt& opt, std::vector<GPath>& pa, double best_a, double second_a,
                    std::vector<GPath>& pb, double best_b, double second_b, const std::string& ref) {
    const size_t na = pa.size(), n = na + pb.size();
    if (!n) return;
    std::vector<AlignJob> jobs(n);
    for (size_t i = 0; i < n; ++i) { jobs[i].q = (i < na ? pa[i] : pb[i - na]).to_string(g); jobs[i].tref = &ref; jobs[i].mode = 1; }
    std::vector<int32_t> dist;
    std::vector<std::vector<uint8_t>> ops;
    gpu_paths(ctx, jobs, dist, ops);
    for (size_t i = 0; i < n; ++i) {
        const double best = i < na ? best_a : best_b, second = i < na ? second_a : second_b;
        const double score_comp = best * ((best == 0.0) ? 0.0 : (1.0 - (second / best)));
        const char c_best = rtk_get_qual(best, 0, opt.max_qual);
        const std::string& ps = jobs[i].q;
        std::string q(ps.length(), rtk_get_qual(score_comp, opt.out_qual, opt.max_qual));
        size_t qp = 0, rp = 0;
        for (const uint8_t op : ops[i]) {
            if (op == 0 || op == 3) { if (ps[qp] == ref[rp]) q[qp] = c_best; ++qp; ++rp; }
            else if (op == 1) ++qp;
            else ++rp;
        }
        (i < na ? pa[i] : pb[i - na]).set_quality(q);
    }
}

// One `explore` step (src/GraphTraversal.cpp:251-304) as a single GPU request: prefix alignment -> window start, burst on the
// rest of the window (K2/K3 + leaf K4), then the qualities of the kept paths (K5).  `sub_out` = the part of the window the burst saw.
Burst explore_subgraph(rtk_ctx* ctx, const rtk_graph_view& g, const TraverseOpt& opt, const std::vector<uint32_t>& pids,
                       const std::string& ref, const std::string& prefix, size_t path_len, size_t max_len_path_total, size_t l_max,
                       const PNode& um, const PNode& um_e, uint32_t level, std::string& sub_out, bool& explored) {
    SubgraphReq rq;
    memset(&rq.call, 0, sizeof(rq.call));
    rq.call.start_unitig = um.unitig; rq.call.start_strand = um.strand;
    if (um_e.empty()) rq.call.end_unitig = RTK_NONE32;
    else { rq.call.end_unitig = um_e.unitig; rq.call.end_strand = um_e.strand; rq.call.end_dist = um_e.dist; }
    rq.call.level = level; rq.call.max_len_path = (uint32_t)l_max; rq.call.min_cov = opt.min_cov_vertices;
    rq.call.max_len_subpath = opt.long_read_correct ? (uint32_t)opt.max_len_subpath() : 0u;
    rq.ref = &ref; rq.prefix = &prefix; rq.path_len = path_len; rq.max_len_path_total = max_len_path_total;
    rq.pids = &pids; rq.wrlf = opt.weak_region_len_factor;
    SubgraphResult res;
    rq.out = &res;
    if (GpuBroker* b = current_broker()) b->submit(&rq);
    else run_subgraph_batch(ctx, std::vector<SubgraphReq*>(1, &rq));
    Burst b;
    explored = res.explored;
    if (!res.explored) return b;
    sub_out = ref.substr(res.end_pos_ref);
    b.t1 = res.scores[0]; b.nt1 = res.scores[2];
    for (int kind = 0; kind < 2; ++kind) {
        for (const auto& nodes : (kind == 0 ? res.terminal : res.nonterminal)) {
            GPath p;
            for (const PNode& n : nodes) p.extend(g, n);
            (kind == 0 ? b.terminal : b.nonterminal).push_back(std::move(p));
        }
    }
    // both quality batches of the burst in one K5 request
    set_qualities2(ctx, g, opt, b.terminal, b.t1, res.scores[1], b.nonterminal, b.nt1, res.scores[3], sub_out);
    return b;
}

// the `explore` lambda shared by explorePathsBFS2 (:251-304) and explorePathsBFS (:43-94)
Burst explore_step(rtk_ctx* ctx, const rtk_graph_view& g, const TraverseOpt& opt, const std::vector<uint32_t>& pids, const std::string& ref,
                   const PNode& um, const GPath& path, size_t max_len_path, uint32_t level, const PNode& um_e) {
    const size_t k = g.k;
    const size_t path_len = path.length();
    const bool non_empty_path = (path_len > ((size_t)um.len + k - 1)) && !um.empty();
    const size_t path_len_prefix = non_empty_path ? (path_len - um.len - k + 1) : 0;
    std::string prefix;
    if (non_empty_path) prefix = path.to_string(g).substr(0, path_len_prefix);
    const size_t l_max = max_len_path - path_len_prefix;
    std::string sub;
    bool explored = false;
    Burst b = explore_subgraph(ctx, g, opt, pids, ref, prefix, path_len, max_len_path, l_max, um, um_e, level - 1, sub, explored);
    if (explored) {
        if (!b.terminal.empty() && b.t1 < opt.min_score) b.terminal.clear();
        if (!b.nonterminal.empty() && b.nt1 < opt.min_score) b.nonterminal.clear();
        if (b.nonterminal.size() > 1) {
            const int id = select_best_substring_alignment(ctx, g, b.nonterminal, sub).first;
            GPath keep = b.nonterminal[(size_t)id];
            b.nonterminal.assign(1, keep);
        }
    }
    return b;
}

void resize_vector(rtk_ctx* ctx, const rtk_graph_view& g, std::vector<GPath>& v, const std::string& ref) {
    if (v.size() <= 1) return;
    std::vector<const GPath*> ptr;
    for (const auto& p : v) ptr.push_back(&p);
    const int id = select_best_prefix_alignment(ctx, g, ptr, ref).first;
    GPath keep = v[(size_t)id];
    v.assign(1, keep);
}

void resize_queue(rtk_ctx* ctx, const rtk_graph_view& g, std::queue<GPath>& q, const std::string& ref) {
    if (q.empty()) return;
    std::vector<GPath> v;
    while (!q.empty()) { v.push_back(std::move(q.front())); q.pop(); }
    resize_vector(ctx, g, v, ref);
    for (auto& p : v) q.push(std::move(p));
}

// p_ext = p + the burst path's vertices with their slice of the burst's quality string
GPath extend_with(const rtk_graph_view& g, const GPath& p, const GPath& burst) {
    GPath e(p);
    const size_t k = g.k;
    size_t j = 0;
    for (const PNode& n : burst.v) {
        e.extend(g, n, burst.qual.substr(std::min(j, burst.qual.length()), (size_t)n.len + k - 1));
        j += n.len;
    }
    return e;
}

}  // namespace

// ---------------------------------------------------------------------------------------------- explorePathsBFS2
std::vector<GPath> explore_paths_bfs2(rtk_ctx* ctx, const rtk_graph_view& g, const TraverseOpt& opt, const std::string& ref,
                                      const std::vector<uint32_t>& all_pids, const PNode& um_s, const PNode& um_e) {
    std::vector<GPath> v, v_tmp;
    const size_t k = g.k, ref_len = ref.length();
    if (!um_s.empty() && !um_e.empty() && has_shared_pids(g, um_s.unitig) && has_shared_pids(g, um_e.unitig)) {
        const uint32_t level = 4;
        const size_t min_len_path = rtk_min_max_length(ref_len - k, opt.weak_region_len_factor).first + k;
        const size_t max_len_path = std::max<size_t>(rtk_min_max_length(ref_len - k, opt.weak_region_len_factor).second, 10UL) + k;
        const size_t max_paths = 1024, max_sz_stck = 512;
        std::queue<GPath> q;
        {
            PNode st = um_s;  // suffix of the start unitig from the anchor k-mer, in traversal orientation
            const uint32_t sz = usize(g, um_s.unitig);
            if (st.strand) { st.dist += st.len - 1; st.len = sz - st.dist - (uint32_t)k + 1; }
            else { st.len = um_s.dist + 1; st.dist = 0; }
            GPath p_start;
            const char qmax = rtk_get_qual(1.0, 0, opt.max_qual);
            if (um_s.unitig == um_e.unitig && um_s.strand == um_e.strand && st.dist <= um_e.dist) {
                const size_t len = ((size_t)st.len + k - 1) - (um_e.strand ? (size_t)sz - um_e.dist - k : (size_t)um_e.dist);
                if (len >= min_len_path && len <= max_len_path) {
                    PNode back = st;
                    if (back.strand) back.len = um_e.dist - back.dist + 1;
                    else { back.dist = um_e.dist; back.len -= um_e.dist; }
                    p_start.extend(g, back, std::string((size_t)back.len + k - 1, qmax));
                    v.push_back(std::move(p_start));
                    p_start.clear();
                }
            }
            p_start.extend(g, st, std::string((size_t)st.len + k - 1, qmax));
            q.push(std::move(p_start));
        }
        while (!q.empty()) {
            GPath p = std::move(q.front());
            q.pop();
            const PNode um = p.back();
            if (p.length() < max_len_path) {
                Burst b = explore_step(ctx, g, opt, all_pids, ref, um, p, max_len_path, level, um_e);
                for (const auto& path : b.terminal) v_tmp.push_back(extend_with(g, p, path));
                for (const auto& path : b.nonterminal) {
                    if ((!opt.long_read_correct && path.size() == level) || (opt.long_read_correct && path.length() >= opt.max_len_subpath())) {
                        q.push(extend_with(g, p, path));
                        if (q.size() >= max_sz_stck) resize_queue(ctx, g, q, ref);
                    }
                }
                if (v_tmp.size() >= max_paths) {
                    for (auto& t : v_tmp) {
                        if (t.length() >= min_len_path && t.length() <= max_len_path) {
                            if (v.size() + 1 >= max_paths) resize_vector(ctx, g, v, ref);
                            v.push_back(std::move(t));
                        }
                    }
                    v_tmp.clear();
                }
            }
        }
        for (auto& t : v_tmp) {
            if (t.length() >= min_len_path && t.length() <= max_len_path) {
                if (v.size() + 1 >= max_paths) resize_vector(ctx, g, v, ref);
                v.push_back(std::move(t));
            }
        }
        v_tmp.clear();
    }
    if (!v.empty()) {
        if (v.size() > 1) { GPath keep = v[(size_t)select_best_alignment(ctx, g, v, ref).first]; v.assign(1, keep); }
        v = fix_repeats(ctx, g, opt, v, ref);
    }
    return v;
}

// ---------------------------------------------------------------------------------------------- explorePathsBFS
std::vector<GPath> explore_paths_bfs(rtk_ctx* ctx, const rtk_graph_view& g, const TraverseOpt& opt, const std::string& ref,
                                     const std::vector<uint32_t>& all_pids, const PNode& um_s) {
    std::vector<GPath> v, v_tmp;
    const size_t k = g.k, ref_len = ref.length();
    if (!um_s.empty() && has_shared_pids(g, um_s.unitig)) {
        const uint32_t level = 4;
        const size_t min_len_path = rtk_min_max_length(ref_len - k, opt.weak_region_len_factor).first + k;
        const size_t max_len_path = std::max<size_t>(rtk_min_max_length(ref_len - k, opt.weak_region_len_factor).second, 10UL) + k;
        const size_t max_paths = 1024, max_sz_stck = 512;
        std::queue<GPath> q;
        {
            PNode st = um_s;
            const uint32_t sz = usize(g, um_s.unitig);
            if (st.strand) { st.dist += st.len - 1; st.len = sz - st.dist - (uint32_t)k + 1; }
            else { st.len = um_s.dist + 1; st.dist = 0; }
            const char qmax = rtk_get_qual(1.0, 0, opt.max_qual);
            GPath p_tmp;
            if (((size_t)st.len + k - 1) >= min_len_path) {
                PNode back = st;
                if (((size_t)back.len + k - 1) > max_len_path) {
                    if (!back.strand) back.dist = back.len - (uint32_t)(max_len_path - k + 1);
                    back.len = (uint32_t)(max_len_path - k + 1);
                }
                p_tmp.extend(g, back, std::string((size_t)back.len + k - 1, qmax));
                v.push_back(std::move(p_tmp));
                p_tmp.clear();
            }
            p_tmp.extend(g, st, std::string((size_t)st.len + k - 1, qmax));
            q.push(std::move(p_tmp));
        }
        while (!q.empty()) {
            GPath p = std::move(q.front());
            q.pop();
            const PNode um = p.back();
            if (p.length() < max_len_path) {
                Burst b = explore_step(ctx, g, opt, all_pids, ref, um, p, max_len_path, level, PNode());
                for (const auto& path : b.nonterminal) {
                    GPath p_ext(p);
                    size_t j = 0;
                    for (const PNode& n : path.v) {
                        p_ext.extend(g, n, path.qual.substr(std::min(j, path.qual.length()), (size_t)n.len + k - 1));
                        if (p_ext.length() >= min_len_path && p_ext.length() <= max_len_path) v_tmp.push_back(p_ext);
                        j += n.len;
                    }
                    if ((!opt.long_read_correct && path.size() == level) || (opt.long_read_correct && path.length() >= opt.max_len_subpath())) {
                        q.push(std::move(p_ext));
                        if (q.size() >= max_sz_stck) resize_queue(ctx, g, q, ref);
                    }
                }
                if (v_tmp.size() >= max_paths) {
                    for (auto& t : v_tmp) { t.prune_prefix(g, max_len_path); v.push_back(std::move(t)); }
                    v_tmp.clear();
                }
            }
        }
        for (auto& t : v_tmp) { t.prune_prefix(g, max_len_path); v.push_back(std::move(t)); }
    }
    if (!v.empty()) {
        if (v.size() > 1) { GPath keep = v[(size_t)select_best_alignment(ctx, g, v, ref).first]; v.assign(1, keep); }
        v = fix_repeats(ctx, g, opt, v, ref);
    }
    return v;
}

// ---------------------------------------------------------------------------------------------- fixRepeats
std::vector<GPath> fix_repeats(rtk_ctx* ctx, const rtk_graph_view& g, const TraverseOpt& opt, const std::vector<GPath>& v_path, const std::string& ref) {
    std::vector<GPath> out;
    const size_t k = g.k;
    // (unitig, traversal strand) -> candidate cycle paths already built for it (m_cycles, keyed by the mapped head k-mer)
    std::unordered_map<uint64_t, std::vector<GPath>> m_cycles;
    const char qmax = rtk_get_qual(1.0, 0, opt.max_qual);
    for (GPath path : v_path) {
        bool any_cycle = false;
        for (const PNode& n : path.v) any_cycle |= is_short_cycle(g, n.unitig);
        if (!any_cycle) { out.push_back(path); continue; }  // the reference's first alignment only seeds comparisons
        std::vector<PNode> v_um = path.v;
        std::string s_qual = path.qual;
        int64_t edit;
        {
            std::vector<AlignJob> j(1);
            j[0].q = path.to_string(g).substr(0, path.length()); j[0].tref = &ref; j[0].mode = 0;
            std::vector<int32_t> d, fe;
            gpu_distances(ctx, j, d, fe);
            edit = d[0];
        }
        // The reference tries the stored cycles of a vertex one by one, each alignment bounded by the best distance so far
        // (k = edit: worse => -1).  The unbounded distances do not depend on that bound, so all candidates of a vertex are
        // aligned in ONE K4 request and the sequential accept / reject scan is replayed on the results.
        auto build_ext = [&](const GPath& repeat, size_t pos) -> GPath {
            GPath ext;
            size_t len_prefix = 0;
            for (size_t j = 0; j < pos; ++j) { ext.extend(g, v_um[j]); len_prefix += v_um[j].len; }
            for (const PNode& n : repeat.v) ext.extend(g, n);
            for (size_t j = pos + 1; j < v_um.size(); ++j) ext.extend(g, v_um[j]);
            std::string lq = s_qual;
            if (len_prefix <= lq.length()) lq.replace(len_prefix, (size_t)v_um[pos].len + k - 1, std::string(repeat.length(), qmax), 0, repeat.length());
            ext.set_quality(lq);
            return ext;
        };
        // candidate cycle paths of vertex i of the current path (built once per (unitig, strand) for interior vertices)
        std::vector<GPath> local_cands;
        auto candidates_of = [&](size_t i) -> const std::vector<GPath>* {
            const PNode um = v_um[i];
            PNode us = um, ue = um;
            us.len = usize(g, um.unitig) - um.dist - (uint32_t)k + 1;
            ue.dist = 0;
            ue.len = um.dist + um.len;
            us.strand = 1; ue.strand = 1;
            std::vector<std::string> cycles;
            {
                const char* cp = g.cyc_pool + g.cyc_off[um.unitig];
                const size_t cl = (size_t)(g.cyc_off[um.unitig + 1] - g.cyc_off[um.unitig]);
                size_t s = 0;
                while (s < cl) { const size_t n = strnlen(cp + s, cl - s); cycles.emplace_back(cp + s, n); s += n + 1; }
            }
            if (i == 0 || i == v_um.size() - 1) {
                local_cands.clear();
                for (const auto& cyc : cycles) {
                    const GPath pe = GPath::from_compact(g, us, cyc, ue);
                    local_cands.push_back(um.strand ? pe : pe.rev_comp());
                }
                return &local_cands;
            }
            const uint64_t key = ((uint64_t)um.unitig << 1) | um.strand;
            auto ins = m_cycles.insert({key, std::vector<GPath>()});
            if (ins.second) {
                for (const auto& cyc : cycles) {
                    GPath pe = GPath::from_compact(g, us, cyc, ue);
                    if (!um.strand) pe = pe.rev_comp();
                    ins.first->second.push_back(pe);
                }
            }
            return &ins.first->second;
        };
        // The reference walks the vertices left to right and re-aligns after every accepted cycle.  Acceptances are rare, so the
        // walk is SPECULATED: every vertex the walk would visit if nothing were accepted is evaluated in one K4 request; the
        // accept / reject scan is replayed in order, and only an acceptance (which changes the path) restarts the speculation
        // from the next vertex.  Same decisions, one request per acceptance instead of one per vertex.
        size_t i = 0;
        while (i < v_um.size()) {
            struct Visit { size_t i; size_t first_job, n_jobs; };
            std::vector<Visit> visits;
            std::vector<GPath> exts;
            std::vector<AlignJob> jobs;
            for (size_t j = i; j < v_um.size(); ++j) {
                if (!is_short_cycle(g, v_um[j].unitig)) continue;
                const std::vector<GPath>* cands = candidates_of(j);
                visits.push_back({j, jobs.size(), cands->size()});
                for (const GPath& cand : *cands) {
                    exts.push_back(build_ext(cand, j));
                    AlignJob aj;
                    aj.q = exts.back().to_string(g).substr(0, exts.back().length()); aj.tref = &ref; aj.mode = 0;
                    jobs.push_back(std::move(aj));
                }
                // no acceptance at j: the following vertices on the same unitig are skipped (:1322-1330)
                size_t jn = j;
                for (size_t t = j + 1; t < v_um.size(); ++t) { if (v_um[j].unitig == v_um[t].unitig) ++jn; else break; }
                j = jn;
            }
            if (visits.empty()) break;
            std::vector<int32_t> d, fe;
            gpu_distances(ctx, jobs, d, fe);
            bool accepted = false;
            for (const Visit& vis : visits) {
                GPath best_ext;
                for (size_t c = 0; c < vis.n_jobs; ++c) {
                    const int32_t dc = d[vis.first_job + c];
                    const int64_t rd = (dc > edit) ? -1 : dc;   // bounded alignment (k = edit): worse than the running best => -1
                    if (rd >= 0 && rd < edit) { edit = rd; best_ext = std::move(exts[vis.first_job + c]); }
                }
                if (best_ext.length() != 0) {
                    const size_t diff = best_ext.size() - path.size();
                    path = std::move(best_ext);
                    v_um = path.v;
                    s_qual = path.qual;
                    i = vis.i + diff;   // the reference: i += diff - 1, then the loop's ++i
                    accepted = true;
                    break;
                }
            }
            if (!accepted) break;
        }
        out.push_back(path);
    }
    return out;
}

}  // namespace rtk
