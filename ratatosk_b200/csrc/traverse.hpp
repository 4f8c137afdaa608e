// traverse.hpp — host orchestration of the bounded path search between anchors:
// explorePathsBFS2 / explorePathsBFS (src/GraphTraversal.cpp:212-454, :3-210), the selectors
// (src/Alignment.cpp:3-147, :967-1015), path qualities (getScorePath, src/GraphTraversal.cpp:722-772) and
// fixRepeats (:1149-1334), restated over the flat graph.  Everything numeric runs in batches on the GPU
// through the C ABI entries of this library: rtk_explore_subgraph_batch (K2/K3+K4), rtk_edlib_batch (K4),
// rtk_edlib_path_batch (K5).  What stays here is the order-defining control flow (queue of partial paths,
// first-wins / ties-kept selections) the survey says to keep on the host until it moves device-side.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/rtk.h"
#include "flat_graph.h"

namespace rtk {

// const_UnitigMap<UnitigData> restricted to what the traversal needs
struct PNode {
    uint32_t unitig = RTK_NONE32, strand = 0, dist = 0, len = 0;
    bool empty() const { return unitig == RTK_NONE32; }
};

// Path<UnitigData> (src/Path.hpp) with explicit vertices instead of {start, succ string, end}: interior
// vertices are always whole unitigs (the reference re-walks them from `succ`, which forgets any partial
// mapping), the first and last may be partial.  `l` is kept exactly as the reference accumulates it.
class GPath {
public:
    std::vector<PNode> v;
    std::string qual;
    size_t l = 0;

    void clear() { v.clear(); qual.clear(); l = 0; }
    size_t size() const { return v.size(); }
    size_t length() const { return l; }
    const PNode& front() const { return v.front(); }
    const PNode& back() const { return v.back(); }
    bool extend(const rtk_graph_view& g, const PNode& um);
    bool extend(const rtk_graph_view& g, const PNode& um, const std::string& qual_s);
    bool merge(const rtk_graph_view& g, const GPath& o);
    void prune_prefix(const rtk_graph_view& g, size_t len);   // Path::prunePrefix: keep the first `len` bases
    void set_quality(const std::string& q) { if (q.length() == l) qual = q; }
    std::string to_string(const rtk_graph_view& g) const;
    GPath rev_comp() const;
    // Path(um_start, ext, um_end): follow `ext` (one base per hop) from um_start, finish on um_end
    static GPath from_compact(const rtk_graph_view& g, const PNode& um_start, const std::string& ext, const PNode& um_end);
};

struct TraverseOpt {
    uint32_t k = 31;
    uint32_t min_cov_vertices = 2;
    int out_qual = 1, max_qual = 40;
    double weak_region_len_factor = 0.25, large_k_factor = 1.5, min_score = 0.0;
    bool long_read_correct = false;   // pass 2: exploreSubGraphLong bursts (bounded by k * large_k_factor bases)
    size_t max_len_subpath() const { return static_cast<size_t>(k * large_k_factor); }   // src/GraphTraversal.cpp:594
};

// the GPU services (thin wrappers over the C ABI of this library, batched)
struct AlignJob {
    std::string q, t;
    uint8_t mode;  // 0 NW, 1 SHW, 2 HW
    const std::string* tref = nullptr;   // target shared by several jobs of a request (e.g. the read window): used instead of `t`, not copied
    const std::string& target() const { return tref ? *tref : t; }
};
void gpu_distances(rtk_ctx* ctx, const std::vector<AlignJob>& jobs, std::vector<int32_t>& dist, std::vector<int32_t>& first_end);
// + the largest end column carrying the distance (the end locations are ascending: first = endLocations[0], last = the final one)
void gpu_distances_fl(rtk_ctx* ctx, const std::vector<AlignJob>& jobs, std::vector<int32_t>& dist, std::vector<int32_t>& first_end,
                      std::vector<int32_t>& last_end);
void gpu_paths(rtk_ctx* ctx, const std::vector<AlignJob>& jobs, std::vector<int32_t>& dist, std::vector<std::vector<uint8_t>>& ops);

// selectors: first candidate wins ties (strict <), src/Alignment.cpp
std::pair<int, int> select_best_alignment(rtk_ctx* ctx, const rtk_graph_view& g, const std::vector<GPath>& cands, const std::string& ref);
std::pair<int, int> select_best_prefix_alignment(rtk_ctx* ctx, const rtk_graph_view& g, const std::vector<const GPath*>& cands, const std::string& ref);
std::pair<int, int> select_best_substring_alignment(rtk_ctx* ctx, const rtk_graph_view& g, const std::vector<GPath>& cands, const std::string& ref);

std::vector<GPath> fix_repeats(rtk_ctx* ctx, const rtk_graph_view& g, const TraverseOpt& opt, const std::vector<GPath>& v_path, const std::string& ref);

// all_pids = WeightsPairID::all_pids (sorted).  um_s / um_e: anchors as getSeeds returns them (len = 1).
std::vector<GPath> explore_paths_bfs2(rtk_ctx* ctx, const rtk_graph_view& g, const TraverseOpt& opt, const std::string& ref,
                                      const std::vector<uint32_t>& all_pids, const PNode& um_s, const PNode& um_e);
std::vector<GPath> explore_paths_bfs(rtk_ctx* ctx, const rtk_graph_view& g, const TraverseOpt& opt, const std::string& ref,
                                     const std::vector<uint32_t>& all_pids, const PNode& um_s);

inline char rtk_get_qual(const double score, const size_t qv_min, const size_t qv_max) {  // getQual, src/Common.hpp:410-418
    const char phred_base_std = static_cast<char>(33);
    const char phred_scale_std = static_cast<char>(qv_max);
    const double qv_score = std::min(score, 1.0) * static_cast<double>(phred_scale_std - qv_min);
    return static_cast<char>(qv_score + phred_base_std + qv_min);
}

inline std::pair<size_t, size_t> rtk_min_max_length(const size_t l, const double len_factor) {  // getMinMaxLength, src/Common.hpp:435-438
    return {static_cast<size_t>(std::max(l - (l * len_factor), 1.0)), static_cast<size_t>(std::max(l + (l * len_factor), 1.0))};
}

}  // namespace rtk
