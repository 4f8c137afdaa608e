// correct.cpp — per-read correction: correctSequence (src/Correction.cpp:159-958) with its `correct` and
// `chooseColors` lambdas, extractSemiWeakPaths (:3-157), generateConsensus (src/Alignment.cpp:309-470),
// fixAmbiguity (:527-844) and getAmbiguityVector (src/GraphTraversal.cpp:966-1055), restated over the flat
// graph.  This is the order-defining host logic of the reference; every k-mer lookup sweep,
// graph burst and alignment it needs is executed by the kernels of this library through rtk_get_seeds
// (K1), explore_paths_bfs* (K2/K3/K4/K5, traverse.cpp), rtk_edlib_batch (K4) and rtk_edlib_path_batch (K5).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "kmer.cuh"
#include "lookup.cuh"
#include "rtk_host_common.hpp"
#include "traverse.hpp"
#include <chrono>

#include "broker.hpp"
#include "subgraph_host.hpp"   // RTK_DFS_MAX_NODES

namespace rtk {

namespace {

typedef rtk_u128 KW;
typedef std::vector<uint32_t> IdSet;  // sorted, unique (PairID)

// ------------------------------------------------------------------ small helpers
inline bool is_dna(char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'a' || c == 'c' || c == 'g' || c == 't'; }
const char ambiguity_c[16] = {'.', 'A', 'C', 'M', 'G', 'R', 'S', 'V', 'T', 'W', 'Y', 'H', 'K', 'D', 'B', 'N'};  // src/Common.hpp:260
struct AmbTables {   // amb_index / reverse complement of every byte, built once (the linear scans were 3 % of the host time)
    uint8_t idx[256];
    char rc[256];
    AmbTables() {
        for (int b = 0; b < 256; ++b) {
            const char c = (char)(b & 0xDF);
            uint8_t i = 0;
            for (uint8_t x = 0; x < 16; ++x) if (ambiguity_c[x] == c) { i = x; break; }
            idx[b] = i;
            // reverse_complement(char), Bifrost/src/Common.hpp:61-90: IUPAC-aware; anything else is returned unchanged
            if (ambiguity_c[i] != c || i == 0) rc[b] = (char)b;
            else rc[b] = ambiguity_c[(uint8_t)(((i & 1) << 3) | ((i & 2) << 1) | ((i & 4) >> 1) | ((i & 8) >> 3))];
        }
    }
};
static const AmbTables g_amb;
inline uint8_t amb_index(char c) { return g_amb.idx[(unsigned char)c]; }
inline char rc_char(char c) { return g_amb.rc[(unsigned char)c]; }
std::string rc_string(const std::string& s) {
    std::string r(s.rbegin(), s.rend());
    for (auto& c : r) c = rc_char(c);
    return r;
}

inline uint32_t usize(const rtk_graph_view& g, uint32_t u) { return (uint32_t)(g.unitig_off[u + 1] - g.unitig_off[u]); }
inline bool is_branching(const rtk_graph_view& g, uint32_t u) { return (g.kmcov[u] >> 63) & 1ULL; }
inline bool is_short_cycle(const rtk_graph_view& g, uint32_t u) { return (g.shared[u] & 0x100ULL) != 0; }
inline double kmer_coverage(const rtk_graph_view& g, uint32_t u) {  // UnitigData::getKmerCoverage, src/UnitigData.hpp:396-399
    const uint64_t w = g.kmcov[u];
    const double cov = (double)((w & 0x7fffffffULL) + ((w >> 31) & 0x7fffffffULL));
    return std::round(cov / (double)(usize(g, u) - g.k + 1));
}

struct Span { const uint32_t* p; uint64_t n; };
inline Span gl_ids(const rtk_graph_view& g, uint32_t u) {
    const uint32_t gs = g.gset_of[u];
    if (gs == RTK_NONE32) return {nullptr, 0};
    return {g.gset_ids + g.gset_off[gs], g.gset_off[gs + 1] - g.gset_off[gs]};
}
inline Span lo_ids(const rtk_graph_view& g, uint32_t u) { return {g.loc_ids + g.loc_off[u], g.loc_off[u + 1] - g.loc_off[u]}; }
inline uint64_t spid_card(const rtk_graph_view& g, uint32_t u) { return gl_ids(g, u).n + lo_ids(g, u).n; }

IdSet set_union(const IdSet& a, const IdSet& b) { IdSet r; std::set_union(a.begin(), a.end(), b.begin(), b.end(), std::back_inserter(r)); return r; }
IdSet set_inter(const IdSet& a, const IdSet& b) { IdSet r; std::set_intersection(a.begin(), a.end(), b.begin(), b.end(), std::back_inserter(r)); return r; }
IdSet set_diff(const IdSet& a, const IdSet& b) { IdSet r; std::set_difference(a.begin(), a.end(), b.begin(), b.end(), std::back_inserter(r)); return r; }
IdSet span_set(const Span& s) { return IdSet(s.p, s.p + s.n); }
uint64_t inter_capped(const uint32_t* a, uint64_t na, const uint32_t* b, uint64_t nb, uint64_t cap) {
    uint64_t i = 0, j = 0, c = 0;
    while (i < na && j < nb && c < cap) { if (a[i] == b[j]) { ++c; ++i; ++j; } else if (a[i] < b[j]) ++i; else ++j; }
    return c;
}
// min(getNumberSharedPairID(spid(u), b, cap), cap)
uint64_t shared_with(const rtk_graph_view& g, uint32_t u, const IdSet& b, uint64_t cap) {
    const Span gs = gl_ids(g, u), ls = lo_ids(g, u);
    uint64_t c = inter_capped(gs.p, gs.n, b.data(), b.size(), cap);
    if (c < cap) c += inter_capped(ls.p, ls.n, b.data(), b.size(), cap - c);
    return c;
}

// ambiguity characters of a mapping, UnitigData::get_ambiguity_char(um) (src/UnitigData.hpp:455-478)
std::vector<std::pair<size_t, char>> ambiguity_of(const rtk_graph_view& g, const PNode& um) {
    std::vector<std::pair<size_t, char>> all, v;
    for (uint64_t i = g.amb_off[um.unitig]; i < g.amb_off[um.unitig + 1]; ++i) all.push_back({g.amb_ids[i] >> 4, ambiguity_c[g.amb_ids[i] & 0xF]});
    std::sort(all.begin(), all.end());
    const size_t sz = (size_t)um.len + g.k - 1, end = um.dist + sz;
    if (um.strand) { for (const auto& p : all) if (p.first >= um.dist && p.first < end) v.push_back({p.first - um.dist, p.second}); }
    else { for (auto it = all.rbegin(); it != all.rend(); ++it) if (it->first >= um.dist && it->first < end) v.push_back({sz - (it->first - um.dist) - 1, rc_char(it->second)}); }
    return v;
}

// getAmbiguityVector(v_um, k), src/GraphTraversal.cpp:966-1035
std::vector<std::pair<size_t, char>> ambiguity_vector(const rtk_graph_view& g, const std::vector<PNode>& v_um) {
    const size_t k = g.k;
    std::vector<std::pair<size_t, char>> v_amb;
    size_t prev_l = 0, pos_prev_l = 0;
    for (const PNode& um : v_um) {
        const std::vector<std::pair<size_t, char>> v_amb_um = ambiguity_of(g, um);
        std::vector<std::pair<size_t, char>> tmp;
        size_t ip = pos_prev_l, ic = 0;
        while (ip < v_amb.size() && ic < v_amb_um.size() && v_amb_um[ic].first < k - 1) {
            const size_t cpos = v_amb_um[ic].first + prev_l;
            if (v_amb[ip].first < cpos) tmp.push_back(v_amb[ip++]);
            else if (v_amb[ip].first > cpos) { tmp.push_back({cpos, v_amb_um[ic].second}); ++ic; }
            else { tmp.push_back({v_amb[ip].first, ambiguity_c[amb_index(v_amb[ip].second) | amb_index(v_amb_um[ic].second)]}); ++ip; ++ic; }
        }
        while (ip < v_amb.size()) tmp.push_back(v_amb[ip++]);
        while (ic < v_amb_um.size()) { tmp.push_back({v_amb_um[ic].first + prev_l, v_amb_um[ic].second}); ++ic; }
        prev_l += um.len;
        v_amb.erase(v_amb.begin() + pos_prev_l, v_amb.end());
        for (auto& p : tmp) { v_amb.push_back(p); pos_prev_l += (size_t)(v_amb.back().first < prev_l); }
    }
    return v_amb;
}

// const_UnitigMap CompactedDBG::findUnitig(s, pos, len) on the host mirror (lookup + extension along the unitig)
struct HostMatch { bool found = false; uint32_t unitig = 0, dist = 0, len = 0, strand = 0; };
HostMatch find_unitig_host(const rtk_graph_view& g, const std::string& s, size_t pos) {
    HostMatch m;
    const size_t k = g.k;
    if (s.size() < k || pos > s.size() - k) return m;
    KW fw = 0, rc = 0;
    for (size_t i = 0; i < k; ++i) {
        const uint32_t c = rtk_base_code((char)(s[pos + i] & 0xDF));
        if (c > 3) return m;
        fw = (fw << 2) | (KW)c;
        rc = (rc >> 2) | ((KW)(3 - c) << (2 * (k - 1)));
    }
    rtk_kmer_hit h;
    if (!rtk_lookup<KW>(g.table, g.n_buckets, g.pool, (int)k, fw, rc, h)) return m;
    m.found = true;
    m.unitig = rtk_unitig_of(g.blk2unitig, g.unitig_off, h.P);
    m.dist = (uint32_t)(h.P - g.unitig_off[m.unitig]);
    m.strand = h.strand;
    m.len = 1;
    const uint32_t usz = usize(g, m.unitig);
    if (usz == k) return m;
    const uint64_t ub = g.unitig_off[m.unitig];
    if (m.strand) {
        size_t up = m.dist + k, sp = pos + k;
        while (sp < s.size() && up < usz && s[sp] == "ACGT"[rtk_pool_base(g.pool, ub + up)]) { ++sp; ++up; ++m.len; }
    } else {
        size_t sp = pos + k;
        int64_t up = (int64_t)m.dist - 1;
        while (sp < s.size() && up >= 0 && s[sp] == "ACGT"[3 - rtk_pool_base(g.pool, ub + (uint64_t)up)]) { ++sp; --up; ++m.len; }
        m.dist -= (m.len - 1);
    }
    return m;
}

// next position >= p whose k-mer is all-ACGT (KmerIterator); npos if none
size_t next_kmer(const std::string& s, size_t p, size_t k) {
    while (p + k <= s.size()) {
        size_t bad = std::string::npos;
        for (size_t i = p + k; i-- > p;) if (!is_dna(s[i])) { bad = i; break; }
        if (bad == std::string::npos) return p;
        p = bad + 1;
    }
    return std::string::npos;
}

// ------------------------------------------------------------------ ResultCorrection (src/ResultCorrection.hpp)
// pos_corrected_old_seq is a set of positions in the reference (one Roaring / std::set entry per base); here a sorted
// vector of unique positions with the same observable behaviour (size, ordered scan, lower_bound)
struct ResultCorrection {
    std::vector<uint32_t> pos;  // pos_corrected_old_seq, ascending, unique
    std::string seq, qual;
    IdSet all_pids;          // WeightsPairID::all_pids (the only part that influences results)
    size_t old_seq_len;
    bool is_corrected = false;
    explicit ResultCorrection(size_t n) : old_seq_len(n) {}
    void add_range(uint64_t a, uint64_t b) {
        if (b <= a) return;
        if (pos.empty() || pos.back() < (uint32_t)a) {   // the usual case: regions are emitted left to right
            pos.reserve(pos.size() + (size_t)(b - a));
            for (uint64_t x = a; x < b; ++x) pos.push_back((uint32_t)x);
            return;
        }
        std::vector<uint32_t> add, merged;
        add.reserve((size_t)(b - a));
        for (uint64_t x = a; x < b; ++x) add.push_back((uint32_t)x);
        std::sort(add.begin(), add.end());   // a 32-bit wrap of the range keeps set semantics
        add.erase(std::unique(add.begin(), add.end()), add.end());
        merged.reserve(pos.size() + add.size());
        std::set_union(pos.begin(), pos.end(), add.begin(), add.end(), std::back_inserter(merged));
        pos.swap(merged);
    }
    size_t nb_corrected() const { return pos.size(); }
    ResultCorrection& reverse_complement() {
        if (seq.length() != 0) {
            for (uint32_t& p : pos) p = (uint32_t)(old_seq_len - p - 1);
            std::sort(pos.begin(), pos.end());   // a plain reversal unless a position lies beyond old_seq_len (wraps, as in the set)
            seq = rc_string(seq);
            std::reverse(qual.begin(), qual.end());
        }
        return *this;
    }
    size_t len_corrected_region(size_t p) const {
        size_t next = p;
        for (auto it = std::lower_bound(pos.begin(), pos.end(), (uint32_t)p); it != pos.end() && *it < old_seq_len && *it == next; ++it) ++next;
        return next - p;
    }
    size_t len_uncorrected_region(size_t p) const {
        if (p >= old_seq_len) return 0;
        auto it = std::lower_bound(pos.begin(), pos.end(), (uint32_t)p);
        if (it == pos.end()) return old_seq_len - p;
        return std::min<size_t>(*it, old_seq_len) - p;
    }
};

// CIGAR-free equivalents: the reference converts edlib's op list to a standard CIGAR (M = match or mismatch,
// I = op 1, D = op 2) and walks it run by run; here runs are formed on the fly.
struct Run { char op; size_t len; };
std::vector<Run> runs_of(const std::vector<uint8_t>& ops) {
    std::vector<Run> r;
    for (uint8_t o : ops) {
        const char c = (o == 1) ? 'I' : (o == 2) ? 'D' : 'M';
        if (!r.empty() && r.back().op == c) ++r.back().len; else r.push_back({c, 1});
    }
    return r;
}

struct Ctx {
    rtk_ctx* ctx;
    const rtk_graph_view& g;
    const rtk_opt& opt;
    TraverseOpt topt;
    bool pass2;
    size_t max_km_cov;
    bool device_regions = true;   // extractSemiWeakPaths through the device-resident region engine (region.cuh)
};

// ------------------------------------------------------------------ chooseColors (src/Correction.cpp:215-429)
struct ColorSide { std::vector<std::pair<uint32_t, bool>> v; };  // (unitig, !isBranching), first occurrence order; keyed by unitig
bool side_insert(ColorSide& s, uint32_t u, bool nb) {
    for (const auto& p : s.v) if (p.first == u) return false;
    s.v.push_back({u, nb});
    return true;
}

IdSet choose_colors(const Ctx& C, const ColorSide& s_pid_s, const ColorSide& s_pid_e, const ColorSide& s_pid_w) {
    const rtk_graph_view& g = C.g;
    const ColorSide* v_s_pid[3] = {&s_pid_w, &s_pid_e, &s_pid_s};
    std::set<uint32_t> s_spid;
    IdSet a_pid[6];
    for (size_t i = 0; i < 3; ++i) {
        for (const auto& it : v_s_pid[i]->v) {
            const int shift = (int)i + (it.second ? 3 : 0);
            const Span gs = gl_ids(g, it.first);
            a_pid[shift] = set_union(a_pid[shift], gs.p ? span_set(gs) : span_set(lo_ids(g, it.first)));
            if (spid_card(g, it.first) >= C.opt.min_cov_vertices) s_spid.insert(it.first);
        }
    }
    IdSet all_pids;
    const IdSet pos0 = set_union(a_pid[0], a_pid[3]), pos1 = set_union(a_pid[1], a_pid[4]), pos2 = set_union(a_pid[2], a_pid[5]);
    const IdSet a01 = set_inter(pos0, pos1), a12 = set_inter(pos1, pos2), a02 = set_inter(pos0, pos2);
    IdSet inter2, inter3, branch;
    IdSet nobranch = set_union(set_union(a_pid[3], a_pid[4]), a_pid[5]);
    const IdSet nobranch_cpy = nobranch;
    const size_t cov = 30;
    size_t nb_unselected = s_spid.size();
    IdSet a2[6];
    struct Sel { uint32_t u; int quota; uint64_t card; };
    std::vector<Sel> v_spids;
    for (uint32_t u : s_spid) v_spids.push_back({u, (int)std::min<uint64_t>(cov, spid_card(g, u)), spid_card(g, u)});
    // the reference sorts pointers by cardinality (ties in pointer-hash order, which does not reproduce); ties by unitig id
    std::stable_sort(v_spids.begin(), v_spids.end(), [](const Sel& a, const Sel& b) { return a.card < b.card; });
    for (int i = 5; i >= 0; --i) {
        if (nb_unselected == 0) break;
        if (i == 5) { inter3 = set_inter(a01, a12); a2[5] = set_inter(nobranch, inter3); }
        else if (i == 4) { inter2 = set_union(set_union(a01, a12), a02); nobranch = set_diff(nobranch, a2[5]); a2[4] = set_inter(nobranch, inter2); }
        else if (i == 3) { nobranch = set_diff(nobranch, a2[4]); a2[3] = nobranch; nobranch.clear(); }
        else if (i == 2) { branch = set_diff(set_union(set_union(a_pid[0], a_pid[1]), a_pid[2]), nobranch_cpy); a2[2] = set_inter(branch, inter3); }
        else if (i == 1) { branch = set_diff(branch, a2[2]); a2[1] = set_inter(branch, inter2); }
        else { branch = set_diff(branch, a2[1]); a2[0] = branch; branch.clear(); }
        if (!a2[i].empty()) {
            nb_unselected = 0;
            IdSet curr = a2[i];
            for (auto& sp : v_spids) {
                if (sp.quota > 0 && (i == 0 || shared_with(g, sp.u, curr, 1) >= 1)) {
                    const uint64_t min_cov = std::min<uint64_t>(cov, sp.card);
                    sp.quota = (int)(min_cov - std::min<uint64_t>(shared_with(g, sp.u, all_pids, min_cov), min_cov));
                    if (sp.quota > 0) {
                        const size_t all_card = all_pids.size();
                        IdSet pid = set_union(set_inter(span_set(gl_ids(g, sp.u)), curr), set_inter(span_set(lo_ids(g, sp.u)), curr));
                        if (pid.size() > (size_t)sp.quota) pid.resize((size_t)sp.quota);
                        all_pids = set_union(all_pids, pid);
                        curr = set_diff(curr, pid);
                        sp.quota -= std::min((int)(all_pids.size() - all_card), sp.quota);
                    }
                }
                nb_unselected += (size_t)(sp.quota > 0);
            }
        }
    }
    return all_pids;
}

// ------------------------------------------------------------------ extractSemiWeakPaths (src/Correction.cpp:3-157)
typedef std::pair<std::vector<GPath>, std::vector<GPath>> PathPair;
inline PNode node_of(const rtk_hit& h) { PNode n; n.unitig = h.unitig; n.strand = h.strand; n.dist = h.dist; n.len = 1; return n; }

// mappedSequenceToString of a path's last vertex, as a comparable key
std::string back_string(const rtk_graph_view& g, const GPath& p) {
    GPath t; t.v.push_back(p.back());
    return t.to_string(g);
}
inline bool same_node(const PNode& a, const PNode& b) { return a.unitig == b.unitig && a.strand == b.strand && a.dist == b.dist && a.len == b.len; }

PathPair extract_semi_weak_paths(const Ctx& C, const std::string& s, const IdSet& all_pids, const rtk_hit& um_start, bool has_end,
                                 const rtk_hit& um_end, size_t end_pos, const std::vector<rtk_hit>& v_w, size_t i_weak) {
    const rtk_graph_view& g = C.g;
    const size_t k = g.k;
    PathPair paths;
    std::vector<std::pair<GPath, size_t>> paths1, paths2;
    const bool no_end = !has_end;
    const size_t pos_um_solid2 = no_end ? s.length() - k : end_pos;
    const size_t len_weak_region = (pos_um_solid2 - um_start.pos) + k;
    const size_t max_len_weak_region = C.pass2 ? C.opt.max_len_weak_region2 : C.opt.max_len_weak_region1;
    const size_t max_paths = 512;
    size_t next_weak_pos = 0;
    bool begin = true, end = false;
    {
        GPath tmp;
        tmp.extend(g, node_of(um_start), std::string(1 + k - 1, rtk_get_qual(1.0, 0, C.opt.max_qual)));
        paths1.push_back({tmp, um_start.pos});
    }
    while (i_weak < v_w.size() && v_w[i_weak].pos < um_start.pos) ++i_weak;
    if (i_weak < v_w.size()) next_weak_pos = std::max<size_t>(v_w[i_weak].pos, um_start.pos + k);
    while (!paths1.empty() && !end) {
        std::vector<GPath> g_prev;
        bool g_prev_ok = false;
        if (i_weak < v_w.size()) {
            while (i_weak < v_w.size() && v_w[i_weak].pos < (pos_um_solid2 - k) && v_w[i_weak].pos < next_weak_pos) ++i_weak;
        } else i_weak = v_w.size();
        // customSort: by the mapped sequence of the last vertex (std::sort in the reference; ties are identical vertices or
        // vertices spelling the same string, which are processed identically, so a stable sort is equivalent)
        std::vector<std::string> keys(paths1.size());
        for (size_t i = 0; i < paths1.size(); ++i) keys[i] = back_string(g, paths1[i].first);
        std::vector<size_t> ord(paths1.size());
        for (size_t i = 0; i < ord.size(); ++i) ord[i] = i;
        std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return keys[a] < keys[b]; });
        { std::vector<std::pair<GPath, size_t>> t; for (size_t i : ord) t.push_back(std::move(paths1[i])); paths1.swap(t); }
        end = (i_weak == v_w.size()) || (v_w[i_weak].pos >= (pos_um_solid2 - k));
        for (size_t i = 0; i < paths1.size(); ++i) {
            const std::pair<GPath, size_t>& p = paths1[i];
            const size_t target_pos = end ? pos_um_solid2 : v_w[i_weak].pos;
            const size_t l_len = (target_pos - p.second) + k;
            if (i == 0 || !same_node(paths1[i].first.back(), paths1[i - 1].first.back())) {
                g_prev.clear(); g_prev_ok = false;
                const PNode um_s = begin ? node_of(um_start) : p.first.back();
                const std::string ref = s.substr(p.second, l_len);
                if (end) {
                    if (no_end) {
                        if (l_len <= (max_len_weak_region / 2)) { g_prev = explore_paths_bfs(C.ctx, g, C.topt, ref, all_pids, um_s); g_prev_ok = true; }
                    } else if (l_len <= max_len_weak_region) { g_prev = explore_paths_bfs2(C.ctx, g, C.topt, ref, all_pids, um_s, node_of(um_end)); g_prev_ok = true; }
                } else if (l_len <= max_len_weak_region) { g_prev = explore_paths_bfs2(C.ctx, g, C.topt, ref, all_pids, um_s, node_of(v_w[i_weak])); g_prev_ok = true; }
            }
            if (g_prev_ok && !g_prev.empty()) {
                for (const auto& pt : g_prev) { GPath t = p.first; t.merge(g, pt); paths2.push_back({std::move(t), target_pos}); }
            } else paths.second.push_back(p.first);
        }
        if (!end) next_weak_pos = v_w[i_weak].pos + k;
        begin = false;
        paths1 = std::move(paths2);
        paths2.clear();
        if (!end && paths1.size() > max_paths) {
            std::vector<const GPath*> ptr;
            for (const auto& p : paths1) ptr.push_back(&p.first);
            const int best = select_best_prefix_alignment(C.ctx, g, ptr, s.substr(um_start.pos, len_weak_region)).first;
            paths2.push_back(paths1[(size_t)best]);
            paths1 = std::move(paths2);
            paths2.clear();
        }
    }
    for (auto& p : paths1) paths.first.push_back(std::move(p.first));
    return paths;
}

// extractSemiWeakPaths as ONE request to the device-resident region engine; calls the engine declines (short-cycle unitigs ->
// fixRepeats, list collapses, alignments above edlib's traceback switch, scratch capacity) run through the request-at-a-time
// orchestration above, which yields the same bytes
PathPair region_paths(const Ctx& C, const std::string& s, const IdSet& all_pids, const rtk_hit& um_start, bool has_end, const rtk_hit& um_end,
                      size_t end_pos, const std::vector<rtk_hit>& v_w, size_t i_weak) {
    GpuBroker* b = C.device_regions ? current_broker() : nullptr;
    if (b) {
        RegionReq r;
        r.opt = &C.opt; r.pass = C.pass2 ? 2 : 1;
        r.s = &s; r.um_start = um_start; r.um_end = um_end; r.has_end = has_end; r.end_pos = end_pos;
        r.v_w = &v_w; r.i_weak = i_weak; r.pids = &all_pids;
        b->submit(&r);
        if (r.status != 2 && r.segs.size() == 1) {
            PathPair pp;
            (r.status == 0 ? pp.first : pp.second).push_back(std::move(r.segs[0].path));
            return pp;
        }
    }
    return extract_semi_weak_paths(C, s, all_pids, um_start, has_end, um_end, end_pos, v_w, i_weak);
}

// selectBestPrefixAlignment(ref, len, vector<Path>, cut_threshold) (src/Alignment.cpp:47-97): {-1,-1} above the threshold
std::pair<int, int> select_prefix_cut(const Ctx& C, const std::vector<GPath>& cands, const std::string& ref, double cut) {
    std::vector<AlignJob> jobs(cands.size());
    for (size_t i = 0; i < cands.size(); ++i) { jobs[i].q = cands[i].to_string(C.g); jobs[i].tref = &ref; jobs[i].mode = 1; }
    std::vector<int32_t> dist, fe;
    gpu_distances(C.ctx, jobs, dist, fe);
    double best = 0.0; int id = -1, endl = -1;
    for (size_t i = 0; i < cands.size(); ++i) {
        const double d = static_cast<double>(dist[i]) / jobs[i].q.length();
        if (i == 0 || (dist[i] >= 0 && d < best)) { best = d; id = (int)i; endl = fe[i]; }
    }
    if (cut > 0.0 && best > cut) return {-1, -1};
    return {id, endl};
}

// ------------------------------------------------------------------ fixAmbiguity (src/Alignment.cpp:527-844), hap_id undetermined
// In two steps around its one alignment (SHW + path of the query with the ambiguity codes filled in against the read window), so
// that a caller can batch the alignments of several regions into one GPU request.
struct FixAmbiguity {
    std::string query_tmp;
    std::unordered_map<size_t, char> safe, all;
    bool active = false;
    // step 1: returns true and fills `job` when an alignment is needed
    bool begin(const Ctx& C, const std::string& query, const std::string& quality, const char* ref_seq, size_t ref_len,
               const std::vector<std::pair<size_t, char>>& v_ambiguity, AlignJob& job) {
        active = false;
        if (v_ambiguity.empty()) return false;
        const char q_min_conf_corr = rtk_get_qual(C.opt.min_confidence_snp_corr, 0, C.opt.max_qual);
        query_tmp = query;
        safe.clear(); all.clear();
        for (const auto& p : v_ambiguity) {
            if (quality[p.first] < q_min_conf_corr) { safe.insert(p); query_tmp[p.first] = p.second; }
        }
        all = safe;
        job.q = query_tmp.substr(0, query.length()); job.t = std::string(ref_seq, ref_len); job.mode = 1; job.tref = nullptr;
        active = true;
        return true;
    }
    void finish(const Ctx& C, std::string& query, std::string& quality, const char* ref_seq, const std::vector<uint8_t>& ops0);
};

void FixAmbiguity::finish(const Ctx& C, std::string& query, std::string& quality, const char* ref_seq, const std::vector<uint8_t>& ops0) {
    if (!active) return;
    const rtk_graph_view& g = C.g;
    const size_t query_len = query.length(), k = g.k;
    const char q_max_corr = rtk_get_qual(1.0, C.opt.out_qual, C.opt.max_qual);
    const char q_min_corr = rtk_get_qual(0.0, C.opt.out_qual, C.opt.max_qual);
    const char q_min_conf_corr = rtk_get_qual(C.opt.min_confidence_snp_corr, 0, C.opt.max_qual);
    const char c_noCorrect = 'X';
    std::vector<std::vector<uint8_t>> ops(1, ops0);
    size_t query_pos = 0, target_pos = 0;  // SHW: startLocations[0] == 0
    auto rev = [](char c, bool* a) { const uint8_t i = amb_index(c); a[0] = i & 1; a[1] = i & 2; a[2] = i & 4; a[3] = i & 8; };
    for (const Run& r : runs_of(ops[0])) {
        if (r.op == 'M') {
            for (size_t q_pos = query_pos, t_pos = target_pos; q_pos < query_pos + r.len; ++q_pos, ++t_pos) {
                if (!is_dna(query_tmp[q_pos])) {
                    if (!is_dna(ref_seq[t_pos])) { auto it = safe.find(q_pos); if (it != safe.end()) it->second = c_noCorrect; }
                    else if (quality[q_pos] >= q_min_corr) {
                        bool aq[4], at[4];
                        rev(ref_seq[t_pos], at); rev(query_tmp[q_pos], aq);
                        if ((aq[0] && at[0]) || (aq[1] && at[1]) || (aq[2] && at[2]) || (aq[3] && at[3])) { auto it = safe.find(q_pos); if (it != safe.end()) it->second = ref_seq[t_pos]; }
                    }
                    auto it = all.find(q_pos);
                    if (it != all.end()) it->second = ref_seq[t_pos];
                } else if (!is_dna(ref_seq[t_pos])) {
                    if (quality[q_pos] < q_min_conf_corr) { safe.insert({q_pos, c_noCorrect}); all.insert({q_pos, ref_seq[t_pos]}); }
                    else {
                        bool aq[4], at[4];
                        rev(ref_seq[t_pos], at); rev(query_tmp[q_pos], aq);
                        if (!(aq[0] && at[0]) && !(aq[1] && at[1]) && !(aq[2] && at[2]) && !(aq[3] && at[3])) { safe.insert({q_pos, c_noCorrect}); all.insert({q_pos, ref_seq[t_pos]}); }
                    }
                }
            }
            query_pos += r.len; target_pos += r.len;
        } else if (r.op == 'I') {
            for (size_t q_pos = query_pos; q_pos < query_pos + r.len; ++q_pos) {
                if (!is_dna(query_tmp[q_pos])) {
                    auto is = safe.find(q_pos); auto ia = all.find(q_pos);
                    if (is != safe.end() && ia != all.end()) { ia->second = is->second; is->second = c_noCorrect; }
                }
            }
            query_pos += r.len;
        } else target_pos += r.len;
    }
    std::set<std::pair<size_t, char>> s_amb;
    for (const auto& p : safe) {
        if (!is_dna(p.second)) continue;
        const size_t pos_buff = (p.first < (k - 1)) ? 0 : (p.first - k + 1);
        const size_t len_buff = std::min(p.first + k, query_len) - pos_buff;
        const size_t pos_snp_buff = p.first - pos_buff;
        std::string q_sub = query.substr(pos_buff, len_buff);
        q_sub[pos_snp_buff] = p.second;
        for (size_t kp = next_kmer(q_sub, 0, k); kp != std::string::npos;) {
            const HostMatch um = find_unitig_host(g, q_sub, kp);
            size_t adv = 1;
            if (um.found) {
                PNode full; full.unitig = um.unitig; full.strand = um.strand; full.dist = 0; full.len = usize(g, um.unitig) - (uint32_t)k + 1;
                GPath t; t.v.push_back(full);
                const std::string unitig_seq = t.to_string(g);
                const std::vector<std::pair<size_t, char>> v_amb = ambiguity_of(g, full);
                size_t pos_snp_unitig = (pos_snp_buff - kp) + um.dist;
                if (!um.strand) pos_snp_unitig = usize(g, um.unitig) - pos_snp_unitig - 1;
                for (const auto& pa : v_amb) {
                    int64_t pos = (int64_t)pa.first;
                    if ((size_t)pos <= pos_snp_unitig) pos = (int64_t)p.first - (int64_t)(pos_snp_unitig - (size_t)pos);
                    else pos = (int64_t)p.first + (int64_t)((size_t)pos - pos_snp_unitig);
                    if (pos >= 0 && (size_t)pos < query_len && (size_t)pos != p.first) {
                        const auto it = safe.find((size_t)pos);
                        if (it != safe.end() && !is_dna(it->second)) s_amb.insert({(size_t)pos, unitig_seq[pa.first]});
                    }
                }
                adv = um.len;
            }
            kp = next_kmer(q_sub, kp + adv, k);
        }
    }
    {
        const std::vector<std::pair<size_t, char>> v(s_amb.begin(), s_amb.end());
        for (int64_t i = 0; i < (int64_t)v.size(); ++i) {
            if ((i == 0 || v[i].first != v[i - 1].first) && (i == (int64_t)v.size() - 1 || v[i].first != v[i + 1].first)) {
                auto is = safe.find(v[i].first);
                if (is != safe.end()) { if (amb_index(is->second) & amb_index(v[i].second)) is->second = v[i].second; }
            }
        }
    }
    for (const auto& p : safe) {
        if (p.second == c_noCorrect || quality[p.first] < q_min_corr) {
            const auto ia = all.find(p.first);
            if (ia != all.end()) { query_tmp[p.first] = ia->second; quality[p.first] = q_max_corr; }  // hap_id undetermined => validHap
        } else if (!is_dna(p.second)) query_tmp[p.first] = query[p.first];
        else query_tmp[p.first] = p.second;
    }
    query = std::move(query_tmp);
}

// ------------------------------------------------------------------ generateConsensus (src/Alignment.cpp:309-470)
std::pair<std::string, std::string> generate_consensus(const Ctx& C, const ResultCorrection* fw_s, const ResultCorrection* bw_s,
                                                        const std::string& ref_seq, double max_norm) {
    if (bw_s->nb_corrected() == 0 && fw_s->nb_corrected() != 0) return {fw_s->seq, fw_s->qual};
    else if (fw_s->nb_corrected() == 0 && bw_s->nb_corrected() != 0) return {bw_s->seq, bw_s->qual};
    else if (fw_s->nb_corrected() + bw_s->nb_corrected() == 0) return {std::string(), std::string()};
    if (bw_s->nb_corrected() > fw_s->nb_corrected()) std::swap(fw_s, bw_s);
    std::vector<AlignJob> j(2);
    j[0].q = fw_s->seq; j[0].tref = &ref_seq; j[0].mode = 0;
    j[1].q = bw_s->seq; j[1].tref = &ref_seq; j[1].mode = 0;
    std::vector<int32_t> d;
    std::vector<std::vector<uint8_t>> ops;
    gpu_paths(C.ctx, j, d, ops);
    const double n_fw = static_cast<double>(d[0]) / std::max(fw_s->seq.length(), ref_seq.length());
    const double n_bw = static_cast<double>(d[1]) / std::max(bw_s->seq.length(), ref_seq.length());
    if (max_norm > 0.0 && (n_fw > max_norm || n_bw > max_norm)) {
        if (n_fw > max_norm && n_bw > max_norm) return {std::string(), std::string()};
        if (n_fw > max_norm) return {bw_s->seq, bw_s->qual};
        return {fw_s->seq, fw_s->qual};
    }
    struct Cur { std::vector<Run> runs; size_t ri = 0, used = 0, qpos = 0, rpos = 0; };  // `used`: nothing partially consumed - runs are atomic like CIGAR tokens
    Cur cf, cb;
    cf.runs = runs_of(ops[0]); cb.runs = runs_of(ops[1]);
    // moveIntoCIGAR (:349-414): tokens are consumed whole; an M token spanning a boundary is NOT consumed
    auto move_into = [](size_t start, size_t end, Cur& c) -> std::pair<std::pair<size_t, size_t>, size_t> {
        size_t read_start = c.qpos, read_end = c.qpos;
        while (c.ri != c.runs.size() && c.rpos < start) {
            const Run& r = c.runs[c.ri];
            if (r.op == 'M') {
                if (c.rpos + r.len > start) { read_start = c.qpos + (start - c.rpos); break; }
                c.qpos += r.len; c.rpos += r.len;
            } else if (r.op == 'I') c.qpos += r.len;
            else c.rpos += r.len;
            read_start = c.qpos;
            ++c.ri;
        }
        read_end = read_start;
        while (c.ri != c.runs.size() && c.rpos < end) {
            const Run& r = c.runs[c.ri];
            if (r.op == 'M') {
                if (c.rpos + r.len > end) return {{read_start, c.qpos + (end - c.rpos)}, end};
                c.qpos += r.len; c.rpos += r.len;
            } else if (r.op == 'I') c.qpos += r.len;
            else c.rpos += r.len;
            read_end = c.qpos;
            ++c.ri;
        }
        return {{read_start, read_end}, c.rpos};
    };
    std::string ss, sq;
    size_t i = 0;
    while (i < ref_seq.length()) {
        int64_t len_fw = (int64_t)fw_s->len_corrected_region(i);
        int64_t len_bw = (int64_t)bw_s->len_corrected_region(i);
        std::pair<std::pair<size_t, size_t>, size_t> pr;
        if ((len_fw + len_bw) <= 0) {
            len_fw = (int64_t)fw_s->len_uncorrected_region(i);
            len_bw = (int64_t)bw_s->len_uncorrected_region(i);
            if (len_fw > len_bw || len_fw <= 0) len_fw = -1;
            else len_bw = -1;
        }
        if (len_fw >= len_bw) {
            pr = move_into(i, i + (size_t)len_fw, cf);
            if (pr.first.second > pr.first.first) { ss += fw_s->seq.substr(pr.first.first, pr.first.second - pr.first.first); sq += fw_s->qual.substr(pr.first.first, pr.first.second - pr.first.first); }
        } else {
            pr = move_into(i, i + (size_t)len_bw, cb);
            if (pr.first.second > pr.first.first) { ss += bw_s->seq.substr(pr.first.first, pr.first.second - pr.first.first); sq += bw_s->qual.substr(pr.first.first, pr.first.second - pr.first.first); }
        }
        i = pr.second;
    }
    if (max_norm > 0.0) {
        // edlibDefaultAlignConfig(): NW distance WITHOUT the IUPAC equalities (src/edlib.cpp:1474-1476)
        std::vector<AlignJob> jj(1);
        jj[0].q = ss; jj[0].t = ref_seq; jj[0].mode = 0 | 4;  // bit 2: no additional equalities
        std::vector<int32_t> dd, fe;
        gpu_distances(C.ctx, jj, dd, fe);
        const double n = static_cast<double>(dd[0]) / std::max(ss.length(), ref_seq.length());
        if (n > max_norm) return {fw_s->seq, fw_s->qual};
    }
    return {ss, sq};
}

// ------------------------------------------------------------------ the `correct` lambda (src/Correction.cpp:431-753)
// One call of the lambda as a job in stages, each stage ending where the reference needs a GPU answer:
//   prepare()  colour selection (chooseColors) and the region's weak anchors            -> `req`, one region-engine request
//   paths()    the path phase on the engine's segments (or, declined, request by request) -> s_corrected / q_corrected / ranges
//   amb_begin() / amb_finish()   fixAmbiguity around its one SHW path alignment
//   trim_begin() / trim_finish() the SHW cut of a not fully corrected region
// A caller may run the stages of several jobs side by side and batch their alignments (run_piece does so for the forward
// and backward attempts of a gap); correct_region() runs one job start to end.
struct InconsistentSegments {};   // thrown when the engine's segments do not replay like the host loop (never seen; kept as a guard)

struct RegionJob {
    const Ctx& C; const std::string& s; const std::string& q; const std::vector<rtk_hit>& v_s; const std::vector<rtk_hit>& v_w;
    const size_t i_s, i_w;
    bool has_end_pt = false;
    rtk_hit um_solid1, um_solid2;
    size_t solid2_pos = 0, len_weak_region = 0;
    const char* s_start = nullptr;
    ResultCorrection res;
    std::string s_corrected, q_corrected;
    std::vector<std::pair<size_t, char>> v_ambiguity;
    std::vector<rtk_hit> l_v_w;
    IdSet all_pids;
    RegionReq req;
    bool has_req = false;
    FixAmbiguity amb;
    RegionJob(const Ctx& C_, const std::string& s_, const std::string& q_, const std::vector<rtk_hit>& v_s_, const std::vector<rtk_hit>& v_w_, size_t i_s_, size_t i_w_)
        : C(C_), s(s_), q(q_), v_s(v_s_), v_w(v_w_), i_s(i_s_), i_w(i_w_), res(0) {}
    void prepare(const IdSet* rc_pids);
    void paths();
    void paths_impl(bool use_device);
    bool amb_begin(AlignJob& j) { return amb.begin(C, s_corrected, q_corrected, s_start, res.old_seq_len, v_ambiguity, j); }
    void amb_finish(const std::vector<uint8_t>& ops) { amb.finish(C, s_corrected, q_corrected, s_start, ops); }
    bool trim_begin(AlignJob& j);
    void trim_finish(int32_t dd, int32_t fe, int32_t le);
    ResultCorrection& done() { res.seq = std::move(s_corrected); res.qual = std::move(q_corrected); return res; }
};

void RegionJob::prepare(const IdSet* rc_pids) {
    const rtk_graph_view& g = C.g;
    const rtk_opt& opt = C.opt;
    const size_t k = g.k;
    has_end_pt = (i_s + 1) < v_s.size();
    um_solid1 = v_s[i_s];
    memset(&um_solid2, 0, sizeof(um_solid2));
    if (has_end_pt) { um_solid2 = v_s[i_s + 1]; solid2_pos = um_solid2.pos; }
    else solid2_pos = s.length() - k;
    len_weak_region = solid2_pos - um_solid1.pos + k;
    const int64_t min_start = (int64_t)((size_t)um_solid1.pos - (size_t)opt.insert_sz);
    const int64_t min_end = (int64_t)(solid2_pos + opt.insert_sz);
    s_start = s.c_str() + um_solid1.pos;
    res = ResultCorrection(len_weak_region);
    const bool rc = rc_pids != nullptr;
    // `v[i].first > min_start` etc. compare a size_t with an int64_t: the signed side converts to unsigned (a negative
    // min_start becomes huge and the loop body never runs) - reproduced with explicit casts
    const uint64_t u_min_start = (uint64_t)min_start, u_min_end = (uint64_t)min_end;
    auto usable = [&](uint32_t u) { return kmer_coverage(g, u) < (double)C.max_km_cov; };
    const size_t v_w_sz = v_w.size();
    if (!rc) {
        ColorSide s_l, s_m, s_r;
        {
            size_t nb_branching = 0;
            for (int64_t i = (int64_t)i_s; i >= 0 && (uint64_t)v_s[(size_t)i].pos > u_min_start; --i) {
                const uint32_t u = v_s[(size_t)i].unitig;
                if (usable(u) && (!is_branching(g, u) || nb_branching < 5)) { const bool unseen = side_insert(s_l, u, !is_branching(g, u)); nb_branching += (size_t)(unseen && is_branching(g, u)); }
            }
            size_t i_w_s = i_w - (size_t)((i_w != 0) && (i_w >= v_w_sz));
            while (i_w_s > 0 && (uint64_t)v_w[i_w_s].pos > u_min_start) --i_w_s;
            for (; i_w_s < v_w_sz && v_w[i_w_s].pos < v_s[i_s].pos; ++i_w_s) {
                const uint32_t u = v_w[i_w_s].unitig;
                if (usable(u) && (!is_branching(g, u) || nb_branching < 5)) { const bool unseen = side_insert(s_l, u, !is_branching(g, u)); nb_branching += (size_t)(unseen && is_branching(g, u)); }
            }
        }
        if (has_end_pt) {
            size_t nb_branching = 0;
            for (size_t i = i_s + 1; i < v_s.size() && (uint64_t)v_s[i].pos < u_min_end; ++i) {
                const uint32_t u = v_s[i].unitig;
                if (usable(u) && (!is_branching(g, u) || nb_branching < 5)) { const bool unseen = side_insert(s_r, u, !is_branching(g, u)); nb_branching += (size_t)(unseen && is_branching(g, u)); }
            }
            size_t i_w_s = i_w - (size_t)((i_w != 0) && (i_w >= v_w_sz));
            while (i_w_s < v_w_sz && v_w[i_w_s].pos < v_s[i_s + 1].pos) ++i_w_s;
            for (; i_w_s < v_w_sz && (uint64_t)v_w[i_w_s].pos < u_min_end; ++i_w_s) {
                const uint32_t u = v_w[i_w_s].unitig;
                if (usable(u) && (!is_branching(g, u) || nb_branching < 5)) { const bool unseen = side_insert(s_r, u, !is_branching(g, u)); nb_branching += (size_t)(unseen && is_branching(g, u)); }
            }
        }
        if (!v_w.empty()) {
            const size_t pos_end = has_end_pt ? v_s[i_s + 1].pos : s.length();
            size_t i_w_s = i_w - (size_t)((i_w != 0) && (i_w >= v_w_sz));
            while (i_w_s < v_w_sz && v_w[i_w_s].pos < v_s[i_s].pos) ++i_w_s;
            for (; i_w_s < v_w_sz && v_w[i_w_s].pos < pos_end; ++i_w_s) {
                l_v_w.push_back(v_w[i_w_s]);
                if (usable(v_w[i_w_s].unitig)) side_insert(s_m, v_w[i_w_s].unitig, !is_branching(g, v_w[i_w_s].unitig));
            }
        }
        all_pids = choose_colors(C, s_l, s_r, s_m);
        res.all_pids = all_pids;
    } else {
        if (!v_w.empty()) {
            const size_t pos_end = has_end_pt ? v_s[i_s + 1].pos : s.length();
            size_t i_w_s = i_w - (size_t)((i_w != 0) && (i_w >= v_w_sz));
            while (i_w_s < v_w_sz && v_w[i_w_s].pos < v_s[i_s].pos) ++i_w_s;
            for (; i_w_s < v_w_sz && v_w[i_w_s].pos < pos_end; ++i_w_s) l_v_w.push_back(v_w[i_w_s]);
        }
        all_pids = *rc_pids;
    }
    // the region-engine request of the path phase (dead ends followed on the device)
    has_req = false;
    if (all_pids.size() >= opt.min_cov_vertices && C.device_regions && current_broker()) {
        req = RegionReq();
        req.opt = &C.opt; req.pass = C.pass2 ? 2 : 1;
        req.s = &s; req.um_start = um_solid1; req.um_end = um_solid2; req.has_end = has_end_pt; req.end_pos = solid2_pos;
        req.v_w = &l_v_w; req.i_weak = 0; req.pids = &all_pids; req.follow_dead_ends = true;
        has_req = true;
    }
}

void RegionJob::paths() {
    if (has_req && req.status != 2) {
        const rtk_hit saved1 = um_solid1;
        const size_t saved_len = len_weak_region;
        try { paths_impl(true); return; }
        catch (const InconsistentSegments&) {   // replay the region request by request
            um_solid1 = saved1; len_weak_region = saved_len;
            s_corrected.clear(); q_corrected.clear(); v_ambiguity.clear();
            res = ResultCorrection(saved_len); res.all_pids = all_pids;
        }
    }
    if (has_req) { if (GpuBroker* b = current_broker()) b->set_express(); }   // declined: hundreds of dependent requests follow
    paths_impl(false);
}

void RegionJob::paths_impl(const bool use_device) {
    const rtk_graph_view& g = C.g;
    const rtk_opt& opt = C.opt;
    const size_t k = g.k;
    const bool lrc = C.pass2;
    const size_t max_len_weak_anchors = lrc ? opt.max_len_weak_region2 : opt.max_len_weak_region1;
    const char q_min = rtk_get_qual(0.0, 0, opt.max_qual);
    // NB: the reference fills q_corrected with len_weak_region (captured by reference, so its CURRENT value) copies of qual
    auto set_uncorrected = [&](size_t pos, size_t len, char qual) {
        s_corrected = s.substr(pos, len);
        q_corrected = lrc ? q.substr(pos, len) : std::string(len_weak_region, qual);
    };
    auto add_uncorrected = [&](size_t pos, size_t len, char qual) {
        s_corrected += s.substr(pos, len);
        q_corrected += lrc ? q.substr(pos, len) : std::string(len_weak_region, qual);
    };
    // the two GPU-backed steps of the loop: extractSemiWeakPaths and the prefix alignment of a dead-end path.  With the engine's
    // answer they read its segments (one per extractSemiWeakPaths call, in order); otherwise they are requests of their own.
    size_t seg = 0;
    std::vector<std::string> seg_str;   // spelled path of the current segment (the engine spells it)
    auto eswp = [&](size_t i_weak_arg, bool initial) -> PathPair {
        // declined by the engine (or no engine): the request-at-a-time orchestration; not another engine request, which would
        // wait for the next bulk wave only to be declined again
        if (!use_device) return has_req ? extract_semi_weak_paths(C, s, all_pids, um_solid1, has_end_pt, um_solid2, solid2_pos, l_v_w, i_weak_arg)
                                        : region_paths(C, s, all_pids, um_solid1, has_end_pt, um_solid2, solid2_pos, l_v_w, i_weak_arg);
        if (!initial) ++seg;
        if (seg >= req.segs.size()) throw InconsistentSegments();
        RegionReq::Seg& S = req.segs[seg];
        if (S.start_weak != (initial ? RTK_NONE32 : (uint32_t)i_weak_arg)) throw InconsistentSegments();
        PathPair pp;
        (S.status == 0 ? pp.first : pp.second).push_back(std::move(S.path));
        return pp;
    };
    auto spelled = [&](const GPath& po) -> std::string { return use_device ? req.segs[seg].seq : po.to_string(g); };
    auto prefix_cut = [&](const std::vector<GPath>& cands, const std::string& ref) -> std::pair<int, int> {
        if (!use_device) return select_prefix_cut(C, cands, ref, opt.weak_region_len_factor);
        const RegionReq::Seg& S = req.segs[seg];
        if (cands.size() != 1 || S.status != 1) throw InconsistentSegments();
        const double best = static_cast<double>(S.shw_dist) / cands[0].length();
        if (opt.weak_region_len_factor > 0.0 && best > opt.weak_region_len_factor) return {-1, -1};
        return {0, S.shw_first_end};
    };
    const size_t card_pids = all_pids.size();
    PathPair paths1;
    if (card_pids >= opt.min_cov_vertices) paths1 = eswp(0, true);
    auto emit_path = [&](const GPath& p, bool offset_amb) {
        const std::vector<std::pair<size_t, char>> v_amb = ambiguity_vector(g, p.v);
        for (const auto& a : v_amb) v_ambiguity.push_back({(offset_amb ? s_corrected.length() : 0) + a.first, a.second});
    };
    if (paths1.first.empty()) {
        size_t i_w_s = 0;
        while (paths1.first.empty() && !paths1.second.empty() && !l_v_w.empty() && card_pids >= opt.min_cov_vertices) {
            const std::pair<int, int> align = prefix_cut(paths1.second, s.substr(um_solid1.pos, len_weak_region));
            if (align.first == -1) break;
            {
                const size_t next_pos = um_solid1.pos + align.second + k;
                while (i_w_s < l_v_w.size() && l_v_w[i_w_s].pos < next_pos) ++i_w_s;
                if (i_w_s >= l_v_w.size() || l_v_w[i_w_s].pos >= solid2_pos - k || (l_v_w[i_w_s].pos - um_solid1.pos) >= max_len_weak_anchors) break;
            }
            const GPath& po = paths1.second[(size_t)align.first];
            emit_path(po, true);
            const size_t gap = l_v_w[i_w_s].pos - um_solid1.pos - align.second - 1;
            s_corrected += spelled(po) + s.substr(um_solid1.pos + align.second + 1, gap);
            q_corrected += po.qual;
            if (lrc) q_corrected += q.substr(um_solid1.pos + align.second + 1, gap);
            else q_corrected += std::string(gap, q_min);
            res.add_range(um_solid1.pos - v_s[i_s].pos, um_solid1.pos + align.second + 1 - v_s[i_s].pos);
            um_solid1 = l_v_w[i_w_s];
            len_weak_region = solid2_pos - um_solid1.pos + k;
            paths1 = eswp(i_w_s, false);
        }
        if (!paths1.first.empty()) {
            // a single candidate wins whatever its distance (and the end location is not used here): no alignment needed
            const std::pair<int, int> pa = paths1.first.size() == 1 ? std::pair<int, int>(0, -1) : select_best_alignment(C.ctx, g, paths1.first, s.substr(um_solid1.pos, len_weak_region));
            const GPath& po = paths1.first[(size_t)pa.first];
            emit_path(po, true);
            s_corrected += spelled(po);
            q_corrected += po.qual;
            res.add_range(um_solid1.pos - v_s[i_s].pos, solid2_pos - v_s[i_s].pos + k);
        } else if (!paths1.second.empty()) {
            const std::pair<int, int> align = prefix_cut(paths1.second, s.substr(um_solid1.pos, len_weak_region));
            if (align.first == -1) add_uncorrected(um_solid1.pos, len_weak_region, q_min);
            else {
                const GPath& po = paths1.second[(size_t)align.first];
                emit_path(po, true);
                const size_t gap = len_weak_region - align.second - 1;
                s_corrected += spelled(po) + s.substr(um_solid1.pos + align.second + 1, gap);
                q_corrected += po.qual;
                if (lrc) q_corrected += q.substr(um_solid1.pos + align.second + 1, gap);
                else q_corrected += std::string(gap, q_min);
                res.add_range(um_solid1.pos - v_s[i_s].pos, um_solid1.pos + align.second + 1 - v_s[i_s].pos);
            }
        } else if (!s_corrected.empty()) add_uncorrected(um_solid1.pos, len_weak_region, q_min);
        else set_uncorrected(v_s[i_s].pos, len_weak_region, q_min);
    } else {
        const std::pair<int, int> pa = paths1.first.size() == 1 ? std::pair<int, int>(0, -1) : select_best_alignment(C.ctx, g, paths1.first, s.substr(um_solid1.pos, len_weak_region));
        const GPath& po = paths1.first[(size_t)pa.first];
        emit_path(po, false);
        s_corrected = spelled(po);
        q_corrected = po.qual;
        res.add_range(0, len_weak_region);
    }
    if (use_device && seg + 1 != req.segs.size()) throw InconsistentSegments();
}

// after fixAmbiguity: fully corrected?  else the SHW alignment of the raw window against the correction that trims it (:727-747)
bool RegionJob::trim_begin(AlignJob& job) {
    const size_t k = C.g.k;
    if (res.nb_corrected() == res.old_seq_len) {
        // Kmer(end of s) == Kmer(end of s_corrected): Kmer::set_kmer maps characters through bit tricks, compare the same way
        bool same = s_corrected.length() >= k;
        for (size_t i = 0; same && i < k; ++i) {
            const char a = s[s.length() - k + i], b = s_corrected[s_corrected.length() - k + i];
            const size_t xa = (a & 4) >> 1, xb = (b & 4) >> 1;
            same = ((xa + ((xa ^ (a & 2)) >> 1)) == (xb + ((xb ^ (b & 2)) >> 1)));
        }
        if (same) res.is_corrected = true;
    }
    if (res.is_corrected) return false;
    job.q = s.substr(v_s[i_s].pos, solid2_pos - v_s[i_s].pos + k);
    job.t = s_corrected; job.mode = 1; job.tref = nullptr;
    return true;
}

void RegionJob::trim_finish(const int32_t dd, const int32_t fe, const int32_t le) {
    // the largest best end is kept (:735-741); the reference scans endLocations[] as size_t, so a first end of -1 (edlib's
    // "position -1") wraps to the maximum and wins
    if (dd >= 0) {
        const size_t end_location = (fe < 0) ? (size_t)fe : (size_t)le;
        s_corrected = s_corrected.substr(0, end_location + 1);
        q_corrected = q_corrected.substr(0, end_location + 1);
    }
}

// the stages of up to two jobs side by side: ONE region-engine request (both path searches), one K5 request (both fixAmbiguity
// alignments), one K4 request (both trims)
void run_region_jobs(const Ctx& C, RegionJob* a, RegionJob* b) {
    RegionJob* jobs[2] = {a, b};
    std::vector<RegionReq*> reqs;
    for (RegionJob* j : jobs) if (j && j->has_req) reqs.push_back(&j->req);
    if (!reqs.empty()) current_broker()->submit(reqs);
    for (RegionJob* j : jobs) if (j) j->paths();
    {
        std::vector<AlignJob> aj;
        RegionJob* who[2]; size_t n = 0;
        for (RegionJob* j : jobs) if (j) { AlignJob x; if (j->amb_begin(x)) { aj.push_back(std::move(x)); who[n++] = j; } }
        if (n) {
            std::vector<int32_t> d;
            std::vector<std::vector<uint8_t>> ops;
            gpu_paths(C.ctx, aj, d, ops);
            for (size_t i = 0; i < n; ++i) who[i]->amb_finish(ops[i]);
        }
    }
    {
        std::vector<AlignJob> aj;
        RegionJob* who[2]; size_t n = 0;
        for (RegionJob* j : jobs) if (j) { AlignJob x; if (j->trim_begin(x)) { aj.push_back(std::move(x)); who[n++] = j; } }
        if (n) {
            std::vector<int32_t> dd, fe, le;
            gpu_distances_fl(C.ctx, aj, dd, fe, le);
            for (size_t i = 0; i < n; ++i) who[i]->trim_finish(dd[i], fe[i], le[i]);
        }
    }
}

ResultCorrection correct_region(const Ctx& C, const std::string& s, const std::string& q, const std::vector<rtk_hit>& v_s,
                                const std::vector<rtk_hit>& v_w, size_t i_s, size_t i_w, const ResultCorrection* rc) {
    RegionJob job(C, s, q, v_s, v_w, i_s, i_w);
    job.prepare(rc ? &rc->all_pids : nullptr);
    run_region_jobs(C, &job, nullptr);
    return std::move(job.done());
}

inline bool has_min_qual(const std::string& s, const std::string& q, size_t start, size_t end, char min_q) {
    bool ok = true;
    for (size_t i = start; i < end && ok; ++i) ok = (q[i] >= min_q) || !is_dna(s[i]);
    return ok;
}

// ------------------------------------------------------------------ correctSequence (src/Correction.cpp:159-958)
// The reference walks the solid anchors left to right and appends one piece of output per step.  Every piece
// depends only on the read, its anchors and three loop variables (i_solid, i_weak, prev_pos) that are pure
// functions of the anchor positions, so the pieces are planned first and computed independently (one broker
// task each), then concatenated in order: same bytes, but a read's regions no longer wait for one another.
struct ReadJob {
    const std::string* s_fw; const std::string* q_fw;
    const std::vector<rtk_hit>* solid; const std::vector<rtk_hit>* weak;
    std::string s_bw, q_bw;
    std::vector<rtk_hit> solid_rev, weak_rev;
    bool trivial = false;
    std::pair<std::string, std::string> trivial_out;
};
struct Piece { uint32_t read; int kind; size_t i_solid, i_weak, prev_pos; std::string s, q; };  // kind 0 leading, 1 gap, 2 trailing

void plan_read(const rtk_graph_view& g, const rtk_opt& opt, bool lrc, uint32_t r, ReadJob& J, std::vector<Piece>& pieces) {
    const size_t k = g.k;
    const std::string& s_fw = *J.s_fw; const std::string& q_fw = *J.q_fw;
    const std::vector<rtk_hit>& v_um_solid = *J.solid; const std::vector<rtk_hit>& v_um_weak = *J.weak;
    if (s_fw.length() <= k || v_um_solid.empty() || v_um_solid.size() == s_fw.length() - k + 1) {
        J.trivial = true;
        if (lrc) J.trivial_out = {s_fw, q_fw};
        else if (v_um_solid.size() == s_fw.length() - k + 1) J.trivial_out = {s_fw, std::string(s_fw.length(), rtk_get_qual(1.0, 0, opt.max_qual))};
        else J.trivial_out = {s_fw, std::string(s_fw.length(), rtk_get_qual(0.0, 0, opt.max_qual))};
        return;
    }
    const size_t seq_len = s_fw.length();
    J.s_bw = rc_string(s_fw);
    J.q_bw = q_fw;
    std::reverse(J.q_bw.begin(), J.q_bw.end());
    J.solid_rev.assign(v_um_solid.rbegin(), v_um_solid.rend());
    J.weak_rev.assign(v_um_weak.rbegin(), v_um_weak.rend());
    for (auto& p : J.solid_rev) { p.pos = (uint32_t)(seq_len - p.pos - k); p.strand = 1 - p.strand; }
    for (auto& p : J.weak_rev) { p.pos = (uint32_t)(seq_len - p.pos - k); p.strand = 1 - p.strand; }
    size_t prev_pos = v_um_solid[0].pos, i_solid = 0, i_weak = 0;
    if (v_um_solid[0].pos != 0) pieces.push_back({r, 0, 0, 0, prev_pos, {}, {}});
    while (i_solid < v_um_solid.size() - 1) {
        while (i_weak < v_um_weak.size() && v_um_weak[i_weak].pos < v_um_solid[i_solid].pos) ++i_weak;
        if (v_um_solid[i_solid].pos != v_um_solid[i_solid + 1].pos - 1) {
            pieces.push_back({r, 1, i_solid, i_weak, prev_pos, {}, {}});
            prev_pos = v_um_solid[i_solid + 1].pos;
        }
        ++i_solid;
    }
    pieces.push_back({r, 2, i_solid, i_weak, prev_pos, {}, {}});
}

void run_piece(const Ctx& C, const ReadJob& J, Piece& P) {
    const rtk_graph_view& g = C.g;
    const rtk_opt& opt = C.opt;
    const size_t k = g.k;
    const bool lrc = C.pass2;
    const std::string& s_fw = *J.s_fw; const std::string& q_fw = *J.q_fw;
    const std::string& s_bw = J.s_bw; const std::string& q_bw = J.q_bw;
    const std::vector<rtk_hit>& v_um_solid = *J.solid; const std::vector<rtk_hit>& v_um_weak = *J.weak;
    const std::vector<rtk_hit>& solid_rev = J.solid_rev; const std::vector<rtk_hit>& weak_rev = J.weak_rev;
    const char q_min = rtk_get_qual(0.0, 0, opt.max_qual), q_max = rtk_get_qual(1.0, 0, opt.max_qual);
    std::string& corrected_s = P.s; std::string& corrected_q = P.q;
    const size_t prev_pos = P.prev_pos;
    size_t i_solid = P.i_solid, i_weak = P.i_weak;
    if (P.kind == 0) {
        if (!lrc || q_fw.length() == 0 || !has_min_qual(s_fw, q_fw, 0, v_um_solid[0].pos + k, q_max)) {
            const size_t i_solid_rev = solid_rev.size() - 1;
            size_t i_weak_rev = weak_rev.size();
            while (i_weak_rev > 0 && weak_rev[i_weak_rev - 1].pos > solid_rev[i_solid_rev].pos) --i_weak_rev;
            ResultCorrection bw = correct_region(C, s_bw, q_fw, solid_rev, weak_rev, i_solid_rev, i_weak_rev, nullptr);
            bw.reverse_complement();
            corrected_s += bw.seq.substr(0, bw.seq.length() - k);
            corrected_q += bw.qual.substr(0, bw.qual.length() - k);
        } else {
            corrected_s += s_fw.substr(0, v_um_solid[0].pos);
            corrected_q += lrc ? q_fw.substr(0, v_um_solid[0].pos) : std::string(v_um_solid[0].pos, q_min);
        }
        return;
    }
    if (P.kind == 1) {
        {
            bool isUncorrected = false;
            const size_t p0 = v_um_solid[i_solid].pos, p1 = v_um_solid[i_solid + 1].pos;
            if (!lrc || q_fw.length() == 0 || !has_min_qual(s_fw, q_fw, p0, p1 + k, q_max)) {
                const rtk_hit& start_um = v_um_solid[i_solid];
                const rtk_hit& end_um = v_um_solid[i_solid + 1];
                bool sameUnitig = (start_um.unitig == end_um.unitig) && (start_um.strand == end_um.strand);
                if (sameUnitig && !is_short_cycle(g, start_um.unitig)) {
                    const size_t min_pos = std::min(start_um.dist, end_um.dist), max_pos = std::max(start_um.dist, end_um.dist);
                    const size_t len_query_km = p1 - p0, len_unitig_km = max_pos - min_pos;
                    const size_t min_len = rtk_min_max_length(len_unitig_km, opt.weak_region_len_factor).first;
                    const size_t max_len = rtk_min_max_length(len_unitig_km, opt.weak_region_len_factor).second;
                    sameUnitig = sameUnitig && ((start_um.strand && start_um.dist < end_um.dist) || (!start_um.strand && start_um.dist > end_um.dist));
                    sameUnitig = sameUnitig && (len_query_km >= min_len) && (len_query_km <= max_len);
                    if (sameUnitig) {
                        PNode um_sub; um_sub.unitig = start_um.unitig; um_sub.strand = start_um.strand; um_sub.dist = (uint32_t)min_pos; um_sub.len = (uint32_t)len_unitig_km + 1;
                        GPath t; t.v.push_back(um_sub);
                        const std::string s_um_sub = t.to_string(g);
                        corrected_s += s_fw.substr(prev_pos, p0 - prev_pos) + s_um_sub.substr(0, s_um_sub.length() - k);
                        if (lrc) {
                            const size_t buff = (s_um_sub.length() >= 2 * k) ? k : (s_um_sub.length() - k);
                            corrected_q += q_fw.substr(prev_pos, p0 - prev_pos + buff);
                            if ((s_um_sub.length() - buff - k) > 0) corrected_q += std::string(s_um_sub.length() - buff - k, q_max);
                        } else corrected_q += std::string((p0 - prev_pos) + (s_um_sub.length() - k), q_max);
                    } else isUncorrected = true;
                } else if (p1 >= p0 + k) {
                    // The backward attempt only runs when the forward one leaves the region not fully corrected - which is the common
                    // case on noisy reads (9 in 10 regions here) - and depends on the forward attempt through its colour set alone, known
                    // before any path search.  Both attempts are therefore prepared together and share every GPU request; a backward
                    // result the forward attempt makes unnecessary is dropped.
                    const size_t i_solid_bw = solid_rev.size() - i_solid - 2;
                    size_t i_weak_bw = weak_rev.size() - i_weak;
                    while (i_weak_bw > 0 && weak_rev[i_weak_bw - 1].pos > solid_rev[i_solid_bw].pos) --i_weak_bw;
                    RegionJob jf(C, s_fw, q_fw, v_um_solid, v_um_weak, i_solid, i_weak);
                    jf.prepare(nullptr);
                    RegionJob jb(C, s_bw, q_bw, solid_rev, weak_rev, i_solid_bw, i_weak_bw);
                    jb.prepare(&jf.all_pids);
                    run_region_jobs(C, &jf, &jb);
                    const ResultCorrection fw = std::move(jf.done());
                    if (fw.is_corrected) {
                        const size_t l_solid = p0 - prev_pos;
                        const std::string sub = s_fw.substr(prev_pos, l_solid) + fw.seq;
                        const std::string subq = (lrc ? q_fw.substr(prev_pos, l_solid) : std::string(l_solid, q_max)) + fw.qual;
                        corrected_s += sub.substr(0, sub.length() - k);
                        corrected_q += subq.substr(0, subq.length() - k);
                    } else {
                        ResultCorrection bw = std::move(jb.done());
                        bw.reverse_complement();
                        if (bw.is_corrected) {
                            const size_t l_solid = (s_bw.length() - solid_rev[i_solid_bw + 1].pos - k) - prev_pos;
                            const std::string sub = s_fw.substr(prev_pos, l_solid) + bw.seq;
                            const std::string subq = (lrc ? q_fw.substr(prev_pos, l_solid) : std::string(l_solid, q_max)) + bw.qual;
                            corrected_s += sub.substr(0, sub.length() - k);
                            corrected_q += subq.substr(0, subq.length() - k);
                        } else {
                            std::string l_ref = s_fw.substr(p0, p1 - p0 + k);
                            std::pair<std::string, std::string> cons = generate_consensus(C, &fw, &bw, l_ref, opt.weak_region_len_factor);
                            if (cons.first.length() == 0) {
                                cons.first = l_ref;
                                if (lrc) cons.second = q_fw.substr(p0, p1 - p0 + k);
                                else cons.second = std::string(k, q_max) + std::string(p1 - p0, q_min);
                            }
                            const size_t l_solid = p0 - prev_pos;
                            const std::string sub = s_fw.substr(prev_pos, l_solid) + cons.first;
                            const std::string subq = (lrc ? q_fw.substr(prev_pos, l_solid) : std::string(l_solid, q_max)) + cons.second;
                            corrected_s += sub.substr(0, sub.length() - k);
                            corrected_q += subq.substr(0, subq.length() - k);
                        }
                    }
                } else isUncorrected = true;
            } else isUncorrected = true;
            if (isUncorrected) {
                corrected_s += s_fw.substr(prev_pos, p1 - prev_pos);
                if (lrc) corrected_q += q_fw.substr(prev_pos, p1 - prev_pos);
                else {
                    corrected_q += std::string(p0 - prev_pos, q_max);
                    if (p1 < p0 + k) corrected_q += std::string(p1 - p0, q_max);
                    else corrected_q += std::string(k, q_max) + std::string(p1 - p0 - k, q_min);
                }
            }
        }
        return;
    }
    // trailing region / tail of the read
    if ((v_um_solid.back().pos < s_fw.length() - k) &&
        (!lrc || q_fw.length() == 0 || !has_min_qual(s_fw, q_fw, v_um_solid.back().pos, s_fw.length(), q_max))) {
        while (i_weak < v_um_weak.size() && v_um_weak[i_weak].pos < v_um_solid[i_solid].pos) ++i_weak;
        const ResultCorrection fw = correct_region(C, s_fw, q_fw, v_um_solid, v_um_weak, i_solid, i_weak, nullptr);
        const size_t l_solid = v_um_solid[i_solid].pos - prev_pos;
        corrected_s += s_fw.substr(prev_pos, l_solid) + fw.seq;
        corrected_q += (lrc ? q_fw.substr(prev_pos, l_solid) : std::string(l_solid, q_max)) + fw.qual;
    } else {
        corrected_s += s_fw.substr(prev_pos);
        corrected_q += lrc ? q_fw.substr(prev_pos)
                           : (std::string(v_um_solid[i_solid].pos - prev_pos + k, q_max) + std::string(s_fw.length() - v_um_solid[i_solid].pos - k, q_min));
    }
}

}  // namespace

// weak regions in flight per batch (= fibers parked on the GPU broker); RTK_CORRECT_INFLIGHT overrides
static unsigned correct_threads() {
    const char* e = getenv("RTK_CORRECT_INFLIGHT");
    const int v = e ? atoi(e) : 131072;
    return (unsigned)std::max(1, std::min(v, 1 << 20));
}

// ------------------------------------------------------------------ batch driver: the per-read body of search() (src/Ratatosk.cpp:808-867)
// one gang: reads [0, n_reads) of its own pools through every correction round, on its own context
static void correct_range(rtk_ctx* ctx, const rtk_opt& opt, int pass, uint32_t n_reads, const char* seq_pool, const uint64_t* seq_off,
                          const char* qual_pool, const uint64_t* qual_off, std::string* out_seq, std::string* out_qual, uint64_t* stats,
                          std::mutex* seeds_turn = nullptr) {
    const rtk_graph_view& g = ctx->host_graph->view;
    const bool pass2 = (pass == 2);
    // exploreSubGraphLong expands a partial path while it spells fewer than k * large_k_factor bases: every unitig adds at least one
    // base, so a burst holds at most k * large_k_factor - k + 2 unitigs; the kernels' per-burst capacity is fixed (subgraph.cuh)
    if (pass2 && (double)opt.k * opt.large_k_factor - (double)opt.k + 2.0 > (double)RTK_DFS_MAX_NODES)
        throw std::invalid_argument("large_k_factor too large for this build: k * large_k_factor - k + 2 must not exceed " + std::to_string(RTK_DFS_MAX_NODES) + " unitigs per burst");
    const size_t max_km_cov = std::max<size_t>(ctx->host_graph->hdr.max_km_cov_graph, opt.max_km_cov);  // src/Ratatosk.cpp:625
    parallel_for(n_reads, [&](size_t rb, size_t re) {
        for (size_t r = rb; r < re; ++r) {
            out_seq[r].assign(seq_pool + seq_off[r], seq_off[r + 1] - seq_off[r]);
            for (auto& c : out_seq[r]) if (c >= 'a' && c <= 'z') c = (char)(c - 'a' + 'A');  // :814 (::toupper in the "C" locale)
            if (qual_pool && qual_off) out_qual[r].assign(qual_pool + qual_off[r], qual_off[r + 1] - qual_off[r]);
            if (!pass2) for (auto& c : out_qual[r]) { if (c < (char)33) c = (char)33; if (c > (char)(33 + opt.max_qual)) c = (char)(33 + opt.max_qual); }  // getStdQual
        }
    });
    const size_t rounds = pass2 ? 1 : std::max<uint32_t>(1, opt.nb_correction_rounds);
    const double step_min_score = 1.00 / static_cast<double>(rounds);
    const double step_wrlf = (rounds == 1) ? 0.0 : ((opt.weak_region_len_factor - 0.10) / static_cast<double>(rounds - 1));
    const size_t step_mlwr1 = opt.max_len_weak_region1 / rounds;
    for (size_t j = 0; j < rounds; ++j) {
        rtk_opt l_opt = opt;
        if (!pass2) {
            l_opt.min_score = 1.00 - (j + 1) * step_min_score;
            l_opt.weak_region_len_factor = opt.weak_region_len_factor - (rounds - j - 1) * step_wrlf;
            l_opt.max_len_weak_region1 = (uint32_t)((j + 1) * step_mlwr1);
        }
        std::vector<uint64_t> off(n_reads + 1, 0);
        for (uint32_t r = 0; r < n_reads; ++r) off[r + 1] = off[r] + out_seq[r].size();
        std::string pool(off[n_reads] + 1, '\0');
        parallel_for(n_reads, [&](size_t rb, size_t re) { for (size_t r = rb; r < re; ++r) memcpy(&pool[off[r]], out_seq[r].data(), out_seq[r].size()); });
        std::vector<std::vector<rtk_hit>> solid, weak;
        auto t_seeds = std::chrono::steady_clock::now();
        if (seeds_turn) {
            // The gangs take turns at the anchor stage, each with ALL host threads: the stage is short and host-heavy, and taking
            // turns staggers the gangs - while one gang's regions are on the device, the next gang's anchors are computed.
            std::lock_guard<std::mutex> turn(*seeds_turn);
            const unsigned mine = thread_budget();
            set_thread_budget(host_threads());
            t_seeds = std::chrono::steady_clock::now();
            get_seeds_host(ctx, l_opt, pass, n_reads, pool.data(), off.data(), solid, weak, stats);
            set_thread_budget(mine);
        } else get_seeds_host(ctx, l_opt, pass, n_reads, pool.data(), off.data(), solid, weak, stats);
        const auto t_broker = std::chrono::steady_clock::now();
        // every region of every read is one broker task: regions run concurrently on host threads, their GPU requests
        // are served in waves (broker.hpp); pieces are concatenated in read order afterwards
        Ctx C{ctx, g, l_opt, TraverseOpt(), pass2, max_km_cov};
        C.topt.k = g.k; C.topt.min_cov_vertices = l_opt.min_cov_vertices; C.topt.out_qual = l_opt.out_qual; C.topt.max_qual = l_opt.max_qual;
        C.topt.weak_region_len_factor = l_opt.weak_region_len_factor; C.topt.large_k_factor = l_opt.large_k_factor; C.topt.min_score = l_opt.min_score;
        C.topt.long_read_correct = pass2;
        C.device_regions = getenv("RTK_NO_REGION_ENGINE") == nullptr;
        std::vector<ReadJob> jobs(n_reads);
        std::vector<Piece> pieces;
        for (uint32_t r = 0; r < n_reads; ++r) {
            jobs[r].s_fw = &out_seq[r]; jobs[r].q_fw = &out_qual[r]; jobs[r].solid = &solid[r]; jobs[r].weak = &weak[r];
            plan_read(g, l_opt, pass2, r, jobs[r], pieces);
        }
        // longest regions first: a region is a chain of dependent GPU requests roughly proportional to its span, and the
        // call ends when the longest chain does (the output order is restored below, so scheduling order is free)
        std::vector<uint32_t> order(pieces.size());
        std::vector<uint32_t> span(pieces.size());
        for (size_t i = 0; i < pieces.size(); ++i) {
            order[i] = (uint32_t)i;
            const Piece& P = pieces[i];
            const std::vector<rtk_hit>& vs = *jobs[P.read].solid;
            const size_t len = jobs[P.read].s_fw->length();
            span[i] = (P.kind == 0) ? vs[0].pos : (P.kind == 1) ? (vs[P.i_solid + 1].pos - vs[P.i_solid].pos) : (uint32_t)(len - vs[P.i_solid].pos);
        }
        if (!getenv("RTK_NO_LPT")) std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return span[a] > span[b]; });
        GpuBroker broker(ctx);
        broker.run(pieces.size(), correct_threads(), [&](size_t i) { run_piece(C, jobs[pieces[order[i]].read], pieces[order[i]]); });
        // pieces were planned read by read: [first_piece[r], first_piece[r + 1]) belong to read r, in order
        std::vector<size_t> first_piece(n_reads + 1, 0);
        for (const Piece& P : pieces) ++first_piece[P.read + 1];
        for (uint32_t r = 0; r < n_reads; ++r) first_piece[r + 1] += first_piece[r];
        parallel_for(n_reads, [&](size_t rb, size_t re) {
            for (size_t r = rb; r < re; ++r) {
                std::string ns, nq;
                if (jobs[r].trivial) { ns = jobs[r].trivial_out.first; nq = jobs[r].trivial_out.second; }
                else {
                    size_t tot = 0;
                    for (size_t x = first_piece[r]; x < first_piece[r + 1]; ++x) tot += pieces[x].s.size();
                    ns.reserve(tot); nq.reserve(tot);
                    for (size_t x = first_piece[r]; x < first_piece[r + 1]; ++x) { ns += pieces[x].s; nq += pieces[x].q; }
                }
                out_seq[r] = std::move(ns);
                out_qual[r] = std::move(nq);
            }
        });
        if (stats) {
            stats[5] += broker.waves; stats[6] += broker.jobs;
            for (int s3 = 0; s3 < 3; ++s3) stats[7 + s3] += broker.kernel_ns[s3];
            stats[16] += broker.kernel_ns[3]; stats[17] += broker.region_calls; stats[18] += broker.region_bails; stats[19] += broker.region_kcells;
            stats[10] += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t_broker - t_seeds).count();
            stats[11] += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_broker).count();
        }
    }
}

// Reads of a batch are independent: the batch is split into GANGS (contiguous read ranges of about equal size in bases),
// each running the whole pipeline (K1 sweeps -> anchors -> regions) on its own forked context and its own share of the host
// threads.  While one gang's regions run on the device as one bulk launch (broker.hpp: the region service launches when every
// live region of the gang is waiting for it), another gang's host stages (anchor logic, colour selection, stitching) run:
// device and host overlap without any cross-gang dependency.  RTK_GANGS overrides the default of 3.
void correct_batch_host(rtk_ctx* ctx, const rtk_opt& opt, int pass, uint32_t n_reads, const char* seq_pool, const uint64_t* seq_off,
                        const char* qual_pool, const uint64_t* qual_off, std::vector<std::string>& out_seq, std::vector<std::string>& out_qual,
                        uint64_t* stats) {
    if (!ctx->has_graph || !ctx->host_graph) throw std::invalid_argument("no graph uploaded to this context");
    if (opt.k != ctx->host_graph->view.k) throw std::invalid_argument("rtk_opt.k does not match the graph's k");
#ifndef RTK_HOSTSIM
    const uint64_t launches0 = g_launches, h2d0 = g_h2d_bytes, d2h0 = g_d2h_bytes;
#endif
    out_seq.assign(n_reads, std::string());
    out_qual.assign(n_reads, std::string());
    const char* e = getenv("RTK_GANGS");
    unsigned n_gangs = e ? (unsigned)std::max(1, atoi(e)) : 3u;
    const uint64_t total_bases = n_reads ? seq_off[n_reads] - seq_off[0] : 0;
    if (total_bases < (4u << 20) || n_reads < 2 * n_gangs) n_gangs = 1;   // small batches: not worth splitting
    n_gangs = std::min(n_gangs, std::max(1u, host_threads()));
    if (n_gangs == 1) {
        correct_range(ctx, opt, pass, n_reads, seq_pool, seq_off, qual_pool, qual_off, out_seq.data(), out_qual.data(), stats);
    } else {
        // contiguous ranges of about equal bases
        std::vector<uint32_t> cut(n_gangs + 1, n_reads);
        cut[0] = 0;
        for (unsigned gi = 1; gi < n_gangs; ++gi) {
            const uint64_t want = seq_off[0] + total_bases * gi / n_gangs;
            cut[gi] = (uint32_t)(std::lower_bound(seq_off, seq_off + n_reads + 1, want) - seq_off);
            cut[gi] = std::max(cut[gi], cut[gi - 1]);
        }
        std::vector<std::vector<uint64_t>> gstats(n_gangs, std::vector<uint64_t>(24, 0));
        std::vector<std::string> errors(n_gangs);
        std::vector<rtk_ctx*> gctx(n_gangs, nullptr);
        for (unsigned gi = 0; gi < n_gangs; ++gi) {
            try { gctx[gi] = fork_acquire(ctx, false); }
            catch (...) { for (rtk_ctx* c : gctx) fork_release(ctx, c, false); throw; }
#ifndef RTK_HOSTSIM
            if (ctx->resident_seq && ctx->resident_n == n_reads) {   // the caller's resident copy of the reads, this gang's slice
                gctx[gi]->resident_seq = ctx->resident_seq; gctx[gi]->resident_off = ctx->resident_off + cut[gi];
                gctx[gi]->resident_n = cut[gi + 1] - cut[gi]; gctx[gi]->resident_total = seq_off[cut[gi + 1]] - seq_off[cut[gi]];
            }
#endif
        }
        const unsigned budget = std::max(1u, host_threads() / n_gangs);
        std::mutex seeds_turn;
        std::vector<std::thread> th;
        for (unsigned gi = 0; gi < n_gangs; ++gi) {
            th.emplace_back([&, gi] {
                set_thread_budget(budget);
                try {
#ifndef RTK_HOSTSIM
                    DeviceBind bind(gctx[gi]);
#endif
                    const uint32_t r0 = cut[gi], n = cut[gi + 1] - cut[gi];
                    if (n) correct_range(gctx[gi], opt, pass, n, seq_pool, seq_off + r0, qual_pool, qual_off ? qual_off + r0 : nullptr, out_seq.data() + r0,
                                         out_qual.data() + r0, gstats[gi].data(), &seeds_turn);
                } catch (const std::exception& ex) { errors[gi] = ex.what(); if (errors[gi].empty()) errors[gi] = "correction gang failed"; }
                catch (...) { errors[gi] = "correction gang failed (unknown exception)"; }
            });
        }
        for (auto& t : th) t.join();
        for (rtk_ctx* c : gctx) fork_release(ctx, c, false);
#ifndef RTK_HOSTSIM
        ctx->resident_seq = nullptr;
#endif
        for (const std::string& er : errors) if (!er.empty()) throw std::runtime_error(er);
        if (stats) {
            // counters add up; stage times are wall-clock per gang (the gangs run concurrently): report the slowest gang
            for (int i = 0; i < 24; ++i) {
                uint64_t sum = 0, mx = 0;
                for (unsigned gi = 0; gi < n_gangs; ++gi) { sum += gstats[gi][i]; mx = std::max(mx, gstats[gi][i]); }
                stats[i] += (i == 3 || i == 10 || i == 11) ? mx : sum;
            }
        }
    }
#ifndef RTK_HOSTSIM
    if (stats) { stats[12] += g_h2d_bytes - h2d0; stats[13] += g_d2h_bytes - d2h0; stats[14] += g_launches - launches0; }
#endif
}

// Both passes of a batch as ONE pipeline (rtk_correct_two_pass_batch).  The stages of a read depend only on that read and on
// the two read-only graphs, so the batch is cut into gangs (contiguous read ranges) and every gang runs pass 1 -> phasing ->
// pass 2 on its own reads with its own forked contexts of both graphs.  The gangs drift apart (they take turns at the
// host-heavy anchor stage), so while one gang's longest second-pass regions occupy a few SMs, another gang's k-mer sweeps,
// first-pass regions or whole-read alignments fill the rest of the device and the host threads: no stage waits for the
// slowest read of the previous stage of the WHOLE batch, as three separate calls would.
void phasing_batch_host(rtk_ctx* ctx, const rtk_opt& opt, uint32_t n, const char* raw_pool, const uint64_t* raw_off, const char* corr_pool,
                        const uint64_t* corr_off, const char* qual_pool, const uint64_t* qual_off, std::vector<std::string>& out_seq,
                        std::vector<std::string>& out_qual);

static void pack_strings(const std::string* v, uint32_t n, std::string& pool, std::vector<uint64_t>& off) {
    off.assign(n + 1, 0);
    for (uint32_t r = 0; r < n; ++r) off[r + 1] = off[r] + v[r].size();
    pool.assign(off[n] + 1, '\0');
    char* dst = &pool[0];
    parallel_for(n, [&](size_t rb, size_t re) { for (size_t r = rb; r < re; ++r) memcpy(dst + off[r], v[r].data(), v[r].size()); });
}

void correct_two_pass_host(rtk_ctx* ctx1, rtk_ctx* ctx2, const rtk_opt& opt1, const rtk_opt& opt2, uint32_t n_reads, const char* seq_pool,
                           const uint64_t* seq_off, const char* qual_pool, const uint64_t* qual_off, std::vector<std::string>& p1_seq,
                           std::vector<std::string>& p1_qual, std::vector<std::string>& out_seq, std::vector<std::string>& out_qual,
                           uint64_t* stats1, uint64_t* stats2, uint64_t* stage_ns) {
    if (!ctx1->has_graph || !ctx1->host_graph || !ctx2->has_graph || !ctx2->host_graph) throw std::invalid_argument("no graph uploaded to a context");
    if (opt1.k != ctx1->host_graph->view.k || opt2.k != ctx2->host_graph->view.k) throw std::invalid_argument("rtk_opt.k does not match the graph's k");
#ifndef RTK_HOSTSIM
    if (ctx1->device != ctx2->device) throw std::invalid_argument("the two contexts must live on the same device");
    const uint64_t launches0 = g_launches, h2d0 = g_h2d_bytes, d2h0 = g_d2h_bytes;
#endif
    p1_seq.assign(n_reads, std::string()); p1_qual.assign(n_reads, std::string());
    out_seq.assign(n_reads, std::string()); out_qual.assign(n_reads, std::string());
    if (!n_reads) return;
    const char* e = getenv("RTK_GANGS2");
    // default 3; a rank with few host threads (several ranks sharing a node's cores) runs fewer, larger gangs: measured with 4
    // threads, 2 gangs 7.6 Mbases/s, 3 gangs 6.9, 1 gang 6.5 (scripts/r2_sweep_lowthreads.sh)
    unsigned n_gangs = e ? (unsigned)std::max(1, atoi(e)) : std::min(3u, std::max(1u, host_threads() / 2));
    const uint64_t total_bases = seq_off[n_reads] - seq_off[0];
    const char* e_min = getenv("RTK_GANGS_MIN_BASES");   // tests force several gangs on small fixtures
    const uint64_t min_bases = e_min ? (uint64_t)std::max(1ll, atoll(e_min)) : (2ull << 20);
    n_gangs = (unsigned)std::min<uint64_t>(n_gangs, std::max<uint64_t>(1, total_bases / min_bases));   // at least ~2 Mbases per gang
    n_gangs = std::min(n_gangs, std::min(std::max(1u, host_threads()), n_reads));
    std::vector<uint32_t> cut(n_gangs + 1, n_reads);
    cut[0] = 0;
    for (unsigned gi = 1; gi < n_gangs; ++gi) {
        const uint64_t want = seq_off[0] + total_bases * gi / n_gangs;
        cut[gi] = (uint32_t)(std::lower_bound(seq_off, seq_off + n_reads + 1, want) - seq_off);
        cut[gi] = std::max(cut[gi], cut[gi - 1]);
    }
    std::vector<std::vector<uint64_t>> gs1(n_gangs, std::vector<uint64_t>(24, 0)), gs2(n_gangs, std::vector<uint64_t>(24, 0));
    std::vector<std::vector<uint64_t>> gns(n_gangs, std::vector<uint64_t>(3, 0));
    std::vector<std::string> errors(n_gangs);
    std::vector<rtk_ctx*> g1(n_gangs, nullptr), g2(n_gangs, nullptr);
    auto release_all = [&] { for (rtk_ctx* c : g1) fork_release(ctx1, c, false); for (rtk_ctx* c : g2) fork_release(ctx2, c, false); };
    try { for (unsigned gi = 0; gi < n_gangs; ++gi) { g1[gi] = fork_acquire(ctx1, false); g2[gi] = fork_acquire(ctx2, false); } }
    catch (...) { release_all(); throw; }
#ifndef RTK_HOSTSIM
    if (ctx1->resident_seq && ctx1->resident_n == n_reads) {   // the caller's resident copy of the reads (rtk_ctx_resident_reads): each gang its slice
        for (unsigned gi = 0; gi < n_gangs; ++gi) {
            g1[gi]->resident_seq = ctx1->resident_seq; g1[gi]->resident_off = ctx1->resident_off + cut[gi];
            g1[gi]->resident_n = cut[gi + 1] - cut[gi]; g1[gi]->resident_total = seq_off[cut[gi + 1]] - seq_off[cut[gi]];
        }
    }
    ctx1->resident_seq = nullptr;
#endif
    const unsigned budget = std::max(1u, host_threads() / n_gangs);
    std::mutex seeds_turn;
    std::vector<std::thread> th;
    for (unsigned gi = 0; gi < n_gangs; ++gi) {
        th.emplace_back([&, gi] {
            set_thread_budget(n_gangs == 1 ? host_threads() : budget);
            try {
#ifndef RTK_HOSTSIM
                DeviceBind bind(g1[gi]);
#endif
                const uint32_t r0 = cut[gi], n = cut[gi + 1] - cut[gi];
                if (!n) return;
                auto t0 = std::chrono::steady_clock::now();
                auto lap = [&](int k) { const auto t = std::chrono::steady_clock::now(); gns[gi][k] = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t - t0).count(); t0 = t; };
                static const bool no_turn = getenv("RTK_NO_SEEDS_TURN") != nullptr;
                std::mutex* turn = (n_gangs > 1 && !no_turn) ? &seeds_turn : nullptr;
                correct_range(g1[gi], opt1, 1, n, seq_pool, seq_off + r0, qual_pool, qual_off ? qual_off + r0 : nullptr, p1_seq.data() + r0, p1_qual.data() + r0,
                              gs1[gi].data(), turn);
                lap(0);
                std::string cp, cq;
                std::vector<uint64_t> co, qo;
                pack_strings(p1_seq.data() + r0, n, cp, co);
                pack_strings(p1_qual.data() + r0, n, cq, qo);
                std::vector<std::string> ps, pq;
                phasing_batch_host(g2[gi], opt2, n, seq_pool, seq_off + r0, cp.data(), co.data(), cq.data(), qo.data(), ps, pq);
                lap(1);
                pack_strings(ps.data(), n, cp, co);
                pack_strings(pq.data(), n, cq, qo);
                correct_range(g2[gi], opt2, 2, n, cp.data(), co.data(), cq.data(), qo.data(), out_seq.data() + r0, out_qual.data() + r0, gs2[gi].data(), turn);
                lap(2);
            } catch (const std::exception& ex) { errors[gi] = ex.what(); if (errors[gi].empty()) errors[gi] = "correction gang failed"; }
            catch (...) { errors[gi] = "correction gang failed (unknown exception)"; }
        });
    }
    for (auto& t : th) t.join();
    release_all();
    for (const std::string& er : errors) if (!er.empty()) throw std::runtime_error(er);
    for (int pass = 0; pass < 2; ++pass) {
        uint64_t* st = pass ? stats2 : stats1;
        if (!st) continue;
        for (int i = 0; i < 24; ++i) {
            uint64_t sum = 0, mx = 0;
            for (unsigned gi = 0; gi < n_gangs; ++gi) { const uint64_t v = (pass ? gs2 : gs1)[gi][i]; sum += v; mx = std::max(mx, v); }
            st[i] += (i == 3 || i == 10 || i == 11) ? mx : sum;
        }
    }
    if (stage_ns) for (int k = 0; k < 3; ++k) { uint64_t sum = 0; for (unsigned gi = 0; gi < n_gangs; ++gi) sum += gns[gi][k]; stage_ns[k] += sum / n_gangs; }
#ifndef RTK_HOSTSIM
    if (stats1) { stats1[12] += g_h2d_bytes - h2d0; stats1[13] += g_d2h_bytes - d2h0; stats1[14] += g_launches - launches0; }
#endif
}

}  // namespace rtk

using namespace rtk;

extern "C" int rtk_correct_batch(rtk_ctx* ctx, const rtk_opt* opt, int pass, uint32_t n_reads, const char* seq_pool, const uint64_t* seq_off,
                                 const char* qual_pool, const uint64_t* qual_off, char** out_seq_pool, char** out_qual_pool,
                                 uint64_t** out_off, uint64_t* stats) {
    return guarded([&] {
        if (!ctx || !opt || !seq_pool || !seq_off || !out_seq_pool || !out_qual_pool || !out_off) throw std::invalid_argument("null argument");
        if (pass != 1 && pass != 2) throw std::invalid_argument("pass must be 1 (k1 graph, short-read colours) or 2 (k2 graph, long-read colours)");
        DeviceBind bind(ctx);
        std::vector<std::string> os, oq;
        correct_batch_host(ctx, *opt, pass, n_reads, seq_pool, seq_off, qual_pool, qual_off, os, oq, stats);
        uint64_t total = 0;
        for (const auto& x : os) total += x.size();
        *out_off = (uint64_t*)malloc((size_t)(n_reads + 1) * 8);
        *out_seq_pool = (char*)malloc(total + 1);
        *out_qual_pool = (char*)malloc(total + 1);
        if (!*out_off || !*out_seq_pool || !*out_qual_pool) throw std::bad_alloc();
        uint64_t t = 0;
        for (uint32_t r = 0; r < n_reads; ++r) {
            (*out_off)[r] = t;
            if (os[r].size() != oq[r].size()) throw std::runtime_error("corrected sequence and quality lengths differ");
            t += os[r].size();
        }
        (*out_off)[n_reads] = t;
        char* sp = *out_seq_pool; char* qp = *out_qual_pool; const uint64_t* oo = *out_off;
        parallel_for(n_reads, [&](size_t rb, size_t re) {
            for (size_t r = rb; r < re; ++r) { memcpy(sp + oo[r], os[r].data(), os[r].size()); memcpy(qp + oo[r], oq[r].data(), oq[r].size()); }
        });
    });
}

static void strings_to_pools(const std::vector<std::string>& os, const std::vector<std::string>& oq, uint32_t n_reads, char** sp_out, char** qp_out, uint64_t** off_out) {
    uint64_t total = 0;
    for (const auto& x : os) total += x.size();
    *off_out = (uint64_t*)malloc((size_t)(n_reads + 1) * 8);
    *sp_out = (char*)malloc(total + 1);
    *qp_out = (char*)malloc(total + 1);
    if (!*off_out || !*sp_out || !*qp_out) { free(*off_out); free(*sp_out); free(*qp_out); *off_out = nullptr; *sp_out = *qp_out = nullptr; throw std::bad_alloc(); }
    uint64_t t = 0;
    for (uint32_t r = 0; r < n_reads; ++r) {
        (*off_out)[r] = t;
        if (os[r].size() != oq[r].size()) throw std::runtime_error("corrected sequence and quality lengths differ");
        t += os[r].size();
    }
    (*off_out)[n_reads] = t;
    char* sp = *sp_out; char* qp = *qp_out; const uint64_t* oo = *off_out;
    parallel_for(n_reads, [&](size_t rb, size_t re) {
        for (size_t r = rb; r < re; ++r) { memcpy(sp + oo[r], os[r].data(), os[r].size()); memcpy(qp + oo[r], oq[r].data(), oq[r].size()); }
    });
}

extern "C" int rtk_correct_two_pass_batch(rtk_ctx* ctx1, rtk_ctx* ctx2, const rtk_opt* opt1, const rtk_opt* opt2, uint32_t n_reads, const char* seq_pool,
                                          const uint64_t* seq_off, const char* qual_pool, const uint64_t* qual_off, char** out_seq_pool,
                                          char** out_qual_pool, uint64_t** out_off, char** p1_seq_pool, char** p1_qual_pool, uint64_t** p1_off,
                                          uint64_t* stats1, uint64_t* stats2, uint64_t* stage_ns) {
    return guarded([&] {
        if (!ctx1 || !ctx2 || !opt1 || !opt2 || !seq_pool || !seq_off || !qual_pool || !qual_off || !out_seq_pool || !out_qual_pool || !out_off)
            throw std::invalid_argument("null argument");
        if ((p1_seq_pool || p1_qual_pool || p1_off) && !(p1_seq_pool && p1_qual_pool && p1_off)) throw std::invalid_argument("give all three pass-1 outputs or none");
        DeviceBind bind(ctx1);
        std::vector<std::string> s1, q1, s2, q2;
        correct_two_pass_host(ctx1, ctx2, *opt1, *opt2, n_reads, seq_pool, seq_off, qual_pool, qual_off, s1, q1, s2, q2, stats1, stats2, stage_ns);
        strings_to_pools(s2, q2, n_reads, out_seq_pool, out_qual_pool, out_off);
        if (p1_off) strings_to_pools(s1, q1, n_reads, p1_seq_pool, p1_qual_pool, p1_off);
    });
}
