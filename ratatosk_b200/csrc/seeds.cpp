// seeds.cpp — anchor extraction: getSeeds (src/Graph.cpp:3-482) and keep_non_overlap
// (src/Alignment.cpp:1017-1199) restated over the flat graph.
//
// The k-mer lookups (>99% of the reference's time in this function) run on the GPU (K1); what
// is left here is the order-defining host logic the survey recommends keeping on the host
// (SURVEY.md §7 "Sort tie-breaks"): sort by (position, mapped k-mer), solid/weak split, overlap
// pruning, colour-consistency of adjacent solid runs.  All graph state is read from the host
// mirror of the slab.
#include <algorithm>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "kmer.cuh"
#include "k1_lookup_layout.h"
#include "rtk_internal.hpp"
#include "rtk_host_common.hpp"

namespace rtk {

namespace {

typedef rtk_u128 KW;  // host side always uses the wide type (k <= 64)

struct Anchor {
    rtk_hit h;
    KW km;       // mapped k-mer in read orientation (mappedSequenceToString as a number)
    bool empty;  // const_UnitigMap::isEmpty
    bool exact;  // found by the exact sweep: the read's own k-mer, solid by construction (km is only fetched when an inexact hit may
                 // share its position, i.e. in pass 1, where it breaks the sort tie of Graph.cpp:9-14)
};

inline KW mapped_kmer(const rtk_graph_view& g, const rtk_hit& h) {
    const KW fw = rtk_pool_kmer<KW>(g.pool, g.unitig_off[h.unitig] + h.dist, (int)g.k);
    return h.strand ? fw : KmerOps<KW>::rc(fw, (int)g.k);
}

inline bool anchor_less(const Anchor& a, const Anchor& b) {  // Graph.cpp:9-14
    if (a.h.pos == b.h.pos) return a.km < b.km;
    return a.h.pos < b.h.pos;
}

inline bool same_hit(const rtk_hit& a, const rtk_hit& b) {
    return a.pos == b.pos && a.unitig == b.unitig && a.dist == b.dist && a.strand == b.strand;
}

inline bool is_branching(const rtk_graph_view& g, uint32_t u) { return (g.kmcov[u] >> 63) & 1ULL; }

struct IdSpan {
    const uint32_t* p;
    uint64_t n;
};
inline IdSpan global_ids(const rtk_graph_view& g, uint32_t u) {
    const uint32_t gs = g.gset_of[u];
    if (gs == RTK_NONE32) return {nullptr, 0};
    return {g.gset_ids + g.gset_off[gs], g.gset_off[gs + 1] - g.gset_off[gs]};
}
inline IdSpan local_ids(const rtk_graph_view& g, uint32_t u) { return {g.loc_ids + g.loc_off[u], g.loc_off[u + 1] - g.loc_off[u]}; }

// |a ∩ b| capped at `cap` (sorted id lists)
uint64_t inter_capped(const uint32_t* a, uint64_t na, const uint32_t* b, uint64_t nb, uint64_t cap) {
    uint64_t i = 0, j = 0, c = 0;
    while (i < na && j < nb && c < cap) {
        if (a[i] == b[j]) { ++c; ++i; ++j; }
        else if (a[i] < b[j]) ++i;
        else ++j;
    }
    return c;
}

// getNumberSharedPairID(SharedPairID, SharedPairID, min) >= min  (src/Common.cpp:51-71): the two
// parts of a SharedPairID are disjoint, so this is |colours(u) ∩ colours(v)| >= min.
bool share_colors(const rtk_graph_view& g, uint32_t u, uint32_t v, uint64_t min_shared) {
    if (min_shared == 0) return true;
    const IdSpan gu = global_ids(g, u), lu = local_ids(g, u), gv = global_ids(g, v), lv = local_ids(g, v);
    uint64_t c = 0;
    if (gu.n && g.gset_of[u] == g.gset_of[v]) c = gu.n;
    else {
        c += inter_capped(gu.p, gu.n, gv.p, gv.n, min_shared);
        if (c < min_shared) c += inter_capped(lu.p, lu.n, gv.p, gv.n, min_shared - c);
        if (c < min_shared) c += inter_capped(gu.p, gu.n, lv.p, lv.n, min_shared - c);
    }
    if (c < min_shared) c += inter_capped(lu.p, lu.n, lv.p, lv.n, min_shared - c);
    return c >= min_shared;
}

void merge_into(std::vector<uint32_t>& acc, const IdSpan& s) {
    if (!s.n) return;
    std::vector<uint32_t> out;
    out.reserve(acc.size() + s.n);
    std::set_union(acc.begin(), acc.end(), s.p, s.p + s.n, std::back_inserter(out));
    acc.swap(out);
}

// keep_non_overlap (src/Alignment.cpp:1017-1199)
std::vector<Anchor> keep_non_overlap(const rtk_graph_view& g, const char* ref, size_t ref_len, const std::vector<Anchor>& v) {
    const size_t k = g.k;
    struct VarInfo {
        size_t pos_s = 0, pos_e = 0;
        bool keep = true;
        std::vector<uint32_t> idx;      // pos_v
        std::set<uint64_t> unitigs;     // s_km: (unitig, strand)
    };
    std::map<uint64_t, VarInfo> m_var;
    std::string q(k, 'A');
    for (size_t i = 0; i < v.size(); ++i) {
        const rtk_hit& h = v[i].h;
        const char* r = ref + h.pos;
        for (size_t t = 0; t < k; ++t) q[t] = "ACGT"[(int)((v[i].km >> (2 * (k - 1 - t))) & 3)];
        auto match = [&](size_t ro, size_t qo) {  // cstrMatch on the two k-length strings
            size_t n = 0;
            while (ro + n < k && qo + n < k && r[ro + n] == q[qo + n]) ++n;
            return n;
        };
        // cstrMatch stops at the terminator of its FIRST argument (km_ref) or at a mismatch;
        // when the second runs out first its terminator mismatches, so min() bounds both.
        const size_t pref = match(0, 0);
        uint8_t type = 0, mis = 0;
        auto amb = [](char c) -> uint8_t { return c == 'A' ? 1 : c == 'C' ? 2 : c == 'G' ? 4 : c == 'T' ? 8 : 0; };
        if (pref + (pref + 1 <= k ? match(pref + 1, pref + 1) : 0) == k - 1) { type = 1; mis = amb(q[pref]); }
        else if (pref + (pref + 1 <= k ? match(pref, pref + 1) : 0) == k - 1) { type = 2; mis = amb(q[pref]); }
        else if (pref + (pref + 1 <= k ? match(pref + 1, pref) : 0) == k - 1) type = 3;
        if (type != 0 && pref != 0 && pref != k - 1) {
            const uint64_t pos = h.pos + pref;
            const uint64_t key = (pos << 16) | ((uint64_t)mis << 8) | type;
            auto ins = m_var.insert({key, VarInfo()});
            VarInfo& vi = ins.first->second;
            if (ins.second) { vi.pos_s = h.pos; vi.pos_e = h.pos + k; }
            else { vi.pos_s = std::min<size_t>(vi.pos_s, h.pos); vi.pos_e = std::max<size_t>(vi.pos_e, h.pos + k); }
            vi.unitigs.insert(((uint64_t)h.unitig << 1) | h.strand);
            vi.idx.push_back((uint32_t)i);
        }
    }
    std::vector<uint32_t> keep_idx;
    for (auto& it1 : m_var) {
        if (!it1.second.keep) continue;
        const uint64_t p1 = it1.first >> 16;
        const uint64_t lo = (p1 < k - 1) ? 0 : (p1 - k + 1);
        const uint64_t hi = ((p1 + k) >= ref_len) ? ref_len : (p1 + k);
        const uint64_t hi_key = (hi << 16) + 0xffffULL;
        for (auto it2 = m_var.lower_bound(lo << 16); it1.second.keep && it2 != m_var.end() && it2->first <= hi_key; ++it2) {
            const uint64_t p2 = it2->first >> 16;
            const bool ov1 = (p1 >= it2->second.pos_s) && (p1 < it2->second.pos_e);
            const bool ov2 = (p2 >= it1.second.pos_s) && (p2 < it1.second.pos_e);
            if (it1.first != it2->first && (ov1 || ov2)) {
                bool same = false;
                for (uint64_t x : it1.second.unitigs) if (it2->second.unitigs.count(x)) { same = true; break; }
                if (!same) { it1.second.keep = false; it2->second.keep = false; }
            }
        }
        if (it1.second.keep) keep_idx.insert(keep_idx.end(), it1.second.idx.begin(), it1.second.idx.end());
    }
    std::sort(keep_idx.begin(), keep_idx.end());
    keep_idx.erase(std::unique(keep_idx.begin(), keep_idx.end()), keep_idx.end());
    std::vector<Anchor> out;
    out.reserve(keep_idx.size());
    for (uint32_t i : keep_idx) out.push_back(v[i]);
    return out;
}

void remove_empty(std::vector<Anchor>& v) {
    v.erase(std::remove_if(v.begin(), v.end(), [](const Anchor& a) { return a.empty; }), v.end());
}

}  // namespace

void get_seeds_host(rtk_ctx* ctx, const rtk_opt& opt, int pass, uint32_t n_reads, const char* seq_pool,
                    const uint64_t* seq_off, std::vector<std::vector<rtk_hit>>& solid_out,
                    std::vector<std::vector<rtk_hit>>& weak_out, uint64_t* stats) {
    if (!ctx->has_graph || !ctx->host_graph) throw std::invalid_argument("no graph uploaded to this context");
    const rtk_graph_view& g = ctx->host_graph->view;
    if (opt.k != g.k) throw std::invalid_argument("rtk_opt.k does not match the graph's k");
    const size_t k = g.k;
    const bool pass2 = (pass == 2);
    solid_out.assign(n_reads, {});
    weak_out.assign(n_reads, {});
    const bool prof = getenv("RTK_BROKER_PROFILE") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    double laps[5] = {0, 0, 0, 0, 0};
    auto lap = [&](int i) { const auto n = std::chrono::steady_clock::now(); laps[i] = std::chrono::duration_cast<std::chrono::microseconds>(n - t_prev).count() / 1e3; t_prev = n; };

    // 1. exact hits of every read (Graph.cpp:97)
    std::vector<std::vector<rtk_hit>> exact;
    search_sequence_host(ctx, n_reads, seq_pool, seq_off, RTK_SEARCH_EXACT, exact, stats);

    lap(0);
    std::vector<std::vector<Anchor>> v_um(n_reads);
    std::vector<std::string> l_s_of(n_reads);  // pass 1: masked copy of each read that needs the inexact sweep
    parallel_for(n_reads, [&](size_t rb, size_t re) {
    for (size_t r = rb; r < re; ++r) {
        const char* s = seq_pool + seq_off[r];
        const size_t slen = seq_off[r + 1] - seq_off[r];
        if (slen <= k) continue;  // Graph.cpp:49
        std::vector<Anchor>& v = v_um[r];
        v.reserve(exact[r].size());
        if (pass2) { for (const rtk_hit& h : exact[r]) v.push_back({h, (KW)0, false, true}); continue; }   // positions are unique: no tie to break
        for (const rtk_hit& h : exact[r]) v.push_back({h, mapped_kmer(g, h), false, true});
        // 2. pass 1: mask the well-anchored stretches, search the rest inexactly (Graph.cpp:100-196)
        std::string& l_s = l_s_of[r];
        l_s.assign(slen, 'N');
        std::sort(v.begin(), v.end(), anchor_less);
        auto unmask = [&](size_t pos, size_t len) {  // string::replace(pos, len, s, pos, len) semantics
            if (pos > slen) throw std::runtime_error("getSeeds: replace out of range");
            len = std::min(len, slen - pos);
            memcpy(&l_s[pos], s + pos, len);
        };
        for (size_t i = 1; i < v.size(); ++i) {
            if (v[i].h.pos == v[i - 1].h.pos + 1) continue;
            const size_t diff = v[i].h.pos - v[i - 1].h.pos;
            if (diff >= opt.insert_sz) unmask(v[i - 1].h.pos + k, diff - k);
            else if (diff >= (opt.insert_sz / 2)) {
                const size_t space = opt.insert_sz - diff;
                const size_t min_left = (v[i - 1].h.pos < space) ? 0 : (v[i - 1].h.pos - space);
                const size_t max_right = v[i].h.pos + space;
                std::vector<uint32_t> pid_left, pid_right;
                const Anchor* prev = nullptr;
                size_t il = i - 1, ir = i;
                while (il > 0 && v[il].h.pos > min_left) {
                    if (!prev || v[il].h.unitig != prev->h.unitig) {
                        if (!is_branching(g, v[il].h.unitig)) { merge_into(pid_left, global_ids(g, v[il].h.unitig)); merge_into(pid_left, local_ids(g, v[il].h.unitig)); }
                        prev = &v[il];
                    }
                    --il;
                }
                prev = nullptr;
                while (ir < v.size() && v[ir].h.pos < max_right) {
                    if (!prev || v[ir].h.unitig != prev->h.unitig) {
                        if (!is_branching(g, v[ir].h.unitig)) { merge_into(pid_right, global_ids(g, v[ir].h.unitig)); merge_into(pid_right, local_ids(g, v[ir].h.unitig)); }
                        prev = &v[ir];
                    }
                    ++ir;
                }
                if (inter_capped(pid_left.data(), pid_left.size(), pid_right.data(), pid_right.size(), opt.min_cov_vertices) < opt.min_cov_vertices)
                    unmask(v[i - 1].h.pos + k, diff - k);
            }
        }
        if (!v.empty()) {
            if (v.front().h.pos >= opt.insert_sz / 2) unmask(0, v.front().h.pos + k - 1);
            if (slen - v.back().h.pos >= opt.insert_sz / 2) unmask(v.back().h.pos + 1, slen - v.back().h.pos - 1);
        }
    }
    });
    std::string masked;             // concatenated l_s of the reads that need the inexact sweep
    std::vector<uint64_t> moff(1, 0);
    std::vector<uint32_t> mread;    // which read each masked string belongs to
    if (!pass2) {
        size_t tot = 0;
        for (uint32_t r = 0; r < n_reads; ++r) tot += l_s_of[r].size();
        masked.reserve(tot);
        for (uint32_t r = 0; r < n_reads; ++r) {
            if (l_s_of[r].empty()) continue;
            mread.push_back(r);
            masked += l_s_of[r];
            moff.push_back(masked.size());
            std::string().swap(l_s_of[r]);
        }
    }

    lap(1);
    // 3. inexact sweep over the masked strings (Graph.cpp:193)
    if (!mread.empty()) {
        std::vector<std::vector<rtk_hit>> inexact;
        search_sequence_host(ctx, (uint32_t)mread.size(), masked.data(), moff.data(),
                             RTK_SEARCH_INS | RTK_SEARCH_DEL | RTK_SEARCH_SUBST | RTK_SEARCH_OR_EXCL, inexact, stats);
        parallel_for(mread.size(), [&](size_t mb, size_t me) {
            for (size_t m = mb; m < me; ++m) {
                std::vector<Anchor>& v = v_um[mread[m]];
                v.reserve(v.size() + inexact[m].size());
                for (const rtk_hit& h : inexact[m]) v.push_back({h, mapped_kmer(g, h), false, false});
            }
        });
    }

    lap(2);
    const double t0 = (double)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
    // 4. per read: sort, split, prune (Graph.cpp:201-372)
    parallel_for(n_reads, [&](size_t rb, size_t re) {
    for (size_t r = rb; r < re; ++r) {
        const char* s = seq_pool + seq_off[r];
        const size_t slen = seq_off[r + 1] - seq_off[r];
        if (slen <= k) continue;
        std::vector<Anchor>& v = v_um[r];
        std::sort(v.begin(), v.end(), anchor_less);
        std::vector<Anchor> solid, weak;
        for (size_t i = 0; i < v.size(); ++i) {
            if (i != 0 && same_hit(v[i].h, v[i - 1].h)) continue;
            bool eq = true;
            if (!v[i].exact) for (size_t t = 0; t < k && eq; ++t) eq = (s[v[i].h.pos + t] == "ACGT"[(int)((v[i].km >> (2 * (k - 1 - t))) & 3)]);
            (eq ? solid : weak).push_back(v[i]);
        }
        v.clear();
        if (solid.size() >= 2) {  // solid k-mers must overlap by k-1 (Graph.cpp:221-239)
            for (size_t i = 1; i < solid.size(); ++i) {
                if (solid[i].h.pos != solid[i - 1].h.pos + 1 && solid[i].h.pos < solid[i - 1].h.pos + k) {
                    solid[i - 1].empty = true;
                    int64_t j = (int64_t)i - 2;
                    while (j >= 0 && solid[j].h.pos == solid[j + 1].h.pos - 1 && solid[i].h.pos < solid[j].h.pos + k) solid[j--].empty = true;
                }
            }
            remove_empty(solid);
        }
        if (!weak.empty()) weak = keep_non_overlap(g, s, slen, weak);
        // adjacent solid k-mers on different unitigs must be linked and share colours (Graph.cpp:328-372)
        for (size_t i = 1; i < solid.size(); ++i) {
            if (solid[i].h.pos - solid[i - 1].h.pos != 1) continue;
            Anchor& L = solid[i - 1];
            Anchor& R = solid[i];
            if (L.empty || R.empty || L.h.unitig == R.h.unitig) continue;
            // tail of the left unitig (read orientation) must overlap the head of the right one by k-1
            const uint64_t lenL = g.unitig_off[L.h.unitig + 1] - g.unitig_off[L.h.unitig];
            const uint64_t lenR = g.unitig_off[R.h.unitig + 1] - g.unitig_off[R.h.unitig];
            const KW headL = rtk_pool_kmer<KW>(g.pool, g.unitig_off[L.h.unitig], (int)k);
            const KW tailL = rtk_pool_kmer<KW>(g.pool, g.unitig_off[L.h.unitig] + lenL - k, (int)k);
            const KW headR = rtk_pool_kmer<KW>(g.pool, g.unitig_off[R.h.unitig], (int)k);
            const KW tailR = rtk_pool_kmer<KW>(g.pool, g.unitig_off[R.h.unitig] + lenR - k, (int)k);
            const KW tl = L.h.strand ? tailL : KmerOps<KW>::rc(headL, (int)k);
            const KW hr = R.h.strand ? headR : KmerOps<KW>::rc(tailR, (int)k);
            const KW m1 = KmerOps<KW>::mask((int)k - 1);
            bool invalid = ((tl & m1) != (hr >> 2));
            if (!invalid) invalid = !share_colors(g, L.h.unitig, R.h.unitig, opt.min_cov_vertices);
            if (invalid) {
                size_t il = i - 1, ir = i + 1;
                il -= (il != 0);
                const uint32_t uL = L.h.unitig, uR = R.h.unitig;
                while (il > 0 && solid[il].h.pos == solid[il + 1].h.pos - 1 && solid[il].h.unitig == uL) { solid[il].empty = true; --il; }
                while (ir < solid.size() && solid[ir].h.pos == solid[ir - 1].h.pos + 1 && solid[ir].h.unitig == uR) { solid[ir].empty = true; ++ir; }
                L.empty = true;
                R.empty = true;
            }
        }
        remove_empty(solid);
        solid_out[r].reserve(solid.size());
        for (const Anchor& a : solid) solid_out[r].push_back(a.h);
        weak_out[r].reserve(weak.size());
        for (const Anchor& a : weak) weak_out[r].push_back(a.h);
    }
    });
    if (stats) stats[4] += (uint64_t)((double)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count() - t0);
    lap(3);
    if (prof) fprintf(stderr, "[getSeeds] pass %d, %u reads: exact sweep %.1f ms, masks %.1f ms, one-edit sweep %.1f ms, anchors %.1f ms\n", pass, n_reads, laps[0], laps[1], laps[2], laps[3]);
}

}  // namespace rtk
