// oracle/rtk_oracle.cpp — CPU RESTATEMENT of the reference algorithms on the hot path.
//
// TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs as the checker; the product (librtk_b200.so) neither
// links nor loads it.  Written for clarity, on plain strings and hash maps, deliberately
// sharing no code with ratatosk_b200/csrc.
//
// Pinning: tests/test_oracle.py checks every function here against the golden vectors under
// tests/golden/ that were recorded from the UNMODIFIED reference (oracle/_ref, built from
// /root/reference by oracle/Makefile; generator tests/golden/make_golden.py).
//
//   orc_search_sequence   CompactedDBG::searchSequence        Bifrost/src/Search.tcc:526-768
//                         findUnitig / jump                   Bifrost/src/CompactedDBG.tcc:4479-4548,
//                                                             Bifrost/src/CompressedSequence.cpp:534-558
//                         getMappedKmer / getKmerMapping      Bifrost/src/UnitigMap.tcc:204-248
//                         KmerIterator                        Bifrost/src/KmerIterator.cpp:6-60
//   orc_edit_distance     edlibAlign, distance + end locations src/edlib.cpp:141-296, 547-931
//                         (restated as the plain O(nm) recurrence Myers' bit-vectors compute;
//                         equality = identity + the 28 IUPAC pairs of src/Common.hpp:262-276,
//                         added only when both letters occur, src/edlib.cpp:64-81)
#include "rtk_oracle.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

struct Loc {
    uint32_t unitig, off, strand;
};

struct OGraph {
    int k;
    std::vector<std::string> unitigs;
    std::unordered_map<std::string, Loc> dict;  // k-mer spelled as it would appear in a read
};

inline bool is_dna(char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'a' || c == 'c' || c == 'g' || c == 't'; }

std::string revcomp(const std::string& s) {
    std::string r(s.rbegin(), s.rend());
    for (auto& c : r) c = (c == 'A') ? 'T' : (c == 'C') ? 'G' : (c == 'G') ? 'C' : (c == 'T') ? 'A' : c;
    return r;
}

struct Hit {
    uint32_t pos, unitig, dist, strand;
};

// A match as findUnitig returns it: the k-mer at s[pos] lies on `unitig`; extended along the unitig
// while the following characters agree (jump); dist = smallest unitig offset, len = #k-mers.
struct Match {
    bool found = false;
    uint32_t unitig = 0, dist = 0, len = 0, strand = 0;
};

Match find_unitig(const OGraph& g, const std::string& s, size_t pos) {
    Match m;
    const size_t k = g.k;
    if (s.size() < k || pos > s.size() - k) return m;
    for (size_t i = pos; i < pos + k; ++i) if (!is_dna(s[i])) return m;
    const auto it = g.dict.find(s.substr(pos, k));
    if (it == g.dict.end()) return m;
    m.found = true;
    m.unitig = it->second.unitig; m.strand = it->second.strand; m.dist = it->second.off; m.len = 1;
    const std::string& u = g.unitigs[m.unitig];
    if (u.size() == k) return m;  // isShort / isAbundant unitigs are not extended
    if (m.strand) {
        size_t up = m.dist + k, sp = pos + k;  // next unitig base / next read base
        while (sp < s.size() && up < u.size() && s[sp] == u[up]) { ++sp; ++up; ++m.len; }
    } else {
        // read runs against the unitig: next read base must be the complement of the base before the k-mer
        size_t sp = pos + k;
        int64_t up = (int64_t)m.dist - 1;
        const auto comp = [](char c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A'; };
        while (sp < s.size() && up >= 0 && s[sp] == comp(u[(size_t)up])) { ++sp; --up; ++m.len; }
        m.dist = m.dist - (m.len - 1);
    }
    return m;
}

// key UnitigMap::getMappedKmer(j) would produce for the match (empty string = empty k-mer)
std::string mapped_kmer_key(const OGraph& g, const Match& m, size_t j) {
    const size_t k = g.k;
    const std::string& u = g.unitigs[m.unitig];
    if (j >= m.len) return std::string();
    if (u.size() == k) { if (j + m.dist != 0) return std::string(); }
    else if (j >= u.size() - k + 1) return std::string();
    const std::string km = (u.size() == k) ? u : u.substr(j, k);
    return m.strand ? km : revcomp(km);
}

}  // namespace

extern "C" {

void* orc_graph_create(int k, uint64_t n, const char* const* unitigs) {
    OGraph* g = new OGraph();
    g->k = k;
    for (uint64_t i = 0; i < n; ++i) g->unitigs.emplace_back(unitigs[i]);
    for (uint32_t u = 0; u < g->unitigs.size(); ++u) {
        const std::string& s = g->unitigs[u];
        for (size_t j = 0; j + k <= s.size(); ++j) {
            const std::string km = s.substr(j, k);
            g->dict[km] = Loc{u, (uint32_t)j, 1};
            g->dict[revcomp(km)] = Loc{u, (uint32_t)j, 0};
        }
    }
    return g;
}

void orc_graph_free(void* h) { delete (OGraph*)h; }
void orc_free(void* p) { free(p); }

int64_t orc_search_sequence(void* h, const char* s_, int exact, int ins, int del, int subst, int or_excl,
                            uint32_t** out, uint64_t* n_lookups) {
    const OGraph& g = *(const OGraph*)h;
    const size_t k = g.k;
    const std::string s(s_);
    std::vector<Hit> v;
    uint64_t lookups = 0;
    if (s.size() < k) { *out = nullptr; if (n_lookups) *n_lookups = 0; return 0; }
    std::set<std::pair<size_t, std::string>> us_pos_km;
    std::set<size_t> rpos;
    static const char alpha[4] = {'A', 'C', 'G', 'T'};

    // next position >= p whose k-mer has only ACGT (KmerIterator), or npos; `bad` = prefix count of non-ACGT
    std::vector<uint32_t> bad;
    auto index_bad = [&](const std::string& t) {
        bad.assign(t.size() + 1, 0);
        for (size_t i = 0; i < t.size(); ++i) bad[i + 1] = bad[i] + (is_dna(t[i]) ? 0 : 1);
    };
    auto next_valid = [&](const std::string& t, size_t p) -> size_t {
        while (p + k <= t.size()) {
            if (bad[p + k] == bad[p]) return p;
            ++p;
        }
        return std::string::npos;
    };

    auto worker = [&](bool w_subst, bool w_ins, bool w_del, size_t shift, std::string s_inexact) {
        for (int i = 0; i != ((w_subst || w_ins) ? 4 : 1); ++i) {
            if (w_ins) { for (size_t j = shift; j < s_inexact.size(); j += k) s_inexact[j] = alpha[i]; }
            else if (w_subst) {
                for (size_t j = shift; j < s_inexact.size(); j += k) s_inexact[j] = (!is_dna(s[j]) || alpha[i] == s[j]) ? 'N' : alpha[i];
            }
            auto map_pos = [&](size_t p) {
                const size_t sp = (p / k) + ((p % k) > shift ? 1 : 0);
                if (w_ins) return p - sp;
                if (w_del) return p + sp;
                return p;
            };
            index_bad(s_inexact);
            size_t p = next_valid(s_inexact, 0);
            while (p != std::string::npos) {
                const size_t l = map_pos(p);
                size_t adv = 1;
                if (l + k - 1 < s.size() && is_dna(s[l]) && is_dna(s[l + k - 1]) &&
                    (!or_excl || (!rpos.count(l) && !us_pos_km.count({l, s_inexact.substr(p, k)})))) {
                    ++lookups;
                    const Match m = find_unitig(g, s_inexact, p);
                    if (m.found) {
                        for (size_t j = m.dist; j < (size_t)m.dist + m.len; ++j) {
                            const size_t pi = m.strand ? (p + j - m.dist) : (p + m.dist + m.len - j - 1);
                            const size_t lp = map_pos(pi);
                            if (lp + k - 1 < s.size() && us_pos_km.insert({lp, mapped_kmer_key(g, m, j)}).second)
                                v.push_back(Hit{(uint32_t)lp, m.unitig, (uint32_t)j, m.strand});
                        }
                        adv = m.len;
                    }
                }
                p = next_valid(s_inexact, p + adv);
            }
        }
    };

    if (exact) {
        index_bad(s);
        size_t p = next_valid(s, 0);
        while (p != std::string::npos) {
            ++lookups;
            const Match m = find_unitig(g, s, p);
            size_t adv = 1;
            if (m.found) {
                for (size_t j = m.dist; j < (size_t)m.dist + m.len; ++j) {
                    const size_t pi = m.strand ? (p + j - m.dist) : (p + m.dist + m.len - j - 1);
                    v.push_back(Hit{(uint32_t)pi, m.unitig, (uint32_t)j, m.strand});
                }
                adv = m.len;
            }
            p = next_valid(s, p + adv);
        }
        if (or_excl && (ins || del || subst)) {
            for (const Hit& x : v) {
                Match m; m.found = true; m.unitig = x.unitig; m.dist = x.dist; m.len = 1; m.strand = x.strand;
                us_pos_km.insert({x.pos, mapped_kmer_key(g, m, x.dist)});
                rpos.insert(x.pos);
            }
        }
    }
    if (subst) for (size_t i = 0; i != k; ++i) worker(true, false, false, i, s);
    if (ins) {
        for (size_t i = 0; i != k; ++i) {
            std::string t(s, 0, i);
            for (size_t j = i, cpt = 0; j < s.size(); ++j, ++cpt) {
                if (cpt % (k - 1) == 0) t.push_back('A');
                t.push_back(s[j]);
            }
            worker(false, true, false, i, t);
        }
    }
    if (del && s.size() >= k + 1) {
        for (size_t i = 0; i != k + 1; ++i) {
            std::string t(s, 0, std::min(i, s.size()));
            for (size_t j = i, cpt = 0; j < s.size(); ++j, ++cpt) if (cpt % (k + 1) != 0) t.push_back(s[j]);
            worker(false, false, true, i, t);
        }
    }
    uint32_t* o = (uint32_t*)malloc(sizeof(uint32_t) * 4 * (v.size() + 1));
    for (size_t i = 0; i < v.size(); ++i) { o[4 * i] = v[i].pos; o[4 * i + 1] = v[i].unitig; o[4 * i + 2] = v[i].dist; o[4 * i + 3] = v[i].strand; }
    *out = o;
    if (n_lookups) *n_lookups = lookups;
    return (int64_t)v.size();
}

// ---------------------------------------------------------------------------- edit distance
static bool iupac_eq(char a, char b, const bool present[256]) {
    if (a == b) return true;
    static const char* pairs[28] = {"MA", "MC", "RA", "RG", "SC", "SG", "VA", "VC", "VG", "WA", "WT", "YC", "YT", "HA",
                                    "HC", "HT", "KG", "KT", "DA", "DG", "DT", "BC", "BG", "BT", "NA", "NC", "NG", "NT"};
    for (int i = 0; i < 28; ++i) {
        const char x = pairs[i][0], y = pairs[i][1];
        if (!present[(unsigned char)x] || !present[(unsigned char)y]) continue;
        if ((a == x && b == y) || (a == y && b == x)) return true;
    }
    return false;
}

int orc_edit_distance(const char* q, int ql, const char* t, int tl, int mode, int kmax, int iupac, int* dist,
                      int** ends, int* n_ends) {
    *ends = nullptr; *n_ends = 0; *dist = -1;
    if (ql == 0) {  // src/edlib.cpp:156-170
        *dist = (mode == 0) ? tl : 0;  // no k check on this path (src/edlib.cpp:161-176)
        *ends = (int*)malloc(sizeof(int));
        (*ends)[0] = (mode == 0) ? tl - 1 : -1;
        *n_ends = 1;
        return 0;
    }
    if (tl == 0) {
        *dist = ql;
        *ends = (int*)malloc(sizeof(int));
        (*ends)[0] = -1;
        *n_ends = 1;
        return 0;
    }
    bool present[256];
    memset(present, 0, sizeof(present));
    for (int i = 0; i < ql; ++i) present[(unsigned char)q[i]] = true;
    for (int j = 0; j < tl; ++j) present[(unsigned char)t[j]] = true;
    // column-wise DP: D[i] = distance of query[0..i) vs a target prefix/substring ending at column j
    std::vector<int> col(ql + 1), prev(ql + 1);
    for (int i = 0; i <= ql; ++i) prev[i] = i;
    int best = -1;
    std::vector<int> locs;
    // The reference pads the query to a multiple of 64 with W wildcards and reads column c as position
    // c - W (src/edlib.cpp:658-692): whenever W > 0 the "position -1" (query against the empty target
    // prefix, score = query length) takes part in the minimum for SHW/HW and is reported first.
    if (mode != 0 && (ql % 64) != 0) { best = ql; locs.push_back(-1); }
    for (int j = 1; j <= tl; ++j) {
        col[0] = (mode == 2) ? 0 : j;  // HW: free start anywhere in the target
        for (int i = 1; i <= ql; ++i) {
            const bool eq = iupac ? iupac_eq(q[i - 1], t[j - 1], present) : (q[i - 1] == t[j - 1]);
            col[i] = std::min(std::min(col[i - 1] + 1, prev[i] + 1), prev[i - 1] + (eq ? 0 : 1));
        }
        if (mode != 0) {  // SHW / HW: every end column with the best score
            const int d = col[ql];
            if (best < 0 || d < best) { best = d; locs.clear(); locs.push_back(j - 1); }
            else if (d == best) locs.push_back(j - 1);
        }
        std::swap(col, prev);
    }
    if (mode == 0) { best = prev[ql]; locs.assign(1, tl - 1); }
    if (kmax >= 0 && best > kmax) return 0;
    *dist = best;
    *n_ends = (int)locs.size();
    *ends = (int*)malloc(sizeof(int) * (locs.size() + 1));
    for (size_t i = 0; i < locs.size(); ++i) (*ends)[i] = locs[i];
    return 0;
}

}  // extern "C"
