// seeds_resolve.cpp — see seeds_resolve.hpp.
#include "seeds_resolve.hpp"

#include <algorithm>
#include <unordered_set>

#include <stdexcept>

#include "k1_lookup_layout.h"
#include "rtk_host_common.hpp"
#include "lookup.cuh"

namespace rtk {

namespace {

struct Decoded {
    uint32_t var, pos_s, unitig, off, strand;
    uint64_t P;
};

inline Decoded decode(const rtk_graph_view& g, const RawHit& r) {
    Decoded d;
    d.pos_s = (uint32_t)(r.a & ((1ULL << RTK_HIT_POS_BITS) - 1));
    d.var = (uint32_t)((r.a >> RTK_HIT_POS_BITS) & ((1ULL << RTK_HIT_VAR_BITS) - 1));
    d.P = r.b & RTK_POS_MASK;
    d.strand = (uint32_t)((r.b >> 40) & 1);
    d.unitig = rtk_unitig_of(g.blk2unitig, g.unitig_off, d.P);
    d.off = (uint32_t)(d.P - g.unitig_off[d.unitig]);
    return d;
}

inline bool is_dna(const char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'a' || c == 'c' || c == 'g' || c == 't'; }

struct PosKm {
    uint64_t pos;
    uint64_t km;  // (P<<1)|strand, or ~0 for the empty k-mer
    bool operator==(const PosKm& o) const { return pos == o.pos && km == o.km; }
};
struct PosKmHash {
    size_t operator()(const PosKm& p) const { return (size_t)rtk_mix64(p.pos * 0x9E3779B97F4A7C15ULL ^ rtk_mix64(p.km)); }
};

// length of the unitig run starting at raw[i]: consecutive pos_s, same unitig, offsets +-1
inline size_t run_length(const std::vector<Decoded>& d, size_t i, size_t end, uint32_t k, const rtk_graph_view& g) {
    const uint64_t usize = g.unitig_off[d[i].unitig + 1] - g.unitig_off[d[i].unitig];
    if (usize == k) return 1;  // short / abundant unitigs are never extended (CompactedDBG.tcc:4489)
    size_t len = 1;
    while (i + len < end) {
        const Decoded& p = d[i + len - 1];
        const Decoded& q = d[i + len];
        if (q.pos_s != p.pos_s + 1 || q.unitig != p.unitig || q.strand != p.strand) break;
        if (p.strand ? (q.off != p.off + 1) : (q.off + 1 != p.off)) break;
        ++len;
    }
    return len;
}

}  // namespace

void resolve_exact(const rtk_graph_view& g, const RawHit* raw, size_t n, std::vector<rtk_hit>& out) {
    std::vector<Decoded> d(n);
    for (size_t i = 0; i < n; ++i) d[i] = decode(g, raw[i]);
    size_t i = 0;
    while (i < n) {
        const size_t len = run_length(d, i, n, g.k, g);
        if (d[i].strand) {
            for (size_t j = 0; j < len; ++j) out.push_back({d[i + j].pos_s, d[i].unitig, d[i + j].off, 1u});
        } else {
            // Search.tcc:700: j ascends over unitig offsets, read position descends
            for (size_t j = len; j-- > 0;) out.push_back({d[i + j].pos_s, d[i].unitig, d[i + j].off, 0u});
        }
        i += len;
    }
}

void resolve_inexact(const rtk_graph_view& g, const char* s, uint32_t slen, bool or_exclusive, const RawHit* raw,
                     size_t n, std::vector<rtk_hit>& out) {
    const uint32_t k = g.k;
    std::vector<Decoded> d(n);
    for (size_t i = 0; i < n; ++i) d[i] = decode(g, raw[i]);
    // us_pos_km (Search.tcc:566): insert-if-absent and membership only, at most one key per raw hit -> a flat open-addressing
    // table reused by the thread (std::unordered_set's node allocations were the bulk of this stage)
    static thread_local std::vector<PosKm> tl_tab;
    size_t cap = 64;
    while (cap < 2 * n + 16) cap <<= 1;
    if (tl_tab.size() < cap) tl_tab.resize(cap);
    PosKm* tab = tl_tab.data();
    for (size_t i = 0; i < cap; ++i) tab[i].pos = ~0ULL;
    const size_t tmask = cap - 1;
    auto tab_contains = [&](const PosKm& key) {
        for (size_t h = PosKmHash()(key) & tmask;; h = (h + 1) & tmask) {
            if (tab[h].pos == ~0ULL) return false;
            if (tab[h] == key) return true;
        }
    };
    auto tab_insert = [&](const PosKm& key) {   // true when the key was not there
        for (size_t h = PosKmHash()(key) & tmask;; h = (h + 1) & tmask) {
            if (tab[h].pos == ~0ULL) { tab[h] = key; return true; }
            if (tab[h] == key) return false;
        }
    };
    size_t gi = 0;
    while (gi < n) {
        size_t ge = gi;
        while (ge < n && d[ge].var == d[gi].var) ++ge;
        // variant -> (type, shift)
        const uint32_t var = d[gi].var;
        uint32_t shift;
        bool ins = false, del = false;
        if (var < 4 * k) shift = var / 4;
        else if (var < 8 * k) { ins = true; shift = (var - 4 * k) / 4; }
        else { del = true; shift = var - 8 * k; }
        // Search.tcc:586-589 / :663-664
        auto map_pos = [&](const uint64_t pos) -> uint64_t {
            const uint64_t sp = (pos / k) + ((pos % k) > shift ? 1 : 0);
            if (ins) return pos - sp;
            if (del) return pos + sp;
            return pos;
        };
        size_t i = gi;
        while (i < ge) {
            const Decoded& h = d[i];
            const uint64_t l_pos_s = map_pos(h.pos_s);
            bool cond = (l_pos_s + k - 1 < slen) && is_dna(s[l_pos_s]) && is_dna(s[l_pos_s + k - 1]);
            if (cond && or_exclusive) {
                // rpos is empty (no exact pass in the same call); key = the variant k-mer itself
                cond = !tab_contains(PosKm{l_pos_s, (h.P << 1) | h.strand});
            }
            if (!cond) { ++i; continue; }
            const size_t len = run_length(d, i, ge, k, g);
            // um.dist = smallest unitig offset of the run, um.len = len; j ascends over offsets
            const uint32_t um_dist = h.strand ? h.off : (uint32_t)(h.off - (len - 1));
            for (size_t jj = 0; jj < len; ++jj) {
                const uint32_t j = um_dist + (uint32_t)jj;
                // strand: pos_s + j - dist ; else pos_s + dist + len - j - 1
                const uint64_t p_inexact = h.strand ? (uint64_t)h.pos_s + jj : (uint64_t)h.pos_s + (len - 1 - jj);
                const uint64_t l_pos_seq = map_pos(p_inexact);
                if (l_pos_seq + k - 1 >= slen) continue;
                // getMappedKmer(j): real k-mer only when j < um.len (and j indexes the unitig)
                const uint64_t usz = g.unitig_off[h.unitig + 1] - g.unitig_off[h.unitig];
                uint64_t key = ~0ULL;
                if (j < len) {
                    if (usz == k) { if (j == 0) key = ((g.unitig_off[h.unitig]) << 1) | h.strand; }  // isShort: pos+dist==0
                    else if (j < usz - k + 1) key = ((g.unitig_off[h.unitig] + j) << 1) | h.strand;
                }
                if (tab_insert(PosKm{l_pos_seq, key})) out.push_back({(uint32_t)l_pos_seq, h.unitig, j, h.strand});
            }
            i += len;
        }
        gi = ge;
    }
}

void resolve_exact_dense(const rtk_graph_view& g, uint32_t n_reads, const uint64_t* seq_off, const uint64_t* dense,
                         std::vector<std::vector<rtk_hit>>& per_read) {
    per_read.assign(n_reads, {});
    const uint32_t k = g.k;
    parallel_for(n_reads, [&](size_t rb, size_t re) {
        for (size_t r = rb; r < re; ++r) {
            const uint64_t len = seq_off[r + 1] - seq_off[r];
            if (len < k) continue;
            const uint64_t* d = dense + (seq_off[r] - seq_off[0]);
            const uint32_t npos = (uint32_t)(len - k + 1);
            size_t n_hits = 0;
            for (uint32_t l = 0; l < npos; ++l) n_hits += (d[l] != ~0ULL);
            std::vector<rtk_hit>& out = per_read[r];
            out.reserve(n_hits);
            uint32_t l = 0;
            while (l < npos) {
                if (d[l] == ~0ULL) { ++l; continue; }
                // a findUnitig run (CompactedDBG.tcc:4479-4548): consecutive read positions on consecutive k-mers of one unitig, same
                // strand; a k-mer one pool position further is always in the same unitig (a k-mer never straddles two unitigs)
                const uint64_t P0 = d[l] & RTK_POS_MASK;
                const uint32_t strand = (uint32_t)((d[l] >> 40) & 1);
                const uint32_t u = rtk_unitig_of(g.blk2unitig, g.unitig_off, P0);
                const uint64_t ub = g.unitig_off[u], usize = g.unitig_off[u + 1] - ub;
                const uint32_t off0 = (uint32_t)(P0 - ub);
                uint32_t run = 1;
                if (usize != k) {   // short / abundant unitigs are never extended (CompactedDBG.tcc:4489)
                    if (strand) while (l + run < npos && d[l + run] == ((P0 + run) | (1ULL << 40))) ++run;
                    else while (l + run < npos && run <= off0 && d[l + run] == (P0 - run)) ++run;
                }
                if (strand) for (uint32_t j = 0; j < run; ++j) out.push_back({l + j, u, off0 + j, 1u});
                else for (uint32_t j = run; j-- > 0;) out.push_back({l + j, u, off0 - j, 0u});   // Search.tcc:700: read position descends
                l += run;
            }
        }
    });
}

void resolve_batch(const rtk_graph_view& hv, uint32_t n_reads, const char* seq_pool, const uint64_t* seq_off, uint32_t flags,
                   RawHitVec& raw, std::vector<std::vector<rtk_hit>>& per_read) {
    per_read.assign(n_reads, {});
    // bucket the raw hits by read (parallel counting sort on the read id: per-chunk histograms, one prefix pass, parallel
    // scatter), then sort + replay each read independently
    const int sh = RTK_HIT_VAR_BITS + RTK_HIT_POS_BITS;
    const size_t n_raw = raw.size();
    const size_t n_chunks = std::max<size_t>(1, std::min<size_t>(host_threads(), n_raw / 65536 + 1));
    const size_t chunk = (n_raw + n_chunks - 1) / n_chunks;
    std::vector<std::vector<uint32_t>> cnt(n_chunks);
    parallel_for(n_chunks, [&](size_t cb, size_t ce) {
        for (size_t c = cb; c < ce; ++c) {
            std::vector<uint32_t>& h = cnt[c];
            h.assign(n_reads, 0);
            const size_t lo = c * chunk, hi = std::min(n_raw, lo + chunk);
            for (size_t i = lo; i < hi; ++i) {
                const uint32_t r = (uint32_t)(raw[i].a >> sh);
                if (r >= n_reads) throw std::runtime_error("corrupt hit record");
                ++h[r];
            }
        }
    });
    std::vector<uint64_t> start(n_reads + 1, 0);
    for (uint32_t r = 0; r < n_reads; ++r) {
        uint64_t t = 0;
        for (size_t c = 0; c < n_chunks; ++c) t += cnt[c][r];
        start[r + 1] = start[r] + t;
    }
    RawHitVec by_read(n_raw);
    {
        // write cursor of chunk c for read r = start[r] + hits of r in the chunks before c
        std::vector<std::vector<uint64_t>> cur(n_chunks);
        for (size_t c = 0; c < n_chunks; ++c) cur[c].resize(n_reads);
        for (uint32_t r = 0; r < n_reads; ++r) {
            uint64_t t = start[r];
            for (size_t c = 0; c < n_chunks; ++c) { cur[c][r] = t; t += cnt[c][r]; }
        }
        parallel_for(n_chunks, [&](size_t cb, size_t ce) {
            for (size_t c = cb; c < ce; ++c) {
                std::vector<uint64_t>& w = cur[c];
                const size_t lo = c * chunk, hi = std::min(n_raw, lo + chunk);
                for (size_t i = lo; i < hi; ++i) by_read[w[(uint32_t)(raw[i].a >> sh)]++] = raw[i];
            }
        });
    }
    const bool exact = flags & RTK_SEARCH_EXACT;
    parallel_for(n_reads, [&](size_t b, size_t e) {
        for (size_t r = b; r < e; ++r) {
            const size_t n = start[r + 1] - start[r];
            if (!n) continue;
            RawHit* p = by_read.data() + start[r];
            std::sort(p, p + n, [](const RawHit& x, const RawHit& y) { return x.a < y.a; });
            if (exact) resolve_exact(hv, p, n, per_read[r]);
            else resolve_inexact(hv, seq_pool + seq_off[r], (uint32_t)(seq_off[r + 1] - seq_off[r]), (flags & RTK_SEARCH_OR_EXCL) != 0,
                                 p, n, per_read[r]);
        }
    });
}

}  // namespace rtk
