// phasing.cpp — phasing() of the reference's second pass (src/Graph.cpp:869-1097), run by the multi-thread branch of
// search() on every read before getSeeds (src/Ratatosk.cpp:832): stretches of the pass-1 corrected read that map to
// unitigs whose long-read colour sets are compatible with NO other stretch of the read further than insert_sz away
// are reverted to the raw read.
//
//   1. exact k-mer sweep of the corrected read (K1 exact kernel, batched over the reads of the call); consecutive hits
//      on one unitig are one findUnitig() run (:893-915)
//   2. a TinyBloomFilter (src/TinyBloomFilter.hpp; wyhash final version 3 of the 8-byte id, double hashing) of each
//      kept run's colour ids, all-pairs "shares >= 85 % of the bits both ways" test in the reference's scan order
//      (:918-983) -> positions to revert (pos2rm)
//   3. NW alignment path raw vs corrected for the WHOLE read (K5, edlib's divide-and-conquer above 1 MiB of DP state),
//      walked run by run like the CIGAR in :986-1052
//   4. exact sweep of the reverted neighbourhoods; bases of the result covered by a graph k-mer get the maximum
//      quality (:1054-1080; K1 exact kernel, batched)
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "broker.hpp"
#include "rtk_host_common.hpp"
#include "traceback_host.hpp"
#include "traverse.hpp"

namespace rtk {
namespace {

// wyhash final version 3 (Wang Yi, public domain; Bifrost/src/wyhash.h) of one 8-byte little-endian key
inline uint64_t wymix(uint64_t a, uint64_t b) {
    const unsigned __int128 r = (unsigned __int128)a * b;
    return (uint64_t)r ^ (uint64_t)(r >> 64);
}
inline uint64_t wyhash8(uint64_t key, uint64_t seed) {
    static const uint64_t s0 = 0xa0761d6478bd642full, s1 = 0xe7037ed1a0b428dbull;
    seed ^= s0;
    const uint64_t lo = key & 0xffffffffull, hi = key >> 32;
    const uint64_t a = (lo << 32) | hi, b = (hi << 32) | lo;
    return wymix(s1 ^ 8ull, wymix(a ^ s1, b ^ seed));
}

inline uint64_t round_up_pow2(uint64_t v) { --v; v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16; v |= v >> 32; return ++v; }

// TinyBloomFilter<size_t>(nb_elem, bits_per_elem): geometry shared by all filters of a read
struct BloomGeom {
    uint64_t nb_h = 0, bits = 0;
    BloomGeom(uint64_t nb_elem, uint64_t bits_per_elem) {
        if (nb_elem == 0 || bits_per_elem == 0) return;
        auto fpp = [](uint64_t b, uint64_t h) { const double bd = (double)b, hd = (double)h; return std::pow(1 - std::exp(-(hd / bd)), hd); };
        nb_h = (uint64_t)(bits_per_elem * std::log(2));
        nb_h += (uint64_t)(fpp(bits_per_elem, nb_h) >= fpp(bits_per_elem, nb_h + 1));
        nb_h &= 0xffull;
        bits = std::max<uint64_t>(round_up_pow2(bits_per_elem * nb_elem), 64);
    }
    size_t words() const { return (size_t)(bits / 64); }
};
inline void bloom_insert(uint64_t* table, const BloomGeom& gm, uint64_t id) {   // insert(): every one of the nb_h positions ends up set
    const uint64_t mask = gm.bits - 1, h2 = wyhash8(id, 1610612741ull);
    uint64_t h1 = wyhash8(id, 49157ull);
    for (uint64_t i = 0; i != gm.nb_h; ++i) { table[(h1 & mask) >> 6] |= 1ull << (h1 & 0x3full); h1 += h2; }
}

struct Run { uint32_t pos, len, unitig; };

inline bool is_branching(const rtk_graph_view& g, uint32_t u) { return (g.kmcov[u] >> 63) & 1ULL; }

}  // namespace

// fixSNPs (src/Alignment.cpp:846-964) for a batch: the positions that are not A/C/G/T (Bifrost isDNA: either case) are listed
// here, the order-dependent resolution runs one warp per read on the device (fixsnps.cuh)
uint64_t fix_snps_batch_host(rtk_ctx* ctx, uint32_t n, char* seq_pool, const uint64_t* seq_off) {
    if (!ctx->has_graph || !ctx->host_graph) throw std::invalid_argument("no graph uploaded to this context");
    std::vector<std::vector<uint32_t>> per(n);
    parallel_for(n, [&](size_t rb, size_t re) {
        for (size_t r = rb; r < re; ++r) {
            const char* s = seq_pool + seq_off[r];
            const uint64_t L = seq_off[r + 1] - seq_off[r];
            if (L >= 0xFFFFFFFFull) throw std::invalid_argument("fixSNPs: read longer than 4 Gbases");
            for (uint64_t i = 0; i < L; ++i) {
                const char c = (char)(s[i] & 0xDF);
                if (c != 'A' && c != 'C' && c != 'G' && c != 'T') per[r].push_back((uint32_t)i);
            }
        }
    });
    std::vector<uint64_t> amb_off(n + 1, 0);
    for (uint32_t r = 0; r < n; ++r) amb_off[r + 1] = amb_off[r] + per[r].size();
    std::vector<uint32_t> amb_pos(amb_off[n]);
    for (uint32_t r = 0; r < n; ++r) if (!per[r].empty()) memcpy(amb_pos.data() + amb_off[r], per[r].data(), per[r].size() * 4);
    uint64_t fixed = 0;
    fix_snps_host(ctx, n, seq_pool, seq_off, amb_pos, amb_off, &fixed);
    return fixed;
}

void phasing_batch_host(rtk_ctx* ctx, const rtk_opt& opt, uint32_t n, const char* raw_pool, const uint64_t* raw_off, const char* corr_pool,
                        const uint64_t* corr_off, const char* qual_pool, const uint64_t* qual_off, std::vector<std::string>& out_seq,
                        std::vector<std::string>& out_qual) {
    if (!ctx->has_graph || !ctx->host_graph) throw std::invalid_argument("no graph uploaded to this context");
    const rtk_graph_view& g = ctx->host_graph->view;
    if (opt.k != g.k) throw std::invalid_argument("rtk_opt.k does not match the graph's k");
    const size_t k = g.k;
    const char q_min = rtk_get_qual(0.0, 0, opt.max_qual), q_max = rtk_get_qual(1.0, 0, opt.max_qual);
    const double t_bits_sim = 0.85;
    const size_t max_limit_nb_pids = 1000, nb_bits_elem_tbf = 14;
    out_seq.assign(n, std::string());
    out_qual.assign(n, std::string());
    const bool prof = getenv("RTK_BROKER_PROFILE") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!prof) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[phasing] %s: %.1f ms\n", what, std::chrono::duration_cast<std::chrono::microseconds>(now - t_prev).count() / 1e3);
        t_prev = now;
    };

    // 0. `-f` (Correct_Opt::force_unres_snp_corr): fixSNPs on the pass-1 read before anything else (src/Ratatosk.cpp:828)
    std::string fixed_pool;
    std::vector<uint64_t> fixed_off;
    if (opt.force_unres_snp_corr) {
        fixed_pool.assign(corr_pool + corr_off[0], (size_t)(corr_off[n] - corr_off[0]));
        fixed_off.resize(n + 1);
        for (uint32_t i = 0; i <= n; ++i) fixed_off[i] = corr_off[i] - corr_off[0];
        fix_snps_batch_host(ctx, n, &fixed_pool[0], fixed_off.data());
        corr_pool = fixed_pool.data();
        corr_off = fixed_off.data();
        lap("fixSNPs");
    }
    // 1. map the corrected reads (one exact sweep for the whole batch)
    std::vector<std::vector<rtk_hit>> hits;
    {
        std::vector<uint64_t> off(n + 1);
        for (uint32_t i = 0; i <= n; ++i) off[i] = corr_off[i] - corr_off[0];
        search_sequence_host(ctx, n, corr_pool + corr_off[0], off.data(), RTK_SEARCH_EXACT, hits, nullptr);
    }

    lap("exact sweep of the corrected reads");
    // 2. positions to revert, per read
    std::vector<std::vector<uint8_t>> pos2rm(n);
    parallel_for(n, [&](size_t rb, size_t re) {
    for (size_t r = rb; r < re; ++r) {
        const size_t clen = (size_t)(corr_off[r + 1] - corr_off[r]);
        pos2rm[r].assign(clen + k + 1, 0);
        std::vector<Run> v_um;
        size_t max_nb_pids = 0;
        // searchSequence lists the k-mers of a reverse-strand run in descending read position (Search.tcc:700): back to read order
        std::vector<rtk_hit> h = hits[r];
        std::sort(h.begin(), h.end(), [](const rtk_hit& x, const rtk_hit& y) { return x.pos < y.pos; });
        for (size_t i = 0; i < h.size();) {   // findUnitig() runs: consecutive k-mers of the read on consecutive k-mers of one unitig
            size_t j = i + 1;
            const bool single = (g.unitig_off[h[i].unitig + 1] - g.unitig_off[h[i].unitig]) == k;   // short / abundant unitigs are never extended (CompactedDBG.tcc:4489)
            while (!single && j < h.size() && h[j].pos == h[j - 1].pos + 1 && h[j].unitig == h[i].unitig && h[j].strand == h[i].strand &&
                   (h[i].strand ? h[j].dist == h[j - 1].dist + 1 : h[j].dist + 1 == h[j - 1].dist)) ++j;
            const uint32_t u = h[i].unitig;
            const uint64_t gs = (g.gset_of[u] == RTK_NONE32) ? 0 : (g.gset_off[g.gset_of[u] + 1] - g.gset_off[g.gset_of[u]]);
            const size_t card = (size_t)(gs + (g.loc_off[u + 1] - g.loc_off[u]));
            if (!is_branching(g, u) && card <= max_limit_nb_pids) {
                v_um.push_back({h[i].pos, (uint32_t)(j - i), u});
                max_nb_pids = std::max(max_nb_pids, card);
            }
            i = j;
        }
        const size_t m = v_um.size();
        const BloomGeom gm(max_nb_pids, nb_bits_elem_tbf);
        const size_t W = gm.words();
        std::vector<uint64_t> tables(m * W, 0);
        std::vector<size_t> nbits(m, 0);
        for (size_t i = 0; i < m && W; ++i) {
            uint64_t* t = tables.data() + i * W;
            const uint32_t u = v_um[i].unitig;
            if (g.gset_of[u] != RTK_NONE32) for (uint64_t x = g.gset_off[g.gset_of[u]]; x < g.gset_off[g.gset_of[u] + 1]; ++x) bloom_insert(t, gm, g.gset_ids[x]);
            for (uint64_t x = g.loc_off[u]; x < g.loc_off[u + 1]; ++x) bloom_insert(t, gm, g.loc_ids[x]);
            size_t c = 0;
            for (size_t w = 0; w < W; ++w) c += (size_t)__builtin_popcountll(t[w]);
            nbits[i] = c;
        }
        std::vector<uint8_t> valid(m, 0), invalid(m, 0);
        for (size_t i = 0; i < m; ++i) {
            if (valid[i]) continue;
            bool found = false, compatible = false;
            const size_t pos_i = v_um[i].pos, nb_bits_i = nbits[i];
            const uint64_t* ti = tables.data() + i * W;
            for (size_t j = 0; j < m; ++j) {
                if (invalid[j]) continue;
                const size_t pos_j = v_um[j].pos;
                const size_t min_pos_j = (pos_j < opt.insert_sz) ? 0 : (pos_j - opt.insert_sz);
                if (pos_i < min_pos_j || pos_i > pos_j + opt.insert_sz) {
                    const uint64_t* tj = tables.data() + j * W;
                    size_t shared = 0;
                    for (size_t w = 0; w < W; ++w) shared += (size_t)__builtin_popcountll(ti[w] & tj[w]);
                    compatible = true;
                    if (shared >= t_bits_sim * nb_bits_i && shared >= t_bits_sim * nbits[j]) {
                        found = true;
                        valid[i] = 1; valid[j] = 1;
                        break;
                    }
                }
            }
            if (!found && compatible) {
                for (size_t j = pos_i; j < pos_i + v_um[i].len + k && j < pos2rm[r].size(); ++j) pos2rm[r][j] = 1;
                invalid[i] = 1;
            }
        }
    }
    });

    lap("colour sketches / compatibility");
    // 3. whole-read NW paths raw (query) vs corrected (target).  The walk below only ever departs from the corrected read at
    // positions marked in pos2rm: a read without any marked position comes out as its corrected self whatever the alignment
    // is, so its (10 kb x 10 kb) alignment is not computed.  (The reference aligns every read, :1001; with one side empty it
    // gets no alignment and emits nothing, which the unaligned walk below reproduces.)
    // The alignment is only READ at marked positions, so sub-problems of edlib's divide-and-conquer whose target columns hold no
    // mark are not solved (traceback_host.hpp: TbNeed); what is solved follows edlib's splits exactly, so the path through the
    // marked stretches is the reference's.
    std::vector<uint64_t> need_off(n + 1, 0);
    for (uint32_t r = 0; r < n; ++r) need_off[r + 1] = need_off[r] + (corr_off[r + 1] - corr_off[r]) + 2;
    std::vector<uint32_t> need_prefix(need_off[n] + 1, 0);
    std::vector<uint8_t> aligned(n, 0);
    parallel_for(n, [&](size_t rb, size_t re) {
        for (size_t r = rb; r < re; ++r) {
            const size_t tl = (size_t)(corr_off[r + 1] - corr_off[r]);
            const bool empty_side = (raw_off[r + 1] == raw_off[r]) || tl == 0;
            uint32_t* pf = need_prefix.data() + need_off[r];
            uint32_t c = 0;
            pf[0] = 0;
            for (size_t p = 0; p <= tl; ++p) { c += (!empty_side && p < pos2rm[r].size() && pos2rm[r][p]) ? 1u : 0u; pf[p + 1] = c; }
            aligned[r] = c != 0;
        }
    });
    std::vector<std::vector<TbRun>> runs;
    {
        const TbNeed tn{need_prefix.data(), need_off.data()};
        float kms = 0.f;
        nw_path_runs_masked(ctx, n, raw_pool, raw_off, corr_pool, corr_off, tn, runs, &kms);
    }
    if (prof) { size_t na = 0; for (uint32_t r = 0; r < n; ++r) na += aligned[r]; fprintf(stderr, "[phasing] %zu of %u reads aligned\n", na, n); }
    lap("whole-read alignments");
    // edlibAlign returns no alignment when one side is empty (:160-176): nothing is emitted for such a read
    for (uint32_t r = 0; r < n; ++r) if (raw_off[r + 1] == raw_off[r] || corr_off[r + 1] == corr_off[r]) runs[r].clear();

    // walk the alignment (:1001-1052) and collect the reverted neighbourhoods
    std::vector<std::string> s_new(n);
    parallel_for(n, [&](size_t rb, size_t re) {
    for (size_t r = rb; r < re; ++r) {
        const char* s_raw = raw_pool + raw_off[r];
        const char* s_corr = corr_pool + corr_off[r];
        const char* q_corr = qual_pool + qual_off[r];
        const std::vector<uint8_t>& rm = pos2rm[r];
        auto to_rm = [&](size_t p) { return p < rm.size() && rm[p]; };
        std::string& s_out = out_seq[r];
        std::string& q_out = out_qual[r];
        std::vector<size_t> new_base_pos;
        size_t target_pos = 0, query_pos = 0;
        s_out.reserve((size_t)(corr_off[r + 1] - corr_off[r]) + 64);
        q_out.reserve((size_t)(corr_off[r + 1] - corr_off[r]) + 64);
        for (const TbRun& run : runs[r]) {   // CIGAR standard: M (match or mismatch), I, D
            const uint8_t kind = run.kind;
            const size_t l = run.len;
            if (kind == 0) {
                for (size_t i = target_pos; i < target_pos + l; ++i) {
                    if (to_rm(i)) {
                        if (s_corr[i] == s_raw[query_pos + i - target_pos]) q_out += q_corr[i];
                        else { q_out += q_min; new_base_pos.push_back(s_out.length()); }
                        s_out += s_raw[query_pos + i - target_pos];
                    } else { s_out += s_corr[i]; q_out += q_corr[i]; }
                }
                query_pos += l; target_pos += l;
            } else if (kind == 1) {
                if (to_rm(target_pos)) {
                    for (size_t i = 0, sl = s_out.length(); i < l; ++i) new_base_pos.push_back(sl + i);
                    s_out.append(s_raw + query_pos, l);
                    q_out += std::string(l, q_min);
                }
                query_pos += l;
            } else {
                for (size_t i = target_pos; i < target_pos + l; ++i)
                    if (!to_rm(i)) { s_out += s_corr[i]; q_out += q_corr[i]; }
                target_pos += l;
            }
        }
        std::string& sn = s_new[r];
        sn.assign(s_out.length(), 'N');
        for (const size_t pos : new_base_pos) {
            const size_t pos_min = (pos < (k - 1)) ? 0 : (pos - k + 1);
            const size_t pos_max = ((pos + k) > s_out.length()) ? s_out.length() : (pos + k);
            memcpy(&sn[pos_min], s_out.data() + pos_min, pos_max - pos_min);
        }
    }
    });

    lap("revert walk");
    // 4. reverted bases that sit in a graph k-mer after all get the maximum quality (:1081-1090)
    {
        std::string pool;
        std::vector<uint64_t> off(1, 0);
        for (uint32_t r = 0; r < n; ++r) { pool += s_new[r]; off.push_back(pool.size()); }
        pool.push_back('\0');
        std::vector<std::vector<rtk_hit>> h2;
        search_sequence_host(ctx, n, pool.data(), off.data(), RTK_SEARCH_EXACT | RTK_SEARCH_SPARSE_HINT, h2, nullptr);   // s_new is 'N' but for the reverted neighbourhoods
        for (uint32_t r = 0; r < n; ++r)
            for (const rtk_hit& h : h2[r])
                for (size_t j = h.pos; j < h.pos + k && j < out_qual[r].size(); ++j)
                    if (out_qual[r][j] == q_min) out_qual[r][j] = q_max;
    }
    lap("second sweep");
}

}  // namespace rtk

using namespace rtk;

extern "C" int rtk_fix_snps_batch(rtk_ctx* ctx, const rtk_opt* opt, uint32_t n_reads, const char* seq_pool, const uint64_t* seq_off,
                                  char** out_seq_pool, uint64_t* n_fixed) {
    return guarded([&] {
        if (!ctx || !opt || !seq_pool || !seq_off || !out_seq_pool) throw std::invalid_argument("null argument");
        if (!ctx->has_graph) throw std::invalid_argument("no graph uploaded to this context");
        if (opt->k != ctx->hdr.k) throw std::invalid_argument("rtk_opt.k does not match the graph's k");
        DeviceBind bind(ctx);
        const uint64_t total = seq_off[n_reads] - seq_off[0];
        char* out = (char*)malloc(total + 1);
        if (!out) throw std::bad_alloc();
        memcpy(out, seq_pool + seq_off[0], total);
        out[total] = 0;
        std::vector<uint64_t> off(n_reads + 1);
        for (uint32_t i = 0; i <= n_reads; ++i) off[i] = seq_off[i] - seq_off[0];
        uint64_t fixed = 0;
        try { fixed = fix_snps_batch_host(ctx, n_reads, out, off.data()); } catch (...) { free(out); throw; }
        *out_seq_pool = out;
        if (n_fixed) *n_fixed = fixed;
    });
}

extern "C" int rtk_phasing_batch(rtk_ctx* ctx, const rtk_opt* opt, uint32_t n_reads, const char* raw_pool, const uint64_t* raw_off,
                                 const char* corr_pool, const uint64_t* corr_off, const char* qual_pool, const uint64_t* qual_off,
                                 char** out_seq_pool, char** out_qual_pool, uint64_t** out_off) {
    return guarded([&] {
        if (!ctx || !opt || !raw_pool || !raw_off || !corr_pool || !corr_off || !qual_pool || !qual_off || !out_seq_pool || !out_qual_pool || !out_off)
            throw std::invalid_argument("null argument");
        for (uint32_t r = 0; r < n_reads; ++r)
            if (qual_off[r + 1] - qual_off[r] != corr_off[r + 1] - corr_off[r]) throw std::invalid_argument("corrected read and quality lengths differ");
        DeviceBind bind(ctx);
        std::vector<std::string> os, oq;
        phasing_batch_host(ctx, *opt, n_reads, raw_pool, raw_off, corr_pool, corr_off, qual_pool, qual_off, os, oq);
        uint64_t total = 0;
        for (const auto& x : os) total += x.size();
        *out_off = (uint64_t*)malloc((size_t)(n_reads + 1) * 8);
        *out_seq_pool = (char*)malloc(total + 1);
        *out_qual_pool = (char*)malloc(total + 1);
        if (!*out_off || !*out_seq_pool || !*out_qual_pool) throw std::bad_alloc();
        uint64_t t = 0;
        for (uint32_t r = 0; r < n_reads; ++r) {
            (*out_off)[r] = t;
            if (os[r].size() != oq[r].size()) throw std::runtime_error("phased sequence and quality lengths differ");
            memcpy(*out_seq_pool + t, os[r].data(), os[r].size());
            memcpy(*out_qual_pool + t, oq[r].data(), oq[r].size());
            t += os[r].size();
        }
        (*out_off)[n_reads] = t;
    });
}
