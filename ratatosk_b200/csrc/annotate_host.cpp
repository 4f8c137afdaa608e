// annotate_host.cpp — drivers of the graph annotation kernels (annotate.cuh) and their C ABI: detectSNPs
// (src/Graph.cpp:484-720) and detectShortCycles (src/Graph.cpp:4660-4854).  Shared by the CUDA build and tests/hostsim; the
// kernels are launched through cycles_run / snp_run (annotate.cu on the device, sim_annotate.cpp on the CPU simulator).
//
// detectSNPs, per coloured unitig (hasSharedPids):
//   1. searchSequence(unitig sequence, substitution only, or_exclusive_match = false): the K1 sweep over the unitigs as a batch
//      of "reads" (search_sequence_host; hits come back in the order of the reference's v_um);
//   2. std::sort by position with the reference's comparator (:491-494, :518) - an unstable sort whose permutation of equal
//      positions decides which candidate is looked at first, so it is the same libstdc++ algorithm on the same sequence;
//   3. hits on the unitig itself dropped (:524); the substituted offset is the first mismatch of the hit k-mer against the
//      unitig (:528; always < k for a one-substitution hit) and its base is the candidate allele;
//   4. the ordered candidates are replayed on the device (rtk_snp_kernel): IUPAC unions, verdict cache, isValidSNPcandidate;
//   5. positions whose final base set holds more than one base become ambiguity ids (pos << 4 | set) (:562-565,
//      UnitigData::add_ambiguity_char src/UnitigData.hpp:448-451).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "annotate.cuh"
#include "rtk_host_common.hpp"

namespace rtk {

namespace {

const uint32_t kCycArena1 = 1024;        // queue entries per warp, first attempt
const uint32_t kCycArena2 = 1u << 21;    // unitigs in tangles, second attempt
const uint32_t kSnpArena1 = 1024;        // visited vertices per walk, first attempt
const uint32_t kSnpArena2 = RTK_SNP_LIMIT + 8;   // the reference's own bound (limit_sz_stack) + the four successors of the last pop

inline uint32_t min_cov_of(const rtk_opt* opt) { return opt ? opt->min_cov_vertices : 2u; }

// RTK_AN_ARENA=<entries>: size of the first-attempt arenas (tests shrink it so that the re-run path is taken)
inline uint32_t arena1(const uint32_t dflt) {
    const char* e = getenv("RTK_AN_ARENA");
    const long v = e ? atol(e) : 0;
    return v > 0 ? (uint32_t)v : dflt;
}

struct CycleOut {
    std::vector<uint8_t> is_cycle;
    std::vector<std::string> blob;
};

// records of one launch -> per-unitig blobs; jobs with status 2 are left for the caller
void gather_cycles(const std::vector<uint32_t>& rec, const std::vector<uint8_t>& status, const uint32_t* list, uint32_t first,
                   CycleOut& out) {
    size_t i = 0;
    while (i < rec.size()) {
        const uint32_t job = rec[i], len = rec[i + 1];
        const size_t words = 2 + (len + 3) / 4;
        if (job >= status.size() || i + words > rec.size()) throw std::runtime_error("corrupt cycle record");
        if (status[job] == 1) {
            const uint32_t u = list ? list[job] : first + job;
            out.blob[u].append((const char*)&rec[i + 2], len);
            out.blob[u].push_back('\0');
        }
        i += words;
    }
}

}  // namespace

void detect_short_cycles_host(rtk_ctx* ctx, const rtk_opt* opt, CycleOut& out, uint64_t* stats, uint64_t range_first = 0, uint64_t range_n = ~0ULL) {
    if (!ctx->has_graph) throw std::invalid_argument("no graph uploaded to this context");
    const uint64_t n = ctx->hdr.n_unitigs;
    if (ctx->hdr.k > 64) throw std::invalid_argument("detectShortCycles: k > 64");
    out.is_cycle.assign(n, 0);
    out.blob.assign(n, std::string());
    const uint32_t min_cov = min_cov_of(opt);
    std::vector<uint32_t> redo;
    const uint64_t chunk = 1u << 22;
    float ms_total = 0.f;
    const uint64_t range_end = (range_first >= n) ? range_first : ((range_n > n - range_first) ? n : range_first + range_n);   // [range_first, range_end) of the unitigs
    for (uint64_t first = range_first; first < range_end; first += chunk) {
        const uint32_t m = (uint32_t)std::min<uint64_t>(chunk, range_end - first);
        std::vector<uint8_t> status;
        std::vector<uint32_t> rec;
        float ms = 0.f;
        cycles_run(ctx, min_cov, nullptr, (uint32_t)first, m, arena1(kCycArena1), status, rec, &ms);
        ms_total += ms;
        gather_cycles(rec, status, nullptr, (uint32_t)first, out);
        for (uint32_t j = 0; j < m; ++j) {
            if (status[j] == 1) out.is_cycle[first + j] = 1;
            else if (status[j] == 2) redo.push_back((uint32_t)first + j);
        }
    }
    if (!redo.empty()) {
        std::vector<uint8_t> status;
        std::vector<uint32_t> rec;
        float ms = 0.f;
        cycles_run(ctx, min_cov, redo.data(), 0, (uint32_t)redo.size(), kCycArena2, status, rec, &ms);
        ms_total += ms;
        for (size_t j = 0; j < redo.size(); ++j) {
            if (status[j] == 2) throw std::runtime_error("detectShortCycles: unitig " + std::to_string(redo[j]) + " has more than 2^21 partial paths");
            if (status[j] == 1) out.is_cycle[redo[j]] = 1;
        }
        gather_cycles(rec, status, redo.data(), 0, out);
    }
    if (stats) { stats[7] += (uint64_t)(ms_total * 1e6); stats[8] += redo.size(); }
}

void detect_snps_host(rtk_ctx* ctx, const rtk_opt* opt, std::vector<std::vector<uint32_t>>& amb, uint64_t* stats, uint64_t range_first = 0, uint64_t range_n = ~0ULL) {
    if (!ctx->has_graph || !ctx->host_graph) throw std::invalid_argument("no graph uploaded to this context");
    const rtk_graph_view& g = ctx->host_graph->view;
    const uint64_t n = g.n_unitigs;
    const uint32_t k = g.k;
    if (k > 64) throw std::invalid_argument("detectSNPs: k > 64");
    const uint32_t min_cov = min_cov_of(opt);
    amb.assign(n, {});
    // batches of coloured unitigs: bounded pool size, read id within the raw-hit label
    const uint64_t max_bases = 256ull << 20, max_reads = (1u << RTK_HIT_READ_BITS) - 1;
    const uint64_t range_end = (range_first >= n) ? range_first : ((range_n > n - range_first) ? n : range_first + range_n);   // unitigs [range_first, range_end)
    uint64_t u0 = range_first;
    std::vector<char> pool;
    std::vector<uint64_t> off;
    std::vector<uint32_t> ids;
    while (u0 < range_end) {
        ids.clear(); off.assign(1, 0);
        uint64_t bases = 0, u = u0;
        for (; u < range_end && ids.size() < max_reads; ++u) {
            if (!(g.shared[u] & 0xffULL)) continue;                      // hasSharedPids (:503 / :590)
            const uint64_t len = g.unitig_off[u + 1] - g.unitig_off[u];
            if (!ids.empty() && bases + len > max_bases) break;
            ids.push_back((uint32_t)u);
            bases += len;
            off.push_back(bases);
        }
        u0 = u;
        if (ids.empty()) continue;
        const uint32_t nr = (uint32_t)ids.size();
        pool.resize(bases + 1);
        parallel_for(nr, [&](size_t b, size_t e) {
            for (size_t r = b; r < e; ++r) {
                const uint64_t ub = g.unitig_off[ids[r]], len = off[r + 1] - off[r];
                char* s = pool.data() + off[r];
                for (uint64_t i = 0; i < len; ++i) s[i] = "ACGT"[rtk_pool_base(g.pool, ub + i)];
            }
        });
        std::vector<std::vector<rtk_hit>> per_read;
        search_sequence_host(ctx, nr, pool.data(), off.data(), RTK_SEARCH_SUBST, per_read, stats);

        // ordered candidates per unitig
        struct Cand { uint32_t x, b, alt; };
        std::vector<std::vector<Cand>> cand(nr);
        parallel_for(nr, [&](size_t rb, size_t re) {
            for (size_t r = rb; r < re; ++r) {
                std::vector<rtk_hit>& v = per_read[r];
                if (v.empty()) continue;
                std::sort(v.begin(), v.end(), [](const rtk_hit& p1, const rtk_hit& p2) { return p1.pos < p2.pos; });
                const uint32_t self = ids[r];
                const char* s = pool.data() + off[r];
                for (const rtk_hit& h : v) {
                    if (h.unitig == self) continue;                      // isSameReferenceUnitig (:524)
                    const uint64_t hb = g.unitig_off[h.unitig] + h.dist;
                    uint32_t j = 0, alt = 0;
                    for (; j < k; ++j) {                                  // cstrMatch(km_snp, seq_ref + p.first) (:528)
                        alt = h.strand ? rtk_pool_base(g.pool, hb + j) : 3u - rtk_pool_base(g.pool, hb + (k - 1 - j));
                        if ("ACGT"[alt] != s[h.pos + j]) break;
                    }
                    if (j < k) cand[r].push_back(Cand{h.pos + j, h.unitig | (h.strand << 31), alt});
                }
            }
        });

        // jobs: slots = distinct positions, verdict slots = distinct candidate unitigs (both numbered in order of first appearance);
        // built per unitig in parallel, laid out by a prefix sum
        struct JobTmp { std::vector<uint32_t> xs, bs; std::vector<rtk_snp_cand> cs; };
        std::vector<JobTmp> tmp(nr);
        parallel_for(nr, [&](size_t rb, size_t re) {
            std::unordered_map<uint32_t, uint32_t> slot_of, bslot_of;
            for (size_t r = rb; r < re; ++r) {
                if (cand[r].empty()) continue;
                JobTmp& T = tmp[r];
                slot_of.clear(); bslot_of.clear();
                T.cs.reserve(cand[r].size());
                for (const Cand& c : cand[r]) {
                    const auto it = slot_of.emplace(c.x, (uint32_t)T.xs.size());
                    if (it.second) T.xs.push_back(c.x);
                    const auto ib = bslot_of.emplace(c.b & 0x7fffffffu, (uint32_t)T.bs.size());
                    if (ib.second) T.bs.push_back(c.b & 0x7fffffffu);
                    T.cs.push_back(rtk_snp_cand{it.first->second, c.b, ib.first->second, c.alt});
                }
            }
        });
        std::vector<rtk_snp_job> jobs;
        std::vector<uint32_t> job_read, job_slot0;
        uint32_t n_bslots = 0;
        uint64_t n_cands = 0, n_slots = 0;
        for (uint32_t r = 0; r < nr; ++r) {
            if (cand[r].empty()) continue;
            rtk_snp_job J;
            J.unitig = ids[r]; J.cand_off = (uint32_t)n_cands; J.n_cand = (uint32_t)tmp[r].cs.size(); J.bslot_off = n_bslots;
            job_slot0.push_back((uint32_t)n_slots);
            n_cands += tmp[r].cs.size(); n_slots += tmp[r].xs.size(); n_bslots += (uint32_t)tmp[r].bs.size();
            jobs.push_back(J);
            job_read.push_back(r);
        }
        if (n_cands >= 0xFFFFFFFFull || n_slots >= 0xFFFFFFFFull) throw std::runtime_error("detectSNPs: batch too large");
        std::vector<rtk_snp_cand> cands(n_cands);
        std::vector<uint8_t> fin(n_slots);
        std::vector<uint32_t> slot_pos(n_slots);
        parallel_for(jobs.size(), [&](size_t jb, size_t je) {
            for (size_t j = jb; j < je; ++j) {
                const uint32_t r = job_read[j], s0 = job_slot0[j];
                const JobTmp& T = tmp[r];
                const char* s = pool.data() + off[r];
                for (size_t i = 0; i < T.xs.size(); ++i) { fin[s0 + i] = (uint8_t)(1u << rtk_base_code(s[T.xs[i]])); slot_pos[s0 + i] = T.xs[i]; }
                rtk_snp_cand* dst = cands.data() + jobs[j].cand_off;
                for (size_t i = 0; i < T.cs.size(); ++i) { dst[i] = T.cs[i]; dst[i].slot += s0; }
            }
        });
        tmp.clear(); tmp.shrink_to_fit();
        job_slot0.push_back((uint32_t)fin.size());
        if (stats) { stats[4] += cands.size(); stats[5] += jobs.size(); }
        if (jobs.empty()) continue;

        const std::vector<uint8_t> fin0 = fin;
        std::vector<uint8_t> status;
        uint64_t walks = 0;
        float ms = 0.f;
        snp_run(ctx, min_cov, jobs, cands, fin, n_bslots, arena1(kSnpArena1), status, &walks, &ms);
        // unitigs whose walks outgrew the small arena: again, alone, with the reference's own bound
        std::vector<uint32_t> redo;
        for (uint32_t j = 0; j < jobs.size(); ++j) if (status[j] == 2) redo.push_back(j);
        if (!redo.empty()) {
            std::vector<rtk_snp_job> jobs2;
            std::vector<rtk_snp_cand> cands2;
            std::vector<uint8_t> fin2;
            std::vector<uint32_t> slot0_2;
            uint32_t nb2 = 0;
            for (const uint32_t j : redo) {
                rtk_snp_job J = jobs[j];
                const uint32_t s0 = job_slot0[j], s1 = job_slot0[j + 1];
                slot0_2.push_back((uint32_t)fin2.size());
                uint32_t maxb = 0;
                for (uint32_t c = 0; c < J.n_cand; ++c) {
                    rtk_snp_cand x = cands[J.cand_off + c];
                    x.slot = x.slot - s0 + (uint32_t)fin2.size();
                    maxb = std::max(maxb, x.bslot + 1);
                    cands2.push_back(x);
                }
                J.cand_off = (uint32_t)(cands2.size() - J.n_cand);
                J.bslot_off = nb2;
                nb2 += maxb;
                fin2.insert(fin2.end(), fin0.begin() + s0, fin0.begin() + s1);
                jobs2.push_back(J);
            }
            std::vector<uint8_t> status2;
            float ms2 = 0.f;
            snp_run(ctx, min_cov, jobs2, cands2, fin2, nb2, kSnpArena2, status2, &walks, &ms2);
            ms += ms2;
            for (size_t i = 0; i < redo.size(); ++i) {
                if (status2[i] == 2) throw std::runtime_error("detectSNPs: traversal arena overflow");
                const uint32_t j = redo[i];
                std::copy(fin2.begin() + slot0_2[i], fin2.begin() + slot0_2[i] + (job_slot0[j + 1] - job_slot0[j]), fin.begin() + job_slot0[j]);
            }
        }
        if (stats) { stats[6] += walks; stats[7] += (uint64_t)(ms * 1e6); stats[8] += redo.size(); }
        parallel_for(jobs.size(), [&](size_t jb, size_t je) {
            for (size_t j = jb; j < je; ++j) {
                std::vector<uint32_t>& a = amb[jobs[j].unitig];
                for (uint32_t sl = job_slot0[j]; sl < job_slot0[j + 1]; ++sl)
                    if (fin[sl] & (fin[sl] - 1)) a.push_back((slot_pos[sl] << 4) | fin[sl]);
                std::sort(a.begin(), a.end());
            }
        });
    }
}

}  // namespace rtk

using namespace rtk;

extern "C" {

int rtk_detect_snps(rtk_ctx* c, const rtk_opt* opt, uint64_t** amb_off, uint32_t** amb_ids, uint64_t* stats) {
    return rtk_detect_snps_range(c, opt, 0, ~0ULL, amb_off, amb_ids, stats);
}

int rtk_detect_snps_range(rtk_ctx* c, const rtk_opt* opt, uint64_t first_unitig, uint64_t n_unitigs, uint64_t** amb_off, uint32_t** amb_ids,
                          uint64_t* stats) {
    return guarded([&] {
        if (!c || !amb_off || !amb_ids) throw std::invalid_argument("null argument");
        DeviceBind bind(c);
        std::vector<std::vector<uint32_t>> amb;
        detect_snps_host(c, opt, amb, stats, first_unitig, n_unitigs);
        uint64_t total = 0;
        for (const auto& v : amb) total += v.size();
        *amb_off = (uint64_t*)malloc((amb.size() + 1) * sizeof(uint64_t));
        *amb_ids = (uint32_t*)malloc((total + 1) * sizeof(uint32_t));
        if (!*amb_off || !*amb_ids) throw std::bad_alloc();
        uint64_t t = 0;
        for (size_t u = 0; u < amb.size(); ++u) {
            (*amb_off)[u] = t;
            if (!amb[u].empty()) memcpy(*amb_ids + t, amb[u].data(), amb[u].size() * sizeof(uint32_t));
            t += amb[u].size();
        }
        (*amb_off)[amb.size()] = t;
    });
}

int rtk_detect_short_cycles(rtk_ctx* c, const rtk_opt* opt, uint8_t** is_cycle, uint64_t** cyc_off, char** cyc_pool, uint64_t* stats) {
    return rtk_detect_short_cycles_range(c, opt, 0, ~0ULL, is_cycle, cyc_off, cyc_pool, stats);
}

int rtk_detect_short_cycles_range(rtk_ctx* c, const rtk_opt* opt, uint64_t first_unitig, uint64_t n_unitigs, uint8_t** is_cycle, uint64_t** cyc_off,
                                  char** cyc_pool, uint64_t* stats) {
    return guarded([&] {
        if (!c || !is_cycle || !cyc_off || !cyc_pool) throw std::invalid_argument("null argument");
        DeviceBind bind(c);
        CycleOut out;
        detect_short_cycles_host(c, opt, out, stats, first_unitig, n_unitigs);
        const size_t n = out.blob.size();
        uint64_t total = 0;
        for (const auto& b : out.blob) total += b.size();
        *is_cycle = (uint8_t*)malloc(n + 1);
        *cyc_off = (uint64_t*)malloc((n + 1) * sizeof(uint64_t));
        *cyc_pool = (char*)malloc(total + 1);
        if (!*is_cycle || !*cyc_off || !*cyc_pool) throw std::bad_alloc();
        uint64_t t = 0;
        for (size_t u = 0; u < n; ++u) {
            (*is_cycle)[u] = out.is_cycle[u];
            (*cyc_off)[u] = t;
            if (!out.blob[u].empty()) memcpy(*cyc_pool + t, out.blob[u].data(), out.blob[u].size());
            t += out.blob[u].size();
        }
        (*cyc_off)[n] = t;
    });
}

int rtk_rtsk_write_annotations(const rtk_host_graph* g, const char* rtsk_in, const char* rtsk_out, const uint64_t* amb_off,
                               const uint32_t* amb_ids, const uint8_t* is_cycle, const uint64_t* cyc_off, const char* cyc_pool) {
    return guarded([&] {
        if (!g || !rtsk_in || !rtsk_out || !amb_off || !amb_ids || !is_cycle || !cyc_off || !cyc_pool) throw std::invalid_argument("null argument");
        patch_rtsk_annotations(g->view, rtsk_in, rtsk_out, amb_off, amb_ids, is_cycle, cyc_off, cyc_pool);
    });
}

int rtk_graph_unitig_annotations(const rtk_host_graph* g, uint32_t u, const uint32_t** amb_ids, uint64_t* n_amb, const char** cyc,
                                 uint64_t* cyc_bytes) {
    if (!g || u >= g->hdr.n_unitigs || !amb_ids || !n_amb || !cyc || !cyc_bytes) { set_error("bad unitig id"); return RTK_EINVAL; }
    *amb_ids = g->view.amb_ids + g->view.amb_off[u];
    *n_amb = g->view.amb_off[u + 1] - g->view.amb_off[u];
    *cyc = g->view.cyc_pool + g->view.cyc_off[u];
    *cyc_bytes = g->view.cyc_off[u + 1] - g->view.cyc_off[u];
    return RTK_OK;
}

}  // extern "C"
