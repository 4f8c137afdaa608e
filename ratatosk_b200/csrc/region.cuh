// region.cuh — device-resident path search of a weak region: extractSemiWeakPaths (src/Correction.cpp:3-157) with
// everything below it — explorePathsBFS2 / explorePathsBFS (src/GraphTraversal.cpp:212-454, :3-210), the `explore`
// step (:251-304), exploreSubGraph / exploreSubGraphLong (:456-587, :589-720), getScorePath (:722-772, :867-909),
// the selectors (src/Alignment.cpp:3-147, :967-1015) and the edlib distance / path alignments they call — chained on
// the device without a host round trip.
//
// Round 1 ran this control flow on host fibers that asked the GPU for one alignment / one burst at a time (10-800
// dependent requests per region, 3.96 M per 67 Mbases): the host saturated and the device idled.  Here ONE WARP owns one
// region for its whole life:
//   * the chain of hops weak anchor -> weak anchor, the BFS queue of partial paths, the result lists and the candidates
//     kept by a burst live in a per-warp arena in HBM (L2-resident in practice), bump-allocated;
//   * the DFS burst walks the explicit adjacency table with the LIFO stack in the arena, successors filtered by the edge
//     flag and the warp-cooperative colour threshold intersection (K3, subgraph.cuh);
//   * every candidate is spelled from the 2-bit pool and scored at once by the wavefront bit-parallel Myers sweep (K4,
//     the step loop of myers.cuh: lane j owns query block j, one shuffle per column); the `>=` / `>` selection of the
//     reference runs on the stream of scores in discovery order;
//   * the per-base qualities of a kept path come from the matrix-storing sweep + edlib-priority traceback (K5) walked in
//     place, writing qualities instead of an op string.
// Rare shapes the kernel does not handle (short-cycle unitigs -> fixRepeats, queue / result-list collapses at 512 / 1024
// entries, alignments above edlib's 1 MiB traceback switch, arena overflow) set a BAIL code; the host then re-runs that one
// call through the request-at-a-time path (traverse.cpp), which produces the same bytes.  Regions are independent: a launch
// is one warp per region, longest first (see rtk_region_kernel).
#pragma once
#ifndef RTK_HOSTSIM
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "../../include/rtk.h"
#include "flat_graph.h"
#include "kmer.cuh"
#include "myers.cuh"
#include "subgraph.cuh"

#define RTK_RG_WARPS 1          /* one region = one warp = one CTA */
#if defined(__CUDACC__)
#define RTK_RG_NOINLINE __noinline__
#else
#define RTK_RG_NOINLINE
#endif
#define RTK_RG_QCAP 512         /* max_sz_stck of the reference: reaching it collapses the queue (bail) */
#define RTK_RG_VCAP 1024        /* max_paths of the reference */
#define RTK_RG_MAX_SEGS 32       /* extractSemiWeakPaths restarts of one region (dead ends followed) */
#define RTK_RG_HSTACK 64         /* open quadrants of edlib's divide-and-conquer traceback */
#define RTK_RG_DROPPED 0xFFFFFFFEu /* queue marker: a path that is dropped when popped (already >= max_len_path) */

// why a region was handed back to the host path
#define RTK_RG_BAIL_CYCLE 1u      /* best path crosses a short-cycle unitig: fixRepeats */
#define RTK_RG_BAIL_QUEUE 2u      /* queue reached 512 entries (selectBestPrefixAlignment collapse) */
#define RTK_RG_BAIL_VLIST 3u      /* result lists reached 1024 entries */
#define RTK_RG_BAIL_ARENA 4u      /* per-warp arena / candidate list full */
#define RTK_RG_BAIL_STRCAP 5u     /* a spelled path or window longer than the string scratch */
#define RTK_RG_BAIL_HIRSCH 6u     /* traceback above edlib's 1 MiB switch (Hirschberg on the host path) */
#define RTK_RG_BAIL_DFS 7u        /* DFS stack / node capacity */
#define RTK_RG_BAIL_CHAIN 8u      /* accumulated path longer than the chain buffers / output pools */
#define RTK_RG_BAIL_LOGIC 9u      /* a state the reference never reaches on this path (kept for safety) */

typedef rtk_path_node rtk_rg_node;            // const_UnitigMap of a path vertex
typedef rtk_region_call_t rtk_rg_task;        // one extractSemiWeakPaths call, prepared by the host (include/rtk.h)
typedef rtk_region_result_t rtk_rg_result;

struct rtk_rg_params {
    // graph (device views)
    const uint64_t* unitig_off;
    const uint64_t* pool;
    const uint64_t* shared;
    const uint32_t* adj;
    const uint32_t* gset_of;
    const uint64_t* gset_off;
    const uint32_t* gset_ids;
    const uint64_t* loc_off;
    const uint32_t* loc_ids;
    const uint64_t* cyc_off;     // UnitigData::compactedCycles: NUL-separated successor strings per unitig
    const char* cyc_pool;
    uint32_t k;
    // work
    const rtk_rg_task* tasks;
    const uint32_t* order;       // task ids, longest window first
    uint32_t n_tasks;
    const char* win_pool;
    const rtk_hit* weak_pool;
    const uint32_t* pid_pool;
    rtk_rg_result* results;
    // outputs, bump-allocated by the finishing warps: out_top[0] nodes used, out_top[1] chars used, out_top[2] segments
    rtk_rg_node* out_nodes;
    char* out_chars;
    rtk_region_seg_t* out_segs;  // out_top[2]
    unsigned long long* out_top;
    uint64_t out_nodes_cap, out_chars_cap, out_segs_cap;
    uint32_t* slot_flags;        // n_slots flags: scratch slots (one per resident CTA) handed out by atomic compare-and-swap
    uint32_t n_slots;
    // per-warp scratch
    unsigned char* scratch;
    uint64_t scratch_per_warp;
    uint32_t str_cap;            // chars per string buffer (windows, spelled candidates, spill row)
    uint32_t mat_cells;          // traceback cells ({Pv,Mv} + anchor each)
    uint32_t tmp_cap;            // bytes per kept-candidate list (terminal / non-terminal)
    uint32_t arena_cap;          // bytes of the hop arena (queue / result paths)
    uint32_t chain_nodes_cap;    // vertices of the accumulated region path
    uint32_t chain_len_cap;      // bases of the accumulated region path
    // options (Correct_Opt, src/Common.hpp:16-158)
    uint32_t min_cov;            // min_cov_vertices
    uint32_t pass2;              // long_read_correct
    uint32_t max_len_weak_region;
    uint32_t max_len_subpath;    // k * large_k_factor (pass 2 bursts)
    int32_t out_qual, max_qual;
    double wrlf;                 // weak_region_len_factor
    double min_score;
    uint64_t tb_limit;           // edlib's direct-traceback limit in bytes of alignment state (1 MiB)
};

// bytes of per-warp scratch for the capacities in p (shared by the host launcher and the kernel's carve-up)
RTK_HD uint64_t rtk_rg_align16(const uint64_t x) { return (x + 15ull) & ~15ull; }
struct rtk_rg_layout {
    uint64_t sA, sB, sC, hb, mat, anc, dfs, dfs_cur, segs, hstack, tmpT, tmpN, arena, q_items, v_items, vt_items, ch_nodes, ch_qual, total;
};
RTK_HD rtk_rg_layout rtk_rg_make_layout(const uint32_t str_cap, const uint32_t mat_cells, const uint32_t tmp_cap, const uint32_t arena_cap,
                                        const uint32_t chain_nodes_cap, const uint32_t chain_len_cap) {
    rtk_rg_layout L;
    uint64_t o = 0;
    L.sA = o; o += rtk_rg_align16((uint64_t)str_cap + 16);
    L.sB = o; o += rtk_rg_align16((uint64_t)str_cap + 16);
    L.sC = o; o += rtk_rg_align16((uint64_t)str_cap + 16);
    L.hb = o; o += rtk_rg_align16((uint64_t)str_cap + 16);
    L.mat = o; o += rtk_rg_align16((uint64_t)mat_cells * 16);
    L.anc = o; o += rtk_rg_align16((uint64_t)mat_cells * 4);
    L.dfs = o; o += rtk_rg_align16((uint64_t)RTK_DFS_STACK * sizeof(rtk_dfs_frame));
    L.dfs_cur = o; o += rtk_rg_align16(sizeof(rtk_dfs_frame));
    L.segs = o; o += rtk_rg_align16((uint64_t)RTK_RG_MAX_SEGS * sizeof(rtk_region_seg_t));
    L.hstack = o; o += rtk_rg_align16((uint64_t)RTK_RG_HSTACK * 16);
    L.tmpT = o; o += rtk_rg_align16(tmp_cap);
    L.tmpN = o; o += rtk_rg_align16(tmp_cap);
    L.arena = o; o += rtk_rg_align16(arena_cap);
    L.q_items = o; o += rtk_rg_align16((uint64_t)RTK_RG_QCAP * 4);
    L.v_items = o; o += rtk_rg_align16((uint64_t)RTK_RG_VCAP * 4);
    L.vt_items = o; o += rtk_rg_align16((uint64_t)RTK_RG_VCAP * 4);
    L.ch_nodes = o; o += rtk_rg_align16((uint64_t)chain_nodes_cap * sizeof(rtk_rg_node));
    L.ch_qual = o; o += rtk_rg_align16((uint64_t)chain_len_cap + 16);
    L.total = rtk_rg_align16(o + 256);
    return L;
}

#if defined(__CUDACC__) || defined(__CUDACC_SIM__)

// ------------------------------------------------------------------------------------------------ per-warp state
struct rg_cand {              // a candidate kept by a burst: header, then str[path_len], qual[path_len] (8-byte padded each)
    uint32_t n_nodes, last_dist, last_len, path_len;
    uint32_t nodes[RTK_DFS_MAX_NODES];   // unitig | traversal strand << 31 ; all but the last are whole unitigs
};
struct rg_path {              // a path in the hop arena: header, then nodes[n], qual[l]
    uint32_t n, l;
};

struct rg_ctx {
    const rtk_rg_params* p;
    uint32_t lane;
    char* sA; char* sB; char* sC;
    int8_t* hb;
    ulonglong2* mat; int32_t* anc;
    rtk_dfs_frame* dfs; rtk_dfs_frame* dfs_cur;
    rtk_region_seg_t* segs;
    uint4* hstack;
    unsigned char* tmpT; unsigned char* tmpN; unsigned char* arena;
    uint32_t* q_items; uint32_t* v_items; uint32_t* vt_items;
    rtk_rg_node* ch_nodes; char* ch_qual;
    uint32_t arena_top;
    uint32_t bail;
    uint32_t n_hops, n_pops, n_cands, n_aligns;
    uint64_t n_cells;            // DP cells swept (|query| x |target| of every alignment), the engine's unit of work
};

__device__ __forceinline__ uint32_t rg_usize(const rg_ctx& C, const uint32_t u) { return (uint32_t)(C.p->unitig_off[u + 1] - C.p->unitig_off[u]); }
__device__ __forceinline__ uint32_t rg_ufull(const rg_ctx& C, const uint32_t u) { return rg_usize(C, u) - C.p->k + 1; }
__device__ __forceinline__ bool rg_has_shared(const rg_ctx& C, const uint32_t u) { return (C.p->shared[u] & 0xffULL) != 0; }
__device__ __forceinline__ bool rg_short_cycle(const rg_ctx& C, const uint32_t u) { return (C.p->shared[u] & 0x100ULL) != 0; }
__device__ __forceinline__ uint32_t rg_pad8(const uint32_t x) { return (x + 7u) & ~7u; }

// getQual (src/Common.hpp:410-418); the engine's translation unit is compiled without FMA contraction so that the
// double arithmetic below rounds like the host's
__device__ __forceinline__ char rg_get_qual(const double score, const int qv_min, const int qv_max) {
    const char phred_base_std = (char)33;
    const char phred_scale_std = (char)qv_max;
    const double qv_score = (score < 1.0 ? score : 1.0) * (double)(phred_scale_std - qv_min);
    return (char)(qv_score + phred_base_std + qv_min);
}
// getMinMaxLength (src/Common.hpp:435-438)
__device__ __forceinline__ void rg_min_max(const uint64_t l, const double f, uint64_t& mn, uint64_t& mx) {
    const double a = (double)l - ((double)l * f), b = (double)l + ((double)l * f);
    mn = (uint64_t)(a > 1.0 ? a : 1.0);
    mx = (uint64_t)(b > 1.0 ? b : 1.0);
}

// spell the oriented mapping of one vertex minus its first `skip` bases into out[0 ..); returns the count (all lanes)
__device__ __forceinline__ uint32_t rg_spell_node(const rg_ctx& C, const uint32_t unitig, const uint32_t strand, const uint32_t dist,
                                                  const uint32_t len, const uint32_t skip, char* out) {
    const uint64_t ub = C.p->unitig_off[unitig];
    const uint32_t mlen = len + C.p->k - 1;
    for (uint32_t t = skip + C.lane; t < mlen; t += 32) {
        const uint32_t pos = strand ? (dist + t) : (dist + (mlen - 1 - t));
        uint32_t b = rtk_pool_base(C.p->pool, ub + pos);
        if (!strand) b = 3 - b;
        out[t - skip] = "ACGT"[b];
    }
    return mlen > skip ? mlen - skip : 0;
}
// Path::toString of a vertex list (interior vertices overlap their predecessor by k-1 bases)
__device__ RTK_RG_NOINLINE uint32_t rg_spell_nodes(const rg_ctx& C, const rtk_rg_node* nd, const uint32_t n, char* out) {
    uint32_t o = 0;
    for (uint32_t i = 0; i < n; ++i) o += rg_spell_node(C, nd[i].unitig, nd[i].strand, nd[i].dist, nd[i].len, i == 0 ? 0u : C.p->k - 1, out + o);
    __syncwarp();
    return o;
}

// ------------------------------------------------------------------------------------------------ K4 in the warp
struct rg_dist { int dist, first, last; };

// edlibAlign distance (modes 0 NW / 1 SHW / 2 HW, IUPAC equalities) of q against t by all 32 lanes: the wavefront sweep of
// myers.cuh with G = 32; returns the distance and the first / last end column carrying it (edlib's endLocations[0] / [n-1])
__device__ RTK_RG_NOINLINE rg_dist rg_myers(rg_ctx& C, const char* __restrict__ q, const int qlen, const char* __restrict__ t, const int tlen, const int mode) {
    rg_dist R;
    ++C.n_aligns;
    C.n_cells += (uint64_t)qlen * (uint64_t)tlen;
    if (qlen == 0 || tlen == 0) {   // edlibAlign's special case (src/edlib.cpp:160-176)
        if (mode == 0) { R.dist = qlen > tlen ? qlen : tlen; R.first = R.last = tlen - 1; }
        else { R.dist = qlen; R.first = R.last = -1; }
        return R;
    }
    constexpr int G = 32;
    const uint32_t lane = C.lane;
    const unsigned gmask = 0xffffffffu;
    const bool plain = false;
    const int nb = (qlen + 63) >> 6;
    const int rounds = (nb + G - 1) / G;
    int8_t* hb = C.hb;
    int32_t* ends = nullptr;
    int first = -1, last = -1;
    int score = qlen, best = 0x7fffffff, n_best = 0;
    const int last_row = (qlen - 1) & 63;
    if (mode != 0 && (qlen & 63) != 0) { best = qlen; n_best = 1; }   // "position -1" (src/edlib.cpp:658-692)
    bool t_amb = false;
    for (int i = (int)lane; i < tlen; i += G) { const char ch = t[i]; t_amb |= (ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T'); }
    t_amb = __any_sync(gmask, t_amb);
    for (int r = 0; r < rounds; ++r) {
        const int b = r * G + (int)lane;
        const bool has = b < nb;
        uint64_t PB0 = 0, PB1 = 0, PB2 = 0, PB3 = 0;
        if (has) {
            const int lo = b << 6;
            const int n = (qlen - lo < 64) ? (qlen - lo) : 64;
            for (int i = 0; i < n; ++i) {
                const uint64_t m = rtk_iupac_mask(q[lo + i]);
                PB0 |= (m & 1) << i; PB1 |= ((m >> 1) & 1) << i; PB2 |= ((m >> 2) & 1) << i; PB3 |= ((m >> 3) & 1) << i;
            }
        }
        uint64_t Pv = ~0ULL, Mv = 0;
        int hout = 0;
        const bool is_last = has && (b == nb - 1);
        const bool spill = (lane == G - 1) && (r + 1 < rounds);
        const bool top_spilled = (lane == 0) && (r != 0);
        const int hin_top = (mode == 2) ? 0 : 1;
        const bool track = is_last && (mode != 0);
        const int steps = tlen + (nb < G ? nb : G) - 1;   // lanes beyond the query's last block never work: no fill / drain steps for them
        char tc_next = (lane == 0 && tlen > 0) ? t[0] : (char)0;
        const bool rare = t_amb || (rounds > 1);
        if (!rare) { RTK_MYERS_STEP_LOOP(false) } else { RTK_MYERS_STEP_LOOP(true) }
        __syncwarp(gmask);
    }
    const int rep = (nb - 1) % G;   // the lane that owned the last block holds the result
    if (mode == 0) { best = score; first = last = tlen - 1; n_best = 1; }
    if (n_best == 0) { first = -1; last = -1; }
    R.dist = __shfl_sync(gmask, best, rep);
    R.first = __shfl_sync(gmask, first, rep);
    R.last = __shfl_sync(gmask, last, rep);
    return R;
}

// ------------------------------------------------------------------------------------------------ K5 in the warp
// edlib's direct traceback is used below 1 MiB of state (src/edlib.cpp:1191-1193), Hirschberg above
__device__ __forceinline__ bool rg_needs_hirschberg(const rg_ctx& C, const uint64_t qlen, const uint64_t tlen) {
    const uint64_t nb = (qlen + 63) / 64;
    return (2ull * 8 + 4) * nb * tlen + 2ull * 4 * tlen >= C.p->tb_limit;
}

struct rg_tb_cell { uint64_t P, M; int32_t A; };
__device__ __forceinline__ rg_tb_cell rg_tb_load(const rg_ctx& C, const int tlen, const int nb, const int last_row, const int b, const int c) {
    rg_tb_cell r;
    if (c < 0) {  // D[i][-1] = i + 1: every vertical delta is +1, anchor = its row + 1
        const int arow = (b == nb - 1) ? last_row : 63;
        r.P = ~0ULL; r.M = 0; r.A = (b << 6) + arow + 1;
        return r;
    }
    const uint64_t idx = (uint64_t)b * (uint64_t)tlen + (uint64_t)c;
    const ulonglong2 cell = C.mat[idx];
    r.P = cell.x; r.M = cell.y; r.A = C.anc[idx];
    return r;
}
__device__ __forceinline__ int rg_tb_row(const rg_tb_cell& c, const int arow, const int r) {
    const uint64_t hi = (arow == 63) ? ~0ULL : ((1ULL << (arow + 1)) - 1ULL);
    const uint64_t lo = (r == 63) ? ~0ULL : ((1ULL << (r + 1)) - 1ULL);
    const uint64_t m = hi & ~lo;
    return c.A - __popcll(c.P & m) + __popcll(c.M & m);
}

// NW traceback of ps[0, qlen) against t[0, tlen) below edlib's 1 MiB switch: matrix-storing sweep (rtk_myers_fill_body<32, false>
// with the matrix in the warp's scratch) + the walk of rtk_traceback_kernel (move priority up > left > diagonal,
// src/edlib.cpp:1023-1134) by lane 0; every diagonal move onto identical characters sets qual_out[i] = best_q.
__device__ RTK_RG_NOINLINE void rg_quality_direct(rg_ctx& C, const char* __restrict__ ps, const int qlen, const char* __restrict__ t, const int tlen,
                                                  const char best_q, char* __restrict__ qual_out) {
    const uint32_t lane = C.lane;
    const int nb = (qlen + 63) >> 6;
    if ((uint64_t)nb * (uint64_t)tlen > (uint64_t)C.p->mat_cells) { C.bail = RTK_RG_BAIL_HIRSCH; return; }
    ++C.n_aligns;
    C.n_cells += (uint64_t)qlen * (uint64_t)tlen;
    constexpr int G = 32;
    const unsigned gmask = 0xffffffffu;
    const char* q = ps;
    const int rounds = (nb + G - 1) / G;
    int8_t* hb = C.hb;
    int nw_dist = 0;
    bool t_amb = false;
    for (int i = (int)lane; i < tlen; i += G) { const char ch = t[i]; t_amb |= (ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T'); }
    t_amb = __any_sync(gmask, t_amb);
    for (int r = 0; r < rounds; ++r) {
        const int b = r * G + (int)lane;
        const bool has = b < nb;
        const int arow = (b == nb - 1) ? ((qlen - 1) & 63) : 63;
        uint64_t PB0 = 0, PB1 = 0, PB2 = 0, PB3 = 0;
        if (has) {
            const int lo = b << 6;
            const int n = (qlen - lo < 64) ? (qlen - lo) : 64;
            for (int i = 0; i < n; ++i) {
                const uint64_t m = rtk_iupac_mask(q[lo + i]);
                PB0 |= (m & 1) << i; PB1 |= ((m >> 1) & 1) << i; PB2 |= ((m >> 2) & 1) << i; PB3 |= ((m >> 3) & 1) << i;
            }
        }
        uint64_t Pv = ~0ULL, Mv = 0;
        int hout = 0;
        int score = (b << 6) + arow + 1;
        const uint64_t base = (uint64_t)b * (uint64_t)tlen;
        const bool spill = (lane == G - 1) && (r + 1 < rounds);
        const int steps = tlen + (nb < G ? nb : G) - 1;   // lanes beyond the query's last block never work: no fill / drain steps for them
        char tc_next = (lane == 0 && tlen > 0) ? t[0] : (char)0;
        const bool top_spilled = (lane == 0) && (r != 0);
        const bool is_last_blk = (b == nb - 1);
        for (int s = 0; s < steps; ++s) {
            const int from_left = __shfl_up_sync(gmask, hout, 1, G);
            const int col = s - (int)lane;
            const char tc = tc_next;
            tc_next = ((unsigned)(col + 1) < (unsigned)tlen) ? t[col + 1] : (char)0;
            const bool active = has && ((unsigned)col < (unsigned)tlen);
            int hin = (lane == 0) ? 1 : from_left;
            if (top_spilled && active) hin = (int)hb[col];
            uint64_t Eq = (tc == 'A') ? PB0 : (tc == 'C') ? PB1 : (tc == 'G') ? PB2 : (tc == 'T') ? PB3 : 0ULL;
            if (t_amb && active && tc != 'A' && tc != 'C' && tc != 'G' && tc != 'T') {
                const uint32_t mt = rtk_iupac_mask(tc);
                if (mt != 0u) Eq = rtk_iupac_eq_word(mt, PB0, PB1, PB2, PB3);   // ambiguity code: bit-parallel, see myers.cuh
                else {
                    const int lo = b << 6;
                    const int n = (qlen - lo < 64) ? (qlen - lo) : 64;
                    for (int i = 0; i < n; ++i) Eq |= (uint64_t)rtk_iupac_eq(q[lo + i], tc) << i;
                }
            }
            const uint64_t neg = (hin < 0) ? 1ULL : 0ULL;
            const uint64_t Xv = Eq | Mv;
            Eq |= neg;
            const uint64_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
            uint64_t Ph = Mv | ~(Xh | Pv);
            uint64_t Mh = Pv & Xh;
            hout = active ? ((int)(Ph >> 63) - (int)(Mh >> 63)) : 0;
            score += active ? ((int)((Ph >> arow) & 1) - (int)((Mh >> arow) & 1)) : 0;
            Ph = (Ph << 1) | ((hin > 0) ? 1ULL : 0ULL);
            Mh = (Mh << 1) | neg;
            const uint64_t nPv = Mh | ~(Xv | Ph), nMv = Ph & Xv;
            Pv = active ? nPv : Pv;
            Mv = active ? nMv : Mv;
            if (active) {
                ulonglong2 cell; cell.x = Pv; cell.y = Mv;
                C.mat[base + col] = cell;
                C.anc[base + col] = score;
            }
            if (active && is_last_blk && col == tlen - 1) nw_dist = score;
            if (spill && active) hb[col] = (int8_t)hout;
        }
        __syncwarp(gmask);
    }
    nw_dist = __shfl_sync(gmask, nw_dist, (nb - 1) % G);
    __syncwarp();
    if (lane == 0) {
        const int last_row = (qlen - 1) & 63;
        int i = qlen - 1, c = tlen - 1, cur_score = nw_dist;
        int b = i >> 6;
        rg_tb_cell cur = rg_tb_load(C, tlen, nb, last_row, b, c);
        rg_tb_cell left = rg_tb_load(C, tlen, nb, last_row, b, c - 1);
        rg_tb_cell left2 = rg_tb_load(C, tlen, nb, last_row, b, c - 2);
        rg_tb_cell left3 = rg_tb_load(C, tlen, nb, last_row, b, c - 3);
        while (i >= 0 && c >= 0) {
            const int r = i & 63;
            const int arow = (b == nb - 1) ? last_row : 63;
            int u, ul;
            if (r > 0) u = cur_score - (int)((cur.P >> r) & 1) + (int)((cur.M >> r) & 1);
            else if (b == 0) u = c + 1;
            else { const rg_tb_cell up = rg_tb_load(C, tlen, nb, last_row, b - 1, c); u = up.A; }
            if (u + 1 == cur_score) {                                   // up: path base unaligned
                --i; cur_score = u;
                if (r == 0 && i >= 0) {
                    --b;
                    cur = rg_tb_load(C, tlen, nb, last_row, b, c);
                    left = rg_tb_load(C, tlen, nb, last_row, b, c - 1);
                    left2 = rg_tb_load(C, tlen, nb, last_row, b, c - 2);
                    left3 = rg_tb_load(C, tlen, nb, last_row, b, c - 3);
                }
                continue;
            }
            const int l = rg_tb_row(left, arow, r);
            if (l + 1 == cur_score) {                                   // left: window base unaligned
                --c; cur_score = l;
                cur = left; left = left2; left2 = left3;
                left3 = rg_tb_load(C, tlen, nb, last_row, b, c - 3);
                continue;
            }
            if (r > 0) ul = l - (int)((left.P >> r) & 1) + (int)((left.M >> r) & 1);
            else if (b == 0) ul = c;
            else { const rg_tb_cell upl = rg_tb_load(C, tlen, nb, last_row, b - 1, c - 1); ul = (c - 1 < 0) ? (b << 6) : upl.A; }
            if (ps[i] == t[c]) qual_out[i] = best_q;                    // M run position with identical characters
            --i; --c; cur_score = ul;
            if (r == 0 && i >= 0) {
                --b;
                cur = rg_tb_load(C, tlen, nb, last_row, b, c);
                left = rg_tb_load(C, tlen, nb, last_row, b, c - 1);
                left2 = rg_tb_load(C, tlen, nb, last_row, b, c - 2);
                left3 = rg_tb_load(C, tlen, nb, last_row, b, c - 3);
            } else {
                cur = left; left = left2; left2 = left3;
                left3 = rg_tb_load(C, tlen, nb, last_row, b, c - 3);
            }
        }
    }
    __syncwarp();
}

// rows[r] = D[r + 1][tlen] of the NW matrix of q against t (both read backwards when rev): the last column edlib's
// divide-and-conquer needs from the forward left half and the reversed right half (src/edlib.cpp:1234-1330).  A distance
// sweep in which every lane keeps the score of its block's anchor row; the rows follow from the final vertical deltas.
__device__ RTK_RG_NOINLINE void rg_lastcol_rows(rg_ctx& C, const char* __restrict__ q, const int qlen, const char* __restrict__ t, const int tlen,
                                                const bool rev, int32_t* __restrict__ rows) {
    const uint32_t lane = C.lane;
    if (tlen == 0) { for (int r = (int)lane; r < qlen; r += 32) rows[r] = r + 1; __syncwarp(); return; }
    ++C.n_aligns;
    C.n_cells += (uint64_t)qlen * (uint64_t)tlen;
    constexpr int G = 32;
    const unsigned gmask = 0xffffffffu;
    const int nb = (qlen + 63) >> 6;
    const int rounds = (nb + G - 1) / G;
    int8_t* hb = C.hb;
    bool t_amb = false;
    for (int i = (int)lane; i < tlen; i += G) { const char ch = t[i]; t_amb |= (ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T'); }
    t_amb = __any_sync(gmask, t_amb);
    for (int r = 0; r < rounds; ++r) {
        const int b = r * G + (int)lane;
        const bool has = b < nb;
        const int arow = (b == nb - 1) ? ((qlen - 1) & 63) : 63;
        uint64_t PB0 = 0, PB1 = 0, PB2 = 0, PB3 = 0;
        if (has) {
            const int lo = b << 6;
            const int n = (qlen - lo < 64) ? (qlen - lo) : 64;
            for (int i = 0; i < n; ++i) {
                const uint64_t m = rtk_iupac_mask(rev ? q[qlen - 1 - (lo + i)] : q[lo + i]);
                PB0 |= (m & 1) << i; PB1 |= ((m >> 1) & 1) << i; PB2 |= ((m >> 2) & 1) << i; PB3 |= ((m >> 3) & 1) << i;
            }
        }
        uint64_t Pv = ~0ULL, Mv = 0;
        int hout = 0;
        int score = (b << 6) + arow + 1;
        const bool spill = (lane == G - 1) && (r + 1 < rounds);
        const int steps = tlen + (nb < G ? nb : G) - 1;
        char tc_next = (lane == 0) ? (rev ? t[tlen - 1] : t[0]) : (char)0;
        const bool top_spilled = (lane == 0) && (r != 0);
        for (int s = 0; s < steps; ++s) {
            const int from_left = __shfl_up_sync(gmask, hout, 1, G);
            const int col = s - (int)lane;
            const char tc = tc_next;
            tc_next = ((unsigned)(col + 1) < (unsigned)tlen) ? (rev ? t[tlen - 2 - col] : t[col + 1]) : (char)0;
            const bool active = has && ((unsigned)col < (unsigned)tlen);
            int hin = (lane == 0) ? 1 : from_left;
            if (top_spilled && active) hin = (int)hb[col];
            uint64_t Eq = (tc == 'A') ? PB0 : (tc == 'C') ? PB1 : (tc == 'G') ? PB2 : (tc == 'T') ? PB3 : 0ULL;
            if (t_amb && active && tc != 'A' && tc != 'C' && tc != 'G' && tc != 'T') {
                const uint32_t mt = rtk_iupac_mask(tc);
                if (mt != 0u) Eq = rtk_iupac_eq_word(mt, PB0, PB1, PB2, PB3);   // ambiguity code: bit-parallel, see myers.cuh
                else {
                    const int lo = b << 6;
                    const int n = (qlen - lo < 64) ? (qlen - lo) : 64;
                    for (int i = 0; i < n; ++i) Eq |= (uint64_t)rtk_iupac_eq(rev ? q[qlen - 1 - (lo + i)] : q[lo + i], tc) << i;
                }
            }
            const uint64_t neg = (hin < 0) ? 1ULL : 0ULL;
            const uint64_t Xv = Eq | Mv;
            Eq |= neg;
            const uint64_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
            uint64_t Ph = Mv | ~(Xh | Pv);
            uint64_t Mh = Pv & Xh;
            hout = active ? ((int)(Ph >> 63) - (int)(Mh >> 63)) : 0;
            score += active ? ((int)((Ph >> arow) & 1) - (int)((Mh >> arow) & 1)) : 0;
            Ph = (Ph << 1) | ((hin > 0) ? 1ULL : 0ULL);
            Mh = (Mh << 1) | neg;
            const uint64_t nPv = Mh | ~(Xv | Ph), nMv = Ph & Xv;
            Pv = active ? nPv : Pv;
            Mv = active ? nMv : Mv;
            if (spill && active) hb[col] = (int8_t)hout;
        }
        if (has) {   // rows of this block from the anchor row upwards: D(r) = D(r + 1) - delta(r + 1)
            int sc = score;
            rows[(b << 6) + arow] = sc;
            for (int rr = arow - 1; rr >= 0; --rr) {
                sc -= (int)((Pv >> (rr + 1)) & 1) - (int)((Mv >> (rr + 1)) & 1);
                rows[(b << 6) + rr] = sc;
            }
        }
        __syncwarp(gmask);
    }
}

// edlib's obtainAlignment for an NW problem of any size (src/edlib.cpp:1164-1399): below the 1 MiB switch the direct
// traceback; above it the target is halved, the last columns of the forward left half and of the reversed right half give
// the first query row at which the two halves add up to the distance (:1330-1356), and both quadrants are solved the same
// way.  Only the matched positions are wanted here, so the quadrants are simply worked off a stack.
__device__ RTK_RG_NOINLINE void rg_nw_quality(rg_ctx& C, const char* __restrict__ ps, const int qlen, const char* __restrict__ t, const int tlen,
                                              const char best_q, char* __restrict__ qual_out) {
    const uint32_t lane = C.lane;
    uint4* st = C.hstack;
    int sp = 0;
    if (lane == 0) st[0] = make_uint4(0u, (uint32_t)qlen, 0u, (uint32_t)tlen);
    sp = 1;
    __syncwarp();
    while (sp > 0 && !C.bail) {
        const uint4 nd = st[--sp];
        __syncwarp();
        const uint32_t qx = nd.x, qy = nd.y, tu = nd.z, tv = nd.w;
        const uint32_t ql = qy - qx, tl = tv - tu;
        if (ql == 0 || tl == 0) continue;
        if (!rg_needs_hirschberg(C, ql, tl)) { rg_quality_direct(C, ps + qx, (int)ql, t + tu, (int)tl, best_q, qual_out + qx); continue; }
        if ((uint64_t)ql * 8 > (uint64_t)C.p->mat_cells * 16 || sp + 2 > RTK_RG_HSTACK) { C.bail = RTK_RG_BAIL_HIRSCH; break; }
        const uint32_t left = tl / 2, right = tl - left;
        int32_t* L = (int32_t*)C.mat;
        int32_t* Rr = L + ql;
        rg_lastcol_rows(C, ps + qx, (int)ql, t + tu, (int)left, false, L);
        rg_lastcol_rows(C, ps + qx, (int)ql, t + tu + left, (int)right, true, Rr);   // Rr[i] = dist(reversed q prefix i + 1, reversed right half)
        const int best = rg_myers(C, ps + qx, (int)ql, t + tu, (int)tl, 0).dist;
        __syncwarp();
        // first row r in [0, ql - 1) with L[r] + R(r + 1) == best, R(j) = Rr[ql - 1 - j] = dist(q[j:], right half)
        int split = -2;
        for (uint32_t base = 0; base + 1 < ql && split == -2; base += 32) {
            const uint32_t r = base + lane;
            const bool hit = (r + 1 < ql) && (L[r] + Rr[ql - 2 - r] == best);
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m) split = (int)(base + (uint32_t)(__ffs(m) - 1));
        }
        if (split == -2 && (int)left + Rr[ql - 1] == best) split = -1;
        if (split == -2 && L[ql - 1] + (int)right == best) split = (int)ql - 1;
        if (split == -2) { C.bail = RTK_RG_BAIL_LOGIC; break; }
        const uint32_t ul_h = (uint32_t)(split + 1);
        __syncwarp();
        if (lane == 0) {
            st[sp] = make_uint4(qx, qx + ul_h, tu, tu + left);
            st[sp + 1] = make_uint4(qx + ul_h, qy, tu + left, tv);
        }
        sp += 2;
        __syncwarp();
    }
}

// getScorePath(opt, path, ref, ref_len, best, second) (src/GraphTraversal.cpp:722-772): per-base quality of a kept path.
// SHW alignment of the spelled path `ps` against the window `t` (edlib PATH task = distance + first end column from the
// distance sweep, then the NW path against that target prefix, src/edlib.cpp:262-279); every path base that sits on an exact
// match of an M run gets `best_q`, the others `base_q`.
__device__ RTK_RG_NOINLINE void rg_path_quality(rg_ctx& C, const char* __restrict__ ps, const int qlen, const char* __restrict__ t, const int tlen_full,
                                                const char base_q, const char best_q, char* __restrict__ qual_out) {
    const uint32_t lane = C.lane;
    for (int i = (int)lane; i < qlen; i += 32) qual_out[i] = base_q;
    __syncwarp();
    if (qlen == 0 || tlen_full == 0) return;
    const rg_dist d = rg_myers(C, ps, qlen, t, tlen_full, 1);
    const int tlen = d.first + 1;       // SHW ending "before the target starts" (position -1): every path base unaligned
    if (tlen <= 0) return;
    rg_nw_quality(C, ps, qlen, t, tlen, best_q, qual_out);
}

// ------------------------------------------------------------------------------------------------ burst (K2/K3 + leaf K4)
struct rg_burst_out {
    uint32_t nT, nN;           // kept terminal / non-terminal candidates (lists in tmpT / tmpN)
    uint32_t topT, topN;       // bytes used
    double t1, t2, nt1, nt2;
};

__device__ __forceinline__ uint32_t rg_cand_bytes(const uint32_t path_len) { return (uint32_t)sizeof(rg_cand) + 2u * rg_pad8(path_len); }
__device__ __forceinline__ char* rg_cand_str(rg_cand* c) { return (char*)(c + 1); }
__device__ __forceinline__ char* rg_cand_qual(rg_cand* c) { return (char*)(c + 1) + rg_pad8(c->path_len); }
__device__ __forceinline__ rg_cand* rg_cand_next(rg_cand* c) { return (rg_cand*)((unsigned char*)c + rg_cand_bytes(c->path_len)); }

// exploreSubGraph / exploreSubGraphLong from vertex (cu, cs) on the window sub[0, sub_len): enumeration as rtk_dfs_kernel
// (subgraph.cuh), each candidate spelled and scored the moment it is found (getScorePath :867-909), the reference's
// selection (`>=` keeps ties in discovery order, `>` restarts the list; :511-523, :536-548) applied to the stream.
__device__ RTK_RG_NOINLINE void rg_burst(rg_ctx& C, const uint32_t cu0, const uint32_t cs0, const bool has_end, const uint32_t end_unitig,
                                         const uint32_t end_strand, const uint32_t end_dist, const uint32_t level, const uint32_t max_len_path,
                                         const char* __restrict__ sub, const uint32_t sub_len, const uint32_t* __restrict__ P, const uint32_t pid_len,
                                         rg_burst_out& B) {
    const rtk_rg_params& p = *C.p;
    const uint32_t lane = C.lane, k = p.k;
    rtk_dfs_frame* st = C.dfs;
    rtk_dfs_frame* f = C.dfs_cur;
    B.nT = B.nN = B.topT = B.topN = 0;
    B.t1 = B.t2 = B.nt1 = B.nt2 = 0.0;
    uint32_t sp = 0;
    if (lane == 0) { st[0].n = 0; st[0].l = level; st[0].path_len = 0; }
    sp = 1;
    __syncwarp();
    while (sp > 0 && !C.bail) {
        --sp;
        __syncwarp();   // every lane is done with the previous frame before it is overwritten
        {   // private copy of the popped frame: pushes below reuse its slot
            const uint32_t* src = (const uint32_t*)&st[sp];
            uint32_t* dst = (uint32_t*)f;
            for (uint32_t i = lane; i < sizeof(rtk_dfs_frame) / 4; i += 32) dst[i] = src[i];
        }
        __syncwarp();
        const uint32_t fn = f->n, fl = f->l, fplen = f->path_len;
        const uint32_t cur = (fn == 0) ? (cu0 | (cs0 << 31)) : f->nodes[fn - 1];
        const uint32_t cu = cur & 0x7fffffffu, cs = cur >> 31;
        const uint64_t shared_w = p.shared[cu];
        for (uint32_t b = 0; b < 4 && !C.bail; ++b) {
            const uint32_t slot = cs ? p.adj[8 * (uint64_t)cu + b] : p.adj[8 * (uint64_t)cu + 4 + (3 - b)];
            if (slot == RTK_NONE32) continue;
            const uint32_t v = slot & 0x7fffffffu;
            const uint32_t vs = cs ? (slot >> 31) : (1u - (slot >> 31));
            const uint64_t bit = cs ? ((uint64_t)(1u << b) << 4) : (uint64_t)(1u << b);
            if (!(shared_w & bit)) continue;
            if (pid_len != 0) {   // K3
                const uint32_t gs = p.gset_of[v];
                uint32_t cnt = 0;
                if (gs != RTK_NONE32) {
                    const uint64_t o = p.gset_off[gs];
                    cnt = rtk_warp_intersect(p.gset_ids + o, (uint32_t)(p.gset_off[gs + 1] - o), P, pid_len, p.min_cov, lane);
                }
                if (cnt < p.min_cov) {
                    const uint64_t o = p.loc_off[v];
                    cnt += rtk_warp_intersect(p.loc_ids + o, (uint32_t)(p.loc_off[v + 1] - o), P, pid_len, p.min_cov - cnt, lane);
                }
                if (cnt < p.min_cov) continue;
            }
            const uint32_t vsize = (uint32_t)(p.unitig_off[v + 1] - p.unitig_off[v]);
            const uint32_t vfull = vsize - k + 1;

            // score one candidate and apply the selection
            auto emit = [&](const bool terminal, const uint32_t ldist, const uint32_t llen, const uint32_t plen) {
                ++C.n_cands;
                if (plen + 8 > p.str_cap) { C.bail = RTK_RG_BAIL_STRCAP; return; }
                // spell into sA: vertex i contributes its oriented mapping minus the k-1 bases shared with the previous one
                uint32_t o = 0;
                for (uint32_t i = 0; i <= fn; ++i) {
                    const uint32_t nd = (i < fn) ? f->nodes[i] : (v | (vs << 31));
                    const uint32_t u = nd & 0x7fffffffu, s = nd >> 31;
                    const uint32_t usz = rg_usize(C, u);
                    o += rg_spell_node(C, u, s, (i < fn) ? 0u : ldist, (i < fn) ? (usz - k + 1) : llen, (i == 0) ? 0u : (k - 1), C.sA + o);
                }
                __syncwarp();
                // getScorePath(opt, path, ref, ref_len, terminal) (:867-909)
                int ed; uint32_t norm;
                if (terminal) { ed = rg_myers(C, C.sA, (int)plen, sub, (int)sub_len, 0).dist; norm = plen; }
                else if (plen >= sub_len) { ed = rg_myers(C, sub, (int)sub_len, C.sA, (int)plen, 2).dist; norm = sub_len; }
                else {
                    const uint64_t want = (uint64_t)((double)plen * (1.0 + p.wrlf));
                    const uint32_t lref = (uint32_t)(want < (uint64_t)sub_len ? want : (uint64_t)sub_len);
                    ed = rg_myers(C, C.sA, (int)plen, sub, (int)lref, 2).dist; norm = plen;
                }
                double sc = 1.0 - ((double)ed / (double)norm);
                sc = sc > 0.0 ? sc : 0.0;
                sc = sc < 1.0 ? sc : 1.0;
                double& s1 = terminal ? B.t1 : B.nt1;
                double& s2 = terminal ? B.t2 : B.nt2;
                uint32_t& cnt = terminal ? B.nT : B.nN;
                uint32_t& top = terminal ? B.topT : B.topN;
                unsigned char* list = terminal ? C.tmpT : C.tmpN;
                if (sc >= s1) {
                    if (sc > s1) { cnt = 0; top = 0; }
                    const uint32_t bytes = rg_cand_bytes(plen);
                    if (top + bytes > p.tmp_cap) { C.bail = RTK_RG_BAIL_ARENA; return; }
                    rg_cand* cd = (rg_cand*)(list + top);
                    if (lane == 0) {
                        cd->n_nodes = fn + 1; cd->last_dist = ldist; cd->last_len = llen; cd->path_len = plen;
                        for (uint32_t i = 0; i < fn; ++i) cd->nodes[i] = f->nodes[i];
                        cd->nodes[fn] = v | (vs << 31);
                    }
                    char* dst = (char*)(cd + 1);
                    for (uint32_t i = lane; i < plen; i += 32) dst[i] = C.sA[i];
                    __syncwarp();
                    top += bytes; ++cnt;
                    s2 = s1; s1 = sc;
                } else if (sc > s2) s2 = sc;
            };

            // (a) terminal: the target unitig reached in the target's strand (:493-526)
            if (has_end && v == end_unitig && vs == end_strand) {
                const uint32_t ldist = vs ? 0u : end_dist;
                const uint32_t llen = vs ? (end_dist + 1) : (vsize - end_dist - k + 1);
                const uint32_t plen = (fn == 0) ? (llen + k - 1) : (fplen + llen);
                if (plen <= max_len_path) emit(true, ldist, llen, plen);
                if (C.bail) break;
            }
            // (b) non-terminal (:530-551)
            const uint32_t nlen = (fn == 0) ? vsize : (fplen + vfull);
            const bool expand = p.pass2 ? (nlen < p.max_len_subpath) : (fl != 0);
            if (expand) {
                if (!(sp < RTK_DFS_STACK && fn + 1 < RTK_DFS_MAX_NODES)) { C.bail = RTK_RG_BAIL_DFS; break; }
                if (lane == 0) {
                    rtk_dfs_frame& nf = st[sp];
                    for (uint32_t i = 0; i < fn; ++i) nf.nodes[i] = f->nodes[i];
                    nf.nodes[fn] = v | (vs << 31);
                    nf.n = fn + 1; nf.l = fl ? fl - 1 : 0; nf.path_len = nlen;
                }
                ++sp;
                __syncwarp();
            } else {
                bool has_succ = false;
                for (uint32_t b2 = 0; b2 < 4; ++b2) has_succ |= (p.adj[8 * (uint64_t)v + (vs ? b2 : 4 + b2)] != RTK_NONE32);
                if (has_succ) emit(false, 0u, vfull, nlen);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ paths in the hop arena
__device__ __forceinline__ rg_path* rg_at(const rg_ctx& C, const uint32_t off) { return (rg_path*)(C.arena + off); }
__device__ __forceinline__ rtk_rg_node* rg_nodes(rg_path* P) { return (rtk_rg_node*)(P + 1); }
__device__ __forceinline__ char* rg_qual(rg_path* P) { return (char*)(rg_nodes(P) + P->n); }
__device__ __forceinline__ uint32_t rg_path_bytes(const uint32_t n, const uint32_t l) { return (uint32_t)sizeof(rg_path) + n * (uint32_t)sizeof(rtk_rg_node) + rg_pad8(l) + 8u; }

// allocate a path with room for n vertices and l quality bytes; RTK_NONE32 (and a bail code) when the arena is full
__device__ __forceinline__ uint32_t rg_alloc_path(rg_ctx& C, const uint32_t n, const uint32_t l) {
    const uint32_t bytes = (rg_path_bytes(n, l) + 15u) & ~15u;
    if (C.arena_top + bytes > C.p->arena_cap) { C.bail = RTK_RG_BAIL_ARENA; return RTK_NONE32; }
    const uint32_t off = C.arena_top;
    C.arena_top += bytes;
    if (C.lane == 0) { rg_path* P = rg_at(C, off); P->n = n; P->l = l; }
    __syncwarp();
    return off;
}

// p_ext = parent + the first `take` vertices of a burst candidate with their slice of the burst's quality string
// (extend loop of explorePathsBFS*, src/GraphTraversal.cpp:377-412 / :140-170; Path::extend src/Path.hpp)
__device__ RTK_RG_NOINLINE uint32_t rg_extend_with(rg_ctx& C, const uint32_t parent_off, rg_cand* cd, const uint32_t take) {
    rg_path* par = rg_at(C, parent_off);
    const uint32_t k = C.p->k;
    const uint32_t pn = par->n, pl = par->l;
    uint32_t add = 0;
    for (uint32_t j = 0; j < take; ++j) {
        const uint32_t u = cd->nodes[j] & 0x7fffffffu;
        add += (j + 1 == cd->n_nodes) ? cd->last_len : rg_ufull(C, u);
    }
    const uint32_t off = rg_alloc_path(C, pn + take, pl + add);
    if (off == RTK_NONE32) return off;
    rg_path* E = rg_at(C, off);
    rtk_rg_node* en = rg_nodes(E);
    const rtk_rg_node* pnod = rg_nodes(par);
    for (uint32_t i = C.lane; i < pn; i += 32) en[i] = pnod[i];
    __syncwarp();
    if (C.lane == 0) {
        if (pn >= 2) { en[pn - 1].dist = 0; en[pn - 1].len = rg_ufull(C, en[pn - 1].unitig); }   // the old end becomes an interior (whole) vertex
        for (uint32_t j = 0; j < take; ++j) {
            rtk_rg_node nd;
            nd.unitig = cd->nodes[j] & 0x7fffffffu; nd.strand = cd->nodes[j] >> 31;
            if (j + 1 == cd->n_nodes) { nd.dist = cd->last_dist; nd.len = cd->last_len; }
            else { nd.dist = 0; nd.len = rg_ufull(C, nd.unitig); }
            en[pn + j] = nd;
        }
    }
    char* eq = rg_qual(E);
    const char* pq = rg_qual(par);
    const char* cq = rg_cand_qual(cd) + (k - 1);
    for (uint32_t i = C.lane; i < pl; i += 32) eq[i] = pq[i];
    for (uint32_t i = C.lane; i < add; i += 32) eq[pl + i] = cq[i];
    __syncwarp();
    return off;
}

// ------------------------------------------------------------------------------------------------ fixRepeats
// spell a vertex list and cut it at `l` bases (Path::toString().substr(0, length()): after fixRepeats' re-extension a last
// vertex may be recorded whole while the path length still counts its partial mapping); returns the length, 0xFFFFFFFF on overflow
__device__ __forceinline__ uint32_t rg_spell_cut(rg_ctx& C, const rtk_rg_node* nd, const uint32_t n, const uint32_t l, char* out) {
    uint64_t total = 0;
    for (uint32_t i = 0; i < n; ++i) total += (i == 0) ? ((uint64_t)nd[i].len + C.p->k - 1) : (uint64_t)nd[i].len;
    if (total + 8 > (uint64_t)C.p->str_cap) { C.bail = RTK_RG_BAIL_STRCAP; return 0xFFFFFFFFu; }
    const uint32_t sp = rg_spell_nodes(C, nd, n, out);
    return sp < l ? sp : l;
}

// fixRepeats (src/GraphTraversal.cpp:1149-1334) on the winning path of a hop: walk its vertices left to right; at a vertex
// on a short-cycle unitig try every stored cycle (UnitigData::getCompactCycles) spliced in at that vertex, each NW-aligned
// against the window with the running best distance as bound; the last strictly improving one is accepted, the walk resumes
// behind the inserted vertices; without an acceptance the following vertices on the same unitig are skipped.
__device__ RTK_RG_NOINLINE uint32_t rg_fix_repeats(rg_ctx& C, const uint32_t path_off, const char* __restrict__ ref, const uint32_t ref_len) {
    const rtk_rg_params& p = *C.p;
    const uint32_t k = p.k, lane = C.lane;
    uint32_t cur = path_off;
    {
        rg_path* Q = rg_at(C, cur);
        const rtk_rg_node* nd = rg_nodes(Q);
        bool cyc = false;
        for (uint32_t i = lane; i < Q->n; i += 32) cyc |= rg_short_cycle(C, nd[i].unitig);
        if (!__any_sync(0xffffffffu, cyc)) return cur;
    }
    int edit;
    {
        rg_path* Q = rg_at(C, cur);
        const uint32_t ql = rg_spell_cut(C, rg_nodes(Q), Q->n, Q->l, C.sB);
        if (C.bail) return RTK_NONE32;
        edit = rg_myers(C, C.sB, (int)ql, ref, (int)ref_len, 0).dist;
    }
    const char qmax = rg_get_qual(1.0, 0, p.max_qual);
    uint32_t i = 0;
    while (!C.bail) {
        rg_path* Q = rg_at(C, cur);
        const uint32_t n = Q->n;
        if (i >= n) break;
        const rtk_rg_node* nd = rg_nodes(Q);
        const rtk_rg_node um = nd[i];
        if (!rg_short_cycle(C, um.unitig)) { ++i; continue; }
        const char* cp = p.cyc_pool + p.cyc_off[um.unitig];
        const uint32_t cl = (uint32_t)(p.cyc_off[um.unitig + 1] - p.cyc_off[um.unitig]);
        uint32_t best_ext = RTK_NONE32;
        uint32_t len_prefix = 0;
        for (uint32_t j = 0; j < i; ++j) len_prefix += nd[j].len;
        uint32_t sidx = 0;
        while (sidx < cl && !C.bail) {
            uint32_t clen = 0;
            while (sidx + clen < cl && cp[sidx + clen] != 0) ++clen;
            const char* cyc = cp + sidx;
            sidx += clen + 1;
            // Path(us, cyc, ue): us = the unitig from the vertex's first k-mer to its end, `cyc` successors, ue = the unitig from
            // its start to the vertex's last k-mer (forward strand); reverse-complemented when the vertex is traversed in reverse
            const uint32_t m = clen + 2;                         // vertices of the repeat
            const uint32_t usz = rg_usize(C, um.unitig);
            rtk_rg_node us = um, ue = um;
            us.len = usz - um.dist - k + 1; us.strand = 1;
            ue.dist = 0; ue.len = um.dist + um.len; ue.strand = 1;
            // the extension: vertices [0, i) + repeat + (i, n)
            const uint32_t en = (n - 1) + m;
            // lengths first (the walk below may fail)
            uint32_t rep_l = us.len + k - 1 + ue.len;
            bool ok = true;
            {
                uint32_t cu = um.unitig, cs = 1;
                for (uint32_t c = 0; c < clen; ++c) {
                    const uint32_t b = rtk_base_code(cyc[c]);
                    uint32_t slot = RTK_NONE32;
                    if (b < 4) slot = cs ? p.adj[8 * (uint64_t)cu + b] : p.adj[8 * (uint64_t)cu + 4 + (3 - b)];
                    if (slot == RTK_NONE32) { ok = false; break; }
                    const uint32_t v = slot & 0x7fffffffu;
                    cs = cs ? (slot >> 31) : (1u - (slot >> 31));
                    cu = v;
                    rep_l += rg_ufull(C, v);
                }
            }
            if (!ok) { C.bail = RTK_RG_BAIL_LOGIC; break; }   // a stored cycle the adjacency cannot follow: host path decides
            const uint32_t new_l = Q->l - (um.len + k - 1) + rep_l;
            if ((uint64_t)len_prefix + um.len + k - 1 > (uint64_t)Q->l) { C.bail = RTK_RG_BAIL_LOGIC; break; }
            const uint32_t eoff = rg_alloc_path(C, en, new_l);
            if (eoff == RTK_NONE32) break;
            Q = rg_at(C, cur); nd = rg_nodes(Q);
            rg_path* E = rg_at(C, eoff);
            rtk_rg_node* ev = rg_nodes(E);
            if (lane == 0) {
                for (uint32_t j = 0; j < i; ++j) ev[j] = nd[j];
                // repeat vertices in traversal order
                rtk_rg_node* rv = ev + i;
                rv[0] = us;
                uint32_t cu = um.unitig, cs = 1;
                for (uint32_t c = 0; c < clen; ++c) {
                    const uint32_t b = rtk_base_code(cyc[c]);
                    const uint32_t slot = cs ? p.adj[8 * (uint64_t)cu + b] : p.adj[8 * (uint64_t)cu + 4 + (3 - b)];
                    const uint32_t v = slot & 0x7fffffffu;
                    cs = cs ? (slot >> 31) : (1u - (slot >> 31));
                    cu = v;
                    rtk_rg_node x; x.unitig = v; x.strand = cs; x.dist = 0; x.len = rg_ufull(C, v);
                    rv[1 + c] = x;
                }
                rv[m - 1] = ue;
                if (!um.strand) {   // Path::rev_comp: reversed order, flipped strands
                    for (uint32_t a = 0, z = m - 1; a < z; ++a, --z) { const rtk_rg_node t = rv[a]; rv[a] = rv[z]; rv[z] = t; }
                    for (uint32_t a = 0; a < m; ++a) rv[a].strand = 1u - rv[a].strand;
                }
                for (uint32_t j = i + 1; j < n; ++j) ev[m + j - 1] = nd[j];
                // Path::extend re-walk: every vertex but the first and the last is recorded whole
                for (uint32_t j = 1; j + 1 < en; ++j) { ev[j].dist = 0; ev[j].len = rg_ufull(C, ev[j].unitig); }
            }
            // quality: the vertex's stretch replaced by max-quality bases for the whole repeat
            {
                char* eq = rg_qual(E);
                const char* oq = rg_qual(Q);
                const uint32_t cut = len_prefix + um.len + k - 1;
                for (uint32_t x = lane; x < len_prefix; x += 32) eq[x] = oq[x];
                for (uint32_t x = lane; x < rep_l; x += 32) eq[len_prefix + x] = qmax;
                for (uint32_t x = lane; x < Q->l - cut; x += 32) eq[len_prefix + rep_l + x] = oq[cut + x];
            }
            __syncwarp();
            const uint32_t ql = rg_spell_cut(C, ev, en, new_l, C.sB);
            if (C.bail) break;
            const int d = rg_myers(C, C.sB, (int)ql, ref, (int)ref_len, 0).dist;
            const int rd = (d > edit) ? -1 : d;       // edlib bounded by k = edit: worse than the running best => -1
            if (rd >= 0 && rd < edit) { edit = rd; best_ext = eoff; }
        }
        if (C.bail) break;
        if (best_ext != RTK_NONE32) {
            const uint32_t diff = rg_at(C, best_ext)->n - n;
            cur = best_ext;
            i = i + diff;
        } else {
            uint32_t jn = i;
            for (uint32_t t = i + 1; t < n; ++t) { if (nd[i].unitig == nd[t].unitig) ++jn; else break; }
            i = jn + 1;
        }
    }
    if (C.bail) return RTK_NONE32;
    return cur;
}

// ------------------------------------------------------------------------------------------------ one hop
// explorePathsBFS2 (has_end) / explorePathsBFS (open end) on the window ref[0, ref_len) from vertex um_s; returns the arena
// offset of the winning path (before fixRepeats, which bails) or RTK_NONE32 when there is none.
__device__ RTK_RG_NOINLINE uint32_t rg_hop(rg_ctx& C, const char* __restrict__ ref, const uint32_t ref_len, const rtk_rg_node um_s, const bool has_end,
                                           const uint32_t e_unitig, const uint32_t e_strand, const uint32_t e_dist, const uint32_t* __restrict__ P,
                                           const uint32_t pid_len) {
    const rtk_rg_params& p = *C.p;
    const uint32_t k = p.k, lane = C.lane;
    ++C.n_hops;
    C.arena_top = 0;
    if (!rg_has_shared(C, um_s.unitig)) return RTK_NONE32;
    if (has_end && !rg_has_shared(C, e_unitig)) return RTK_NONE32;
    if (ref_len + 8 > p.str_cap) { C.bail = RTK_RG_BAIL_STRCAP; return RTK_NONE32; }
    const uint32_t level = 4;
    uint64_t mn, mx;
    rg_min_max((uint64_t)ref_len - k, p.wrlf, mn, mx);
    const uint64_t min_len_path = mn + k;
    const uint64_t max_len_path = (mx > 10 ? mx : 10) + k;
    const char qmax = rg_get_qual(1.0, 0, p.max_qual);
    uint32_t nv = 0, nvt = 0, qh = 0, qn = 0;     // result list, temp list, queue head / size
    {
        rtk_rg_node st = um_s;   // suffix of the start unitig from the anchor k-mer, in traversal orientation
        const uint32_t sz = rg_usize(C, um_s.unitig);
        if (st.strand) { st.dist += st.len - 1; st.len = sz - st.dist - k + 1; }
        else { st.len = um_s.dist + 1; st.dist = 0; }
        if (has_end) {
            if (um_s.unitig == e_unitig && um_s.strand == e_strand && st.dist <= e_dist) {
                const uint64_t len = ((uint64_t)st.len + k - 1) - (e_strand ? (uint64_t)sz - e_dist - k : (uint64_t)e_dist);
                if (len >= min_len_path && len <= max_len_path) {
                    rtk_rg_node back = st;
                    if (back.strand) back.len = e_dist - back.dist + 1;
                    else { back.dist = e_dist; back.len -= e_dist; }
                    const uint32_t l = back.len + k - 1;
                    const uint32_t off = rg_alloc_path(C, 1, l);
                    if (off == RTK_NONE32) return off;
                    rg_path* Q = rg_at(C, off);
                    if (lane == 0) rg_nodes(Q)[0] = back;
                    __syncwarp();
                    char* qq = rg_qual(Q);
                    for (uint32_t i = lane; i < l; i += 32) qq[i] = qmax;
                    __syncwarp();
                    C.v_items[nv++] = off;   // every lane writes the same value
                }
            }
        } else if (((uint64_t)st.len + k - 1) >= min_len_path) {
            rtk_rg_node back = st;
            if (((uint64_t)back.len + k - 1) > max_len_path) {
                if (!back.strand) back.dist = back.len - (uint32_t)(max_len_path - k + 1);
                back.len = (uint32_t)(max_len_path - k + 1);
            }
            const uint32_t l = back.len + k - 1;
            const uint32_t off = rg_alloc_path(C, 1, l);
            if (off == RTK_NONE32) return off;
            rg_path* Q = rg_at(C, off);
            if (lane == 0) rg_nodes(Q)[0] = back;
            __syncwarp();
            char* qq = rg_qual(Q);
            for (uint32_t i = lane; i < l; i += 32) qq[i] = qmax;
            __syncwarp();
            C.v_items[nv++] = off;
        }
        const uint32_t l = st.len + k - 1;
        // a queued path that already spells max_len_path bases is popped and dropped (:373): it only counts towards the
        // queue size, so it is queued as a marker without storage
        if ((uint64_t)l >= max_len_path) { C.q_items[(qh + qn) % RTK_RG_QCAP] = RTK_RG_DROPPED; ++qn; }
        else {
            const uint32_t off = rg_alloc_path(C, 1, l);
            if (off == RTK_NONE32) return off;
            rg_path* Q = rg_at(C, off);
            if (lane == 0) rg_nodes(Q)[0] = st;
            __syncwarp();
            char* qq = rg_qual(Q);
            for (uint32_t i = lane; i < l; i += 32) qq[i] = qmax;
            __syncwarp();
            C.q_items[(qh + qn) % RTK_RG_QCAP] = off; ++qn;
        }
    }
    __syncwarp();
    while (qn > 0 && !C.bail) {
        const uint32_t poff = C.q_items[qh];
        qh = (qh + 1) % RTK_RG_QCAP; --qn;
        if (poff == RTK_RG_DROPPED) continue;
        rg_path* pp = rg_at(C, poff);
        const uint32_t pn = pp->n, plen = pp->l;
        if ((uint64_t)plen >= max_len_path) continue;
        ++C.n_pops;
        const rtk_rg_node um = rg_nodes(pp)[pn - 1];
        // ---- the `explore` step (:251-304)
        const bool non_empty_path = plen > (um.len + k - 1);
        const uint32_t prefix_len = non_empty_path ? (plen - um.len - k + 1) : 0u;
        const uint64_t l_max = max_len_path - prefix_len;
        uint32_t end_pos_ref = 0;
        if (prefix_len != 0) {
            if (plen + 8 > p.str_cap) { C.bail = RTK_RG_BAIL_STRCAP; break; }
            rg_spell_nodes(C, rg_nodes(pp), pn, C.sB);
            const rg_dist d = rg_myers(C, C.sB, (int)prefix_len, ref, (int)ref_len, 1);
            end_pos_ref = (uint32_t)(d.first + 1);
        }
        if (!(end_pos_ref <= ref_len && (ref_len - end_pos_ref) != 0)) continue;   // nothing left of the window: no burst
        const char* sub = ref + end_pos_ref;
        const uint32_t sub_len = ref_len - end_pos_ref;
        rg_burst_out B;
        rg_burst(C, um.unitig, um.strand, has_end, e_unitig, e_strand, e_dist, level - 1, (uint32_t)(l_max < 0xffffffffull ? l_max : 0xffffffffull), sub, sub_len, P,
                 pid_len, B);
        if (C.bail) break;
        if (B.nT != 0 && B.t1 < p.min_score) { B.nT = 0; B.topT = 0; }
        if (B.nN != 0 && B.nt1 < p.min_score) { B.nN = 0; B.topN = 0; }
        // more than one non-terminal survivor: selectBestSubstringAlignment keeps the first strictly best (HW, normalised by
        // the candidate's length; src/Alignment.cpp:967-1015)
        rg_cand* keepN = (rg_cand*)C.tmpN;
        if (B.nN > 1) {
            rg_cand* cd = (rg_cand*)C.tmpN;
            double best = 0.0;
            for (uint32_t i = 0; i < B.nN; ++i) {
                const rg_dist d = rg_myers(C, rg_cand_str(cd), (int)cd->path_len, sub, (int)sub_len, 2);
                const double dn = (double)d.dist / (double)cd->path_len;
                if (i == 0 || (d.dist >= 0 && dn < best)) { best = dn; keepN = cd; }
                cd = rg_cand_next(cd);
            }
            B.nN = 1;
        }
        // qualities of the survivors (getScorePath :722-772)
        {
            const double scT = B.t1 * ((B.t1 == 0.0) ? 0.0 : (1.0 - (B.t2 / B.t1)));
            const char baseT = rg_get_qual(scT, p.out_qual, p.max_qual), bestT = rg_get_qual(B.t1, 0, p.max_qual);
            rg_cand* cd = (rg_cand*)C.tmpT;
            for (uint32_t i = 0; i < B.nT && !C.bail; ++i) {
                rg_path_quality(C, rg_cand_str(cd), (int)cd->path_len, sub, (int)sub_len, baseT, bestT, rg_cand_qual(cd));
                cd = rg_cand_next(cd);
            }
            if (B.nN != 0 && !C.bail) {
                const double scN = B.nt1 * ((B.nt1 == 0.0) ? 0.0 : (1.0 - (B.nt2 / B.nt1)));
                const char baseN = rg_get_qual(scN, p.out_qual, p.max_qual), bestN = rg_get_qual(B.nt1, 0, p.max_qual);
                rg_path_quality(C, rg_cand_str(keepN), (int)keepN->path_len, sub, (int)sub_len, baseN, bestN, rg_cand_qual(keepN));
            }
        }
        if (C.bail) break;
        // ---- consume the burst
        if (has_end) {
            rg_cand* cd = (rg_cand*)C.tmpT;
            for (uint32_t i = 0; i < B.nT; ++i) {
                if (nvt >= RTK_RG_VCAP - 1) { C.bail = RTK_RG_BAIL_VLIST; break; }
                const uint32_t off = rg_extend_with(C, poff, cd, cd->n_nodes);
                if (off == RTK_NONE32) break;
                C.vt_items[nvt++] = off;
                cd = rg_cand_next(cd);
            }
            if (C.bail) break;
        }
        if (B.nN != 0) {
            rg_cand* cd = keepN;
            if (!has_end) {   // open end: every prefix of the extension inside the length window is a result candidate (:140-170)
                uint32_t acc = plen;
                for (uint32_t j = 0; j < cd->n_nodes; ++j) {
                    acc += (j + 1 == cd->n_nodes) ? cd->last_len : rg_ufull(C, cd->nodes[j] & 0x7fffffffu);
                    if ((uint64_t)acc >= min_len_path && (uint64_t)acc <= max_len_path) {
                        if (nvt >= RTK_RG_VCAP - 1) { C.bail = RTK_RG_BAIL_VLIST; break; }
                        const uint32_t off = rg_extend_with(C, poff, cd, j + 1);
                        if (off == RTK_NONE32) break;
                        C.vt_items[nvt++] = off;
                    }
                }
                if (C.bail) break;
            }
            const bool requeue = p.pass2 ? ((uint64_t)cd->path_len >= (uint64_t)p.max_len_subpath) : (cd->n_nodes == level);
            if (requeue) {
                uint32_t off = RTK_RG_DROPPED;
                if ((uint64_t)plen + (uint64_t)(cd->path_len - (k - 1)) < max_len_path) {   // else: dropped when popped, queued as a marker
                    off = rg_extend_with(C, poff, cd, cd->n_nodes);
                    if (off == RTK_NONE32) break;
                }
                C.q_items[(qh + qn) % RTK_RG_QCAP] = off; ++qn;
                if (qn >= RTK_RG_QCAP) { C.bail = RTK_RG_BAIL_QUEUE; break; }
            }
        }
        __syncwarp();
    }
    if (C.bail) return RTK_NONE32;
    // results: terminal paths inside the length window (bfs2) / all collected prefixes (bfs; prunePrefix is the identity here)
    for (uint32_t i = 0; i < nvt; ++i) {
        const uint32_t off = C.vt_items[i];
        const uint32_t l = rg_at(C, off)->l;
        if (!has_end || ((uint64_t)l >= min_len_path && (uint64_t)l <= max_len_path)) {
            if (nv + 1 >= RTK_RG_VCAP) { C.bail = RTK_RG_BAIL_VLIST; return RTK_NONE32; }
            C.v_items[nv++] = off;
        }
    }
    __syncwarp();
    if (nv == 0) return RTK_NONE32;
    uint32_t best_off = C.v_items[0];
    if (nv > 1) {   // selectBestAlignment (src/Alignment.cpp:3-45): NW normalised by max(|cand|, |ref|), first strictly best
        double best = 0.0;
        for (uint32_t i = 0; i < nv && !C.bail; ++i) {
            rg_path* Q = rg_at(C, C.v_items[i]);
            if (Q->l + 8 > p.str_cap) { C.bail = RTK_RG_BAIL_STRCAP; break; }
            const uint32_t ql = rg_spell_nodes(C, rg_nodes(Q), Q->n, C.sB);
            const rg_dist d = rg_myers(C, C.sB, (int)ql, ref, (int)ref_len, 0);
            const uint32_t norm = ql > ref_len ? ql : ref_len;
            const double dn = (double)d.dist / (double)norm;
            if (i == 0 || (d.dist >= 0 && dn < best)) { best = dn; best_off = C.v_items[i]; }
        }
        if (C.bail) return RTK_NONE32;
    }
    // fixRepeats (:1149-1334): stored short cycles spliced into the winner where that lowers its distance to the window
    best_off = rg_fix_repeats(C, best_off, ref, ref_len);
    if (C.bail) return RTK_NONE32;
    return best_off;
}

// ------------------------------------------------------------------------------------------------ the region chain
// extractSemiWeakPaths (src/Correction.cpp:3-157) from the anchor (s_pos, s_unitig, s_dist, s_strand), weak anchors considered
// from index i_weak0 on.  Every hop returns at most one path, so the reference's path lists never hold more than one entry:
// the chain is a single accumulated path (Path::merge) in C.ch_nodes / C.ch_qual that either reaches the right anchor / the
// end of the read (returns 0) or dead-ends at a weak anchor (returns 1).  cn / cl = its vertices / spelled length.
__device__ __forceinline__ uint32_t rg_eswp(rg_ctx& C, const rtk_rg_task& T, const uint32_t s_pos, const uint32_t s_unitig, const uint32_t s_dist,
                                            const uint32_t s_strand, const uint64_t i_weak0, uint32_t& cn, uint32_t& cl) {
    const rtk_rg_params& p = *C.p;
    const uint32_t k = p.k, lane = C.lane;
    const char* win = p.win_pool + T.win_off;
    const rtk_hit* vw = p.weak_pool + T.weak_off;
    const uint32_t* P = p.pid_pool + T.pid_off;
    const bool no_end = !T.has_end;
    const uint64_t pos_um_solid2 = no_end ? (uint64_t)T.s_len - k : (uint64_t)T.end_pos;
    const uint64_t max_len_weak_region = p.max_len_weak_region;
    const uint64_t n_weak = T.n_weak;
    uint64_t i_weak = i_weak0, next_weak_pos = 0;
    bool begin = true, end = false, alive = true;
    // chain = the start anchor's k-mer
    cn = 1; cl = k;
    uint64_t chain_pos = s_pos;
    const char qmax = rg_get_qual(1.0, 0, p.max_qual);
    __syncwarp();
    if (lane == 0) { rtk_rg_node s; s.unitig = s_unitig; s.strand = s_strand; s.dist = s_dist; s.len = 1; C.ch_nodes[0] = s; }
    for (uint32_t i = lane; i < k; i += 32) C.ch_qual[i] = qmax;
    __syncwarp();
    while (i_weak < n_weak && (uint64_t)vw[i_weak].pos < (uint64_t)s_pos) ++i_weak;
    if (i_weak < n_weak) { const uint64_t a = vw[i_weak].pos, b = (uint64_t)s_pos + k; next_weak_pos = a > b ? a : b; }
    while (alive && !end && !C.bail) {
        if (i_weak < n_weak) {
            while (i_weak < n_weak && (uint64_t)vw[i_weak].pos < (pos_um_solid2 - k) && (uint64_t)vw[i_weak].pos < next_weak_pos) ++i_weak;
        } else i_weak = n_weak;
        end = (i_weak == n_weak) || ((uint64_t)vw[i_weak].pos >= (pos_um_solid2 - k));
        const uint64_t target_pos = end ? pos_um_solid2 : (uint64_t)vw[i_weak].pos;
        if (target_pos < chain_pos) { C.bail = RTK_RG_BAIL_LOGIC; break; }
        const uint64_t l_len = (target_pos - chain_pos) + k;
        rtk_rg_node um_s = C.ch_nodes[cn - 1];
        if (begin) { um_s.unitig = s_unitig; um_s.strand = s_strand; um_s.dist = s_dist; um_s.len = 1; }
        const uint64_t woff = chain_pos - T.start_pos;
        if (woff + l_len > (uint64_t)T.win_len) { C.bail = RTK_RG_BAIL_LOGIC; break; }
        const char* ref = win + woff;
        uint32_t got = RTK_NONE32;
        if (end) {
            if (no_end) { if (l_len <= (max_len_weak_region / 2)) got = rg_hop(C, ref, (uint32_t)l_len, um_s, false, 0, 0, 0, P, T.pid_len); }
            else if (l_len <= max_len_weak_region) got = rg_hop(C, ref, (uint32_t)l_len, um_s, true, T.end_unitig, T.end_strand, T.end_dist, P, T.pid_len);
        } else if (l_len <= max_len_weak_region) got = rg_hop(C, ref, (uint32_t)l_len, um_s, true, vw[i_weak].unitig, vw[i_weak].strand, vw[i_weak].dist, P, T.pid_len);
        if (C.bail) break;
        if (got != RTK_NONE32) {   // Path::merge (src/Path.hpp)
            rg_path* O = rg_at(C, got);
            const rtk_rg_node* on = rg_nodes(O);
            const uint32_t onn = O->n, ol = O->l;
            const rtk_rg_node lastn = C.ch_nodes[cn - 1];
            if (ol == 0 || lastn.unitig != on[0].unitig || lastn.strand != on[0].strand) { C.bail = RTK_RG_BAIL_LOGIC; break; }
            if (cn + onn > p.chain_nodes_cap || (uint64_t)cl + ol > (uint64_t)p.chain_len_cap) { C.bail = RTK_RG_BAIL_CHAIN; break; }
            __syncwarp();
            if (lane == 0) {
                rtk_rg_node e = lastn;
                if (!e.strand) e.dist = on[0].dist;
                e.len += on[0].len - 1;
                if (cn != 1 && onn >= 2) { e.dist = 0; e.len = rg_ufull(C, e.unitig); }
                C.ch_nodes[cn - 1] = e;
            }
            for (uint32_t i = 1 + lane; i < onn; i += 32) C.ch_nodes[cn - 1 + i] = on[i];
            const char* oq = rg_qual(O);
            for (uint32_t i = k + lane; i < ol; i += 32) C.ch_qual[cl + (i - k)] = oq[i];
            __syncwarp();
            cn += onn - 1;
            cl += ol - k;
            chain_pos = target_pos;
        } else alive = false;
        if (!end) next_weak_pos = (uint64_t)vw[i_weak].pos + k;
        begin = false;
    }
    return alive ? 0u : 1u;
}

// The path phase of the `correct` lambda (src/Correction.cpp:613-651): extractSemiWeakPaths from the left anchor; when it
// dead-ends (and the call asks for it: rtk_region_call_t::reserved bit 0), the best prefix alignment of the dead-end path
// against the rest of the window (selectBestPrefixAlignment with the weak_region_len_factor cut, src/Alignment.cpp:47-97)
// says how far the path is trusted; the search restarts from the next weak anchor behind that point, until a path reaches
// the end, no anchor is left or a prefix fails the cut.  Each extractSemiWeakPaths call yields one SEGMENT (path, status,
// prefix alignment); the host stitches them exactly like the reference loop.
__device__ __forceinline__ void rg_region(rg_ctx& C, const rtk_rg_task& T, rtk_rg_result& R) {
    const rtk_rg_params& p = *C.p;
    const uint32_t k = p.k, lane = C.lane;
    const rtk_hit* vw = p.weak_pool + T.weak_off;
    const uint64_t pos_um_solid2 = T.has_end ? (uint64_t)T.end_pos : (uint64_t)T.s_len - k;
    const bool follow = (T.reserved & 1u) != 0;
    uint32_t s_pos = T.start_pos, s_unitig = T.start_unitig, s_dist = T.start_dist, s_strand = T.start_strand;
    uint32_t start_idx = RTK_NONE32;   // weak-anchor index the current segment starts from (none: the left anchor)
    uint64_t i_w_s = 0;
    uint32_t n_segs = 0;
    rtk_region_seg_t last;
    last.status = 2; last.start_weak = RTK_NONE32; last.n_nodes = 0; last.len = 0; last.node_off = 0; last.str_off = 0; last.shw_dist = -1; last.shw_first_end = -1;
    while (!C.bail) {
        uint32_t cn = 0, cl = 0;
        const uint32_t status = rg_eswp(C, T, s_pos, s_unitig, s_dist, s_strand, start_idx == RTK_NONE32 ? 0ull : i_w_s, cn, cl);
        if (C.bail) break;
        if (n_segs >= RTK_RG_MAX_SEGS) { C.bail = RTK_RG_BAIL_CHAIN; break; }
        // publish the segment's path: vertices + spelled string + quality string, bump-allocated from the output pools
        unsigned long long no = 0, so = 0;
        if (lane == 0) {
            no = atomicAdd(&p.out_top[0], (unsigned long long)cn);
            so = atomicAdd(&p.out_top[1], (unsigned long long)(2ull * rg_pad8(cl)));
        }
        no = __shfl_sync(0xffffffffu, no, 0);
        so = __shfl_sync(0xffffffffu, so, 0);
        if (no + cn > p.out_nodes_cap || so + 2ull * rg_pad8(cl) > p.out_chars_cap) { C.bail = RTK_RG_BAIL_CHAIN; break; }
        for (uint32_t i = lane; i < cn; i += 32) p.out_nodes[no + i] = C.ch_nodes[i];
        char* ostr = p.out_chars + so;
        rg_spell_nodes(C, C.ch_nodes, cn, ostr);
        char* oq = ostr + rg_pad8(cl);
        for (uint32_t i = lane; i < cl; i += 32) oq[i] = C.ch_qual[i];
        __syncwarp();
        last.status = status; last.start_weak = start_idx; last.n_nodes = cn; last.len = cl; last.node_off = no; last.str_off = so;
        last.shw_dist = -1; last.shw_first_end = -1;
        bool again = false;
        if (status == 1 && follow) {
            // select_prefix_cut: SHW of the dead-end path against the window from this segment's start (first end location)
            const uint64_t woff = (uint64_t)s_pos - T.start_pos;
            const uint32_t wlen = (uint32_t)(pos_um_solid2 - s_pos + k);
            if ((uint64_t)cl + 8 > p.str_cap || (uint64_t)wlen + 8 > p.str_cap) { C.bail = RTK_RG_BAIL_STRCAP; break; }
            const rg_dist d = rg_myers(C, ostr, (int)cl, p.win_pool + T.win_off + woff, (int)wlen, 1);
            last.shw_dist = d.dist; last.shw_first_end = d.first;
            const double dn = (double)d.dist / (double)cl;
            const bool cut_fail = (p.wrlf > 0.0) && (dn > p.wrlf);
            if (!cut_fail && T.n_weak != 0) {
                const uint64_t next_pos = (uint64_t)s_pos + (uint64_t)(int64_t)d.first + k;
                while (i_w_s < T.n_weak && (uint64_t)vw[i_w_s].pos < next_pos) ++i_w_s;
                if (!(i_w_s >= T.n_weak || (uint64_t)vw[i_w_s].pos >= pos_um_solid2 - k || ((uint64_t)vw[i_w_s].pos - s_pos) >= (uint64_t)p.max_len_weak_region)) {
                    again = true;
                }
            }
        }
        if (lane == 0) C.segs[n_segs] = last;
        ++n_segs;
        if (!again) break;
        start_idx = (uint32_t)i_w_s;
        s_pos = vw[i_w_s].pos; s_unitig = vw[i_w_s].unitig; s_dist = vw[i_w_s].dist; s_strand = vw[i_w_s].strand;
    }
    R.n_hops = C.n_hops; R.n_pops = C.n_pops; R.n_cands = C.n_cands; R.n_aligns = C.n_aligns;
    R.seg_off = 0; R.n_segs = 0;
    { const uint64_t kc = C.n_cells >> 10; R.reserved = kc > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)kc; }   // kilo-cells
    if (!C.bail) {   // the segment descriptors, contiguous
        unsigned long long seg_base = 0;
        if (lane == 0) seg_base = atomicAdd(&p.out_top[2], (unsigned long long)n_segs);
        seg_base = __shfl_sync(0xffffffffu, seg_base, 0);
        if (seg_base + n_segs > p.out_segs_cap) C.bail = RTK_RG_BAIL_CHAIN;
        else {
            __syncwarp();
            for (uint32_t i = lane; i < n_segs; i += 32) p.out_segs[seg_base + i] = C.segs[i];
            R.seg_off = seg_base;
        }
    }
    if (C.bail) { R.status = 2; R.bail = C.bail; R.n_nodes = 0; R.len = 0; R.node_off = 0; R.str_off = 0; return; }
    R.status = last.status; R.bail = 0; R.n_nodes = last.n_nodes; R.len = last.len; R.node_off = last.node_off; R.str_off = last.str_off;
    R.n_segs = n_segs;
}

// One region per warp, ONE WARP PER CTA, grid = all regions of the batch (longest first).  The grid is not persistent and a
// CTA is a single warp: a region that finishes gives its SM slot back at once (regions differ 100x in duration; with several
// regions per CTA the finished warps sat at the final barrier - 39 % of the stall samples of the first bulk launches,
// profiles/r2_region_bulk_a.md), and the short high-priority kernels of the other services never wait behind a batch of
// regions.  Scratch is a pool of per-warp slots (one per warp that can be resident), acquired on entry, released on exit.
__global__ void __launch_bounds__(32) rtk_region_kernel(const rtk_rg_params p) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t ti = blockIdx.x;
    if (ti >= p.n_tasks) return;
    uint32_t slot = 0;
    if (lane == 0) {
        slot = blockIdx.x % p.n_slots;
        while (atomicCAS(&p.slot_flags[slot], 0u, 1u) != 0u) slot = (slot + 1 == p.n_slots) ? 0u : slot + 1;
        __threadfence();
    }
    slot = __shfl_sync(0xffffffffu, slot, 0);
    {
        const rtk_rg_layout L = rtk_rg_make_layout(p.str_cap, p.mat_cells, p.tmp_cap, p.arena_cap, p.chain_nodes_cap, p.chain_len_cap);
        unsigned char* S = p.scratch + (uint64_t)slot * p.scratch_per_warp;
        rg_ctx C;
        C.p = &p; C.lane = lane;
        C.sA = (char*)(S + L.sA); C.sB = (char*)(S + L.sB); C.sC = (char*)(S + L.sC); C.hb = (int8_t*)(S + L.hb);
        C.mat = (ulonglong2*)(S + L.mat); C.anc = (int32_t*)(S + L.anc);
        C.dfs = (rtk_dfs_frame*)(S + L.dfs); C.dfs_cur = (rtk_dfs_frame*)(S + L.dfs_cur); C.segs = (rtk_region_seg_t*)(S + L.segs); C.hstack = (uint4*)(S + L.hstack);
        C.tmpT = S + L.tmpT; C.tmpN = S + L.tmpN; C.arena = S + L.arena;
        C.q_items = (uint32_t*)(S + L.q_items); C.v_items = (uint32_t*)(S + L.v_items); C.vt_items = (uint32_t*)(S + L.vt_items);
        C.ch_nodes = (rtk_rg_node*)(S + L.ch_nodes); C.ch_qual = (char*)(S + L.ch_qual);
        const uint32_t id = p.order ? p.order[ti] : ti;
        C.arena_top = 0; C.bail = 0; C.n_hops = C.n_pops = C.n_cands = C.n_aligns = 0; C.n_cells = 0;
        rtk_rg_result R;
        rg_region(C, p.tasks[id], R);
        __syncwarp();
        if (lane == 0) p.results[id] = R;
    }
    __syncwarp();
    if (lane == 0) { __threadfence(); atomicExch(&p.slot_flags[slot], 0u); }
}

#endif  // __CUDACC__ || __CUDACC_SIM__
