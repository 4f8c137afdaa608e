// myers.cuh — K4: batched bit-parallel edit distance (Myers 1999 / Hyyro 2001 block recurrence) on
// sm_100a, replacing edlibAlign's distance modes (src/edlib.cpp:141-296, 547-931) as the reference
// configures them everywhere on the hot path: IUPAC equalities (src/Common.hpp:262-276), modes NW /
// SHW / HW, result = edit distance + every end column carrying it (ascending), -1 when above k.
//
// The reference walks a query's 64-row blocks sequentially per target column, with Ukkonen banding
// and k-doubling; banding only prunes cells that cannot matter, so the result is a function of the
// plain DP and is reproduced here without it.  Mapping to the GPU: G = 1..32 lanes cooperate on one
// alignment, lane j owns query block j (Pv/Mv and the four base profiles live in registers) and the
// lanes sweep the DP matrix as an anti-diagonal WAVEFRONT - at step s lane j processes column s-j,
// taking its horizontal input from lane j-1's previous step through one warp shuffle.  All lanes of
// a group work every step (except the G-1 fill/drain steps); a warp packs 32/G short alignments.
// Queries longer than 64*G rows are swept in ROUNDS of G blocks, the bottom lane spilling its
// horizontal deltas (one int8 per column) to a scratch row that feeds the next round's top lane.
#pragma once
#ifndef RTK_HOSTSIM
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "kmer.cuh"

#define RTK_MYERS_THREADS 128

struct rtk_myers_params {
    const char* q_pool;
    const uint64_t* q_beg;   // [alignment] start of the query in q_pool
    const uint32_t* q_len;
    const char* t_pool;
    const uint64_t* t_beg;
    const uint32_t* t_len;
    const uint8_t* mode;     // 0 NW, 1 SHW, 2 HW
    const int32_t* kmax;     // -1 = unbounded
    const uint32_t* order;   // alignment ids handled by this launch (sorted by target length)
    uint32_t n;
    int32_t* dist;           // [alignment]
    int32_t* n_ends;         // [alignment]
    int32_t* ends;           // capacity: ends_off[a+1]-ends_off[a] = tlen + 1 entries
    const uint64_t* ends_off;
    int8_t* hbound;          // scratch: hb_off[a] .. + tlen, horizontal deltas between rounds
    const uint64_t* hb_off;
    // lean mode (ends == nullptr): only the first and the last end column carrying the distance are reported
    int32_t* first_end = nullptr;   // [alignment]
    int32_t* last_end = nullptr;    // [alignment]
    // fused launch (rtk_myers_fused_kernel): class c = lane groups of G = 32 >> c lanes; its blocks are
    // [cls_blk[c], cls_blk[c+1]) and its alignment ids order[cls_ord[c] .. cls_ord[c+1])
    uint32_t cls_blk[7] = {0, 0, 0, 0, 0, 0, 0};
    uint32_t cls_ord[7] = {0, 0, 0, 0, 0, 0, 0};
};

// IUPAC membership mask of a character: A1 C2 G4 T8, ambiguity codes = union (the index of the letter
// in ambiguity_c[], src/Common.hpp:260); 0 for anything else.
// Branch-free: the 25 letters 'A'..'Y' index two packed nibble tables.  (A switch here compiled to a branch tree executed once per
// query character of every alignment: 15 % of the region engine's instructions and 44 % of its stall samples, profiles/r2_region_lines.md.)
RTK_HD uint32_t rtk_iupac_mask(const char c) {
    const uint32_t i = (uint32_t)(unsigned char)c - 65u;                 // 'A' -> 0
    const uint64_t lo = 0x00F30C00B400D2E1ULL;                            // P O N M L K J I H G F E D C B A
    const uint64_t hi = 0x0000000A09708650ULL;                            //               Y X W V U T S R Q
    const uint64_t w = (i < 16u) ? lo : hi;
    const uint32_t m = (uint32_t)(w >> ((i & 15u) << 2)) & 15u;
    return (i < 25u) ? m : 0u;
}

// edlib equality: identical characters, or {ambiguity code, one of its bases} in either order
// (EqualityDefinition, src/edlib.cpp:50-90, with the 28 pairs of edlib_iupac_alpha)
RTK_HD bool rtk_iupac_eq(const char a, const char b) {
    if (a == b) return true;
    const uint32_t ma = rtk_iupac_mask(a), mb = rtk_iupac_mask(b);
    const bool base_a = ma && !(ma & (ma - 1)), base_b = mb && !(mb & (mb - 1));
    return (base_a != base_b) && (ma & mb);  // exactly one side is a plain base and the code contains it
}

// Eq word of a block against an AMBIGUITY CODE in the target (mask mt with several bits): the code equals the plain query bases it
// contains and itself (rtk_iupac_eq), both read off the block's four profile words - plain = exactly one plane set, itself = exactly
// the code's planes.  Bits past the query's end are clear in every plane, so they stay clear.
RTK_HD uint64_t rtk_iupac_eq_word(const uint32_t mt, const uint64_t PB0, const uint64_t PB1, const uint64_t PB2, const uint64_t PB3) {
    const uint64_t multi = (PB0 & PB1) | (PB0 & PB2) | (PB0 & PB3) | (PB1 & PB2) | (PB1 & PB3) | (PB2 & PB3);
    const uint64_t cover = ((mt & 1u) ? PB0 : 0ULL) | ((mt & 2u) ? PB1 : 0ULL) | ((mt & 4u) ? PB2 : 0ULL) | ((mt & 8u) ? PB3 : 0ULL);
    const uint64_t same = ((mt & 1u) ? PB0 : ~PB0) & ((mt & 2u) ? PB1 : ~PB1) & ((mt & 4u) ? PB2 : ~PB2) & ((mt & 8u) ? PB3 : ~PB3);
    return (cover & ~multi) | same;
}

// One wavefront sweep of a round (see rtk_myers_body): RARE = false drops the loop-invariant rare branches at compile time.
#define RTK_MYERS_STEP_LOOP(RARE) \
        for (int s = 0; s < steps; ++s) { \
            const int from_left = __shfl_up_sync(gmask, hout, 1, G); \
            const int col = s - (int)lane; \
            const char tc = tc_next; \
            tc_next = ((unsigned)(col + 1) < (unsigned)tlen) ? t[col + 1] : (char)0; \
            const bool active = has && ((unsigned)col < (unsigned)tlen); \
            int hin = (lane == 0) ? hin_top : from_left; \
            if (RARE && top_spilled && active) hin = (int)hb[col]; \
            uint64_t Eq = (tc == 'A') ? PB0 : (tc == 'C') ? PB1 : (tc == 'G') ? PB2 : (tc == 'T') ? PB3 : 0ULL; \
            if (RARE && t_amb && active && tc != 'A' && tc != 'C' && tc != 'G' && tc != 'T') { \
                const uint32_t mt = rtk_iupac_mask(tc); \
                if (mt != 0u && !plain) { \
                    Eq = rtk_iupac_eq_word(mt, PB0, PB1, PB2, PB3); \
                } else { \
                    const int lo = b << 6; \
                    const int n = (qlen - lo < 64) ? (qlen - lo) : 64; \
                    for (int i = 0; i < n; ++i) Eq |= (uint64_t)(plain ? (q[lo + i] == tc) : rtk_iupac_eq(q[lo + i], tc)) << i; \
                } \
            } \
            const uint64_t neg = (hin < 0) ? 1ULL : 0ULL; \
            const uint64_t Xv = Eq | Mv; \
            Eq |= neg; \
            const uint64_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq; \
            uint64_t Ph = Mv | ~(Xh | Pv); \
            uint64_t Mh = Pv & Xh; \
            hout = active ? ((int)(Ph >> 63) - (int)(Mh >> 63)) : 0; \
            const int dscore = (int)((Ph >> last_row) & 1) - (int)((Mh >> last_row) & 1); \
            Ph = (Ph << 1) | ((hin > 0) ? 1ULL : 0ULL); \
            Mh = (Mh << 1) | neg; \
            const uint64_t nPv = Mh | ~(Xv | Ph), nMv = Ph & Xv; \
            Pv = active ? nPv : Pv; \
            Mv = active ? nMv : Mv; \
            score += (active && is_last) ? dscore : 0; \
            const bool upd = active && track; \
            const bool better = upd && (score < best); \
            best = better ? score : best; \
            n_best = better ? 0 : n_best; \
            const bool eq = upd && (score == best); \
            if (RARE && ends != nullptr && eq) ends[n_best] = col; \
            first = (eq && n_best == 0) ? col : first; \
            last = eq ? col : last; \
            n_best += eq ? 1 : 0; \
            if (RARE && spill && active) hb[col] = (int8_t)hout; \
        }

#if defined(__CUDACC__) || defined(__CUDACC_SIM__)
// one alignment per group of G lanes; `order` / `n` = the alignments of this class, `blk` = block index within the class
template <int G>
__device__ __forceinline__ void rtk_myers_body(const rtk_myers_params& p, const uint32_t* __restrict__ order, const uint32_t n, const uint32_t blk) {
    const uint32_t lane = threadIdx.x & (G - 1);
    const uint32_t grp = (blk * blockDim.x + threadIdx.x) / G;
    if (grp >= n) return;  // whole groups drop out together; shuffles below are group-scoped
    const uint32_t wl = threadIdx.x & 31;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (wl & ~(uint32_t)(G - 1)));

    const uint32_t a = order[grp];
    const char* q = p.q_pool + p.q_beg[a];
    const char* t = p.t_pool + p.t_beg[a];
    const int qlen = (int)p.q_len[a];
    const int tlen = (int)p.t_len[a];
    const bool plain = (p.mode[a] & 4) != 0;  // bit 2: no additional equalities (edlibDefaultAlignConfig)
    const int mode = p.mode[a] & 3;
    const int nb = (qlen + 63) >> 6;
    const int rounds = (nb + G - 1) / G;
    int8_t* hb = p.hbound + p.hb_off[a];
    int32_t* ends = p.ends ? p.ends + p.ends_off[a] : nullptr;
    int first = -1, last = -1;   // first / last end column with the best score (lean mode reports only these)

    // bottom-row score of the column before the first: D[m][-1] = m ; tracked by the lane owning the last block
    int score = qlen, best = 0x7fffffff, n_best = 0;
    const int last_row = (qlen - 1) & 63;
    // The reference pads the query to a multiple of 64 with W wildcards and reads column c as position
    // c - W (src/edlib.cpp:658-692): when W > 0, "position -1" (the empty target prefix, score qlen) takes
    // part in the SHW/HW minimum and is reported first.  Only the reporting lane's copy is ever read.
    if (mode != 0 && (qlen & 63) != 0) { best = qlen; n_best = 1; if (ends && (int)lane == (nb - 1) % G) ends[0] = -1; }

    // does the target hold anything but A/C/G/T?  (group-cooperative scan; decides whether the step loop needs its slow path)
    bool t_amb = false;
    for (int i = (int)lane; i < tlen; i += G) { const char ch = t[i]; t_amb |= (ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T'); }
    t_amb = __any_sync(gmask, t_amb);
    for (int r = 0; r < rounds; ++r) {
        const int b = r * G + (int)lane;
        const bool has = b < nb;
        // base profiles of this lane's 64 query rows: bit i of PB[x] = query row accepts base x
        uint64_t PB0 = 0, PB1 = 0, PB2 = 0, PB3 = 0;
        if (has) {
            const int lo = b << 6;
            const int n = (qlen - lo < 64) ? (qlen - lo) : 64;
            for (int i = 0; i < n; ++i) {
                uint64_t m = rtk_iupac_mask(q[lo + i]);
                if (plain && (m & (m - 1))) m = 0;  // ambiguity codes only match themselves
                PB0 |= (m & 1) << i; PB1 |= ((m >> 1) & 1) << i; PB2 |= ((m >> 2) & 1) << i; PB3 |= ((m >> 3) & 1) << i;
            }
        }
        uint64_t Pv = ~0ULL, Mv = 0;  // first column: D[i][-1] = i  => all vertical deltas +1
        int hout = 0;
        const bool is_last = has && (b == nb - 1);
        const bool spill = (lane == G - 1) && (r + 1 < rounds);
        const bool top_spilled = (lane == 0) && (r != 0);            // this round's top lane reads the previous round's spill row
        const int hin_top = (mode == 2) ? 0 : 1;                     // D[0][j] = 0 (HW) or j
        const bool track = is_last && (mode != 0);                   // SHW / HW: remember the columns that carry the minimum
        const int steps = tlen + G - 1;
        char tc_next = (lane == 0 && tlen > 0) ? t[0] : (char)0;   // the target is read one step ahead of its use
        // The step is STRAIGHT-LINE code: every decision is a select on values, never a branch.  One warp runs one to 32
        // alignments with nothing else to hide latency, and the first version (an `if` around the active range, a `switch`
        // on the target character, nested `if`s for the end columns) spent 42 % of its stall samples resolving branches
        // (profiles/r1_myers_ncu_full.md).  Only loop-invariant, rare cases keep a branch: ambiguity codes in the target
        // (`t_amb`), the spill row of multi-round queries, and the end-column store of the non-lean mode.
        // two copies of the loop: the common one carries none of the loop-invariant rare branches (ambiguity codes in the
        // target, spill rows of multi-round queries, the end-column store of the non-lean mode) - even a never-taken branch
        // costs a warp with nothing else to run ~20 cycles per step to resolve
        const bool rare = t_amb || (rounds > 1) || (ends != nullptr);   // uniform over the lane group: its lanes shuffle with each other inside the loop
        if (!rare) { RTK_MYERS_STEP_LOOP(false) } else { RTK_MYERS_STEP_LOOP(true) }
        __syncwarp(gmask);  // the next round's top lane reads what this round's bottom lane spilled
    }
    // the lane that owned the last block reports
    if ((int)lane == (nb - 1) % G) {
        if (mode == 0) { best = score; if (ends) ends[0] = tlen - 1; first = last = tlen - 1; n_best = 1; }
        const int kmax = p.kmax ? p.kmax[a] : -1;
        if (kmax >= 0 && best > kmax) { best = -1; n_best = 0; }
        p.dist[a] = best;
        p.n_ends[a] = n_best;
        if (p.first_end) { p.first_end[a] = n_best ? first : -1; p.last_end[a] = n_best ? last : -1; }
    }
}

template <int G>
__global__ void __launch_bounds__(RTK_MYERS_THREADS) rtk_myers_kernel(const rtk_myers_params p) {
    rtk_myers_body<G>(p, p.order, p.n, blockIdx.x);
}

// every lane-group class of a batch in ONE launch (classes in block order 32, 16, 8, 4, 2, 1 lanes: the longest queries
// start first): one launch instead of six launches + twelve event calls per batch - the per-batch driver calls of all service
// threads serialise on the context, which capped the whole correction step
template <int UNUSED>   // a template so that the header can be included by several translation units
__global__ void __launch_bounds__(RTK_MYERS_THREADS) rtk_myers_fused_kernel(const rtk_myers_params p) {
    const uint32_t b = blockIdx.x;
    if (b < p.cls_blk[1]) rtk_myers_body<32>(p, p.order + p.cls_ord[0], p.cls_ord[1] - p.cls_ord[0], b - p.cls_blk[0]);
    else if (b < p.cls_blk[2]) rtk_myers_body<16>(p, p.order + p.cls_ord[1], p.cls_ord[2] - p.cls_ord[1], b - p.cls_blk[1]);
    else if (b < p.cls_blk[3]) rtk_myers_body<8>(p, p.order + p.cls_ord[2], p.cls_ord[3] - p.cls_ord[2], b - p.cls_blk[2]);
    else if (b < p.cls_blk[4]) rtk_myers_body<4>(p, p.order + p.cls_ord[3], p.cls_ord[4] - p.cls_ord[3], b - p.cls_blk[3]);
    else if (b < p.cls_blk[5]) rtk_myers_body<2>(p, p.order + p.cls_ord[4], p.cls_ord[5] - p.cls_ord[4], b - p.cls_blk[4]);
    else rtk_myers_body<1>(p, p.order + p.cls_ord[5], p.cls_ord[6] - p.cls_ord[5], b - p.cls_blk[5]);
}
#endif  // __CUDACC__ || __CUDACC_SIM__
