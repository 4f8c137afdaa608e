// traceback.cuh — K5: alignment path (edit operations) for NW alignments, replacing edlib's
// obtainAlignmentTraceback (src/edlib.cpp:945-1134) as reached from edlibAlign with EDLIB_TASK_PATH
// (:262-279; SHW paths are NW paths against the target prefix ending at the first best end location).
//
// Two kernels.  rtk_myers_fill_kernel<G> is the K4 wavefront (myers.cuh) in NW mode that additionally
// stores, for every (query block, target column), the vertical delta words Pv/Mv after the column and the
// score of the block's anchor row (row 63, or the query's last row in the last block) - the same state
// edlib's AlignmentData keeps (Ps, Ms, scores), laid out [block][column] so a lane streams its own row.
// rtk_traceback_kernel walks back from (|q|-1, |t|-1): any cell's score is the anchor score minus a
// popcount over the delta bits between the cell and the anchor, so each step costs O(1); moves are tried
// in edlib's priority - up (op 1, query base unaligned) > left (op 2, target base unaligned) > diagonal
// (0 match / 3 mismatch) - which is what makes the path, not just its cost, identical.
// edlib's Ukkonen band cannot change the walk: a move is only ever taken into a cell whose score continues
// an optimal path, and such cells are inside the band with exact scores.
#pragma once
#ifndef RTK_HOSTSIM
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "myers.cuh"

struct rtk_fill_params {
    const char* q_pool;
    const uint64_t* q_beg;
    const uint32_t* q_len;
    const char* t_pool;
    const uint64_t* t_beg;
    const uint32_t* t_len;
    const uint32_t* order;
    uint32_t n;
    const uint64_t* mat_off;   // [alignment] first (block 0, column 0) cell of its matrix, in cells
    ulonglong2* mat;           // {Pv, Mv} per (block, column): index mat_off + block * t_len + column
    int32_t* anchor;           // anchor-row score, same indexing
    int32_t* dist;             // [alignment] NW distance
    int8_t* hbound;            // scratch between rounds: hb_off[a] .. + t_len
    const uint64_t* hb_off;
    // fused launch: slot j = lane groups of G = 32 >> j lanes; its blocks are [cls_blk[j], cls_blk[j+1]) and its alignment ids
    // order[cls_ord[j] .. + cls_cnt[j])
    uint32_t cls_blk[7] = {0, 0, 0, 0, 0, 0, 0};
    uint32_t cls_ord[6] = {0, 0, 0, 0, 0, 0};
    uint32_t cls_cnt[6] = {0, 0, 0, 0, 0, 0};
};

#if defined(__CUDACC__) || defined(__CUDACC_SIM__)

// LASTCOL = false: store every (block, column) cell (direct traceback).  LASTCOL = true: store only the last
// column, at mat_off + block (Hirschberg's split needs one column of the forward and of the reversed problem).
// Queries longer than 64*G rows are swept in rounds of G blocks like K4 (hbound spill between rounds).
template <int G, bool LASTCOL>
__device__ __forceinline__ void rtk_myers_fill_body(const rtk_fill_params& p, const uint32_t* __restrict__ order, const uint32_t n, const uint32_t blk) {
    const uint32_t lane = threadIdx.x & (G - 1);
    const uint32_t grp = (blk * blockDim.x + threadIdx.x) / G;
    if (grp >= n) return;
    const uint32_t wl = threadIdx.x & 31;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (wl & ~(uint32_t)(G - 1)));
    const uint32_t a = order[grp];
    const char* q = p.q_pool + p.q_beg[a];
    const char* t = p.t_pool + p.t_beg[a];
    const int qlen = (int)p.q_len[a], tlen = (int)p.t_len[a];
    const int nb = (qlen + 63) >> 6;
    const int rounds = (nb + G - 1) / G;
    int8_t* hb = p.hbound + p.hb_off[a];
    bool t_amb = false;   // anything but A/C/G/T in the target?
    for (int i = (int)lane; i < tlen; i += G) { const char ch = t[i]; t_amb |= (ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T'); }
    t_amb = __any_sync(gmask, t_amb);
    for (int r = 0; r < rounds; ++r) {
        const int b = r * G + (int)lane;
        const bool has = b < nb;
        const int arow = (b == nb - 1) ? ((qlen - 1) & 63) : 63;  // anchor row of this block
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
        int score = (b << 6) + arow + 1;   // D[anchor row][-1]
        const uint64_t base = LASTCOL ? (p.mat_off[a] + (uint64_t)b) : (p.mat_off[a] + (uint64_t)b * (uint64_t)tlen);
        const bool spill = (lane == G - 1) && (r + 1 < rounds);
        const int steps = tlen + G - 1;
        char tc_next = (lane == 0 && tlen > 0) ? t[0] : (char)0;   // the target is read one step ahead of its use
        const bool top_spilled = (lane == 0) && (r != 0);
        const bool is_last_blk = (b == nb - 1);
        // straight-line step, selects instead of branches (see myers.cuh); the cell stores are predicated
        for (int s = 0; s < steps; ++s) {
            const int from_left = __shfl_up_sync(gmask, hout, 1, G);
            const int col = s - (int)lane;
            const char tc = tc_next;
            tc_next = ((unsigned)(col + 1) < (unsigned)tlen) ? t[col + 1] : (char)0;
            const bool active = has && ((unsigned)col < (unsigned)tlen);
            int hin = (lane == 0) ? 1 : from_left;   // NW: D[0][j] = j
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
            if (active && (!LASTCOL || col == tlen - 1)) {
                ulonglong2 cell; cell.x = Pv; cell.y = Mv;
                const uint64_t idx = LASTCOL ? base : (base + col);
                p.mat[idx] = cell;
                p.anchor[idx] = score;
            }
            if (active && is_last_blk && col == tlen - 1) p.dist[a] = score;
            if (spill && active) hb[col] = (int8_t)hout;
        }
        __syncwarp(gmask);
    }
}

template <int G, bool LASTCOL>
__global__ void __launch_bounds__(RTK_MYERS_THREADS) rtk_myers_fill_kernel(const rtk_fill_params p) {
    rtk_myers_fill_body<G, LASTCOL>(p, p.order, p.n, blockIdx.x);
}

// every lane-group class of a batch in one launch, longest queries first (see rtk_myers_fused_kernel)
template <bool LASTCOL>
__global__ void __launch_bounds__(RTK_MYERS_THREADS) rtk_myers_fill_fused_kernel(const rtk_fill_params p) {
    const uint32_t b = blockIdx.x;
    if (b < p.cls_blk[1]) rtk_myers_fill_body<32, LASTCOL>(p, p.order + p.cls_ord[0], p.cls_cnt[0], b - p.cls_blk[0]);
    else if (b < p.cls_blk[2]) rtk_myers_fill_body<16, LASTCOL>(p, p.order + p.cls_ord[1], p.cls_cnt[1], b - p.cls_blk[1]);
    else if (b < p.cls_blk[3]) rtk_myers_fill_body<8, LASTCOL>(p, p.order + p.cls_ord[2], p.cls_cnt[2], b - p.cls_blk[2]);
    else if (b < p.cls_blk[4]) rtk_myers_fill_body<4, LASTCOL>(p, p.order + p.cls_ord[3], p.cls_cnt[3], b - p.cls_blk[3]);
    else if (b < p.cls_blk[5]) rtk_myers_fill_body<2, LASTCOL>(p, p.order + p.cls_ord[4], p.cls_cnt[4], b - p.cls_blk[4]);
    else rtk_myers_fill_body<1, LASTCOL>(p, p.order + p.cls_ord[5], p.cls_cnt[5], b - p.cls_blk[5]);
}

struct rtk_tb_params {
    const uint32_t* q_len;
    const uint32_t* t_len;
    const uint32_t* ids;       // alignment ids of this launch
    uint32_t n;
    const uint64_t* mat_off;
    const ulonglong2* mat;
    const int32_t* anchor;
    const int32_t* dist;
    const uint64_t* ops_off;   // [alignment] capacity q_len + t_len each
    uint8_t* ops;              // written right-aligned inside the capacity window
    uint32_t* ops_len;         // [alignment]
};

// One stored column block: delta words + anchor score.  Boundary column -1 is synthesised (all deltas +1).
struct rtk_tb_cell { uint64_t P, M; int32_t A; };

__device__ __forceinline__ rtk_tb_cell rtk_tb_load(const rtk_tb_params& p, const uint64_t moff, const int tlen, const int nb,
                                                  const int last_row, const int b, const int c) {
    rtk_tb_cell r;
    if (c < 0) {  // D[i][-1] = i + 1: every vertical delta is +1, anchor = its row + 1
        const int arow = (b == nb - 1) ? last_row : 63;
        r.P = ~0ULL; r.M = 0; r.A = (b << 6) + arow + 1;
        return r;
    }
    const uint64_t idx = moff + (uint64_t)b * (uint64_t)tlen + (uint64_t)c;
    const ulonglong2 cell = p.mat[idx];
    r.P = cell.x; r.M = cell.y; r.A = p.anchor[idx];
    return r;
}

// score of row r (0..63) of a block whose anchor row is arow
__device__ __forceinline__ int rtk_tb_row(const rtk_tb_cell& c, const int arow, const int r) {
    const uint64_t hi = (arow == 63) ? ~0ULL : ((1ULL << (arow + 1)) - 1ULL);
    const uint64_t lo = (r == 63) ? ~0ULL : ((1ULL << (r + 1)) - 1ULL);
    const uint64_t m = hi & ~lo;   // rows r+1 .. arow
    return c.A - __popcll(c.P & m) + __popcll(c.M & m);
}

// The walk keeps the blocks it can touch next in registers: (b, c) `cur`, (b, c-1) `left`, and prefetches (b, c-2) .. (b, c-5):
// a thread walks ONE alignment, a step is ~60 cycles of dependent work and a cell load ~400 cycles from L2, so the loads are
// issued four columns ahead of their use; only crossing into the block above reloads the window.
__global__ void __launch_bounds__(128) rtk_traceback_kernel(const rtk_tb_params p) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.n) return;
    const uint32_t a = p.ids[g];
    const int qlen = (int)p.q_len[a], tlen = (int)p.t_len[a];
    const int nb = (qlen + 63) >> 6, last_row = (qlen - 1) & 63;
    const uint64_t moff = p.mat_off[a];
    uint8_t* out = p.ops + p.ops_off[a];
    uint32_t w = (uint32_t)(qlen + tlen);   // next write position (exclusive), walking backwards
    int i = qlen - 1, c = tlen - 1, cur_score = p.dist[a];
    int b = i >> 6;
    rtk_tb_cell cur = rtk_tb_load(p, moff, tlen, nb, last_row, b, c);
    rtk_tb_cell left = rtk_tb_load(p, moff, tlen, nb, last_row, b, c - 1);
    rtk_tb_cell left2 = rtk_tb_load(p, moff, tlen, nb, last_row, b, c - 2);
    rtk_tb_cell left3 = rtk_tb_load(p, moff, tlen, nb, last_row, b, c - 3);
    rtk_tb_cell left4 = rtk_tb_load(p, moff, tlen, nb, last_row, b, c - 4);
    rtk_tb_cell left5 = rtk_tb_load(p, moff, tlen, nb, last_row, b, c - 5);
    while (i >= 0 && c >= 0) {
        const int r = i & 63;
        const int arow = (b == nb - 1) ? last_row : 63;
        // score above: same column, row i-1 (D[-1][c] = c + 1 above the first row)
        int u, ul;
        if (r > 0) { u = cur_score - (int)((cur.P >> r) & 1) + (int)((cur.M >> r) & 1); }
        else if (b == 0) u = c + 1;
        else u = -0x3fffffff;   // first row of a block above block 0: resolved below by loading the block above
        if (r == 0 && b > 0) {
            const rtk_tb_cell up = rtk_tb_load(p, moff, tlen, nb, last_row, b - 1, c);
            u = up.A;            // anchor row of a non-last block is its row 63
        }
        if (u + 1 == cur_score) {                                   // up: query base unaligned
            out[--w] = 1; --i; cur_score = u;
            if (r == 0 && i >= 0) {
                --b;
                cur = rtk_tb_load(p, moff, tlen, nb, last_row, b, c);
                left = rtk_tb_load(p, moff, tlen, nb, last_row, b, c - 1);
                left2 = rtk_tb_load(p, moff, tlen, nb, last_row, b, c - 2);
                left3 = rtk_tb_load(p, moff, tlen, nb, last_row, b, c - 3);
                left4 = rtk_tb_load(p, moff, tlen, nb, last_row, b, c - 4);
                left5 = rtk_tb_load(p, moff, tlen, nb, last_row, b, c - 5);
            }
            continue;
        }
        const int l = rtk_tb_row(left, arow, r);                   // D[i][c-1] (column -1 synthesised)
        if (l + 1 == cur_score) {                                   // left: target base unaligned
            out[--w] = 2; --c; cur_score = l;
            cur = left; left = left2; left2 = left3; left3 = left4; left4 = left5;
            left5 = rtk_tb_load(p, moff, tlen, nb, last_row, b, c - 5);
            continue;
        }
        // diagonal: D[i-1][c-1]
        if (r > 0) ul = l - (int)((left.P >> r) & 1) + (int)((left.M >> r) & 1);
        else if (b == 0) ul = c;                                     // D[-1][c-1] = c
        else { const rtk_tb_cell upl = rtk_tb_load(p, moff, tlen, nb, last_row, b - 1, c - 1); ul = (c - 1 < 0) ? ((b << 6)) : upl.A; }
        out[--w] = (ul == cur_score) ? 0 : 3;
        --i; --c; cur_score = ul;
        if (r == 0 && i >= 0) {
            --b;
            cur = rtk_tb_load(p, moff, tlen, nb, last_row, b, c);
            left = rtk_tb_load(p, moff, tlen, nb, last_row, b, c - 1);
            left2 = rtk_tb_load(p, moff, tlen, nb, last_row, b, c - 2);
            left3 = rtk_tb_load(p, moff, tlen, nb, last_row, b, c - 3);
            left4 = rtk_tb_load(p, moff, tlen, nb, last_row, b, c - 4);
            left5 = rtk_tb_load(p, moff, tlen, nb, last_row, b, c - 5);
        } else {
            cur = left; left = left2; left2 = left3; left3 = left4; left4 = left5;
            left5 = rtk_tb_load(p, moff, tlen, nb, last_row, b, c - 5);
        }
    }
    while (c >= 0) { out[--w] = 2; --c; }
    while (i >= 0) { out[--w] = 1; --i; }
    p.ops_len[a] = (uint32_t)(qlen + tlen) - w;
}

#endif  // __CUDACC__ || __CUDACC_SIM__
