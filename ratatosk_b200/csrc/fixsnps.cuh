// fixsnps.cuh — fixSNPs (src/Alignment.cpp:846-964), the optional first step of the second pass (`-f`,
// Correct_Opt::force_unres_snp_corr; src/Ratatosk.cpp:672 / :828): an IUPAC code that pass 1 left in a read is replaced by a
// base when exactly one of its bases puts a k-mer of the (k = 63) graph over that position.
//
// One warp per read.  The host lists the non-ACGT positions of each read (CSR); the warp visits them in ascending order,
// exactly like the reference's scan, because a resolved position changes the windows of the positions behind it:
//   window  = out[max(0, i-k+1), min(i+k, L))            (s_sub)
//   v_amb   = the still unresolved codes inside [window start, window start + window length]   -- the closed upper bound
//             of the reference's map::upper_bound is kept: a code just past the window counts and is enumerated
//   product of their base counts, stopping at the first partial product >= 64; nothing is tried at >= 64
//   j = 0 .. 4*|v_amb|-1 while at most one base is known to work: code t takes base (j >> 2t) & 3 (invalid combination
//             when the code does not contain it); a combination whose base at i is already known to work is skipped;
//             it "works" when any k-mer of the substituted window is in the graph
//   exactly one working base -> the read takes it and the code leaves the list.
// The lanes share the k-mer start positions of a window (at most k of them); membership is the K1 lookup of lookup.cuh.
#pragma once
#ifndef RTK_HOSTSIM
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "flat_graph.h"
#include "kmer.cuh"
#include "lookup.cuh"
#include "myers.cuh"   // rtk_iupac_mask

#define RTK_FS_WARPS 4
#define RTK_FS_MAXK 64

struct rtk_fs_params {
    const uint64_t* table;
    uint64_t n_buckets;
    const uint64_t* pool;
    uint32_t k;
    uint32_t n_reads;
    char* seq;                  // reads, modified in place
    const uint64_t* seq_off;    // n_reads + 1
    const uint32_t* amb_pos;    // positions (within the read) of the non-ACGT characters, ascending per read
    const uint64_t* amb_off;    // n_reads + 1
    uint8_t* amb_done;          // per listed position: 1 once resolved (zero-initialised)
    unsigned long long* n_fixed;
};

#if defined(__CUDACC__) || defined(__CUDACC_SIM__)

// getAmbiguityIndex (src/Common.hpp:358-382): the code's base set A=1 C=2 G=4 T=8, case-insensitive, 0 for anything else
__device__ __forceinline__ uint32_t rtk_fs_mask(const char c) { return rtk_iupac_mask((char)(c & 0xDF)); }

template <typename KT>
__global__ void __launch_bounds__(RTK_FS_WARPS * 32) rtk_fixsnps_kernel(const rtk_fs_params p) {
    __shared__ char s_win[RTK_FS_WARPS][2 * RTK_FS_MAXK];
    const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t r = blockIdx.x * RTK_FS_WARPS + w;
    if (r >= p.n_reads) return;
    const uint64_t a0 = p.amb_off[r], a1 = p.amb_off[r + 1];
    if (a0 == a1) return;
    char* s = p.seq + p.seq_off[r];
    const uint64_t L = p.seq_off[r + 1] - p.seq_off[r];
    const uint32_t k = p.k;
    if (L < k) return;
    char* win = s_win[w];
    const KT kmask = KmerOps<KT>::mask((int)k);
    for (uint64_t e = a0; e < a1; ++e) {
        const uint64_t i = p.amb_pos[e];
        const uint64_t min_pos = (i < (uint64_t)(k - 1)) ? 0 : (i - k + 1);
        const uint64_t wlen = ((i + k < L) ? (i + k) : L) - min_pos;
        // v_amb: unresolved codes at positions [min_pos, min_pos + wlen] in ascending order; at most 6 matter (product < 64)
        uint32_t vpos[8], vmask[8], nv = 0, self = 0;
        uint64_t prod = 1;
        uint64_t f = e;
        while (f > a0 && p.amb_pos[f - 1] >= min_pos) --f;
        bool too_many = false;
        for (; f < a1 && p.amb_pos[f] <= min_pos + wlen; ++f) {
            if (p.amb_done[f]) continue;
            const uint32_t m = rtk_fs_mask(s[p.amb_pos[f]]);
            prod *= (uint64_t)__popc(m);
            if (nv < 8) { vpos[nv] = (uint32_t)(p.amb_pos[f] - min_pos); vmask[nv] = m; if (f == e) self = nv; }
            ++nv;
            if (prod >= 64) { too_many = true; break; }
        }
        // a code without bases ('.' or a foreign character) zeroes the product: no combination is ever valid
        if (too_many || prod == 0 || nv == 0 || nv > 8) continue;
        uint32_t cand = 0, n_cand = 0;
        for (uint32_t j = 0; j < nv * 4 && n_cand <= 1; ++j) {
            bool valid = true;
            uint32_t sub[8];
            for (uint32_t t = 0; t < nv && valid; ++t) {
                sub[t] = (j >> ((t << 1) & 63)) & 3u;
                valid = (vmask[t] >> sub[t]) & 1u;
            }
            if (!valid || ((cand >> sub[self]) & 1u)) continue;
            // the substituted window, staged once for the warp
            __syncwarp();
            for (uint32_t x = lane; x < (uint32_t)wlen; x += 32) win[x] = s[min_pos + x];
            __syncwarp();
            if (lane == 0) for (uint32_t t = 0; t < nv; ++t) if (vpos[t] < (uint32_t)wlen) win[vpos[t]] = "ACGT"[sub[t]];
            __syncwarp();
            bool hit = false;
            const uint32_t n_km = (uint32_t)wlen - k + 1;
            for (uint32_t st = lane; st < n_km; st += 32) {
                KT fw = 0;
                bool ok = true;
                for (uint32_t x = 0; x < k; ++x) {
                    const uint32_t c = rtk_base_code(win[st + x]);
                    ok &= (c < 4);
                    fw = ((fw << 2) | (KT)(c & 3u)) & kmask;
                }
                if (ok) {
                    rtk_kmer_hit h;
                    hit |= rtk_lookup<KT>(p.table, p.n_buckets, p.pool, (int)k, fw, KmerOps<KT>::rc(fw, (int)k), h);
                }
            }
            if (__any_sync(0xffffffffu, hit)) { cand |= 1u << sub[self]; ++n_cand; }
        }
        if (n_cand == 1) {
            if (lane == 0) {
                s[i] = "ACGT"[__ffs((int)cand) - 1];
                p.amb_done[e] = 1;
                atomicAdd(p.n_fixed, 1ULL);
            }
            __syncwarp();
        }
    }
}

#endif
