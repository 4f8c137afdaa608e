// subgraph.cuh — K2/K3: bounded DFS between anchor unitigs with colour-set threshold intersection, the
// enumeration half of exploreSubGraph (src/GraphTraversal.cpp:456-587).  The scoring half reuses K4
// (myers.cuh): every candidate path this kernel emits is spelled from the 2-bit pool into a scratch
// string and aligned against the read window (NW for terminal paths, HW for non-terminal ones,
// getScorePath src/GraphTraversal.cpp:867-909); the `>=` / `>` selection of equal-best paths in
// discovery order runs afterwards on the scored list (subgraph_host.hpp).
//
// One warp per call.  The reference expands a LIFO stack of partial paths (std::stack, :466-478): pop a
// path, visit the successors of its last unitig in A,C,G,T order (NeighborIterator.tcc:25-47 - here the
// explicit `adj` table, reading the predecessor slots complemented when the unitig is traversed in
// reverse), keep a successor iff
//   (K3) its colour set shares >= min_cov_vertices ids with the region's set (getNumberSharedPairID,
//        src/Common.cpp:73-83; skipped when the region set is empty) - a warp-cooperative threshold
//        intersection of sorted u32 lists (lanes binary-search the larger list), and
//   the 1-bit edge flag of the CURRENT unitig for that strand/base is set (UnitigData::getSharedPids,
//        src/UnitigData.hpp:275-284);
// a kept successor is (a) a TERMINAL candidate if it is the target unitig in the target's strand and the
// path is not longer than max_len_path, and (b) pushed for further expansion while depth remains, else
// a NON-TERMINAL candidate if it has at least one successor.  The stack lives in shared memory and the
// control flow is warp-uniform; the reference's per-call memo (l_m_pid) only caches the deterministic
// colour test, so it is not reproduced.
// The kernel runs twice: COUNT (sizes only) and WRITE (descriptors + spelled strings at the offsets the
// host derived from the counts), so scratch is sized exactly.
#pragma once
#ifndef RTK_HOSTSIM
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "../../include/rtk.h"
#include "flat_graph.h"
#include "kmer.cuh"

#define RTK_DFS_WARPS 2
#define RTK_DFS_MAX_NODES 40  /* pass 1: level + 1 <= 5; pass 2 (long mode): a partial path of < 1.5 k bases holds <= k/2 + 1 unitigs */
#define RTK_DFS_STACK 100     /* LIFO depth <= 3 * (nodes per partial path) + 1 */

typedef rtk_subgraph_call_t rtk_subgraph_call;  // include/rtk.h

struct rtk_cand {
    uint32_t call;
    uint32_t terminal;
    uint32_t n_nodes;
    uint32_t nodes[RTK_DFS_MAX_NODES];            // unitig | traversal strand << 31
    uint32_t last_dist, last_len;                 // mapping of the last node (terminal: the prefix up to the target k-mer)
    uint32_t path_len;                            // spelled length
    uint64_t str_off;                             // spelled string in the char scratch
};

struct rtk_dfs_params {
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
    uint32_t k;
    // calls
    const rtk_subgraph_call* calls;
    const uint32_t* pid_pool;
    uint32_t n_calls;
    // COUNT outputs / WRITE inputs
    uint32_t* n_cand;          // [call]
    uint64_t* n_chars;         // [call]
    const uint64_t* cand_off;  // [call] (WRITE)
    const uint64_t* char_off;  // [call] (WRITE)
    rtk_cand* cands;           // (WRITE)
    char* chars;               // (WRITE)
    uint32_t* overflow;        // set when a call would exceed RTK_DFS_MAX_NODES / RTK_DFS_STACK (the host then fails the batch)
};

struct rtk_dfs_frame {
    uint32_t nodes[RTK_DFS_MAX_NODES];
    uint32_t n, l, path_len;
};

#if defined(__CUDACC__) || defined(__CUDACC_SIM__)  // kernels: device compiler or the CPU simulator only
// |X ∩ P| >= need ?  X, P sorted; all 32 lanes participate.
__device__ __forceinline__ uint32_t rtk_warp_intersect(const uint32_t* __restrict__ X, const uint32_t nx,
                                                       const uint32_t* __restrict__ P, const uint32_t np,
                                                       const uint32_t need, const uint32_t lane) {
    if (nx == 0 || np == 0 || need == 0) return 0;
    const uint32_t* S = nx <= np ? X : P;   // stride over the smaller list,
    const uint32_t* B = nx <= np ? P : X;   // binary-search the bigger one
    const uint32_t ns = nx <= np ? nx : np, nbig = nx <= np ? np : nx;
    uint32_t found = 0;
    for (uint32_t base = 0; base < ns && found < need; base += 32) {
        const uint32_t i = base + lane;
        bool hit = false;
        if (i < ns) {
            const uint32_t v = S[i];
            uint32_t lo = 0, hi = nbig;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (B[mid] < v) lo = mid + 1; else hi = mid; }
            hit = (lo < nbig) && (B[lo] == v);
        }
        found += __popc(__ballot_sync(0xffffffffu, hit));
    }
    return found;
}

template <bool WRITE>
__global__ void __launch_bounds__(RTK_DFS_WARPS * 32) rtk_dfs_kernel(const rtk_dfs_params p) {
    __shared__ rtk_dfs_frame s_stack[RTK_DFS_WARPS][RTK_DFS_STACK];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t ci = blockIdx.x * RTK_DFS_WARPS + w;
    if (ci >= p.n_calls) return;
    const rtk_subgraph_call c = p.calls[ci];
    const uint32_t k = p.k;
    const uint32_t* P = p.pid_pool + c.pid_off;
    rtk_dfs_frame* st = s_stack[w];
    uint32_t sp = 0, ncand = 0;
    uint64_t nchars = 0;
    if (lane == 0) { st[0].n = 0; st[0].l = c.level; st[0].path_len = 0; }
    sp = 1;
    __syncwarp();

    while (sp > 0) {
        const rtk_dfs_frame f = st[--sp];  // every lane takes a private copy (uniform)
        __syncwarp();
        const uint32_t cur = (f.n == 0) ? (c.start_unitig | (c.start_strand << 31)) : f.nodes[f.n - 1];
        const uint32_t cu = cur & 0x7fffffffu, cs = cur >> 31;
        const uint64_t shared_w = p.shared[cu];
        for (uint32_t b = 0; b < 4; ++b) {
            // successor through base b in traversal orientation
            const uint32_t slot = cs ? p.adj[8 * (uint64_t)cu + b] : p.adj[8 * (uint64_t)cu + 4 + (3 - b)];
            if (slot == RTK_NONE32) continue;
            const uint32_t v = slot & 0x7fffffffu;
            const uint32_t vs = cs ? (slot >> 31) : (1u - (slot >> 31));
            // edge flag of the current unitig: fw mask in bits 4..7, bw mask in bits 0..3, base index A1 C2 G4 T8
            const uint64_t bit = cs ? ((uint64_t)(1u << b) << 4) : (uint64_t)(1u << b);
            if (!(shared_w & bit)) continue;
            // K3: colour threshold
            if (c.pid_len != 0) {
                const uint32_t gs = p.gset_of[v];
                uint32_t cnt = 0;
                if (gs != RTK_NONE32) {
                    const uint64_t o = p.gset_off[gs];
                    cnt = rtk_warp_intersect(p.gset_ids + o, (uint32_t)(p.gset_off[gs + 1] - o), P, c.pid_len, c.min_cov, lane);
                }
                if (cnt < c.min_cov) {
                    const uint64_t o = p.loc_off[v];
                    cnt += rtk_warp_intersect(p.loc_ids + o, (uint32_t)(p.loc_off[v + 1] - o), P, c.pid_len, c.min_cov - cnt, lane);
                }
                if (cnt < c.min_cov) continue;
            }
            const uint32_t vsize = (uint32_t)(p.unitig_off[v + 1] - p.unitig_off[v]);
            const uint32_t vfull = vsize - k + 1;  // k-mers in the whole unitig

            // emit one candidate: descriptor by lane 0, spelling by all lanes
            auto emit = [&](const uint32_t terminal, const uint32_t ldist, const uint32_t llen, const uint32_t plen) {
                if (WRITE) {
                    const uint64_t so = p.char_off[ci] + nchars;
                    if (lane == 0) {
                        rtk_cand& d = p.cands[p.cand_off[ci] + ncand];
                        d.call = ci; d.terminal = terminal; d.n_nodes = f.n + 1;
                        for (uint32_t i = 0; i < f.n; ++i) d.nodes[i] = f.nodes[i];
                        d.nodes[f.n] = v | (vs << 31);
                        d.last_dist = ldist; d.last_len = llen; d.path_len = plen; d.str_off = so;
                    }
                    // spell: node i contributes its oriented mapping minus the k-1 bases shared with the previous node
                    uint64_t o = so;
                    for (uint32_t i = 0; i <= f.n; ++i) {
                        const uint32_t nd = (i < f.n) ? f.nodes[i] : (v | (vs << 31));
                        const uint32_t u = nd & 0x7fffffffu, s = nd >> 31;
                        const uint64_t ub = p.unitig_off[u];
                        const uint32_t usz = (uint32_t)(p.unitig_off[u + 1] - ub);
                        const uint32_t dist = (i < f.n) ? 0u : ldist;
                        const uint32_t mlen = ((i < f.n) ? (usz - k + 1) : llen) + k - 1;  // oriented mapping length
                        const uint32_t skip = (i == 0) ? 0u : (k - 1);
                        for (uint32_t t = skip + lane; t < mlen; t += 32) {
                            const uint32_t pos = s ? (dist + t) : (dist + (mlen - 1 - t));
                            uint32_t base = rtk_pool_base(p.pool, ub + pos);
                            if (!s) base = 3 - base;
                            p.chars[o + (t - skip)] = "ACGT"[base];
                        }
                        o += mlen - skip;
                    }
                }
                ++ncand;
                nchars += plen;
            };

            // (a) terminal: the target unitig reached in the target's strand (:493-526)
            if (c.end_unitig != RTK_NONE32 && v == c.end_unitig && vs == c.end_strand) {
                const uint32_t ldist = vs ? 0u : c.end_dist;
                const uint32_t llen = vs ? (c.end_dist + 1) : (vsize - c.end_dist - k + 1);
                const uint32_t plen = (f.n == 0) ? (llen + k - 1) : (f.path_len + llen);
                if (plen <= c.max_len_path) emit(1u, ldist, llen, plen);
            }
            // (b) non-terminal (:530-551)
            const uint32_t nlen = (f.n == 0) ? vsize : (f.path_len + vfull);
            // pass 1: expand while depth remains (:530); pass 2: while the partial path spells < max_len_subpath bases (:660)
            const bool expand = c.max_len_subpath ? (nlen < c.max_len_subpath) : (f.l != 0);
            if (expand) {
                if (!(sp < RTK_DFS_STACK && f.n + 1 < RTK_DFS_MAX_NODES)) { if (lane == 0) *p.overflow = 1u; }
                else {
                    if (lane == 0) {
                        rtk_dfs_frame& nf = st[sp];
                        for (uint32_t i = 0; i < f.n; ++i) nf.nodes[i] = f.nodes[i];
                        nf.nodes[f.n] = v | (vs << 31);
                        nf.n = f.n + 1; nf.l = f.l ? f.l - 1 : 0; nf.path_len = nlen;
                    }
                    ++sp;
                    __syncwarp();
                }
            } else {
                bool has_succ = false;
                for (uint32_t b2 = 0; b2 < 4; ++b2) has_succ |= (p.adj[8 * (uint64_t)v + (vs ? b2 : 4 + b2)] != RTK_NONE32);
                if (has_succ) emit(0u, 0u, vfull, nlen);
            }
        }
    }
    if (!WRITE && lane == 0) { p.n_cand[ci] = ncand; p.n_chars[ci] = nchars; }
}
#endif  // __CUDACC__ || __CUDACC_SIM__
