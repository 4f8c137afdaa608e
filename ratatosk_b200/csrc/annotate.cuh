// annotate.cuh — graph annotation kernels: detectShortCycles (src/Graph.cpp:4660-4854) and the candidate
// validation of detectSNPs (src/Graph.cpp:484-720 with isValidSNPcandidate, src/GraphTraversal.cpp:1057-1147).  Both walk the
// coloured graph itself (adjacency table, 1-bit edge flags, colour lists of the slab) with one warp per unitig; unitigs are
// independent, the walk of one unitig is sequential by definition (discovery order is the output order), and the lanes of the
// warp share the colour-set intersections (K3, rtk_warp_intersect) that dominate the work.
//
// detectShortCycles: breadth-first enumeration of the paths that leave a unitig in forward orientation and come back to it in
// forward orientation within k+1 further k-mers.  The reference copies a Path object per queue entry; here a queue entry is
// {vertex, parent entry, k-mers so far, base it was reached by} in a per-warp arena (the queue never frees, so the parent
// links spell every path).  A closed path is reported when its inner vertices are pairwise distinct (no shorter cycle inside)
// and at least min_cov_vertices read pairs colour the start unitig and every inner vertex; the report is the string of the
// bases the inner vertices were reached by (Path::getMiddleCompactedPath, src/Path.hpp:805; Path::extend :306-330).
//
// detectSNPs: the one-substitution k-mer hits of a unitig against the rest of the graph come from the K1 sweep
// (k1_lookup.cuh, substitution only) and are ordered on the host exactly like the reference's std::sort; this kernel replays
// the ordered candidates of a unitig: IUPAC union per position (seq_final / seq_tried as 4-bit base sets), the per-unitig
// verdict cache (s_valid_unitigs / s_invalid_unitigs) and isValidSNPcandidate, whose two local traversals (forward and
// backward from the unitig, through vertices sharing >= min_cov read pairs with it) keep their visited list and queue across
// the candidates of the unitig, as the reference's local_graph_traversal objects do.
//
// Arena overflow (a unitig in a tangle) is reported per unitig; the host re-runs those unitigs with a larger arena.
//
// rtk_edge_flags_kernel (end of file) is the per-unitig step of addCoverage that produces the edge flags both walks read:
// postProcessUnitigs (src/Graph.cpp:1986-2023), used by the long-read colouring (color_host.cpp).
#pragma once
#ifndef RTK_HOSTSIM
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "flat_graph.h"
#include "kmer.cuh"
#include "subgraph.cuh"   // rtk_warp_intersect

#define RTK_AN_WARPS 4
#define RTK_AN_MAXCHAIN 72           /* a short-cycle path holds at most k + 3 vertices (k <= 64) */
#define RTK_SNP_LIMIT 65536u         /* limit_sz_stack of isValidSNPcandidate (src/GraphTraversal.hpp:71) */

struct rtk_an_graph {
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
    uint32_t min_cov;
};

struct rtk_cyc_params {
    rtk_an_graph g;
    const uint32_t* list;       // unitigs to process (null: first, first + 1, ...)
    uint32_t first, n;
    uint32_t* arena;            // 3 x u32 per entry, arena_cap entries per resident warp
    uint32_t arena_cap;
    uint8_t* status;            // per job: 0 no cycle, 1 cycles reported, 2 arena overflow (nothing reported is valid)
    uint32_t* out;              // records: job, length, chars padded to a multiple of 4
    unsigned long long* out_used;   // u32 words reserved (keeps counting past out_cap)
    uint64_t out_cap;
};

struct rtk_snp_job {
    uint32_t unitig;
    uint32_t cand_off, n_cand;
    uint32_t bslot_off;         // first verdict slot of the job
};

struct rtk_snp_cand {
    uint32_t slot;              // index into final / tried (one slot per distinct unitig position of the batch)
    uint32_t b;                 // candidate unitig | strand << 31
    uint32_t bslot;             // verdict slot (one per distinct candidate unitig of the job), relative to the job
    uint32_t alt;               // base of the candidate k-mer at the substituted position (0..3)
};

struct rtk_snp_params {
    rtk_an_graph g;
    const rtk_snp_job* jobs;
    uint32_t n_jobs;
    const rtk_snp_cand* cands;
    uint8_t* fin;               // seq_final as base sets (A1 C2 G4 T8), initialised to the unitig's own bases
    uint8_t* tried;             // seq_tried
    uint8_t* verdict;           // 0 unknown, 1 valid, 2 invalid
    uint32_t* arena;            // per resident warp: 4 x arena_cap u32 (visited fw, queue fw, visited bw, queue bw)
    uint32_t arena_cap;
    uint8_t* status;            // per job: 0 done, 2 arena overflow
    unsigned long long* n_walks;    // isValidSNPcandidate calls (statistics)
};

struct rtk_edge_params {
    const uint32_t* adj;
    const uint64_t* col_off;    // n + 1
    const uint32_t* col_ids;    // sorted per unitig
    uint64_t* kmcov;            // in / out: bit 63 = isBranching
    uint64_t* shared;           // in / out: bits 0..7 = edge flags
    uint32_t n, min_cov;
};

#if defined(__CUDACC__) || defined(__CUDACC_SIM__)

struct rtk_colset {
    const uint32_t* g; uint32_t ng;
    const uint32_t* l; uint32_t nl;
};

__device__ __forceinline__ rtk_colset rtk_an_colours(const rtk_an_graph& G, const uint32_t u) {
    rtk_colset c;
    const uint32_t gs = G.gset_of[u];
    if (gs != RTK_NONE32) { const uint64_t o = G.gset_off[gs]; c.g = G.gset_ids + o; c.ng = (uint32_t)(G.gset_off[gs + 1] - o); }
    else { c.g = nullptr; c.ng = 0; }
    const uint64_t o = G.loc_off[u];
    c.l = G.loc_ids + o; c.nl = (uint32_t)(G.loc_off[u + 1] - o);
    return c;
}

// |A ∩ B| >= need (getNumberSharedPairID, src/Common.cpp:51-73; global and local part of a SharedPairID are disjoint)
__device__ __forceinline__ bool rtk_an_share(const rtk_colset& A, const rtk_colset& B, const uint32_t need, const uint32_t lane) {
    if (need == 0) return true;
    uint32_t cnt = 0;
    if (A.ng && A.g == B.g) cnt = A.ng;
    else cnt = rtk_warp_intersect(A.g, A.ng, B.g, B.ng, need, lane);
    if (cnt >= need) return true;
    if (!(A.ng && A.g == B.g)) {
        cnt += rtk_warp_intersect(A.g, A.ng, B.l, B.nl, need - cnt, lane);
        if (cnt >= need) return true;
        cnt += rtk_warp_intersect(A.l, A.nl, B.g, B.ng, need - cnt, lane);
        if (cnt >= need) return true;
    }
    cnt += rtk_warp_intersect(A.l, A.nl, B.l, B.nl, need - cnt, lane);
    return cnt >= need;
}

__device__ __forceinline__ bool rtk_an_has(const uint32_t* __restrict__ X, const uint32_t n, const uint32_t v) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (X[mid] < v) lo = mid + 1; else hi = mid; }
    return lo < n && X[lo] == v;
}

// successor of (cu, cs) through base b in traversal orientation (NeighborIterator.tcc:25-47 over the adj table), with the
// 1-bit edge flag of the current unitig (UnitigData::getSharedPids, src/UnitigData.hpp:275-284); RTK_NONE32 when absent / unflagged
__device__ __forceinline__ uint32_t rtk_an_succ(const rtk_an_graph& G, const uint32_t cu, const uint32_t cs, const uint64_t shared_w,
                                                const uint32_t b) {
    const uint32_t slot = cs ? G.adj[8 * (uint64_t)cu + b] : G.adj[8 * (uint64_t)cu + 4 + (3 - b)];
    if (slot == RTK_NONE32) return RTK_NONE32;
    const uint64_t bit = cs ? ((uint64_t)(1u << b) << 4) : (uint64_t)(1u << b);
    if (!(shared_w & bit)) return RTK_NONE32;
    const uint32_t vs = cs ? (slot >> 31) : (1u - (slot >> 31));
    return (slot & 0x7fffffffu) | (vs << 31);
}

// head k-mer of the unitig == reverse complement of its tail k-mer: the two orientations of such a unitig have the same mapped head
// (UnitigMap::getMappedHead), which is what the reference's visited sets are keyed by
__device__ __forceinline__ bool rtk_an_self_rc(const rtk_an_graph& G, const uint32_t u) {
    const uint64_t ub = G.unitig_off[u];
    const uint32_t sz = (uint32_t)(G.unitig_off[u + 1] - ub);
    for (uint32_t i = 0; i < G.k; ++i)
        if (rtk_pool_base(G.pool, ub + i) != 3u - rtk_pool_base(G.pool, ub + (sz - 1 - i))) return false;
    return true;
}

__device__ __forceinline__ bool rtk_an_same_head(const rtk_an_graph& G, const uint32_t x, const uint32_t y) {
    if (x == y) return true;
    if ((x ^ y) != 0x80000000u) return false;
    return rtk_an_self_rc(G, x & 0x7fffffffu);
}

// ------------------------------------------------------------------------------------------------ detectShortCycles
__global__ void __launch_bounds__(RTK_AN_WARPS * 32) rtk_cycles_kernel(const rtk_cyc_params p) {
    __shared__ uint32_t s_chain[RTK_AN_WARPS][RTK_AN_MAXCHAIN];   // vertices of the closed path, start first
    __shared__ uint8_t s_base[RTK_AN_WARPS][RTK_AN_MAXCHAIN];     // base each vertex was reached by
    const rtk_an_graph& G = p.g;
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * RTK_AN_WARPS + w, nw = gridDim.x * RTK_AN_WARPS;
    uint32_t* Q = p.arena + (uint64_t)gw * 3u * p.arena_cap;
    const uint32_t k = G.k, min_cov = G.min_cov;
    uint32_t* chain = s_chain[w];
    uint8_t* cbase = s_base[w];

    for (uint32_t job = gw; job < p.n; job += nw) {
        const uint32_t u = p.list ? p.list[job] : p.first + job;
        const uint32_t start = u | 0x80000000u;
        const rtk_colset Cs = rtk_an_colours(G, u);
        uint32_t head = 0, tail = 1, found = 0;
        bool overflow = false;
        __syncwarp();
        if (lane == 0) { Q[0] = start; Q[1] = RTK_NONE32; Q[2] = 0; }
        __syncwarp();
        while (head < tail && !overflow) {
            const uint32_t ei = head++;
            const uint32_t cur = Q[3 * ei], clen = Q[3 * ei + 2] & 0x0fffffffu;
            const uint32_t cu = cur & 0x7fffffffu, cs = cur >> 31;
            const uint64_t shared_w = G.shared[cu];
            if (!(shared_w & 0xffULL)) continue;                       // no flagged edge at all
            // the unitig the path stands on must be read-compatible with the start unitig (:4694; independent of the successor)
            const rtk_colset Cc = rtk_an_colours(G, cu);
            if (!rtk_an_share(Cc, Cs, min_cov, lane)) continue;
            for (uint32_t b = 0; b < 4 && !overflow; ++b) {
                const uint32_t v = rtk_an_succ(G, cu, cs, shared_w, b);
                if (v == RTK_NONE32) continue;
                if (v == start) {
                    // closed: vertices start, m_1 .. m_j (, start); walk the parent links back to the start
                    uint32_t n = 0;
                    for (uint32_t e = ei; e != RTK_NONE32; e = Q[3 * e + 1]) ++n;       // vertices up to cur, start included
                    if (lane == 0) {
                        uint32_t i = n;
                        for (uint32_t e = ei; e != RTK_NONE32; e = Q[3 * e + 1]) { --i; chain[i] = Q[3 * e]; cbase[i] = (uint8_t)(Q[3 * e + 2] >> 28); }
                    }
                    __syncwarp();
                    const uint32_t j = n - 1;                                             // inner vertices chain[1..j]
                    // no inner vertex twice (:4707-4711)
                    bool dup = false;
                    for (uint32_t i = 1 + lane; i <= j; i += 32)
                        for (uint32_t t = 1; t < i && !dup; ++t) dup = rtk_an_same_head(G, chain[i], chain[t]);
                    const bool no_shorter = !__any_sync(0xffffffffu, dup);
                    bool supported = false;
                    if (no_shorter) {
                        // read pairs colouring the start AND every inner vertex (:4715-4719)
                        uint32_t cnt = 0;
                        for (uint32_t part = 0; part < 2 && cnt < min_cov; ++part) {
                            const uint32_t* X = part ? Cs.l : Cs.g;
                            const uint32_t nx = part ? Cs.nl : Cs.ng;
                            for (uint32_t base = 0; base < nx && cnt < min_cov; base += 32) {
                                const uint32_t i = base + lane;
                                bool in = i < nx;
                                if (in) {
                                    const uint32_t id = X[i];
                                    for (uint32_t t = 1; t <= j && in; ++t) {
                                        const rtk_colset Cm = rtk_an_colours(G, chain[t] & 0x7fffffffu);
                                        in = rtk_an_has(Cm.g, Cm.ng, id) || rtk_an_has(Cm.l, Cm.nl, id);
                                    }
                                }
                                cnt += __popc(__ballot_sync(0xffffffffu, in));
                            }
                        }
                        supported = cnt >= min_cov;
                    }
                    if (supported) {
                        const uint32_t words = 2 + (j + 3) / 4;
                        unsigned long long o = 0;
                        if (lane == 0) o = atomicAdd(p.out_used, (unsigned long long)words);
                        o = __shfl_sync(0xffffffffu, o, 0);
                        if (o + words <= p.out_cap) {
                            if (lane == 0) { p.out[o] = job; p.out[o + 1] = j; }
                            for (uint32_t i = lane; i < (j + 3) / 4; i += 32) {
                                uint32_t x = 0;
                                for (uint32_t t = 0; t < 4; ++t) {
                                    const uint32_t c = 4 * i + t;
                                    if (c < j) x |= (uint32_t)(unsigned char)"ACGT"[cbase[1 + c]] << (8 * t);
                                }
                                p.out[o + 2 + i] = x;
                            }
                        }
                        ++found;
                    }
                    __syncwarp();
                } else if (k - 1 + clen < 2 * k) {                                       // only short cycles are followed (:4723)
                    if (tail >= p.arena_cap) { overflow = true; break; }
                    const uint32_t vu = v & 0x7fffffffu;
                    const uint32_t vlen = (uint32_t)(G.unitig_off[vu + 1] - G.unitig_off[vu]) - k + 1;
                    if (lane == 0) { Q[3 * tail] = v; Q[3 * tail + 1] = ei; Q[3 * tail + 2] = ((clen + vlen) & 0x0fffffffu) | (b << 28); }
                    ++tail;
                }
            }
            __syncwarp();
        }
        if (lane == 0) p.status[job] = overflow ? 2 : (found ? 1 : 0);
    }
}

// ------------------------------------------------------------------------------------------------ detectSNPs
struct rtk_snp_walk {     // local_graph_traversal (src/GraphTraversal.hpp:30-34): visited list (m_km) + FIFO queue (q_um)
    uint32_t* vis;
    uint32_t* que;
    uint32_t nv, qh, qt;
};

// exploreLocalGraph (src/GraphTraversal.cpp:1061-1103).  a = the unitig in the direction of this walk, Cb = colours of the candidate.
// Returns 0 / 1, or 2 when the arena is too small.
__device__ __forceinline__ uint32_t rtk_snp_explore(const rtk_an_graph& G, rtk_snp_walk& W, const uint32_t cap, const uint32_t a,
                                                    const rtk_colset& Ca, const rtk_colset& Cb, const uint32_t lane) {
    const uint32_t min_cov = G.min_cov;
    if (Ca.ng + Ca.nl < min_cov || Cb.ng + Cb.nl < min_cov) return 0;
    if (W.nv == 0) {
        if (lane == 0) { W.vis[0] = a; W.que[0] = a; }
        W.nv = 1; W.qt = 1;
        __syncwarp();
    } else if (W.nv >= RTK_SNP_LIMIT) return 1;
    while (W.qh < W.qt) {
        const uint32_t cur = W.que[W.qh++];
        const uint32_t cu = cur & 0x7fffffffu, cs = cur >> 31;
        const uint64_t shared_w = G.shared[cu];
        for (uint32_t b = 0; b < 4; ++b) {
            const uint32_t v = rtk_an_succ(G, cu, cs, shared_w, b);
            if (v == RTK_NONE32) continue;
            // m_km.insert(mapped head of the successor).second
            bool seen = false;
            for (uint32_t i = lane; i < W.nv && !seen; i += 32) {
                const uint32_t x = W.vis[i];
                seen = (x == v) || ((x ^ v) == 0x80000000u && rtk_an_self_rc(G, v & 0x7fffffffu));
            }
            if (__any_sync(0xffffffffu, seen)) continue;
            if (W.nv >= cap) return 2;
            if (lane == 0) W.vis[W.nv] = v;
            ++W.nv;
            __syncwarp();
            const rtk_colset Cv = rtk_an_colours(G, v & 0x7fffffffu);
            if (rtk_an_share(Cv, Ca, min_cov, lane)) {
                if (rtk_an_share(Cv, Cb, min_cov, lane)) return 1;        // the rest of cur's successors is never looked at again
                if (W.qt >= cap) return 2;
                if (lane == 0) W.que[W.qt] = v;
                ++W.qt;
                __syncwarp();
            }
        }
        if (W.nv >= RTK_SNP_LIMIT) return 1;
    }
    return 0;
}

// one direction of isValidSNPcandidate (:1107-1118 / :1122-1143): any vertex already visited shares enough read pairs with the
// candidate, else the walk goes on from where it stopped
__device__ __forceinline__ uint32_t rtk_snp_side(const rtk_an_graph& G, rtk_snp_walk& W, const uint32_t cap, const uint32_t a,
                                                 const rtk_colset& Ca, const rtk_colset& Cb, const uint32_t lane) {
    for (uint32_t i = 0; i < W.nv; ++i) {
        const rtk_colset Cv = rtk_an_colours(G, W.vis[i] & 0x7fffffffu);
        if (rtk_an_share(Cv, Cb, G.min_cov, lane)) return 1;
    }
    return rtk_snp_explore(G, W, cap, a, Ca, Cb, lane);
}

__global__ void __launch_bounds__(RTK_AN_WARPS * 32) rtk_snp_kernel(const rtk_snp_params p) {
    const rtk_an_graph& G = p.g;
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * RTK_AN_WARPS + w, nw = gridDim.x * RTK_AN_WARPS;
    uint32_t* A = p.arena + (uint64_t)gw * 4u * p.arena_cap;
    unsigned long long walks = 0;
    for (uint32_t job = gw; job < p.n_jobs; job += nw) {
        const rtk_snp_job J = p.jobs[job];
        const rtk_colset Ca = rtk_an_colours(G, J.unitig);
        rtk_snp_walk fw, bw;
        fw.vis = A; fw.que = A + p.arena_cap; fw.nv = fw.qh = fw.qt = 0;
        bw.vis = A + 2 * (uint64_t)p.arena_cap; bw.que = A + 3 * (uint64_t)p.arena_cap; bw.nv = bw.qh = bw.qt = 0;
        bool overflow = false;
        __syncwarp();
        for (uint32_t ci = 0; ci < J.n_cand && !overflow; ++ci) {
            const rtk_snp_cand c = p.cands[J.cand_off + ci];
            const uint8_t mk = (uint8_t)(1u << c.alt);
            const uint8_t cf = p.fin[c.slot] | mk, ct = p.tried[c.slot] | mk;
            if (p.tried[c.slot] == ct) continue;                       // that base was tried at this position before (:532)
            __syncwarp();
            if (lane == 0) p.tried[c.slot] = ct;
            uint8_t vd = p.verdict[J.bslot_off + c.bslot];
            if (vd == 0) {
                ++walks;
                const rtk_colset Cb = rtk_an_colours(G, c.b & 0x7fffffffu);
                uint32_t r = rtk_snp_side(G, fw, p.arena_cap, J.unitig | 0x80000000u, Ca, Cb, lane);
                if (r == 1) r = rtk_snp_side(G, bw, p.arena_cap, J.unitig, Ca, Cb, lane);
                if (r == 2) { overflow = true; break; }
                vd = r ? 1 : 2;
                if (lane == 0) p.verdict[J.bslot_off + c.bslot] = vd;
            }
            if (vd == 1 && lane == 0) p.fin[c.slot] = cf;
            __syncwarp();
        }
        if (lane == 0) p.status[job] = overflow ? 2 : 0;
    }
    if (lane == 0 && walks && p.n_walks) atomicAdd(p.n_walks, walks);
}

// ------------------------------------------------------------------------------------------------ edge flags (addCoverage)
// postProcessUnitigs of addCoverage (src/Graph.cpp:1986-2023): a unitig is branching when it has more than one predecessor or
// successor; the 1-bit flag of an edge (forward successors in bits 4..7, successors of the reversed unitig in bits 0..3, base
// index A1 C2 G4 T8 = last base of the neighbour's mapped head) is set when the two unitigs share >= min_cov read ids.
// One warp per unitig, the colours given as one sorted list per unitig (CSR).
__global__ void __launch_bounds__(RTK_AN_WARPS * 32) rtk_edge_flags_kernel(const rtk_edge_params p) {
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * RTK_AN_WARPS + w, nw = gridDim.x * RTK_AN_WARPS;
    for (uint32_t u = gw; u < p.n; u += nw) {
        const uint32_t* X = p.col_ids + p.col_off[u];
        const uint32_t nx = (uint32_t)(p.col_off[u + 1] - p.col_off[u]);
        uint32_t flags = 0, n_fw = 0, n_bw = 0;
        for (uint32_t b = 0; b < 4; ++b) {
            for (uint32_t dir = 0; dir < 2; ++dir) {
                // dir 0: successor of the forward unitig through base b; dir 1: successor of the reversed unitig through base b
                const uint32_t slot = dir ? p.adj[8 * (uint64_t)u + 4 + (3 - b)] : p.adj[8 * (uint64_t)u + b];
                if (slot == RTK_NONE32) continue;
                if (dir) ++n_bw; else ++n_fw;
                const uint32_t v = slot & 0x7fffffffu;
                const uint32_t* Y = p.col_ids + p.col_off[v];
                const uint32_t ny = (uint32_t)(p.col_off[v + 1] - p.col_off[v]);
                uint32_t cnt;
                if (v == u) cnt = nx;
                else cnt = rtk_warp_intersect(X, nx, Y, ny, p.min_cov, lane);
                if (cnt >= p.min_cov) flags |= dir ? (1u << b) : ((1u << b) << 4);
            }
        }
        if (lane == 0) {
            p.shared[u] = (p.shared[u] & ~0xffULL) | flags;
            p.kmcov[u] = (p.kmcov[u] & 0x7fffffffffffffffULL) | ((uint64_t)((n_fw > 1) || (n_bw > 1)) << 63);
        }
    }
}

#endif  // __CUDACC__ || __CUDACC_SIM__
