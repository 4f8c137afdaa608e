// lookup.cuh — k-mer -> (unitig, offset, strand) lookup over the flat slab.
// Replaces CompactedDBG::find(Kmer) (Bifrost/src/CompactedDBG.tcc:751-872, :999-1119):
// same answer (the unitig holding the k-mer, the k-mer's offset in the unitig's forward
// orientation, and whether the query equals the forward k-mer), different index
// (lookup semantics are pure membership).
#pragma once
#include "flat_graph.h"
#include "kmer.cuh"

struct rtk_kmer_hit {
    uint64_t P;       // pool position of the k-mer (forward orientation of the unitig)
    uint32_t strand;  // 1: query == forward k-mer, 0: query == reverse complement
};

RTK_HD void rtk_load_bucket(const uint64_t* __restrict__ table, const uint64_t b, uint64_t e[4]) {
#if defined(__CUDA_ARCH__)
    // one 256-bit read-only load = the whole bucket = one 32-byte sector (SASS: LDG.E.256.CONSTANT)
    asm("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(e[0]), "=l"(e[1]), "=l"(e[2]), "=l"(e[3]) : "l"(table + 4 * b));
#else
    e[0] = table[4 * b]; e[1] = table[4 * b + 1]; e[2] = table[4 * b + 2]; e[3] = table[4 * b + 3];
#endif
}

// Slow path: verify tag matches against the pool, follow bumped buckets.  Rare (tag false positives are
// 4 * 2^-24 per bucket, bumped buckets a few %), so it is kept out of line to keep the sweep loops tight.
template <typename KT>
#if defined(__CUDA_ARCH__)
__device__ __noinline__
#else
inline
#endif
bool rtk_lookup_slow(const uint64_t* __restrict__ table, const uint64_t n_buckets, const uint64_t* __restrict__ pool,
                     const int k, const KT fw, const KT rc, uint64_t b, const uint32_t tag8, rtk_kmer_hit& out) {
    for (uint64_t probes = 0; probes < n_buckets; ++probes) {
        uint64_t e[4];
        rtk_load_bucket(table, b, e);
        for (int i = 0; i < 4; ++i) {
            if ((((uint32_t)(e[i] >> 32)) ^ tag8) < 256u) {
                const uint64_t P = e[i] & RTK_POS_MASK;
                const KT km = rtk_pool_kmer<KT>(pool, P, k);
                if (km == fw) { out.P = P; out.strand = 1; return true; }
                if (km == rc) { out.P = P; out.strand = 0; return true; }
            }
        }
        if (!(e[0] & RTK_BUMP_BIT)) return false;
        b = (b + 1 == n_buckets) ? 0 : b + 1;
    }
    return false;
}

// fw/rc: the query k-mer and its reverse complement. Returns true on hit.
template <typename KT>
RTK_HD bool rtk_lookup(const uint64_t* __restrict__ table, const uint64_t n_buckets,
                       const uint64_t* __restrict__ pool, const int k, const KT fw, const KT rc,
                       rtk_kmer_hit& out) {
    const KT canon = fw < rc ? fw : rc;
    const uint64_t h = rtk_hash_kmer<KT>(canon);
    const uint64_t b = rtk_bucket_of(h, n_buckets);
    const uint32_t tag8 = rtk_tag_of(h) << 8;
    uint64_t e[4];
    rtk_load_bucket(table, b, e);
    const uint32_t h0 = (uint32_t)(e[0] >> 32), h1 = (uint32_t)(e[1] >> 32), h2 = (uint32_t)(e[2] >> 32), h3 = (uint32_t)(e[3] >> 32);
    // fast path: no entry carries the tag and nothing was ever bumped past this bucket -> miss
    const bool any = ((h0 ^ tag8) < 256u) | ((h1 ^ tag8) < 256u) | ((h2 ^ tag8) < 256u) | ((h3 ^ tag8) < 256u) | ((h0 & 0x80u) != 0);
    if (!any) return false;
    return rtk_lookup_slow<KT>(table, n_buckets, pool, k, fw, rc, b, tag8, out);
}

// pool position -> unitig id (unitigs are >= k >= 2 bases so at most 64 start inside one block)
RTK_HD uint32_t rtk_unitig_of(const uint32_t* __restrict__ blk2unitig, const uint64_t* __restrict__ unitig_off,
                              const uint64_t P) {
    uint32_t u = blk2unitig[P >> 7];
    while (unitig_off[u + 1] <= P) ++u;
    return u;
}
