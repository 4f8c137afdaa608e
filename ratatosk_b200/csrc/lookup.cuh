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

// fw/rc: the query k-mer and its reverse complement. Returns true on hit.
template <typename KT>
RTK_HD bool rtk_lookup(const uint64_t* __restrict__ table, const uint64_t n_buckets,
                       const uint64_t* __restrict__ pool, const int k, const KT fw, const KT rc,
                       rtk_kmer_hit& out) {
    const KT canon = fw < rc ? fw : rc;
    const uint64_t h = rtk_hash_kmer<KT>(canon);
    uint64_t b = rtk_mulhi64(h, n_buckets);
    const uint64_t tag = rtk_tag_of(h);
    for (uint64_t probes = 0; probes < n_buckets; ++probes) {
#if defined(__CUDA_ARCH__)
        const ulonglong2 e01 = __ldg(reinterpret_cast<const ulonglong2*>(table + 4 * b));
        const ulonglong2 e23 = __ldg(reinterpret_cast<const ulonglong2*>(table + 4 * b) + 1);
        const uint64_t e[4] = {e01.x, e01.y, e23.x, e23.y};
#else
        const uint64_t e[4] = {table[4 * b], table[4 * b + 1], table[4 * b + 2], table[4 * b + 3]};
#endif
        bool has_empty = false;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (e[i] == 0) { has_empty = true; continue; }
            if ((e[i] >> RTK_POS_BITS) == tag) {
                const uint64_t P = e[i] & RTK_POS_MASK;
                const KT km = rtk_pool_kmer<KT>(pool, P, k);
                if (km == fw) { out.P = P; out.strand = 1; return true; }
                if (km == rc) { out.P = P; out.strand = 0; return true; }
            }
        }
        if (has_empty) return false;
        b = (b + 1 == n_buckets) ? 0 : b + 1;
    }
    return false;
}

// pool position -> unitig id (unitigs are >= k >= 2 bases so at most 64 start inside one block)
RTK_HD uint32_t rtk_unitig_of(const uint32_t* __restrict__ blk2unitig, const uint64_t* __restrict__ unitig_off,
                              const uint64_t P) {
    uint32_t u = blk2unitig[P >> 7];
    while (unitig_off[u + 1] <= P) ++u;
    return u;
}
