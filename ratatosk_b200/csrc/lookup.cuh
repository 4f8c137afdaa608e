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

// bucket b = 8 x u32 at table32 + 8*b: hi[0..3] then lo[0..3]
RTK_HD void rtk_load_hi(const uint32_t* __restrict__ t32, const uint64_t b, uint32_t hi[4]) {
#if defined(__CUDA_ARCH__)
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(t32 + 8 * b));
    hi[0] = v.x; hi[1] = v.y; hi[2] = v.z; hi[3] = v.w;
#else
    for (int i = 0; i < 4; ++i) hi[i] = t32[8 * b + i];
#endif
}

// Slow path: verify tag matches against the pool, follow bumped buckets.  Rare (tag false positives are
// 4 * 2^-24 per bucket, bumped buckets of the key's class ~1%), so it is kept out of line to keep the
// sweep loops tight; it recomputes the hash rather than carrying it in registers through the fast path.
template <typename KT>
#if defined(__CUDA_ARCH__)
__device__ __noinline__
#else
inline
#endif
bool rtk_lookup_slow(const uint64_t* __restrict__ table, const uint64_t n_buckets, const uint64_t* __restrict__ pool,
                     const int k, const KT fw, const KT rc, rtk_kmer_hit& out) {
    const uint32_t* t32 = reinterpret_cast<const uint32_t*>(table);
    const KT canon = fw < rc ? fw : rc;
    const uint64_t h = rtk_hash_kmer<KT>(canon);
    uint64_t b = rtk_bucket_of(h, n_buckets);
    const uint32_t tag8 = rtk_tag_of(h) << 8, cls = rtk_class_of(h);
    for (uint64_t probes = 0; probes < n_buckets; ++probes) {
        uint32_t hi[4];
        rtk_load_hi(t32, b, hi);
        for (int i = 0; i < 4; ++i) {
            if ((hi[i] ^ tag8) < 256u) {
                const uint64_t P = (((uint64_t)(hi[i] & 0xFu)) << 32) | (uint64_t)t32[8 * b + 4 + i];
                const KT km = rtk_pool_kmer<KT>(pool, P, k);
                if (km == fw) { out.P = P; out.strand = 1; return true; }
                if (km == rc) { out.P = P; out.strand = 0; return true; }
            }
        }
        if (!((hi[0] >> (4 + cls)) & 1u)) return false;
        b = (b + 1 == n_buckets) ? 0 : b + 1;
    }
    return false;
}

// A probe in flight: hash -> one 128-bit load of the bucket's high words.  Splitting issue from check
// lets a thread keep several independent probes outstanding (the sweeps issue the 3-4 letters of a
// slot together).  5 registers per probe.
struct rtk_probe {
    uint32_t hi[4];
    uint32_t tagc;  // tag<<8 | class
};

template <typename KT>
RTK_HD void rtk_probe_issue(const uint64_t* __restrict__ table, const uint64_t n_buckets, const KT fw, const KT rc, rtk_probe& q) {
    const KT canon = fw < rc ? fw : rc;
    const uint64_t h = rtk_hash_kmer<KT>(canon);
    q.tagc = (rtk_tag_of(h) << 8) | rtk_class_of(h);
    rtk_load_hi(reinterpret_cast<const uint32_t*>(table), rtk_bucket_of(h, n_buckets), q.hi);
}

// true if the bucket MAY hold the key or the key may have been bumped further (-> slow path)
RTK_HD bool rtk_probe_maybe(const rtk_probe& q) {
    const uint32_t tag8 = q.tagc & 0xFFFFFF00u, cls = q.tagc & 3u;
    uint32_t m = q.hi[0] ^ tag8;
    const uint32_t m1 = q.hi[1] ^ tag8, m2 = q.hi[2] ^ tag8, m3 = q.hi[3] ^ tag8;
    m = m < m1 ? m : m1; m = m < m2 ? m : m2; m = m < m3 ? m : m3;
    return (m < 256u) | (((q.hi[0] >> (4 + cls)) & 1u) != 0);
}

// fw/rc: the query k-mer and its reverse complement. Returns true on hit.
template <typename KT>
RTK_HD bool rtk_lookup(const uint64_t* __restrict__ table, const uint64_t n_buckets,
                       const uint64_t* __restrict__ pool, const int k, const KT fw, const KT rc,
                       rtk_kmer_hit& out) {
    rtk_probe q;
    rtk_probe_issue<KT>(table, n_buckets, fw, rc, q);
    if (!rtk_probe_maybe(q)) return false;
    return rtk_lookup_slow<KT>(table, n_buckets, pool, k, fw, rc, out);
}

// pool position -> unitig id (unitigs are >= k >= 2 bases so at most 64 start inside one block)
RTK_HD uint32_t rtk_unitig_of(const uint32_t* __restrict__ blk2unitig, const uint64_t* __restrict__ unitig_off,
                              const uint64_t P) {
    uint32_t u = blk2unitig[P >> 7];
    while (unitig_off[u + 1] <= P) ++u;
    return u;
}
