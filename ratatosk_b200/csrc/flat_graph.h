// flat_graph.h — the read-only graph as ONE contiguous slab of SoA/CSR arrays.
//
// This is the HBM layout of everything the correction hot path reads from
// CompactedDBG<UnitigData>.  One slab => one H2D copy
// on rank 0 and one NCCL broadcast to the other GPUs; all section offsets are
// 256-byte aligned and relative to the slab base, so the same bytes are valid on the
// host, on any device and in a cache file.
//
//   section          element              replaces (reference)
//   unitig_off       u64[n_unitigs+1]     v_unitigs / km_unitigs / h_kmers_ccov (CompactedDBG.hpp:809-813)
//   pool             u64[pool_words]      CompressedSequence 2-bit data (first base most significant)
//   table            u64[4*n_buckets]     MinimizerIndex hmap_min_unitigs (k-mer -> unitig,pos,strand)
//   blk2unitig       u32[pool_bases/128+1] pool position -> unitig id (start of the 128-base block)
//   kmcov            u64[n_unitigs]       UnitigData::kmCov_cardBranches (UnitigData.hpp:576)
//   shared           u64[n_unitigs]       UnitigData::shared_pids (UnitigData.hpp:577)
//   adj              u32[8*n_unitigs]     getSuccessors()/getPredecessors() in A,C,G,T order
//                                         (NeighborIterator.tcc:25-47); id | strand<<31, ~0u = none
//   gset_of          u32[n_unitigs]       SharedPairID global set id (~0u = none)
//   gset_off/ids     u64[n_gsets+1]/u32[] global PairID sets, de-duplicated by content (Graph.cpp:756-769)
//   loc_off/ids      u64[n_unitigs+1]/u32[] SharedPairID local PairID, sorted
//   amb_off/ids      u64[n_unitigs+1]/u32[] UnitigData::ambiguity_ids ((pos<<4)|iupac idx), sorted
//   hap_off/ids      u64[n_unitigs+1]/u32[] UnitigData::hap_ids
//   cyc_off/pool     u64[n_unitigs+1]/char[] UnitigData::compactedCycles blobs
#pragma once
#include <stdint.h>

#define RTK_SLAB_MAGIC 0x42544B5352544B31ULL /* "1KTRSKTB" */
#define RTK_SLAB_VERSION 1u
#define RTK_NONE32 0xFFFFFFFFu

struct rtk_slab_header {
    uint64_t magic;
    uint32_t version;
    uint32_t k;
    uint64_t total_bytes;
    uint64_t n_unitigs;
    uint64_t n_kmers;
    uint64_t pool_bases;
    uint64_t pool_words;
    uint64_t n_buckets;
    uint64_t n_gsets;
    uint64_t max_km_cov_graph;  // getMaxKmerCoverage(dbg, 0.001) (src/Graph.cpp:825), before max(.,opt.max_km_cov)
    // section offsets (bytes from slab base)
    uint64_t off_unitig_off, off_pool, off_table, off_blk2unitig, off_kmcov, off_shared, off_adj;
    uint64_t off_gset_of, off_gset_off, off_gset_ids, off_loc_off, off_loc_ids;
    uint64_t off_amb_off, off_amb_ids, off_hap_off, off_hap_ids, off_cyc_off, off_cyc_pool;
    uint64_t reserved[8];
};

// Pointer view over a slab (host or device, depending on `base`).
struct rtk_graph_view {
    const unsigned char* base;
    uint32_t k;
    uint64_t n_unitigs, n_kmers, pool_bases, n_buckets, n_gsets;
    const uint64_t* unitig_off;
    const uint64_t* pool;
    const uint64_t* table;
    const uint32_t* blk2unitig;
    const uint64_t* kmcov;
    const uint64_t* shared;
    const uint32_t* adj;
    const uint32_t* gset_of;
    const uint64_t* gset_off;
    const uint32_t* gset_ids;
    const uint64_t* loc_off;
    const uint32_t* loc_ids;
    const uint64_t* amb_off;
    const uint32_t* amb_ids;
    const uint64_t* hap_off;
    const uint32_t* hap_ids;
    const uint64_t* cyc_off;
    const char* cyc_pool;
};

static inline rtk_graph_view rtk_make_view(const void* base_, const rtk_slab_header& h) {
    const unsigned char* b = (const unsigned char*)base_;
    rtk_graph_view v;
    v.base = b;
    v.k = h.k;
    v.n_unitigs = h.n_unitigs; v.n_kmers = h.n_kmers; v.pool_bases = h.pool_bases;
    v.n_buckets = h.n_buckets; v.n_gsets = h.n_gsets;
    v.unitig_off = (const uint64_t*)(b + h.off_unitig_off);
    v.pool = (const uint64_t*)(b + h.off_pool);
    v.table = (const uint64_t*)(b + h.off_table);
    v.blk2unitig = (const uint32_t*)(b + h.off_blk2unitig);
    v.kmcov = (const uint64_t*)(b + h.off_kmcov);
    v.shared = (const uint64_t*)(b + h.off_shared);
    v.adj = (const uint32_t*)(b + h.off_adj);
    v.gset_of = (const uint32_t*)(b + h.off_gset_of);
    v.gset_off = (const uint64_t*)(b + h.off_gset_off);
    v.gset_ids = (const uint32_t*)(b + h.off_gset_ids);
    v.loc_off = (const uint64_t*)(b + h.off_loc_off);
    v.loc_ids = (const uint32_t*)(b + h.off_loc_ids);
    v.amb_off = (const uint64_t*)(b + h.off_amb_off);
    v.amb_ids = (const uint32_t*)(b + h.off_amb_ids);
    v.hap_off = (const uint64_t*)(b + h.off_hap_off);
    v.hap_ids = (const uint32_t*)(b + h.off_hap_ids);
    v.cyc_off = (const uint64_t*)(b + h.off_cyc_off);
    v.cyc_pool = (const char*)(b + h.off_cyc_pool);
    return v;
}
