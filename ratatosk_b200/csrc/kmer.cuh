// kmer.cuh — 2-bit k-mer arithmetic shared by host (graph build) and device (K1 kernels).
//
// Encoding follows the reference so that unitig / k-mer order relations carry over:
//   A=0 C=1 G=2 T=3 (Bifrost/src/Kmer.cpp:92-107), first base most significant.
// A k-mer value is right-aligned in 2k bits: v = sum base_i << 2(k-1-i).
// KT = uint64_t for k <= 32 (pass 1, k=31), unsigned __int128 for k <= 64 (pass 2, k=63).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RTK_HD __host__ __device__ __forceinline__
#else
#define RTK_HD inline
#endif

typedef unsigned __int128 rtk_u128;

// A0 C1 G2 T3; anything else (N, IUPAC, lower case is upper-cased by the caller) -> 4
RTK_HD uint32_t rtk_base_code(const char c) {
    switch (c) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        default: return 4;
    }
}

RTK_HD uint64_t rtk_rev2_64(uint64_t x) {  // reverse the order of the 32 2-bit groups
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFULL) | ((x & 0x00FF00FF00FF00FFULL) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFULL) | ((x & 0x0000FFFF0000FFFFULL) << 16);
    return (x >> 32) | (x << 32);
}

template <typename KT> struct KmerOps;

template <> struct KmerOps<uint64_t> {
    static constexpr int MAXK = 32;
    RTK_HD static uint64_t mask(const int k) { return (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1ULL); }
    RTK_HD static uint64_t rc(const uint64_t v, const int k) { return rtk_rev2_64(~v) >> (64 - 2 * k); }
    RTK_HD static uint64_t lo64(const uint64_t v) { return v; }
    RTK_HD static uint64_t hi64(const uint64_t) { return 0; }
};

template <> struct KmerOps<rtk_u128> {
    static constexpr int MAXK = 64;
    RTK_HD static rtk_u128 mask(const int k) { return (k >= 64) ? ~(rtk_u128)0 : ((((rtk_u128)1) << (2 * k)) - 1); }
    RTK_HD static rtk_u128 rc(const rtk_u128 v, const int k) {
        const uint64_t hi = (uint64_t)(v >> 64), lo = (uint64_t)v;
        const rtk_u128 r = (((rtk_u128)rtk_rev2_64(~lo)) << 64) | (rtk_u128)rtk_rev2_64(~hi);
        return r >> (128 - 2 * k);
    }
    RTK_HD static uint64_t lo64(const rtk_u128 v) { return (uint64_t)v; }
    RTK_HD static uint64_t hi64(const rtk_u128 v) { return (uint64_t)(v >> 64); }
};

RTK_HD uint64_t rtk_mix64(uint64_t x) {  // murmur3 finaliser (host-side content hashing only)
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

// Index hash: ONE 64-bit multiply (3 IMADs on the device).  The bucket comes from the high word by
// a 32-bit multiply-high range reduction, the 24-bit tag from a second, independent 32-bit mix, so a
// probe costs ~10 integer instructions before its single 32-byte load.
template <typename KT> RTK_HD uint64_t rtk_hash_kmer(const KT canon) {
    const uint64_t lo = KmerOps<KT>::lo64(canon), hi = KmerOps<KT>::hi64(canon);
    const uint64_t x = lo ^ ((hi << 29) | (hi >> 35)) ^ (hi * 0xD6E8FEB86659FD93ULL);
    return x * 0x9E3779B97F4A7C15ULL;
}

RTK_HD uint32_t rtk_mulhi32(const uint32_t a, const uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

RTK_HD uint64_t rtk_bucket_of(const uint64_t h, const uint64_t n_buckets) {  // n_buckets < 2^32
    return (uint64_t)rtk_mulhi32((uint32_t)(h >> 32), (uint32_t)n_buckets);
}

RTK_HD uint64_t rtk_mulhi64(const uint64_t a, const uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((rtk_u128)a * (rtk_u128)b) >> 64);
#endif
}

// ---- 2-bit pool: base i lives in word i/32 at bits 62-2*(i%32) (first base most significant) ----
// Extract the k-mer starting at pool base position P (right-aligned 2k bits).
template <typename KT> RTK_HD KT rtk_pool_kmer(const uint64_t* __restrict__ pool, const uint64_t P, const int k);

template <> RTK_HD uint64_t rtk_pool_kmer<uint64_t>(const uint64_t* __restrict__ pool, const uint64_t P, const int k) {
    const uint64_t w = P >> 5;
    const int sh = (int)(P & 31) * 2;
    const uint64_t hi = pool[w];
    uint64_t x = hi << sh;
    if (sh != 0 && sh + 2 * k > 64) x |= pool[w + 1] >> (64 - sh);
    return x >> (64 - 2 * k);
}

template <> RTK_HD rtk_u128 rtk_pool_kmer<rtk_u128>(const uint64_t* __restrict__ pool, const uint64_t P, const int k) {
    const uint64_t w = P >> 5;
    const int sh = (int)(P & 31) * 2;
    const uint64_t a = pool[w];
    const uint64_t b = (sh + 2 * k > 64) ? pool[w + 1] : 0;
    const uint64_t c = (sh + 2 * k > 128) ? pool[w + 2] : 0;
    rtk_u128 x = (((rtk_u128)a) << 64) | (rtk_u128)b;
    if (sh != 0) x = (x << sh) | (rtk_u128)(c >> (64 - sh));
    return x >> (128 - 2 * k);
}

RTK_HD uint32_t rtk_pool_base(const uint64_t* __restrict__ pool, const uint64_t P) {
    return (uint32_t)((pool[P >> 5] >> (62 - 2 * (int)(P & 31))) & 3ULL);
}

// ---- k-mer index: 32-byte buckets of four 8-byte entries {tag:24 | aux:4 | pool position:36} ----
// stored split: the four high words (tag | aux | pos[35:32]) first, then the four low words, so the miss
// path needs one 128-bit load of the high words only (4 registers per probe in flight).
// One bucket = one DRAM/L2 sector.  Entries are placed by linear probing at
// bucket granularity.  Whenever an insertion finds a bucket full and moves on, it records that in the
// bucket's 4-bit "bumped" filter (aux nibble of entry 0), one bit per hash class of the bumped key.
// A lookup that sees no matching tag, in a bucket whose bit for ITS class is clear, is a definite
// miss: no "empty slot" test, full-but-never-bumped buckets terminate at once, so at the default fill
// ~99% of the misses cost exactly one sector and take the branch-free fast path.
#define RTK_TAG_BITS 24
#define RTK_POS_BITS 36
#define RTK_POS_MASK ((1ULL << RTK_POS_BITS) - 1ULL)
#define RTK_AUX_SHIFT 36
#define RTK_TAG_SHIFT 40
#define RTK_BUCKET_ENTRIES 4

RTK_HD uint32_t rtk_class_of(const uint64_t h) { return ((uint32_t)h) >> 30; }  // 2 bits, independent of tag and bucket

RTK_HD uint32_t rtk_tag_of(const uint64_t h) {
    const uint32_t t = (((uint32_t)(h >> 32) * 0x85EBCA6Bu) ^ (uint32_t)h) >> 8;
    return t ? t : 1u;
}
