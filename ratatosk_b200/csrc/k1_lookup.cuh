// k1_lookup.cuh — K1: batched k-mer lookup sweep (exact + 1-edit inexact) on sm_100a.
//
// Replaces the inner loops of CompactedDBG::searchSequence (Bifrost/src/Search.tcc:526-768):
//   exact        : every k-mer of the read                                   (:685-705)
//   substitution : k shifts x 4 letters, s_inexact[j]=letter for j=shift+m*k  (:717-725, :616-623)
//   insertion    : k shifts x 4 letters, one slot every k-1 read bases        (:727-746, :612-615)
//   deletion     : k+1 shifts, one read base dropped every k+1                (:748-765)
// The reference materialises each of the 9k+1 variant strings and walks it with a rolling
// minimizer; here every (variant string, position) pair is an independent work item whose
// k-mer is derived with a few bit operations from the read's own k-mer at the matching read
// position, then probed in the bucketed k-mer index (lookup.cuh).  Only hits are written out
// (they are ~1e-3 of the probes); the order-dependent bookkeeping of the reference (run
// extension, `us_pos_km` de-duplication) is replayed afterwards on the sparse hit list
// (seeds_resolve.cpp), which needs the hits labelled with (variant order, pos_s).
//
// Work decomposition: one CTA per (read, tile of RTK_K1_TILE read positions).  The tile's
// bases (+k+2 look-ahead) are staged once in shared memory, every thread builds the forward
// and reverse-complement k-mer of one read position into shared memory, then the CTA loops
// over the 3k+1 (type, shift) groups; in a group each thread owns one variant-string
// position, maps it to a read position inside the tile and issues up to 4 independent probes.
#pragma once
#ifndef RTK_HOSTSIM
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "kmer.cuh"
#include "lookup.cuh"

#ifndef RTK_K1_THREADS
#define RTK_K1_THREADS 256
#define RTK_K1_TILE 224 /* read positions per CTA: leaves room for the 1/(k-1) stretch of insertion strings */
#endif

#include "k1_lookup_layout.h"
struct rtk_raw_hit {
    uint64_t a;
    uint64_t b;
};

struct rtk_k1_params {
    const uint64_t* table;
    uint64_t n_buckets;
    const uint64_t* pool;
    int k;
    const char* seq;           // concatenated reads (upper case)
    const uint64_t* seq_off;   // n_reads+1
    const uint32_t* tiles;     // 2*n_tiles: read id, tile start
    uint32_t n_tiles;
    uint32_t tile;             // read positions per tile (<= RTK_K1_TILE; smaller for tiny k)
    uint32_t do_subst, do_ins, do_del;
    rtk_raw_hit* hits;
    unsigned long long* n_hits;
    uint64_t hit_cap;
    unsigned long long* n_probes;  // optional probe counter (may be null)
};

__device__ __forceinline__ void rtk_emit_hit(const rtk_k1_params& p, const uint32_t read, const uint32_t var,
                                             const uint32_t pos_s, const rtk_kmer_hit& h) {
    const unsigned long long idx = atomicAdd(p.n_hits, 1ULL);
    if (idx < p.hit_cap) {
        rtk_raw_hit r;
        r.a = ((uint64_t)read << (RTK_HIT_VAR_BITS + RTK_HIT_POS_BITS)) | ((uint64_t)var << RTK_HIT_POS_BITS) | (uint64_t)pos_s;
        r.b = h.P | ((uint64_t)h.strand << 40);
        p.hits[idx] = r;
    }
}

// ---------------------------------------------------------------- exact pass
// One thread per read position; warp-aggregated append (hits are dense on good reads).
template <typename KT>
__global__ void __launch_bounds__(RTK_K1_THREADS) rtk_k1_exact_kernel(const rtk_k1_params p) {
    __shared__ uint8_t s_code[RTK_K1_THREADS + 64 + 2];
    const int k = p.k;
    for (uint32_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const uint32_t read = p.tiles[2 * tile], t0 = p.tiles[2 * tile + 1];
        const uint64_t base = p.seq_off[read];
        const uint32_t slen = (uint32_t)(p.seq_off[read + 1] - base);
        __syncthreads();
        for (int i = threadIdx.x; i < RTK_K1_THREADS + k; i += blockDim.x) {
            const uint32_t pos = t0 + i;
            s_code[i] = (pos < slen) ? (uint8_t)rtk_base_code(p.seq[base + pos]) : (uint8_t)4;
        }
        __syncthreads();
        const uint32_t l = t0 + threadIdx.x;
        bool hit = false;
        rtk_kmer_hit h;
        if (l + k <= slen) {
            KT fw = 0, rc = 0;
            uint32_t bad = 0;
            for (int i = 0; i < k; ++i) {
                const uint32_t c = s_code[threadIdx.x + i];
                bad |= (c >> 2);
                fw = (fw << 2) | (KT)(c & 3);
                rc = (rc >> 2) | ((KT)(3 - (c & 3)) << (2 * (k - 1)));
            }
            if (!bad) hit = rtk_lookup<KT>(p.table, p.n_buckets, p.pool, k, fw, rc, h);
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m) {
            const int lane = threadIdx.x & 31;
            unsigned long long basei = 0;
            if (lane == (__ffs(m) - 1)) basei = atomicAdd(p.n_hits, (unsigned long long)__popc(m));
            basei = __shfl_sync(0xffffffffu, basei, __ffs(m) - 1);
            if (hit) {
                const unsigned long long idx = basei + __popc(m & ((1u << lane) - 1));
                if (idx < p.hit_cap) {
                    rtk_raw_hit r;
                    r.a = ((uint64_t)read << (RTK_HIT_VAR_BITS + RTK_HIT_POS_BITS)) | (uint64_t)l;
                    r.b = h.P | ((uint64_t)h.strand << 40);
                    p.hits[idx] = r;
                }
            }
        }
    }
}

// ---------------------------------------------------------------- inexact sweep
template <typename KT>
__global__ void __launch_bounds__(RTK_K1_THREADS) rtk_k1_inexact_kernel(const rtk_k1_params p) {
    // W(l) for l in [t0, t0+TILE] (TILE+1 positions), codes for [t0, t0+TILE+k+2)
    __shared__ KT s_fw[RTK_K1_TILE + 2];
    __shared__ KT s_rc[RTK_K1_TILE + 2];
    __shared__ uint8_t s_nbad[RTK_K1_TILE + 2];          // # non-ACGT in [l, l+k), saturated at 255
    __shared__ uint8_t s_code[RTK_K1_TILE + 64 + 8];     // 0..3 base, 4 = non-ACGT / past the end
    const int k = p.k;
    const KT kmask = KmerOps<KT>::mask(k);
    unsigned long long probes = 0;

    for (uint32_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const uint32_t read = p.tiles[2 * tile], t0 = p.tiles[2 * tile + 1];
        const uint64_t base = p.seq_off[read];
        const uint32_t slen = (uint32_t)(p.seq_off[read + 1] - base);
        __syncthreads();
        const uint32_t TILE = p.tile;
        for (int i = threadIdx.x; i < (int)TILE + k + 4; i += blockDim.x) {
            const uint32_t pos = t0 + i;
            s_code[i] = (pos < slen) ? (uint8_t)rtk_base_code(p.seq[base + pos]) : (uint8_t)4;
        }
        __syncthreads();
        if (threadIdx.x < TILE + 2) {
            KT fw = 0, rc = 0;
            uint32_t nb = 0;
            for (int i = 0; i < k; ++i) {
                const uint32_t c = s_code[threadIdx.x + i];
                nb += (c >> 2);
                fw = (fw << 2) | (KT)(c & 3);
                rc = (rc >> 2) | ((KT)(3 - (c & 3)) << (2 * (k - 1)));
            }
            s_fw[threadIdx.x] = fw;
            s_rc[threadIdx.x] = rc;
            s_nbad[threadIdx.x] = (uint8_t)(nb > 255 ? 255 : nb);
        }
        __syncthreads();
        const uint32_t t1 = t0 + TILE;  // tile = read positions [t0, t1)

        // ---------------- substitution: pos_s == l, slot offset o = (shift - l) mod k
        if (p.do_subst && threadIdx.x < TILE) {
            const uint32_t l = t0 + threadIdx.x;
            if (l + k <= slen && s_nbad[threadIdx.x] == 0) {
                const KT fw0 = s_fw[threadIdx.x], rc0 = s_rc[threadIdx.x];
                uint32_t shift = l % k;  // o = 0 first
                for (int o = 0; o < k; ++o) {
                    const uint32_t c = s_code[threadIdx.x + o];
                    const int sh_fw = 2 * (k - 1 - o), sh_rc = 2 * o;
#pragma unroll
                    for (uint32_t a = 0; a < 4; ++a) {
                        if (a == c) continue;  // Search.tcc:620 writes 'N' where the read already has the letter
                        const KT fw = fw0 ^ ((KT)(a ^ c) << sh_fw);
                        const KT rc = rc0 ^ ((KT)((3 - a) ^ (3 - c)) << sh_rc);
                        rtk_kmer_hit h;
                        ++probes;
                        if (rtk_lookup<KT>(p.table, p.n_buckets, p.pool, k, fw, rc, h)) rtk_emit_hit(p, read, shift * 4 + a, l, h);
                    }
                    shift = (shift + 1 == (uint32_t)k) ? 0 : shift + 1;
                }
            }
        }

        // ---------------- insertion: variant string i has a slot at inexact index i+m*k
        if (p.do_ins) {
            for (uint32_t i = 0; i < (uint32_t)k; ++i) {
                if (i >= slen) break;  // string would have no slot: identical to s, Search.tcc:735 loop still runs; see below
                // inexact length and the pos_s range whose first read base falls inside the tile
                const uint32_t n_after = slen - i;
                const uint32_t len_i = slen + (n_after + (k - 2)) / (k - 1);
                if (len_i < (uint32_t)k) continue;
                const uint32_t last_pos = len_i - k;
                // g(x) = x - (#slots with index < x) = read index of first read base at/after x
                auto g = [&](const uint32_t x) -> uint32_t {
                    const uint32_t a = x / k, b = x % k;
                    return x - (a + (b > i ? 1u : 0u));
                };
                uint32_t lo = t0 + t0 / (k - 1);
                lo = lo > 2 ? lo - 2 : 0;
                while (lo <= last_pos && g(lo) < t0) ++lo;
                const uint32_t x = lo + threadIdx.x;
                if (x > last_pos) continue;
                const uint32_t l = g(x);
                if (l >= t1) continue;
                const uint32_t b = x % k;
                const uint32_t o = (i + k - b) % k;  // slot offset inside the window
                const uint32_t li = l - t0;
                // window = read[l, l+o) + letter + read[l+o, l+k-1): needs k-1 read bases
                const uint32_t nb = (uint32_t)s_nbad[li] - (uint32_t)(s_code[li + k - 1] >> 2);
                if (s_nbad[li] == 255 || nb != 0) continue;
                const KT W = s_fw[li], R = s_rc[li];
                // fw: top o bases of W, letter at offset o, then W's bases o..k-2 moved one to the right
                const KT lowmask = (o == 0) ? kmask : (((KT)1 << (2 * (k - o))) - 1);  // bases o..k-1
                const KT fw_base = (W & ~lowmask & kmask) | ((W & lowmask) >> 2 & (lowmask >> 2));
                // rc: drop R's first base, complement letter at offset k-1-o, keep R's last o bases
                const KT tailmask = (o == 0) ? (KT)0 : (((KT)1 << (2 * o)) - 1);
                const KT rc_base = (((R << 2) & kmask) & ~(((KT)1 << (2 * (o + 1))) - 1)) | (R & tailmask);
#pragma unroll
                for (uint32_t a = 0; a < 4; ++a) {
                    const KT fw = fw_base | ((KT)a << (2 * (k - 1 - o)));
                    const KT rc = rc_base | ((KT)(3 - a) << (2 * o));
                    rtk_kmer_hit h;
                    ++probes;
                    if (rtk_lookup<KT>(p.table, p.n_buckets, p.pool, k, fw, rc, h)) rtk_emit_hit(p, read, 4 * k + i * 4 + a, x, h);
                }
            }
        }

        // ---------------- deletion: variant string i drops read positions i + m*(k+1)
        if (p.do_del && slen >= (uint32_t)k + 1) {
            for (uint32_t i = 0; i <= (uint32_t)k; ++i) {
                const uint32_t n_after = (i < slen) ? slen - i : 0;
                const uint32_t n_del = (n_after + k) / (k + 1);
                const uint32_t len_i = slen - n_del;
                if (len_i < (uint32_t)k) continue;
                const uint32_t last_pos = len_i - k;
                // r0(x) = read index of inexact index x
                auto r0 = [&](const uint32_t x) -> uint32_t {
                    if (x < i) return x;
                    const uint32_t y = x - i;
                    return i + (y / k) * (k + 1) + (y % k) + 1;
                };
                uint32_t lo = t0 - t0 / (k + 1);
                lo = lo > 2 ? lo - 2 : 0;
                while (lo <= last_pos && r0(lo) < t0) ++lo;
                const uint32_t x = lo + threadIdx.x;
                if (x > last_pos) continue;
                const uint32_t R0 = r0(x);
                if (R0 >= t1) continue;
                // offset of the junction inside the window (k = none: exact k-mer of the read)
                uint32_t o;
                if (x < i) o = (i - x < (uint32_t)k) ? (i - x) : (uint32_t)k;
                else { const uint32_t ym = (x - i) % k; o = (ym == 0) ? (uint32_t)k : (uint32_t)k - ym; }
                const uint32_t li = R0 - t0;
                KT fw, rc;
                if (o == (uint32_t)k) {
                    if (s_nbad[li] != 0) continue;
                    fw = s_fw[li]; rc = s_rc[li];
                } else {
                    // bases c_0..c_k of the read from R0, c_o deleted
                    const uint32_t ck = s_code[li + k];
                    const uint32_t nb = (uint32_t)s_nbad[li] + (ck >> 2) - (uint32_t)(s_code[li + o] >> 2);
                    if (s_nbad[li] == 255 || nb != 0) continue;
                    const KT W = s_fw[li], R = s_rc[li];
                    const KT lowmask = ((KT)1 << (2 * (k - o))) - 1;  // bases o..k-1
                    fw = (W & ~lowmask & kmask) | ((((W << 2) | (KT)(ck & 3))) & lowmask);
                    const KT tailmask = ((KT)1 << (2 * o)) - 1;       // last o bases
                    const KT midmask = ((((KT)1 << (2 * (k - 1))) - 1)) & ~tailmask;
                    rc = ((KT)(3 - (ck & 3)) << (2 * (k - 1))) | ((R >> 2) & midmask) | (R & tailmask);
                }
                rtk_kmer_hit h;
                ++probes;
                if (rtk_lookup<KT>(p.table, p.n_buckets, p.pool, k, fw, rc, h)) rtk_emit_hit(p, read, 8 * k + i, x, h);
            }
        }
    }
    if (p.n_probes) {
        // block-level reduction of the probe counter (one atomic per warp)
        for (int off = 16; off > 0; off >>= 1) probes += __shfl_down_sync(0xffffffffu, probes, off);
        if ((threadIdx.x & 31) == 0 && probes) atomicAdd(p.n_probes, probes);
    }
}
