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
    uint64_t* dense;               // exact sweep only, optional: dense[seq_off[read] - seq_off[0] + l] = P | strand << 40 of position l's k-mer,
                                   // ~0 when it is not in the graph (pre-set by the caller); `hits` is then unused.  On corrected
                                   // reads (second pass, phasing) nearly every position hits: a dense array in read order needs
                                   // neither labels nor a sort, and the runs of findUnitig fall out of consecutive entries
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
        bool hit = false, probed = false;
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
            probed = !bad;
            if (probed) hit = rtk_lookup<KT>(p.table, p.n_buckets, p.pool, k, fw, rc, h);
        }
        if (p.n_probes) {
            const unsigned pm = __ballot_sync(0xffffffffu, probed);
            if ((threadIdx.x & 31) == 0 && pm) atomicAdd(p.n_probes, (unsigned long long)__popc(pm));
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (p.dense) {
            if (hit) p.dense[base - p.seq_off[0] + l] = h.P | ((uint64_t)h.strand << 40);   // relative to the first read of this launch
            if ((threadIdx.x & 31) == 0 && m) atomicAdd(p.n_hits, (unsigned long long)__popc(m));
        } else if (m) {
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
// Thread t of a CTA owns read position l = t0 + t of the tile and keeps that position's forward /
// reverse-complement k-mer in registers.  Every variant-string window whose first read base is l
// is generated from those two registers; the mapping (l, shift) -> variant-string position is
// closed-form with ONE division per thread per tile (the per-shift terms are incremental), which
// keeps the kernel off the XU pipe that runtime integer division lives on.  The 3-4 letter probes of
// one slot are issued together (independent 256-bit loads in flight) and checked afterwards.
template <typename KT>
__device__ __forceinline__ void rtk_finish_probe(const rtk_k1_params& p, const int k, const rtk_probe& q, const KT fw, const KT rc,
                                                 const uint32_t read, const uint32_t var, const uint32_t x) {
    if (rtk_probe_maybe(q)) {
        rtk_kmer_hit h;
        if (rtk_lookup_slow<KT>(p.table, p.n_buckets, p.pool, k, fw, rc, h)) rtk_emit_hit(p, read, var, x, h);
    }
}

template <typename KT>
__device__ __forceinline__ void rtk_probe4_ins(const rtk_k1_params& p, const int k, const KT kmask, const KT W, const KT R,
                                               const uint32_t o, const uint32_t read, const uint32_t var0,
                                               const uint32_t x, uint32_t& probes) {
    // window = read[l, l+o) + letter + read[l+o, l+k-1)
    // fw: top o bases of W, letter at offset o, then W's bases o..k-2 moved one to the right
    const KT lowmask = (o == 0) ? kmask : (((KT)1 << (2 * (k - o))) - 1);  // bases o..k-1
    const KT fw_base = (W & ~lowmask & kmask) | ((W & lowmask) >> 2 & (lowmask >> 2));
    // rc: drop R's first base, complement letter at offset k-1-o, keep R's last o bases
    const KT tailmask = (o == 0) ? (KT)0 : (((KT)1 << (2 * o)) - 1);
    const KT rc_base = (((R << 2) & kmask) & ~(((KT)1 << (2 * (o + 1))) - 1)) | (R & tailmask);
    const KT ufw = (KT)1 << (2 * (k - 1 - o)), urc = (KT)1 << (2 * o);
    rtk_probe q[4];
#pragma unroll
    for (uint32_t a = 0; a < 4; ++a) rtk_probe_issue<KT>(p.table, p.n_buckets, fw_base | ((KT)a * ufw), rc_base | ((KT)(3 - a) * urc), q[a]);
#pragma unroll
    for (uint32_t a = 0; a < 4; ++a) rtk_finish_probe<KT>(p, k, q[a], fw_base | ((KT)a * ufw), rc_base | ((KT)(3 - a) * urc), read, var0 + a, x);
    probes += 4;
}

template <typename KT>
__global__ void __launch_bounds__(RTK_K1_THREADS, 3) rtk_k1_inexact_kernel(const rtk_k1_params p) {
    __shared__ uint8_t s_code[RTK_K1_THREADS + 64 + 8];  // 0..3 base, 4 = non-ACGT / past the end
    const int k = p.k;
    const KT kmask = KmerOps<KT>::mask(k);
    uint32_t probes = 0;  // per thread; summed into the 64-bit global counter at the end
    const uint32_t tid = threadIdx.x;

    for (uint32_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const uint32_t read = p.tiles[2 * tile], t0 = p.tiles[2 * tile + 1];
        const uint64_t base = p.seq_off[read];
        const uint32_t slen = (uint32_t)(p.seq_off[read + 1] - base);
        __syncthreads();
        for (int i = tid; i < (int)p.tile + k + 4; i += blockDim.x) {
            const uint32_t pos = t0 + i;
            s_code[i] = (pos < slen) ? (uint8_t)rtk_base_code(p.seq[base + pos]) : (uint8_t)4;
        }
        __syncthreads();
        const uint32_t l = t0 + tid;
        if (tid >= p.tile || l + k - 1 > slen) continue;  // no window starts here (pads count as non-ACGT below)
        // this position's k-mer, its reverse complement, and the number of non-ACGT bases in [l, l+k)
        KT W = 0, R = 0;
        uint32_t nbad = 0;
        for (int i = 0; i < k; ++i) {
            const uint32_t c = s_code[tid + i];
            nbad += (c >> 2);
            W = (W << 2) | (KT)(c & 3);
            R = (R >> 2) | ((KT)(3 - (c & 3)) << (2 * (k - 1)));
        }
        const uint32_t bad_last = s_code[tid + k - 1] >> 2;  // read[l+k-1]
        const uint32_t ck = s_code[tid + k];                  // read[l+k]

        // ---------------- substitution: pos_s == l, slot offset o, shift = (l + o) mod k
        if (p.do_subst && nbad == 0) {
            uint32_t shift = l % k;
            KT dfw = (KT)1 << (2 * (k - 1));  // unit of the base at offset o in W ...
            KT drc = (KT)1;                   // ... and of its complement in R
            for (int o = 0; o < k; ++o) {
                const uint32_t c = s_code[tid + o];
                // the three other letters a = c ^ d (no lane skips a probe: Search.tcc:620 writes 'N' where the
                // read already has the letter, i.e. exactly the d == 0 case); (3-a)^(3-c) == a^c == d
                rtk_probe q[3];
#pragma unroll
                for (uint32_t d = 1; d < 4; ++d) rtk_probe_issue<KT>(p.table, p.n_buckets, W ^ ((KT)d * dfw), R ^ ((KT)d * drc), q[d - 1]);
#pragma unroll
                for (uint32_t d = 1; d < 4; ++d) rtk_finish_probe<KT>(p, k, q[d - 1], W ^ ((KT)d * dfw), R ^ ((KT)d * drc), read, shift * 4 + (c ^ d), l);
                probes += 3;
                dfw >>= 2; drc <<= 2;
                shift = (shift + 1 == (uint32_t)k) ? 0 : shift + 1;
            }
        }

        // ---------------- insertion: string i = read with a slot before read positions i + m(k-1);
        // in string coordinates the slots sit at i + m*k.  With l = a(k-1) + c the windows whose first
        // read base is l start at  a*k + c (slot offset i-c, needs c <= i)  or  a*k + c + 1 (slot offset
        // k-1-(c-i), needs c >= i), plus a*k - 1 when c == 0, i == k-1 (slot first, previous period).
        if (p.do_ins && nbad - bad_last == 0) {  // needs read[l, l+k-1) only
            const uint32_t a_ = l / (uint32_t)(k - 1), c_ = l - a_ * (uint32_t)(k - 1);
            const uint32_t xk = a_ * (uint32_t)k + c_;
            const bool room_last = (l + k <= slen);  // a slot at offset k-1 exists only before an existing read base
            for (uint32_t i = 0; i < (uint32_t)k; ++i) {
                const uint32_t var0 = 4 * k + i * 4;
                const bool first = (c_ <= i);
                // w = 0: the regular window; w = 1, 2: the rare second windows (one shared, inlined probe site)
                const uint32_t nw = (c_ == i || (c_ == 0 && i == (uint32_t)k - 1)) ? 3u : 1u;
                for (uint32_t w = 0; w < nw; ++w) {
                    uint32_t o, x;
                    bool ok;
                    if (w == 0) { o = first ? (i - c_) : ((uint32_t)k - 1 - (c_ - i)); x = xk + (first ? 0u : 1u); ok = (o != (uint32_t)k - 1 || room_last); }
                    else if (w == 1) { o = (uint32_t)k - 1; x = xk + 1; ok = (c_ == i && room_last); }
                    else { o = 0; x = xk - 1; ok = (c_ == 0 && i == (uint32_t)k - 1 && a_ >= 1); }
                    if (ok) rtk_probe4_ins<KT>(p, k, kmask, W, R, o, read, var0, x, probes);
                }
            }
        }

        // ---------------- deletion: string i drops read positions i + m(k+1).  A window whose first read
        // base is R0 = l exists iff l is not itself dropped; with z = l - i - 1 = q(k+1) + r (r < k) it starts
        // at string position i + q*k + r and the junction sits r == 0 ? nowhere : k - r bases into it.
        if (p.do_del && slen >= (uint32_t)k + 1) {
            // z for i = 0, then decremented as i grows (no division inside the loop)
            int32_t q = (int32_t)(l == 0 ? 0 : (l - 1) / (uint32_t)(k + 1));
            int32_t r = (int32_t)(l == 0 ? 0 : (l - 1) - (uint32_t)q * (uint32_t)(k + 1));
            // derive the window of shift i (advances q,r); returns false if there is none / it is not all-ACGT
            auto derive = [&](const uint32_t i, KT& fw, KT& rc, uint32_t& x) -> bool {
                uint32_t o = 0;
                bool exists = (i <= (uint32_t)k);
                x = 0;
                if (l < i) { x = l; o = (i - l < (uint32_t)k) ? (i - l) : (uint32_t)k; }
                else if (l == i) exists = false;
                else {
                    if (r == k) exists = false;  // l is one of the dropped positions
                    x = i + (uint32_t)q * (uint32_t)k + (uint32_t)r;
                    o = (r == 0) ? (uint32_t)k : (uint32_t)(k - r);
                }
                if (l > i) { if (--r < 0) { r = k; --q; } }  // z -> z - 1 for the next shift
                if (!exists) return false;
                if (o == (uint32_t)k) {
                    fw = W; rc = R;
                    return nbad == 0;
                }
                // bases c_0..c_k of the read from l, c_o dropped
                const KT lowmask = ((KT)1 << (2 * (k - o))) - 1;  // bases o..k-1
                fw = (W & ~lowmask & kmask) | ((((W << 2) | (KT)(ck & 3))) & lowmask);
                const KT tailmask = ((KT)1 << (2 * o)) - 1;       // last o bases
                const KT midmask = ((((KT)1 << (2 * (k - 1))) - 1)) & ~tailmask;
                rc = ((KT)(3 - (ck & 3)) << (2 * (k - 1))) | ((R >> 2) & midmask) | (R & tailmask);
                return nbad + (ck >> 2) - (uint32_t)(s_code[tid + o] >> 2) == 0;
            };
            for (uint32_t i = 0; i <= (uint32_t)k; i += 2) {  // two shifts per round: two loads in flight
                KT fw0 = 0, rc0 = 0, fw1 = 0, rc1 = 0;
                uint32_t x0, x1;
                const bool v0 = derive(i, fw0, rc0, x0);
                const bool v1 = derive(i + 1, fw1, rc1, x1);
                rtk_probe q0, q1;
                if (v0) rtk_probe_issue<KT>(p.table, p.n_buckets, fw0, rc0, q0);
                if (v1) rtk_probe_issue<KT>(p.table, p.n_buckets, fw1, rc1, q1);
                if (v0) rtk_finish_probe<KT>(p, k, q0, fw0, rc0, read, 8 * k + i, x0);
                if (v1) rtk_finish_probe<KT>(p, k, q1, fw1, rc1, read, 8 * k + i + 1, x1);
                probes += (v0 ? 1u : 0u) + (v1 ? 1u : 0u);
            }
        }
    }
    if (p.n_probes) {
        // warp-level reduction of the probe counter (one atomic per warp)
        unsigned long long pr = probes;
        for (int off = 16; off > 0; off >>= 1) pr += __shfl_down_sync(0xffffffffu, pr, off);
        if ((threadIdx.x & 31) == 0 && pr) atomicAdd(p.n_probes, pr);
    }
}
