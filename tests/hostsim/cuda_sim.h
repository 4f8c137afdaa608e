// cuda_sim.h — TEST INFRASTRUCTURE: runs the repo's CUDA kernel SOURCE on the CPU.
//
// There is no GPU in the build container, so the `-m "not gpu"` tests compile the very same
// .cuh kernel sources with g++ against this shim and execute them with one FIBER per CUDA
// thread: a block's fibers are multiplexed on one host thread and switch at every collective
// (__syncthreads, __syncwarp, shuffles, votes), so a barrier costs a few user-level context
// switches instead of futex round trips; blocks run in parallel on a small pool of host
// threads (`__shared__` variables are thread_local statics = one copy per simulated block).  This is a debugging vehicle for index
// arithmetic and bit manipulation, NOT a product path: nothing here is linked into
// librtk_b200.so, and the product fails loudly without a CUDA device (rtk_ctx_create).
#pragma once
#define RTK_HOSTSIM 1
#include <pthread.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <functional>
#include <thread>
#include <vector>

struct sim_dim3 { unsigned x = 1, y = 1, z = 1; };
struct ulonglong2 { unsigned long long x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

extern thread_local sim_dim3 threadIdx, blockIdx;
extern sim_dim3 blockDim, gridDim;
// cooperative barrier among the fibers of one block (single host thread: no atomics needed)
struct sim_bar { unsigned count = 0, gen = 0; };
void sim_yield();                       // switch to the block's scheduler (cuda_sim.cpp)
extern thread_local unsigned long long sim_progress;
static inline void sim_bar_wait(sim_bar& b, unsigned n) {
    if (n <= 1) return;
    const unsigned gen = b.gen;
    ++sim_progress;   // an arrival is progress (the deadlock detector looks for rounds without any)
    if (++b.count == n) { b.count = 0; ++b.gen; return; }
    while (b.gen == gen) sim_yield();
}
extern thread_local sim_bar sim_block_barrier;
struct sim_warp_area {
    sim_bar bar;                      // all lanes of the warp
    sim_bar gbar[5][32];              // aligned sub-groups of width 1<<w (w = 0..4), indexed by first lane
    unsigned long long slot[32];
    unsigned nl = 32;                 // live lanes of this warp (the last warp of a block may be partial)
};
extern thread_local sim_warp_area* sim_warps;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static thread_local
#define __launch_bounds__(...)
#define __CUDACC_SIM__ 1

static inline void __syncthreads() { sim_bar_wait(sim_block_barrier, blockDim.x); }
// barrier over the lanes named by `mask` (full warp or one aligned power-of-two sub-group)
static inline void sim_mask_barrier(unsigned mask) {
    sim_warp_area& w = sim_warps[threadIdx.x >> 5];
    if (mask == 0xffffffffu) { sim_bar_wait(w.bar, w.nl); return; }
    const int n = __builtin_popcount(mask);
    const int first = __builtin_ctz(mask);
    if (n == 1) return;
    sim_bar_wait(w.gbar[__builtin_ctz((unsigned)n)][first], (unsigned)n);
}
static inline void __syncwarp(unsigned mask = 0xffffffffu) { sim_mask_barrier(mask); }
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }

// All 32 lanes of a warp must call these (full-mask use only, as in the kernels).
static inline unsigned __ballot_sync(unsigned, int pred) {
    sim_warp_area& w = sim_warps[threadIdx.x >> 5];
    w.slot[threadIdx.x & 31] = pred ? 1 : 0;
    sim_bar_wait(w.bar, w.nl);
    unsigned m = 0;
    for (unsigned i = 0; i < w.nl; ++i) m |= (unsigned)(w.slot[i] & 1) << i;
    sim_bar_wait(w.bar, w.nl);
    return m;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) {
    sim_warp_area& w = sim_warps[threadIdx.x >> 5];
    unsigned long long x = 0;
    memcpy(&x, &v, sizeof(T) <= 8 ? sizeof(T) : 8);
    w.slot[threadIdx.x & 31] = x;
    sim_bar_wait(w.bar, w.nl);
    const unsigned long long y = w.slot[src & 31];
    sim_bar_wait(w.bar, w.nl);
    T r;
    memcpy(&r, &y, sizeof(T) <= 8 ? sizeof(T) : 8);
    return r;
}
// group-scoped shuffle-up (mask = the aligned sub-group of `width` lanes the caller belongs to)
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, int d, int width = 32) {
    sim_warp_area& w = sim_warps[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    unsigned long long x = 0;
    memcpy(&x, &v, sizeof(T) <= 8 ? sizeof(T) : 8);
    w.slot[lane] = x;
    sim_mask_barrier(mask);
    const int src = lane - d;
    const bool ok = (lane & (width - 1)) >= d;
    const unsigned long long y = ok ? w.slot[src] : x;
    sim_mask_barrier(mask);
    T r;
    memcpy(&r, &y, sizeof(T) <= 8 ? sizeof(T) : 8);
    return r;
}
// group-scoped vote (mask = the aligned sub-group the caller belongs to)
static inline int __any_sync(unsigned mask, int pred) {
    sim_warp_area& w = sim_warps[threadIdx.x >> 5];
    w.slot[threadIdx.x & 31] = pred ? 1 : 0;
    sim_mask_barrier(mask);
    int r = 0;
    for (int i = 0; i < 32; ++i) if ((mask >> i) & 1u) r |= (int)(w.slot[i] & 1);
    sim_mask_barrier(mask);
    return r;
}
template <typename T> static inline T __shfl_down_sync(unsigned m, T v, int d) {
    const int lane = threadIdx.x & 31;
    const T r = __shfl_sync(m, v, (lane + d) & 31);
    return (lane + d < 32) ? r : v;
}
template <typename T> static inline T __shfl_xor_sync(unsigned m, T v, int d) { return __shfl_sync(m, v, (threadIdx.x & 31) ^ d); }

static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicCAS(unsigned* p, unsigned cmp, unsigned v) {
    __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}
static inline unsigned atomicExch(unsigned* p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline unsigned atomicMax(unsigned* p, unsigned v) {
    unsigned o = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (o < v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return o;
}

// Launch `body` (a lambda calling the kernel) on grid x block host threads.
void sim_launch(unsigned grid, unsigned block, const std::function<void()>& body);
