// rtk_internal.hpp — private declarations shared by the translation units of librtk_b200.so.
#pragma once
#ifndef RTK_HOSTSIM
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include <sched.h>

#include <atomic>
#include <chrono>
#include <time.h>
#include <mutex>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rtk.h"
#include "flat_graph.h"
#include "graph_build.hpp"
#include "seeds_resolve.hpp"

struct rtk_snp_job;    // annotate.cuh
struct rtk_snp_cand;

struct rtk_host_graph {
    rtk::rtk_slab slab;
    rtk_slab_header hdr;
    rtk_graph_view view;  // host pointers
};

namespace rtk {

void set_error(const std::string& msg);

}  // namespace rtk

#ifndef RTK_HOSTSIM
namespace rtk {
struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define RTK_CUDA(expr)                                                                                       \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess)                                                                               \
            throw rtk::CudaError(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" __FILE__ ":" +    \
                                 std::to_string(__LINE__) + ")");                                            \
    } while (0)

// process-wide tallies reported through the stats of rtk_correct_batch: kernels launched, bytes copied H2D / D2H
extern std::atomic<uint64_t> g_launches, g_h2d_bytes, g_d2h_bytes;
inline cudaError_t counted_memcpy_async(void* dst, const void* src, size_t n, cudaMemcpyKind kind, cudaStream_t st) {
    (kind == cudaMemcpyHostToDevice ? g_h2d_bytes : g_d2h_bytes) += n;
    return cudaMemcpyAsync(dst, src, n, kind, st);
}

// grow-only device / pinned-host buffers
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        if (p) RTK_CUDA(cudaFree(p));
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 4 + 256;
        RTK_CUDA(cudaMalloc(&p, want));
        cap = want;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return (T*)p; }
};

struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        if (p) RTK_CUDA(cudaFreeHost(p));
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 4 + 256;
        RTK_CUDA(cudaMallocHost(&p, want));
        cap = want;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return (T*)p; }
};

}  // namespace rtk
#endif  // !RTK_HOSTSIM

#ifdef RTK_HOSTSIM
// tests/hostsim: the kernels run on the CPU simulator, the context only carries the graph
struct rtk_ctx {
    bool has_graph = false;
    rtk_slab_header hdr;
    const rtk_host_graph* host_graph = nullptr;
};
#else
struct rtk_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // side streams: the lane-group classes of one K4/K5 batch are independent, latency-bound launches; they run
    // concurrently, forked from `stream` after the uploads and joined before the downloads (fan_out / fan_in)
    cudaStream_t side[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // graph
    const unsigned char* d_slab = nullptr;
    bool owns_slab = false;
    bool has_graph = false;
    rtk_slab_header hdr;
    rtk_graph_view dview;                 // device pointers
    const rtk_host_graph* host_graph = nullptr;  // host mirror (needed by the host-side anchor logic)
    rtk::rtk_slab host_copy;              // when the graph was adopted from device memory
    rtk_host_graph host_graph_owned;
    // scratch
    rtk::DevBuf d_seq, d_seq_off, d_tiles, d_hits, d_counters, d_aux[8], d_sub[8];
    rtk::PinBuf h_pin[16];   // pinned landing zones of the D2H copies (one slot per copy site, see PinnedD2H)
    rtk::DevBuf d_rg[7];     // region engine: [0] packed inputs, [1] results, [2] out vertices, [3] out chars, [4] counters, [5] per-warp scratch, [6] out segments
    rtk::PinBuf h_rg[3];     // region engine: [0] packed upload, [1] results + counters, [2] output pools
    rtk::DevBuf d_fs;        // fixSNPs: packed reads + ambiguity lists
    rtk::PinBuf h_fs;
    rtk::PinBuf h_dense;     // exact sweeps: dense per-position answers (8 B per read base)
    int sm_count = 148;
    // forked contexts kept for re-use (fork_acquire / fork_release): the broker's service contexts and the correction gangs are
    // needed again by every batch; re-creating them per call means re-growing their device / pinned buffers every time, and a
    // cudaMalloc / cudaFree / cudaMallocHost synchronises with whatever the device is running (a bulk region launch: 0.5 s)
    std::vector<rtk_ctx*> fork_cache[2];   // [0] default / low-priority streams, [1] high priority
    std::mutex fork_mu;
    // reads of the next exact sweep already resident in HBM (rtk_correct_batch_resident); consumed once
    const char* resident_seq = nullptr;
    const uint64_t* resident_off = nullptr;
    uint32_t resident_n = 0;
    uint64_t resident_total = 0;
};
#endif

namespace rtk {

// Make ctx's GPU the calling thread's current device for the lifetime of the guard and restore the caller's device
// afterwards (several contexts may live in one process; an entry point may be called from any host thread).  No-op on
// the CPU simulator.
struct DeviceBind {
#ifndef RTK_HOSTSIM
    int prev = -1;
    explicit DeviceBind(const rtk_ctx* c) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != c->device) RTK_CUDA(cudaSetDevice(c->device)); else prev = -1;
    }
    ~DeviceBind() { if (prev >= 0) cudaSetDevice(prev); }
#else
    explicit DeviceBind(const rtk_ctx*) {}
#endif
    DeviceBind(const DeviceBind&) = delete;
    DeviceBind& operator=(const DeviceBind&) = delete;
};

// sorted raw hits of a batch -> per-read hit lists in reference order (shared by product and hostsim)
void resolve_batch(const rtk_graph_view& hv, uint32_t n_reads, const char* seq_pool, const uint64_t* seq_off, uint32_t flags,
                   RawHitVec& raw, std::vector<std::vector<rtk_hit>>& per_read);

// K1 driver: runs the exact and/or inexact kernels over reads resident on the device and leaves the
// raw labelled hits in ctx->d_hits.  Returns raw hit count; *n_probes / *kernel_ms optional.
uint64_t k1_launch(rtk_ctx* ctx, uint32_t n_reads, const char* d_seq, const uint64_t* d_seq_off,
                   const uint64_t* h_seq_off, uint32_t flags, uint64_t* n_probes, float* kernel_ms, const char* h_seq = nullptr, bool dense = false, bool sparse_tiles = false);

#ifndef RTK_HOSTSIM
// Wait for a stream without monopolising a core: the service threads of the correction broker outnumber the spare cores, and
// a thread spinning inside cudaStreamSynchronize keeps a region worker off the CPU.  Poll + yield keeps the wake-up latency
// of spinning (no interrupt round trip) but hands the core to any runnable worker.  RTK_SPIN_SYNC=1: plain spinning.
inline void stream_wait(cudaStream_t st) {
    static const bool spin = getenv("RTK_SPIN_SYNC") != nullptr;
    if (spin) { RTK_CUDA(cudaStreamSynchronize(st)); return; }
    // Every cudaStreamQuery takes the context lock that the other service threads need for their launches and copies: a
    // thread waiting on a long kernel (a bulk region launch runs for hundreds of ms) must not hammer it.  Yield-poll for the
    // first ~30 us (short alignment batches finish within that), then sleep with a doubling period capped at 200 us.
    static const bool no_backoff = getenv("RTK_NO_WAIT_BACKOFF") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    long sleep_ns = 0;
    for (;;) {
        const cudaError_t e = cudaStreamQuery(st);
        if (e == cudaSuccess) return;
        if (e != cudaErrorNotReady) RTK_CUDA(e);
        if (sleep_ns == 0) {
            sched_yield();
            if (!no_backoff && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(30)) sleep_ns = 10000;
        } else {
            struct timespec ts = {0, sleep_ns};
            nanosleep(&ts, nullptr);
            if (sleep_ns < 200000) sleep_ns *= 2;
        }
    }
}

// Device -> host copies land in pinned memory and are copied out after the stream synchronisation.  A cudaMemcpyAsync
// into PAGEABLE memory blocks inside the driver until the preceding kernels have finished; with five service threads
// sharing one context that serialised the services on each other's kernel latency.
struct PinnedD2H {
    rtk_ctx* c;
    cudaStream_t st;
    struct Item { void* dst; const void* pin; size_t n; };
    Item items[4];
    int n_items = 0;
    PinnedD2H(rtk_ctx* ctx, cudaStream_t s) : c(ctx), st(s) {}
    void copy(int slot, void* dst, const void* src_dev, size_t n) {
        if (!n) return;
        c->h_pin[slot].reserve(n + 16);
        RTK_CUDA(counted_memcpy_async(c->h_pin[slot].p, src_dev, n, cudaMemcpyDeviceToHost, st));
        items[n_items++] = Item{dst, c->h_pin[slot].p, n};
    }
    void sync() {
        stream_wait(st);
        for (int i = 0; i < n_items; ++i) memcpy(items[i].dst, items[i].pin, items[i].n);
        n_items = 0;
    }
};

// everything enqueued on the side streams after fan_out sees the work enqueued on c->stream before it;
// fan_in(k) makes c->stream wait for side stream k
inline void fan_out(rtk_ctx* c) { RTK_CUDA(cudaEventRecord(c->ev_fork, c->stream)); }
inline cudaStream_t side_stream(rtk_ctx* c, int k) { RTK_CUDA(cudaStreamWaitEvent(c->side[k], c->ev_fork, 0)); return c->side[k]; }
inline void fan_in(rtk_ctx* c, int k) {
    RTK_CUDA(cudaEventRecord(c->ev_join[k], c->side[k]));
    RTK_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join[k], 0));
}

// K4 over pools already resident on the device.  Host arrays describe the n alignments (begin/length into
// the pools, mode, kmax).  dist gets n entries; when want_ends, *ends / *ends_off are malloc'd dense lists.
struct MyersJobs {
    uint32_t n;
    const uint64_t* q_beg; const uint32_t* q_len;
    const uint64_t* t_beg; const uint32_t* t_len;
    const uint8_t* mode; const int32_t* kmax;   // kmax may be null (= -1 everywhere)
};
void myers_run(rtk_ctx* c, const char* d_qpool, const char* d_tpool, const MyersJobs& j, int32_t* dist, bool want_ends,
               int32_t** ends, uint64_t** ends_off, float* kernel_ms);
// host pools in, distance + first / last best end out (the broker's K4 service); job i = q_pool[q_beg[i], +q_len[i]) vs
// t_pool[t_beg[i], +t_len[i]) (jobs may share target bytes); stats[2] += kernel ns
void dist_batch_lean(rtk_ctx* c, uint32_t n, const char* q_pool, uint64_t q_bytes, const uint64_t* q_beg, const uint32_t* q_len, const char* t_pool,
                     uint64_t t_bytes, const uint64_t* t_beg, const uint32_t* t_len, const uint8_t* mode, int32_t* dist, int32_t* first_end,
                     int32_t* last_end, uint64_t* stats);
// lean variant for internal callers: distance + first / last best end column; pools either resident (d_*) or on the host
// (h_* != nullptr: packed into the same single upload as the descriptors).  kmax is unbounded.
void myers_run_lean(rtk_ctx* c, const char* d_qpool, const char* d_tpool, const char* h_qpool, uint64_t q_bytes, const char* h_tpool,
                    uint64_t t_bytes, const MyersJobs& j, int32_t* dist, int32_t* first_end, int32_t* last_end, float* kernel_ms);
#endif

// rtk_ctx_fork with the streams at the highest / lowest priority of the device (no-op distinction on the CPU simulator)
int ctx_fork_priority(const rtk_ctx* parent, bool high, rtk_ctx** out);

// a fork of `parent` from its cache (created on first use) / back into the cache; cached forks die with the parent and are
// dropped when the parent's graph changes.  On the CPU simulator: plain fork / destroy.
rtk_ctx* fork_acquire(rtk_ctx* parent, bool high);
void fork_release(rtk_ctx* parent, rtk_ctx* child, bool high);

// device-resident region engine (region.cu / tests/hostsim/sim_region.cpp): n extractSemiWeakPaths calls in one launch
struct RegionBatchOut {
    std::vector<rtk_region_result_t> results;
    std::vector<rtk_path_node> nodes;
    std::vector<char> chars;
    std::vector<rtk_region_seg_t> segs;
    float kernel_ms = 0.f;
};
void region_batch_run(rtk_ctx* c, const rtk_opt& opt, int pass, uint32_t n_calls, const rtk_region_call_t* calls, const char* win_pool,
                      uint64_t win_bytes, const rtk_hit* weak_pool, uint64_t n_weak, const uint32_t* pid_pool, uint64_t n_pids, RegionBatchOut& out);

// phasing()'s whole-read NW paths as runs, only where the caller reads them (traceback.cu / tests/hostsim/sim_traceback.cpp)
struct TbNeed;
struct TbRun;
void nw_path_runs_masked(rtk_ctx* c, uint32_t n, const char* q_pool, const uint64_t* q_off, const char* t_pool, const uint64_t* t_off,
                         const TbNeed& need, std::vector<std::vector<TbRun>>& runs, float* kernel_ms);

// internal flag for search_sequence_host: the reads are mostly masked ('N'), hits will be rare -> labelled hit list instead of the
// dense per-position answer, tiles without a valid window skipped
#define RTK_SEARCH_SPARSE_HINT (1u << 16)
// full searchSequence for a host batch -> per read ordered hits
void search_sequence_host(rtk_ctx* ctx, uint32_t n_reads, const char* seq_pool, const uint64_t* seq_off, uint32_t flags,
                          std::vector<std::vector<rtk_hit>>& per_read, uint64_t* stats);

// fixSNPs (src/Alignment.cpp:846) on the device (fixsnps.cu / tests/hostsim/sim_fixsnps.cpp): reads rewritten in place;
// amb_pos / amb_off = positions of the non-ACGT characters of each read (ascending, CSR over the reads)
void fix_snps_host(rtk_ctx* c, uint32_t n_reads, char* seq_pool, const uint64_t* seq_off, const std::vector<uint32_t>& amb_pos,
                   const std::vector<uint64_t>& amb_off, uint64_t* n_fixed);
// lists the non-ACGT positions and runs fix_snps_host (phasing.cpp); returns the number of codes replaced
uint64_t fix_snps_batch_host(rtk_ctx* ctx, uint32_t n_reads, char* seq_pool, const uint64_t* seq_off);

// graph annotation (annotate.cu / tests/hostsim/sim_annotate.cpp launch the kernels of annotate.cuh; annotate_host.cpp drives them)
// detectShortCycles over unitigs list[0..n) (list == nullptr: first .. first + n): status per job (0 none, 1 cycles, 2 arena
// overflow) and the cycle records {job, length, chars padded to 4} in discovery order per job
void cycles_run(rtk_ctx* c, uint32_t min_cov, const uint32_t* list, uint32_t first, uint32_t n, uint32_t arena_cap,
                std::vector<uint8_t>& status, std::vector<uint32_t>& records, float* kernel_ms);
// detectSNPs candidate replay: fin (in: the unitigs' own bases as sets, out: seq_final) per slot
void snp_run(rtk_ctx* c, uint32_t min_cov, const std::vector<rtk_snp_job>& jobs, const std::vector<rtk_snp_cand>& cands,
             std::vector<uint8_t>& fin, uint32_t n_bslots, uint32_t arena_cap, std::vector<uint8_t>& status, uint64_t* n_walks,
             float* kernel_ms);

// addCoverage's postProcessUnitigs: branching bit (kmcov bit 63) and edge flags (shared bits 0..7) from per-unitig colour lists
void edge_flags_run(rtk_ctx* c, uint32_t min_cov, const uint64_t* col_off, const uint32_t* col_ids, uint64_t* kmcov, uint64_t* shared,
                    float* kernel_ms);

// per-read correction of a batch (correct.cpp)
void correct_batch_host(rtk_ctx* ctx, const rtk_opt& opt, int pass, uint32_t n_reads, const char* seq_pool, const uint64_t* seq_off,
                        const char* qual_pool, const uint64_t* qual_off, std::vector<std::string>& out_seq, std::vector<std::string>& out_qual,
                        uint64_t* stats);

// getSeeds host logic over the hit lists (seeds.cpp)
void get_seeds_host(rtk_ctx* ctx, const rtk_opt& opt, int pass, uint32_t n_reads, const char* seq_pool,
                    const uint64_t* seq_off, std::vector<std::vector<rtk_hit>>& solid,
                    std::vector<std::vector<rtk_hit>>& weak, uint64_t* stats);

}  // namespace rtk
