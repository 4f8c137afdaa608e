// broker.hpp — asynchronous batching of the GPU services behind the host orchestration.
//
// The correction logic of a region (correct.cpp / traverse.cpp) is sequential code that needs a GPU answer
// every few lines (one alignment, one graph burst, one traceback).  Issued one by one these calls are
// launch-latency bound.  The broker runs tens of thousands of regions concurrently as FIBERS (stackful
// user-level contexts, <ucontext.h>) multiplexed on a few worker threads (one per host core): a fiber that
// needs the GPU queues its request at the service of that kind and switches back to its worker's scheduler,
// which resumes another ready fiber or starts a new region.  Each service (K4 distances, K5 paths, K2/K3
// graph bursts) has its own thread(s) and forked context (stream + scratch): whenever it is free it takes
// EVERYTHING queued so far, runs it as ONE batched C-ABI call (rtk_edlib_batch, rtk_edlib_path_batch,
// rtk_explore_subgraph_batch), scatters the answers and hands the fibers back to their workers' ready lists.
// There is no global barrier: host logic, the three services and their H2D/D2H copies all overlap, and
// batches size themselves to the service latency (requests accumulate while the previous batch runs).  The
// per-region logic is untouched (and stays byte-identical to the reference).  Fibers never migrate between
// worker threads, so thread-local state (current_broker) stays valid across a park.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rtk.h"
#include "traverse.hpp"

namespace rtk {

struct DistReq {   // K4
    const std::vector<AlignJob>* jobs;
    std::vector<int32_t>* dist;
    std::vector<int32_t>* first_end;   // smallest end column carrying the distance (edlib's endLocations[0]), -1 if none
    std::vector<int32_t>* last_end;    // largest one (endLocations[numLocations-1])
};
struct PathReq {   // K5
    const std::vector<AlignJob>* jobs;
    std::vector<int32_t>* dist;
    std::vector<std::vector<uint8_t>>* ops;
};
struct SubgraphResult {
    double scores[4];
    std::vector<std::vector<PNode>> terminal, nonterminal;
    size_t end_pos_ref = 0;   // where the burst's read window starts in `ref` (after the prefix alignment, if any)
    bool explored = false;    // false: the window was empty or the path already too long (no burst was run)
};
// One step of the `explore` lambda of explorePathsBFS* (src/GraphTraversal.cpp:251-304) as ONE request: the SHW alignment of
// the path's prefix against the read window (its first end location + 1 = where the uncovered part of the window starts), then,
// if something is left to cover, the exploreSubGraph burst on that suffix of the window.
struct SubgraphReq {   // K4 (prefix) + K2/K3 + K4 (leaves)
    rtk_subgraph_call_t call;   // ref_off / ref_len / pid_off are filled by the service; max_len_path = the burst's own bound
    const std::string* ref;     // the whole read window
    const std::string* prefix;  // spelled prefix of the path being extended (nullptr / empty: the window is used from position 0)
    size_t path_len = 0, max_len_path_total = 0;   // the burst runs only if path_len < max_len_path_total and the window is not used up
    const std::vector<uint32_t>* pids;
    double wrlf;
    SubgraphResult* out;
};

// One extractSemiWeakPaths call (src/Correction.cpp:3-157) served by the device-resident region engine (region.cuh): the whole
// chain of hops, BFS queues, bursts, leaf alignments and path qualities of a weak region in ONE request.
struct RegionReq {
    const rtk_opt* opt; int pass;       // options of the correction round this request belongs to
    const std::string* s;               // the read, in the orientation the region is corrected in
    rtk_hit um_start, um_end;
    bool has_end;
    size_t end_pos;                     // pos_um_solid2 when has_end
    const std::vector<rtk_hit>* v_w;    // weak anchors of the region
    size_t i_weak;                      // first weak anchor to consider
    const std::vector<uint32_t>* pids;  // WeightsPairID::all_pids
    bool follow_dead_ends = false;      // run the dead-end restarts of the `correct` lambda on the device too (one segment each)
    // answer
    uint32_t status = 2, bail = 0;      // of the last segment: 0 complete path, 1 dead-end path; 2 = declined (use the request-at-a-time path)
    struct Seg {
        uint32_t status, start_weak;    // start_weak: index into v_w (from i_weak = 0) of the anchor the segment starts from, RTK_NONE32 = um_start
        GPath path;
        std::string seq;                // the path spelled (Path::toString)
        int32_t shw_dist, shw_first_end;
    };
    std::vector<Seg> segs;
};

// direct (un-brokered) execution of a set of requests: one batched call per kind
void run_dist_batch(rtk_ctx* ctx, const std::vector<DistReq*>& reqs);
void run_path_batch(rtk_ctx* ctx, const std::vector<PathReq*>& reqs);
void run_subgraph_batch(rtk_ctx* ctx, const std::vector<SubgraphReq*>& reqs);
void run_region_batch(rtk_ctx* ctx, const std::vector<RegionReq*>& reqs, uint64_t* kernel_ns);

class GpuBroker {
public:
    explicit GpuBroker(rtk_ctx* c);
    ~GpuBroker();
    // run task(i) for i in [0, n) with up to `inflight` tasks alive at once (as fibers on the host worker threads);
    // returns when all tasks have finished; throws the first service / task error
    void run(size_t n, unsigned inflight, const std::function<void(size_t)>& task);
    // called from inside a task (through the thread-local current broker): queue the request, resume when served
    void submit(DistReq* r);
    void submit(PathReq* r);
    void submit(SubgraphReq* r);
    void submit(RegionReq* r);
    void submit(const std::vector<RegionReq*>& rs);   // several region requests of one task, one park
    // From now on the calling task's K2-K5 requests go to the express services (own threads and streams, served at once in
    // small batches).  For the rare region the engine declines: it falls back to hundreds of DEPENDENT requests, which must
    // not queue behind the bulk batches of the other regions.
    void set_express();
    uint64_t waves = 0, jobs = 0;   // batched service calls issued / requests served
    uint64_t kernel_ns[4] = {0, 0, 0, 0};   // GPU kernel time (CUDA events) of the dist / path / subgraph / region services
    uint64_t region_calls = 0, region_bails = 0, region_bail_reason[16] = {0};
    uint64_t region_kcells = 0;   // DP cells swept inside the engine / 1024

    struct Worker;    // one host thread: scheduler context, live fibers, ready list
    struct Fiber;
    struct Service;   // one GPU service thread: request queue + forked context
    void fiber_body(Fiber* f);   // fiber entry (called by the makecontext trampoline)

private:
    void park(int kind, void* req);
    void worker_main(Worker* w);
    void service_main(Service* s);
    rtk_ctx* ctx;
    std::vector<Service*> services[4];   // 0 dist (K4), 1 path (K5), 2 subgraph (K2/K3+K4), 3 region engine
    std::vector<Service*> express[3];    // low-latency twins of services 0-2 for declined regions
    std::vector<Worker*> workers;
    std::string task_error;    // first exception that escaped a task
    // per run()
    std::atomic<size_t> live_total{0};   // fibers started and not finished, all workers
    size_t cap_total = 1;                // in-flight cap, all workers
    unsigned run_budget = 0;             // host threads of the gang that runs this broker
    bool bulk_regions = true;            // the region service launches when every live fiber waits for it (one bulk launch per wave)
    bool bulk_all = false;               // every service launches only when all live fibers are parked (global waves)
    std::atomic<size_t> parked_total{0}; // fibers parked at any service and not yet handed back
    std::chrono::steady_clock::time_point t_run_begin;
    size_t n_tasks = 0, cap_per_worker = 1;
    size_t next_task = 0;      // guarded by mu_task (tasks are handed out in index order)
    std::mutex mu_task;
    const std::function<void(size_t)>* task_fn = nullptr;
};

GpuBroker* current_broker();   // thread-local: set while a worker thread runs a task

}  // namespace rtk
