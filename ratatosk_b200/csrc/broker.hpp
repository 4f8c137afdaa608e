// broker.hpp — wave batching of the GPU services behind the host orchestration.
//
// The correction logic of a read (correct.cpp / traverse.cpp) is sequential code that needs a GPU answer
// every few lines (one alignment, one graph burst, one traceback).  Issued one by one these calls are
// launch-latency bound.  The broker runs many reads concurrently, each on its own host thread; a thread
// that needs the GPU parks its request and blocks; when EVERY live thread is parked, the broker thread
// concatenates all parked requests of a kind into ONE batched C-ABI call (rtk_edlib_batch,
// rtk_edlib_path_batch, rtk_explore_subgraph_batch), scatters the answers and wakes the threads.  The
// per-read logic is untouched (and stays byte-identical to the reference); the GPU sees batches of
// hundreds to thousands of independent jobs per launch instead of one.
#pragma once
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
    std::vector<std::vector<int32_t>>* ends;   // every end location, ascending
};
struct PathReq {   // K5
    const std::vector<AlignJob>* jobs;
    std::vector<int32_t>* dist;
    std::vector<std::vector<uint8_t>>* ops;
};
struct SubgraphResult {
    double scores[4];
    std::vector<std::vector<PNode>> terminal, nonterminal;
};
struct SubgraphReq {   // K2/K3 + K4
    rtk_subgraph_call_t call;   // ref_off / pid_off are filled by the broker
    const std::string* ref;
    const std::vector<uint32_t>* pids;
    double wrlf;
    SubgraphResult* out;
};

// direct (un-brokered) execution of a set of requests: one batched call per kind
void run_dist_batch(rtk_ctx* ctx, const std::vector<DistReq*>& reqs);
void run_path_batch(rtk_ctx* ctx, const std::vector<PathReq*>& reqs);
void run_subgraph_batch(rtk_ctx* ctx, const std::vector<SubgraphReq*>& reqs);

class GpuBroker {
public:
    explicit GpuBroker(rtk_ctx* c);
    ~GpuBroker();
    // run task(i) for i in [0, n) on `threads` worker threads; the CALLING thread serves the GPU until all tasks finished
    void run(size_t n, unsigned threads, const std::function<void(size_t)>& task);
    // called from worker threads (through the thread-local current broker)
    void submit(DistReq* r);
    void submit(PathReq* r);
    void submit(SubgraphReq* r);
    uint64_t waves = 0, jobs = 0;
    uint64_t ns_sub = 0, ns_dist = 0, ns_path = 0, ns_wait = 0, ns_serve = 0, n_sub = 0, n_dist = 0, n_path = 0;

private:
    template <typename R> void park(std::vector<R*>& q, R* r);
    rtk_ctx* ctx;
    rtk_ctx* lane[2] = {nullptr, nullptr};   // forks of ctx: the three services of a wave run concurrently
    std::mutex mu;
    std::condition_variable cv_broker, cv_worker;
    size_t active = 0, waiting = 0;
    uint64_t epoch = 0;
    std::vector<DistReq*> q_dist;
    std::vector<PathReq*> q_path;
    std::vector<SubgraphReq*> q_sub;
    std::string error;
};

GpuBroker* current_broker();   // thread-local: set while a worker thread runs a task

}  // namespace rtk
