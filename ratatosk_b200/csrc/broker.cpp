// broker.cpp — see broker.hpp.
#include "broker.hpp"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <thread>

namespace rtk {

static thread_local GpuBroker* tl_broker = nullptr;
GpuBroker* current_broker() { return tl_broker; }

// ------------------------------------------------------------------ batched execution
void run_dist_batch(rtk_ctx* ctx, const std::vector<DistReq*>& reqs) {
    std::string qp, tp;
    std::vector<uint64_t> qo(1, 0), to(1, 0);
    std::vector<uint8_t> mode;
    for (const DistReq* r : reqs)
        for (const AlignJob& j : *r->jobs) { qp += j.q; qo.push_back(qp.size()); tp += j.t; to.push_back(tp.size()); mode.push_back(j.mode); }
    const uint32_t n = (uint32_t)mode.size();
    std::vector<int32_t> dist(n + 1, -1), kmax(n + 1, -1);
    int32_t* ends = nullptr;
    uint64_t* eoff = nullptr;
    if (n) {
        qp.push_back('\0'); tp.push_back('\0');
        if (rtk_edlib_batch(ctx, n, qp.data(), qo.data(), tp.data(), to.data(), mode.data(), kmax.data(), dist.data(), &ends, &eoff, nullptr) != RTK_OK)
            throw std::runtime_error(std::string("rtk_edlib_batch: ") + rtk_last_error());
    }
    uint32_t a = 0;
    for (const DistReq* r : reqs) {
        const size_t m = r->jobs->size();
        r->dist->assign(m, -1);
        r->ends->assign(m, {});
        for (size_t i = 0; i < m; ++i, ++a) {
            (*r->dist)[i] = dist[a];
            (*r->ends)[i].assign(ends + eoff[a], ends + eoff[a + 1]);
        }
    }
    rtk_free(ends);
    rtk_free(eoff);
}

void run_path_batch(rtk_ctx* ctx, const std::vector<PathReq*>& reqs) {
    std::string qp, tp;
    std::vector<uint64_t> qo(1, 0), to(1, 0);
    std::vector<uint8_t> mode;
    for (const PathReq* r : reqs)
        for (const AlignJob& j : *r->jobs) { qp += j.q; qo.push_back(qp.size()); tp += j.t; to.push_back(tp.size()); mode.push_back(j.mode); }
    const uint32_t n = (uint32_t)mode.size();
    std::vector<int32_t> dist(n + 1, -1), end(n + 1, -1);
    std::vector<uint8_t> flags(n + 1, 0);
    uint8_t* o = nullptr;
    uint64_t* ooff = nullptr;
    if (n) {
        qp.push_back('\0'); tp.push_back('\0');
        if (rtk_edlib_path_batch(ctx, n, qp.data(), qo.data(), tp.data(), to.data(), mode.data(), dist.data(), end.data(), &o, &ooff, flags.data(), nullptr) != RTK_OK)
            throw std::runtime_error(std::string("rtk_edlib_path_batch: ") + rtk_last_error());
    }
    uint32_t a = 0;
    for (const PathReq* r : reqs) {
        const size_t m = r->jobs->size();
        r->dist->assign(m, -1);
        r->ops->assign(m, {});
        for (size_t i = 0; i < m; ++i, ++a) {
            (*r->dist)[i] = dist[a];
            (*r->ops)[i].assign(o + ooff[a], o + ooff[a + 1]);
        }
    }
    rtk_free(o);
    rtk_free(ooff);
}

void run_subgraph_batch(rtk_ctx* ctx, const std::vector<SubgraphReq*>& reqs) {
    if (reqs.empty()) return;
    // requests may carry different weak_region_len_factor values (multi-round correction): one call per value
    std::vector<bool> done(reqs.size(), false);
    for (size_t first = 0; first < reqs.size(); ++first) {
        if (done[first]) continue;
        const double wrlf = reqs[first]->wrlf;
        std::vector<size_t> idx;
        for (size_t i = first; i < reqs.size(); ++i) if (!done[i] && reqs[i]->wrlf == wrlf) { idx.push_back(i); done[i] = true; }
        std::string refs;
        std::vector<uint32_t> pids;
        std::vector<rtk_subgraph_call_t> calls;
        for (size_t i : idx) {
            rtk_subgraph_call_t c = reqs[i]->call;
            c.ref_off = refs.size(); c.ref_len = (uint32_t)reqs[i]->ref->size();
            c.pid_off = pids.size(); c.pid_len = (uint32_t)reqs[i]->pids->size();
            refs += *reqs[i]->ref;
            pids.insert(pids.end(), reqs[i]->pids->begin(), reqs[i]->pids->end());
            calls.push_back(c);
        }
        rtk_subgraph_out out;
        const uint32_t dummy = 0;
        refs.push_back('\0');
        if (rtk_explore_subgraph_batch(ctx, (uint32_t)calls.size(), calls.data(), refs.data(), refs.size() - 1, pids.empty() ? &dummy : pids.data(),
                                       pids.size(), wrlf, &out, nullptr) != RTK_OK)
            throw std::runtime_error(std::string("rtk_explore_subgraph_batch: ") + rtk_last_error());
        for (size_t ci = 0; ci < idx.size(); ++ci) {
            SubgraphResult& res = *reqs[idx[ci]]->out;
            for (int s = 0; s < 4; ++s) res.scores[s] = out.scores[4 * ci + s];
            res.terminal.clear(); res.nonterminal.clear();
            for (uint64_t pi = out.path_off[ci]; pi < out.path_off[ci + 1]; ++pi) {
                std::vector<PNode> nodes;
                for (uint64_t j = out.node_off[pi]; j < out.node_off[pi + 1]; ++j) {
                    PNode n; n.unitig = out.nodes[j].unitig; n.strand = out.nodes[j].strand; n.dist = out.nodes[j].dist; n.len = out.nodes[j].len;
                    nodes.push_back(n);
                }
                ((pi - out.path_off[ci]) < out.n_terminal[ci] ? res.terminal : res.nonterminal).push_back(std::move(nodes));
            }
        }
        rtk_subgraph_out_free(&out);
    }
}

// ------------------------------------------------------------------ broker
GpuBroker::GpuBroker(rtk_ctx* c) : ctx(c) {
    for (int i = 0; i < 2; ++i)
        if (rtk_ctx_fork(ctx, &lane[i]) != RTK_OK) throw std::runtime_error(std::string("rtk_ctx_fork: ") + rtk_last_error());
}
GpuBroker::~GpuBroker() { for (int i = 0; i < 2; ++i) rtk_ctx_destroy(lane[i]); }

template <typename R> void GpuBroker::park(std::vector<R*>& q, R* r) {
    std::unique_lock<std::mutex> lk(mu);
    q.push_back(r);
    ++waiting;
    const uint64_t my_epoch = epoch;
    if (waiting == active) cv_broker.notify_one();
    cv_worker.wait(lk, [&] { return epoch != my_epoch; });
    if (!error.empty()) throw std::runtime_error(error);
}
void GpuBroker::submit(DistReq* r) { park(q_dist, r); }
void GpuBroker::submit(PathReq* r) { park(q_path, r); }
void GpuBroker::submit(SubgraphReq* r) { park(q_sub, r); }

void GpuBroker::run(size_t n, unsigned threads, const std::function<void(size_t)>& task) {
    if (n == 0) return;
    threads = (unsigned)std::min<size_t>(std::max(1u, threads), n);
    std::atomic<size_t> next(0);
    {
        std::lock_guard<std::mutex> lk(mu);
        active = threads; waiting = 0; error.clear();
    }
    std::vector<std::thread> pool;
    std::string task_error;
    std::mutex err_mu;
    for (unsigned t = 0; t < threads; ++t) {
        pool.emplace_back([&] {
            tl_broker = this;
            try {
                for (;;) {
                    const size_t i = next.fetch_add(1);
                    if (i >= n) break;
                    task(i);
                }
            } catch (const std::exception& e) {
                std::lock_guard<std::mutex> g(err_mu);
                if (task_error.empty()) task_error = e.what();
                next.store(n);  // stop handing out work
            }
            tl_broker = nullptr;
            std::lock_guard<std::mutex> lk(mu);
            --active;
            if (waiting == active) cv_broker.notify_one();
        });
    }
    // serve the GPU from this thread
    auto t_last = std::chrono::steady_clock::now();
    for (;;) {
        std::vector<DistReq*> d;
        std::vector<PathReq*> p;
        std::vector<SubgraphReq*> s;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv_broker.wait(lk, [&] { return waiting == active; });
            if (active == 0) break;
            d.swap(q_dist); p.swap(q_path); s.swap(q_sub);
        }
        std::string err;
        const auto t0 = std::chrono::steady_clock::now();
        {
            // the three services are independent within a wave: each runs on its own context (stream + scratch)
            std::string e_sub, e_dist, e_path;
            auto timed = [](uint64_t& acc, std::string& e, const std::function<void()>& f) {
                const auto a = std::chrono::steady_clock::now();
                try { f(); } catch (const std::exception& ex) { e = ex.what(); }
                acc += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - a).count();
            };
            std::thread th_sub, th_path;
#ifdef RTK_HOSTSIM   // the CPU simulator runs one launch at a time
            if (!s.empty()) timed(ns_sub, e_sub, [&] { run_subgraph_batch(lane[0], s); });
            if (!p.empty()) timed(ns_path, e_path, [&] { run_path_batch(lane[1], p); });
#else
            if (!s.empty()) th_sub = std::thread([&] { timed(ns_sub, e_sub, [&] { run_subgraph_batch(lane[0], s); }); });
            if (!p.empty()) th_path = std::thread([&] { timed(ns_path, e_path, [&] { run_path_batch(lane[1], p); }); });
#endif
            if (!d.empty()) timed(ns_dist, e_dist, [&] { run_dist_batch(ctx, d); });
            if (th_sub.joinable()) th_sub.join();
            if (th_path.joinable()) th_path.join();
            n_sub += s.size(); n_dist += d.size(); n_path += p.size();
            err = !e_sub.empty() ? e_sub : (!e_dist.empty() ? e_dist : e_path);
        }
        ns_wait += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t0 - t_last).count();
        t_last = std::chrono::steady_clock::now();
        ns_serve += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t_last - t0).count();
        ++waves;
        jobs += d.size() + p.size() + s.size();
        {
            std::lock_guard<std::mutex> lk(mu);
            if (!err.empty()) error = err;
            waiting = 0;
            ++epoch;
        }
        cv_worker.notify_all();
    }
    for (auto& th : pool) th.join();
    if (getenv("RTK_BROKER_PROFILE"))
        fprintf(stderr, "[broker] waves=%llu  subgraph: %llu reqs %.1f ms | dist: %llu reqs %.1f ms | path: %llu reqs %.1f ms | serving %.1f ms | waiting for workers %.1f ms\n",
                (unsigned long long)waves, (unsigned long long)n_sub, ns_sub / 1e6, (unsigned long long)n_dist, ns_dist / 1e6,
                (unsigned long long)n_path, ns_path / 1e6, ns_serve / 1e6, ns_wait / 1e6);
    if (!task_error.empty()) throw std::runtime_error(task_error);
}

}  // namespace rtk
