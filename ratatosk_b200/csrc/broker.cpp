// broker.cpp — see broker.hpp.
#include "broker.hpp"
#include "rtk_host_common.hpp"

#include <malloc.h>
#include <sys/mman.h>
#include <ucontext.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <thread>

namespace rtk {

#ifndef RTK_HOSTSIM
extern std::atomic<uint64_t> g_myers_prof[6];
#endif
static thread_local GpuBroker* tl_broker = nullptr;
static thread_local GpuBroker::Worker* tl_worker = nullptr;
GpuBroker* current_broker() { return tl_broker; }

// ------------------------------------------------------------------ batched execution
// RTK_BROKER_PROFILE: host time spent assembling a batch, inside the C-ABI call, scattering the answers; GPU kernel time
static std::atomic<uint64_t> g_prof[4][4];
static std::atomic<uint64_t> g_region_kcells{0};
static std::atomic<uint64_t> g_region_stats[18];   // [0] calls, [1] bails, [2 + reason] bails by reason
static std::atomic<uint64_t> g_mix[2][3][2];   // [dist|path][mode NW/SHW/HW][requests with one job | with several]: requests, jobs
struct ProfTimer {
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(int kind, int slot) {
        const auto n = std::chrono::steady_clock::now();
        g_prof[kind][slot] += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(n - t).count();
        t = n;
    }
};
// K4 over host pools: distance + first / last best end column (product: the lean fused path; CPU simulator: the public entry).
// Job i = qp[q_beg[i], +q_len[i]) vs tp[t_beg[i], +t_len[i]); several jobs may point at the same target bytes.
static void dist_core(rtk_ctx* ctx, uint32_t n, std::string& qp, const std::vector<uint64_t>& q_beg, const std::vector<uint32_t>& q_len, std::string& tp,
                      const std::vector<uint64_t>& t_beg, const std::vector<uint32_t>& t_len, const std::vector<uint8_t>& mode,
                      std::vector<int32_t>& dist, std::vector<int32_t>& first, std::vector<int32_t>& last, uint64_t* st) {
    dist.assign(n + 1, -1); first.assign(n + 1, -1); last.assign(n + 1, -1);
    if (!n) return;
#ifdef RTK_HOSTSIM   // the CPU simulator goes through the public entry (contiguous pools, every end location) and keeps the two ends the callers use
    std::string q2, t2;
    std::vector<uint64_t> qo(1, 0), to(1, 0);
    for (uint32_t a = 0; a < n; ++a) { q2.append(qp, q_beg[a], q_len[a]); qo.push_back(q2.size()); t2.append(tp, t_beg[a], t_len[a]); to.push_back(t2.size()); }
    q2.push_back('\0'); t2.push_back('\0');
    std::vector<int32_t> kmax(n + 1, -1);
    int32_t* ends = nullptr;
    uint64_t* eoff = nullptr;
    if (rtk_edlib_batch(ctx, n, q2.data(), qo.data(), t2.data(), to.data(), mode.data(), kmax.data(), dist.data(), &ends, &eoff, st) != RTK_OK)
        throw std::runtime_error(std::string("rtk_edlib_batch: ") + rtk_last_error());
    for (uint32_t a = 0; a < n; ++a) if (eoff[a + 1] > eoff[a]) { first[a] = ends[eoff[a]]; last[a] = ends[eoff[a + 1] - 1]; }
    rtk_free(ends);
    rtk_free(eoff);
#else
    dist_batch_lean(ctx, n, qp.data(), qp.size(), q_beg.data(), q_len.data(), tp.data(), tp.size(), t_beg.data(), t_len.data(), mode.data(), dist.data(),
                    first.data(), last.data(), st);
#endif
}

void run_dist_batch(rtk_ctx* ctx, const std::vector<DistReq*>& reqs) {
    ProfTimer pt;
    std::string qp, tp;
    std::vector<uint64_t> qb, tb;
    std::vector<uint32_t> ql, tl;
    std::vector<uint8_t> mode;
    for (const DistReq* r : reqs) {
        if (!r->jobs->empty()) g_mix[0][(*r->jobs)[0].mode % 3][r->jobs->size() > 1] += 1;
        const std::string* prev = nullptr;   // consecutive jobs of a request that share one target string upload it once
        for (const AlignJob& j : *r->jobs) {
            qb.push_back(qp.size()); ql.push_back((uint32_t)j.q.size()); qp += j.q;
            if (j.tref && j.tref == prev) { tb.push_back(tb.back()); tl.push_back(tl.back()); }
            else { const std::string& t = j.target(); tb.push_back(tp.size()); tl.push_back((uint32_t)t.size()); tp += t; }
            prev = j.tref;
            mode.push_back(j.mode);
        }
    }
    const uint32_t n = (uint32_t)mode.size();
    std::vector<int32_t> dist, first, last;
    uint64_t st[8] = {0};
    pt.lap(0, 0);
    dist_core(ctx, n, qp, qb, ql, tp, tb, tl, mode, dist, first, last, st);
    pt.lap(0, 1);
    g_prof[0][3] += st[2];
    uint32_t a = 0;
    for (const DistReq* r : reqs) {
        const size_t m = r->jobs->size();
        r->dist->assign(dist.begin() + a, dist.begin() + a + m);
        r->first_end->assign(first.begin() + a, first.begin() + a + m);
        r->last_end->assign(last.begin() + a, last.begin() + a + m);
        a += (uint32_t)m;
    }
    pt.lap(0, 2);
}

void run_path_batch(rtk_ctx* ctx, const std::vector<PathReq*>& reqs) {
    ProfTimer pt;
    std::string qp, tp;
    std::vector<uint64_t> qo(1, 0), to(1, 0);
    std::vector<uint8_t> mode;
    for (const PathReq* r : reqs) {
        if (!r->jobs->empty()) g_mix[1][(*r->jobs)[0].mode % 3][r->jobs->size() > 1] += 1;
        for (const AlignJob& j : *r->jobs) { qp += j.q; qo.push_back(qp.size()); tp += j.target(); to.push_back(tp.size()); mode.push_back(j.mode); }
    }
    const uint32_t n = (uint32_t)mode.size();
    std::vector<int32_t> dist(n + 1, -1), end(n + 1, -1);
    std::vector<uint8_t> flags(n + 1, 0);
    uint8_t* o = nullptr;
    uint64_t* ooff = nullptr;
    uint64_t st[8] = {0};
    pt.lap(1, 0);
    if (n) {
        qp.push_back('\0'); tp.push_back('\0');
        if (rtk_edlib_path_batch(ctx, n, qp.data(), qo.data(), tp.data(), to.data(), mode.data(), dist.data(), end.data(), &o, &ooff, flags.data(), st) != RTK_OK)
            throw std::runtime_error(std::string("rtk_edlib_path_batch: ") + rtk_last_error());
    }
    pt.lap(1, 1);
    g_prof[1][3] += st[2];
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
    pt.lap(1, 2);
}

void run_subgraph_batch(rtk_ctx* ctx, const std::vector<SubgraphReq*>& reqs) {
    if (reqs.empty()) return;
    ProfTimer pt;
    // phase A: prefix alignments (SHW, first end location) of the requests that extend a non-trivial path
    {
        std::string qp, tp;
        std::vector<uint64_t> qb, tb;
        std::vector<uint32_t> ql, tl;
        std::vector<uint8_t> mode;
        std::vector<size_t> who;
        for (size_t i = 0; i < reqs.size(); ++i) {
            reqs[i]->out->end_pos_ref = 0;
            reqs[i]->out->explored = false;
            if (reqs[i]->prefix && !reqs[i]->prefix->empty()) {
                qb.push_back(qp.size()); ql.push_back((uint32_t)reqs[i]->prefix->size()); qp += *reqs[i]->prefix;
                tb.push_back(tp.size()); tl.push_back((uint32_t)reqs[i]->ref->size()); tp += *reqs[i]->ref;
                mode.push_back(1);
                who.push_back(i);
            }
        }
        if (!who.empty()) {
            std::vector<int32_t> dist, first, last;
            uint64_t st[8] = {0};
            dist_core(ctx, (uint32_t)who.size(), qp, qb, ql, tp, tb, tl, mode, dist, first, last, st);
            g_prof[2][3] += st[2];
            for (size_t x = 0; x < who.size(); ++x) reqs[who[x]]->out->end_pos_ref = (size_t)(first[x] + 1);
        }
    }
    // phase B: the bursts, on the uncovered suffix of each window (:276)
    std::vector<SubgraphReq*> live;
    for (SubgraphReq* r : reqs) {
        SubgraphResult& res = *r->out;
        for (int s4 = 0; s4 < 4; ++s4) res.scores[s4] = 0.0;
        res.terminal.clear(); res.nonterminal.clear();
        if (res.end_pos_ref <= r->ref->size() && (r->ref->size() - res.end_pos_ref) != 0 && r->path_len < r->max_len_path_total) { res.explored = true; live.push_back(r); }
    }
    // requests may carry different weak_region_len_factor values (multi-round correction): one call per value
    std::vector<bool> done(live.size(), false);
    for (size_t first = 0; first < live.size(); ++first) {
        if (done[first]) continue;
        const double wrlf = live[first]->wrlf;
        std::vector<size_t> idx;
        for (size_t i = first; i < live.size(); ++i) if (!done[i] && live[i]->wrlf == wrlf) { idx.push_back(i); done[i] = true; }
        std::string refs;
        std::vector<uint32_t> pids;
        std::vector<rtk_subgraph_call_t> calls;
        for (size_t i : idx) {
            rtk_subgraph_call_t c = live[i]->call;
            const size_t e = live[i]->out->end_pos_ref;
            c.ref_off = refs.size(); c.ref_len = (uint32_t)(live[i]->ref->size() - e);
            c.pid_off = pids.size(); c.pid_len = (uint32_t)live[i]->pids->size();
            refs.append(*live[i]->ref, e, std::string::npos);
            pids.insert(pids.end(), live[i]->pids->begin(), live[i]->pids->end());
            calls.push_back(c);
        }
        rtk_subgraph_out out;
        const uint32_t dummy = 0;
        refs.push_back('\0');
        uint64_t st[8] = {0};
        pt.lap(2, 0);
        if (rtk_explore_subgraph_batch(ctx, (uint32_t)calls.size(), calls.data(), refs.data(), refs.size() - 1, pids.empty() ? &dummy : pids.data(),
                                       pids.size(), wrlf, &out, st) != RTK_OK)
            throw std::runtime_error(std::string("rtk_explore_subgraph_batch: ") + rtk_last_error());
        pt.lap(2, 1);
        g_prof[2][3] += st[2] + st[3];
        for (size_t ci = 0; ci < idx.size(); ++ci) {
            SubgraphResult& res = *live[idx[ci]]->out;
            for (int s = 0; s < 4; ++s) res.scores[s] = out.scores[4 * ci + s];
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
        pt.lap(2, 2);
    }
}

void run_region_batch(rtk_ctx* ctx, const std::vector<RegionReq*>& reqs, uint64_t* kernel_ns) {
    if (reqs.empty()) return;
    ProfTimer pt;
    // requests of one broker share the options of their correction round; split defensively if they do not
    std::vector<bool> done(reqs.size(), false);
    for (size_t first = 0; first < reqs.size(); ++first) {
        if (done[first]) continue;
        const rtk_opt* opt = reqs[first]->opt;
        const int pass = reqs[first]->pass;
        std::vector<size_t> idx;
        for (size_t i = first; i < reqs.size(); ++i) if (!done[i] && reqs[i]->opt == opt && reqs[i]->pass == pass) { idx.push_back(i); done[i] = true; }
        const uint32_t k = opt->k;
        const size_t n = idx.size();
        std::vector<rtk_region_call_t> calls(n);
        // pool offsets first (serial prefix sums), then the pools are filled in parallel: a bulk wave holds 10^5 regions
        std::vector<uint64_t> wo(n + 1, 0), ko(n + 1, 0), po(n + 1, 0);
        for (size_t x = 0; x < n; ++x) {
            const RegionReq& r = *reqs[idx[x]];
            const size_t pos2 = r.has_end ? r.end_pos : r.s->length() - k;
            wo[x + 1] = wo[x] + (pos2 - r.um_start.pos + k);
            ko[x + 1] = ko[x] + (r.v_w->size() - std::min(r.i_weak, r.v_w->size()));
            po[x + 1] = po[x] + r.pids->size();
        }
        std::string wins(wo[n], '\0');
        std::vector<rtk_hit> weak(ko[n]);
        std::vector<uint32_t> pids(po[n]);
        parallel_for(n, [&](size_t xb, size_t xe) {
            for (size_t x = xb; x < xe; ++x) {
                const RegionReq& r = *reqs[idx[x]];
                rtk_region_call_t& c = calls[x];
                memset(&c, 0, sizeof(c));
                c.win_off = wo[x]; c.win_len = (uint32_t)(wo[x + 1] - wo[x]);
                memcpy(&wins[wo[x]], r.s->data() + r.um_start.pos, c.win_len);
                c.weak_off = ko[x]; c.n_weak = (uint32_t)(ko[x + 1] - ko[x]);
                if (c.n_weak) memcpy(&weak[ko[x]], r.v_w->data() + r.i_weak, (size_t)c.n_weak * sizeof(rtk_hit));
                c.pid_off = po[x]; c.pid_len = (uint32_t)r.pids->size();
                if (c.pid_len) memcpy(&pids[po[x]], r.pids->data(), (size_t)c.pid_len * 4);
                c.start_pos = r.um_start.pos; c.start_unitig = r.um_start.unitig; c.start_dist = r.um_start.dist; c.start_strand = r.um_start.strand;
                c.has_end = r.has_end ? 1u : 0u;
                if (r.has_end) { c.end_pos = (uint32_t)r.end_pos; c.end_unitig = r.um_end.unitig; c.end_dist = r.um_end.dist; c.end_strand = r.um_end.strand; }
                c.s_len = (uint32_t)r.s->length();
                c.reserved = r.follow_dead_ends ? 1u : 0u;
            }
        });
        pt.lap(3, 0);
        RegionBatchOut out;
        region_batch_run(ctx, *opt, pass, (uint32_t)calls.size(), calls.data(), wins.data(), wins.size(), weak.data(), weak.size(), pids.data(), pids.size(), out);
        pt.lap(3, 1);
        g_prof[3][3] += (uint64_t)(out.kernel_ms * 1e6);
        if (kernel_ns) *kernel_ns += (uint64_t)(out.kernel_ms * 1e6);
        std::atomic<uint64_t> n_bail(0);
        std::atomic<uint64_t> by_reason[16];
        for (auto& a : by_reason) a = 0;
        parallel_for(n, [&](size_t xb, size_t xe) {
            for (size_t x = xb; x < xe; ++x) {
                RegionReq& r = *reqs[idx[x]];
                const rtk_region_result_t& R = out.results[x];
                r.status = R.status; r.bail = R.bail;
                if (R.status == 2) { n_bail += 1; by_reason[std::min<uint32_t>(R.bail, 15u)] += 1; continue; }
                r.segs.resize(R.n_segs);
                for (uint32_t si = 0; si < R.n_segs; ++si) {
                    const rtk_region_seg_t& S = out.segs[R.seg_off + si];
                    RegionReq::Seg& D = r.segs[si];
                    D.status = S.status; D.start_weak = S.start_weak; D.shw_dist = S.shw_dist; D.shw_first_end = S.shw_first_end;
                    D.path.clear();
                    D.path.v.resize(S.n_nodes);
                    for (uint32_t i = 0; i < S.n_nodes; ++i) {
                        const rtk_path_node& nd = out.nodes[S.node_off + i];
                        D.path.v[i].unitig = nd.unitig; D.path.v[i].strand = nd.strand; D.path.v[i].dist = nd.dist; D.path.v[i].len = nd.len;
                    }
                    D.path.l = S.len;
                    const uint64_t pad = ((uint64_t)S.len + 7) & ~7ull;
                    D.seq.assign(out.chars.data() + S.str_off, S.len);
                    D.path.qual.assign(out.chars.data() + S.str_off + pad, S.len);
                }
            }
        });
        { uint64_t kc = 0; for (size_t i = 0; i < out.results.size(); ++i) kc += out.results[i].reserved; g_region_kcells += kc; }
        g_region_stats[0] += n;
        g_region_stats[1] += n_bail.load();
        for (int j = 0; j < 16; ++j) g_region_stats[2 + j] += by_reason[j].load();
        pt.lap(3, 2);
    }
}

// ------------------------------------------------------------------ broker
// Context switch between a worker's scheduler and its fibers.  glibc's swapcontext saves / restores the signal mask with a
// system call on every switch (5 % of the host time in the sampling profile); the fibers never touch the mask, so on x86-64 a
// switch only has to exchange the callee-saved registers and the stack pointer.  Other targets keep <ucontext.h>.
#if defined(__x86_64__) && !defined(RTK_FIBER_UCONTEXT)
#define RTK_FIBER_ASM 1
extern "C" void rtk_fiber_switch(void** save_sp, void* new_sp);
__asm__(
    ".text\n"
    ".globl rtk_fiber_switch\n"
    ".type rtk_fiber_switch,@function\n"
    "rtk_fiber_switch:\n"
    "    pushq %rbp\n"
    "    pushq %rbx\n"
    "    pushq %r12\n"
    "    pushq %r13\n"
    "    pushq %r14\n"
    "    pushq %r15\n"
    "    movq %rsp, (%rdi)\n"
    "    movq %rsi, %rsp\n"
    "    popq %r15\n"
    "    popq %r14\n"
    "    popq %r13\n"
    "    popq %r12\n"
    "    popq %rbx\n"
    "    popq %rbp\n"
    "    ret\n"
    ".size rtk_fiber_switch,.-rtk_fiber_switch\n");
#endif

struct GpuBroker::Fiber {
#ifdef RTK_FIBER_ASM
    void* sp = nullptr;               // saved stack pointer while the fiber is switched out
#else
    ucontext_t uc;
#endif
    char* stack = nullptr;
    size_t task = 0;
    bool done = false;
    bool express = false;
    GpuBroker* broker = nullptr;
    Worker* owner = nullptr;
    std::string error;   // set by the service that failed this fiber's request
};

struct GpuBroker::Worker {
    std::thread th;
#ifdef RTK_FIBER_ASM
    void* sched_sp = nullptr;         // the worker thread's own context (scheduler loop), saved while a fiber runs
#else
    ucontext_t sched;                 // the worker thread's own context (scheduler loop)
#endif
    Fiber* current = nullptr;
    std::mutex mu;
    std::condition_variable cv;
    std::vector<Fiber*> ready;        // served fibers, pushed by the service threads
    std::vector<Fiber*> pool;         // finished fibers (stack kept) for reuse
    char* slab = nullptr;             // one mapping holding all fiber stacks of this worker
    size_t slab_bytes = 0, stacks_used = 0;
    unsigned rr = 0;                  // round-robin over the service threads of a kind
    uint64_t ns_idle = 0, n_resumes = 0;
};

struct GpuBroker::Service {
    int kind = 0;
    rtk_ctx* ctx = nullptr;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::vector<std::pair<void*, Fiber*>> q;
    bool is_express = false;
    size_t n_riders = 0;   // queued requests that carry no fiber (extra requests of a task that parked once)
    bool stop = false;
    uint64_t batches = 0, reqs = 0, ns_busy = 0;
};

static const size_t kGuardBytes = 4096;   // one inaccessible page below every fiber stack
static size_t fiber_stack_bytes() {
    const char* e = getenv("RTK_FIBER_STACK_KB");
    const long kb = e ? atol(e) : 256;
    return (size_t)std::max(64L, std::min(kb, 8192L)) << 10;
}
// The region logic allocates and frees millions of short-lived strings / vectors per second from ~20 threads and the
// services allocate multi-megabyte staging vectors per batch: keep freed memory in the heap instead of returning it to
// the kernel after every batch (glibc trims the heap top and unmaps large blocks by default -> sbrk / mmap / page-fault
// churn was 15 % of the host time in the sampling profile).
static void tune_allocator() {
    static std::once_flag once;
    std::call_once(once, [] {
        if (getenv("RTK_NO_MALLOPT")) return;
        mallopt(M_MMAP_THRESHOLD, 32 << 20);
        mallopt(M_TRIM_THRESHOLD, 1 << 30);
        mallopt(M_TOP_PAD, 64 << 20);
    });
}

GpuBroker::GpuBroker(rtk_ctx* c) : ctx(c) {
    tune_allocator();
    // service threads per kind (each with its own forked context = stream + scratch): RTK_SERVICE_THREADS="d,p,s"
    unsigned cnt[4] = {2, 2, 2, 2};
    if (const char* e = getenv("RTK_SERVICE_THREADS")) sscanf(e, "%u,%u,%u,%u", &cnt[0], &cnt[1], &cnt[2], &cnt[3]);
#ifdef RTK_HOSTSIM   // the CPU simulator runs one launch at a time
    cnt[0] = cnt[1] = cnt[2] = cnt[3] = 1;
#endif
    bulk_regions = getenv("RTK_NO_BULK_REGIONS") == nullptr;
    bulk_all = getenv("RTK_BULK_ALL") != nullptr;
    if (bulk_regions) cnt[3] = 1;   // one bulk launch per wave: a single region service thread
    for (int k = 0; k < 4; ++k)
        for (unsigned i = 0; i < std::max(1u, std::min(cnt[k], 8u)); ++i) {
            Service* s = new Service();
            s->kind = k;
            try { s->ctx = fork_acquire(ctx, /*high=*/k != 3); } catch (...) { delete s; throw; }
            services[k].push_back(s);
        }
    for (int k = 0; k < 3; ++k) {
        Service* s = new Service();
        s->kind = k; s->is_express = true;
        try { s->ctx = fork_acquire(ctx, true); } catch (...) { delete s; throw; }
        express[k].push_back(s);
    }
}
GpuBroker::~GpuBroker() {
    for (int k = 0; k < 4; ++k)
        for (Service* s : services[k]) { fork_release(ctx, s->ctx, /*high=*/k != 3); delete s; }
    for (int k = 0; k < 3; ++k)
        for (Service* s : express[k]) { fork_release(ctx, s->ctx, true); delete s; }
}

// called on a fiber: queue the request at a service thread and hand control back to the worker's scheduler.  The
// service may answer before this fiber has switched out; that is safe because only this worker thread resumes it,
// and it only looks at its ready list from the scheduler context.
void GpuBroker::park(int kind, void* req) {
    Worker* w = tl_worker;
    if (!w || !w->current) throw std::logic_error("GpuBroker::submit called outside a broker task");
    Fiber* f = w->current;
    Service* s = (f->express && kind < 3) ? express[kind][0] : services[kind][(w->rr++) % services[kind].size()];
    {
        std::lock_guard<std::mutex> lk(s->mu);
        s->q.emplace_back(req, f);
    }
    parked_total.fetch_add(1, std::memory_order_acq_rel);
    s->cv.notify_one();
#ifdef RTK_FIBER_ASM
    rtk_fiber_switch(&f->sp, w->sched_sp);
#else
    swapcontext(&f->uc, &w->sched);
#endif
    if (!f->error.empty()) { std::string e; e.swap(f->error); throw std::runtime_error(e); }
}
void GpuBroker::set_express() {
    Worker* w = tl_worker;
    if (w && w->current) w->current->express = true;
}
void GpuBroker::submit(DistReq* r) { park(0, r); }
void GpuBroker::submit(PathReq* r) { park(1, r); }
void GpuBroker::submit(SubgraphReq* r) { park(2, r); }
void GpuBroker::submit(RegionReq* r) { park(3, r); }
// all requests go to one region service; the fiber rides on the last one and is resumed when the batch that holds them all is done
void GpuBroker::submit(const std::vector<RegionReq*>& rs) {
    if (rs.empty()) return;
    Worker* w = tl_worker;
    if (!w || !w->current) throw std::logic_error("GpuBroker::submit called outside a broker task");
    if (rs.size() > 1) {
        Service* s = services[3][0];
        std::lock_guard<std::mutex> lk(s->mu);
        for (size_t i = 0; i + 1 < rs.size(); ++i) s->q.emplace_back((void*)rs[i], (Fiber*)nullptr);
        s->q.emplace_back((void*)rs.back(), w->current);
        parked_total.fetch_add(1, std::memory_order_acq_rel);
        s->n_riders += rs.size() - 1;
    } else { park(3, rs[0]); return; }
    Service* s = services[3][0];
    s->cv.notify_one();
    Fiber* f = w->current;
#ifdef RTK_FIBER_ASM
    rtk_fiber_switch(&f->sp, w->sched_sp);
#else
    swapcontext(&f->uc, &w->sched);
#endif
    if (!f->error.empty()) { std::string e; e.swap(f->error); throw std::runtime_error(e); }
}

#ifdef RTK_HOSTSIM
static std::mutex g_sim_launch_mu;   // the CPU simulator runs one launch at a time
#endif

void GpuBroker::service_main(Service* s) {
    DeviceBind bind(s->ctx);   // a new thread starts on device 0: every CUDA call of this service belongs to its context's device
    set_thread_budget(run_budget);   // parallel_for inside a service (bulk pack / unpack) stays within the gang's share of the cores
    std::vector<std::pair<void*, Fiber*>> batch;
    const char* e_mb = getenv("RTK_SERVICE_MIN_BATCH");
    const char* e_lg = getenv("RTK_SERVICE_LINGER_US");
    const size_t min_batch = e_mb ? (size_t)atol(e_mb) : 256;
    const long linger_us = e_lg ? atol(e_lg) : 150;
    for (;;) {
        batch.clear();
        if (bulk_all && !s->is_express) {
            // global waves: a service launches when every live fiber is parked somewhere (this queue or another service's)
            std::unique_lock<std::mutex> lk(s->mu);
            for (;;) {
                if (s->stop && s->q.empty()) break;
                if (!s->q.empty()) {
                    const size_t live = live_total.load(std::memory_order_acquire);
                    bool none_to_start;
                    { std::lock_guard<std::mutex> g(mu_task); none_to_start = next_task >= n_tasks; }
                    if (parked_total.load(std::memory_order_acquire) >= live && (none_to_start || live >= cap_total)) break;
                }
                s->cv.wait_for(lk, std::chrono::microseconds(200));
            }
            if (s->q.empty()) break;
            batch.swap(s->q);
            s->n_riders = 0;
        } else if (s->kind == 3 && bulk_regions) {
            // Bulk mode of the region engine: a launch is efficient when it holds thousands of regions (one warp each; its
            // duration is set by the longest region, not by the count).  Regions of a gang reach their region request in
            // waves (first the forward regions of all pieces, then the backward regions of the uncorrected ones, then the
            // restarts after dead ends), so the service waits until EVERY live fiber is parked here - nothing else can make
            // progress - and serves the whole wave with one launch.
            std::unique_lock<std::mutex> lk(s->mu);
            for (;;) {
                if (s->stop && s->q.empty()) break;
                if (!s->q.empty()) {
                    const size_t live = live_total.load(std::memory_order_acquire);
                    bool none_to_start;
                    { std::lock_guard<std::mutex> g(mu_task); none_to_start = next_task >= n_tasks; }
                    if (s->q.size() - s->n_riders >= live && (none_to_start || live >= cap_total)) break;
                }
                s->cv.wait_for(lk, std::chrono::microseconds(500));
            }
            if (s->q.empty()) break;
            batch.swap(s->q);
            s->n_riders = 0;
        } else {
            std::unique_lock<std::mutex> lk(s->mu);
            s->cv.wait(lk, [&] { return s->stop || !s->q.empty(); });
            if (s->q.empty()) break;   // stop requested and nothing left
            // a batch costs a fixed launch + copy latency: linger briefly for more requests when only a few are queued
            if (!s->is_express && s->q.size() < min_batch && linger_us > 0) {
                const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(linger_us);
                s->cv.wait_until(lk, deadline, [&] { return s->stop || s->q.size() >= min_batch; });
            }
            batch.swap(s->q);
            s->n_riders = 0;
        }
        if (batch.empty()) break;
        const auto t0 = std::chrono::steady_clock::now();
        std::string err;
        try {
#ifdef RTK_HOSTSIM
            std::lock_guard<std::mutex> sim(g_sim_launch_mu);
#endif
            if (s->kind == 0) { std::vector<DistReq*> v; for (auto& b : batch) v.push_back((DistReq*)b.first); run_dist_batch(s->ctx, v); }
            else if (s->kind == 1) { std::vector<PathReq*> v; for (auto& b : batch) v.push_back((PathReq*)b.first); run_path_batch(s->ctx, v); }
            else if (s->kind == 2) { std::vector<SubgraphReq*> v; for (auto& b : batch) v.push_back((SubgraphReq*)b.first); run_subgraph_batch(s->ctx, v); }
            else { std::vector<RegionReq*> v; for (auto& b : batch) v.push_back((RegionReq*)b.first); run_region_batch(s->ctx, v, nullptr); }
        } catch (const std::exception& e) { err = e.what(); if (err.empty()) err = "GPU service failed"; }
        catch (...) { err = "GPU service failed (unknown exception)"; }
        s->ns_busy += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
        ++s->batches; s->reqs += batch.size();
        if (!err.empty()) {   // no new tasks after a service error
            std::lock_guard<std::mutex> g(mu_task);
            next_task = n_tasks;
        }
        if (getenv("RTK_WAVE_LOG") && (s->kind == 3 || bulk_all))
            fprintf(stderr, "[wave] broker %p kind %d: %zu requests, t = %.1f ms, service call %.1f ms\n", (void*)this, s->kind, batch.size(),
                    std::chrono::duration_cast<std::chrono::microseconds>(t0 - t_run_begin).count() / 1e3,
                    std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count() / 1e3);
        {   // requests without a fiber belong to a task that rides on another request of the same batch
            size_t w2 = 0;
            for (size_t i = 0; i < batch.size(); ++i) if (batch[i].second) batch[w2++] = batch[i];
            batch.resize(w2);
        }
        parked_total.fetch_sub(batch.size(), std::memory_order_acq_rel);
        // hand the fibers back, grouped per owner so that each worker is locked / woken once
        std::sort(batch.begin(), batch.end(), [](const std::pair<void*, Fiber*>& a, const std::pair<void*, Fiber*>& b) { return a.second->owner < b.second->owner; });
        for (size_t i = 0; i < batch.size();) {
            Worker* w = batch[i].second->owner;
            size_t j = i;
            {
                std::lock_guard<std::mutex> lk(w->mu);
                for (; j < batch.size() && batch[j].second->owner == w; ++j) {
                    if (!err.empty()) batch[j].second->error = err;
                    w->ready.push_back(batch[j].second);
                }
            }
            w->cv.notify_one();
            i = j;
        }
    }
}

#ifdef RTK_FIBER_ASM
// first activation of a fiber: reached by the `ret` of rtk_fiber_switch on a freshly prepared stack
static void fiber_trampoline() {
    GpuBroker::Worker* w = tl_worker;
    GpuBroker::Fiber* f = w->current;
    f->broker->fiber_body(f);
    rtk_fiber_switch(&f->sp, w->sched_sp);   // done: back to the scheduler for good
    __builtin_trap();
}
#endif

#ifndef RTK_FIBER_ASM
static void fiber_entry(unsigned lo, unsigned hi) {
    GpuBroker::Fiber* f = (GpuBroker::Fiber*)(((uintptr_t)hi << 32) | (uintptr_t)lo);
    f->broker->fiber_body(f);
    // returning switches to uc_link = the worker's scheduler context
}
#endif

void GpuBroker::fiber_body(Fiber* f) {
    try {
        (*task_fn)(f->task);
    } catch (const std::exception& e) {
        std::lock_guard<std::mutex> g(mu_task);
        if (task_error.empty()) task_error = e.what();
        next_task = n_tasks;   // stop handing out work
    } catch (...) {
        std::lock_guard<std::mutex> g(mu_task);
        if (task_error.empty()) task_error = "unknown exception in a correction task";
        next_task = n_tasks;
    }
    f->done = true;
}

// scheduler loop of a worker thread: resume served fibers, start new tasks while there is room, sleep when every
// live fiber is waiting for the GPU; exits when it has no live fiber and no task is left
void GpuBroker::worker_main(Worker* w) {
    DeviceBind bind(ctx);   // fibers may call into the device directly (declined regions): same device as the broker's context
    tl_broker = this;
    tl_worker = w;
    const size_t stack_bytes = fiber_stack_bytes();
    size_t live = 0;
    auto enter = [&](Fiber* f) {
        w->current = f;
#ifdef RTK_FIBER_ASM
        rtk_fiber_switch(&w->sched_sp, f->sp);
#else
        swapcontext(&w->sched, &f->uc);
#endif
        w->current = nullptr;
        if (f->done) { w->pool.push_back(f); --live; live_total.fetch_sub(1, std::memory_order_acq_rel); }
    };
    std::vector<Fiber*> resume;
    bool tasks_left = true;
    for (;;) {
        resume.clear();
        {
            std::lock_guard<std::mutex> lk(w->mu);
            resume.swap(w->ready);
        }
        for (Fiber* f : resume) enter(f);
        w->n_resumes += resume.size();
        size_t started = 0;
        while (tasks_left && live < cap_per_worker && started < 64) {   // in small chunks, so served fibers are resumed promptly
            size_t i;
            {
                std::lock_guard<std::mutex> g(mu_task);
                if (next_task >= n_tasks) { tasks_left = false; break; }
                i = next_task++;
            }
            Fiber* f;
            if (!w->pool.empty()) { f = w->pool.back(); w->pool.pop_back(); }
            else {
                f = new Fiber();
                // stacks are carved from one mapping with a PROT_NONE guard page below each: an overflow faults instead of
                // silently overwriting the neighbouring fiber's saved registers
                char* base = w->slab + (w->stacks_used++) * (stack_bytes + kGuardBytes);
                if (cap_total <= 16384) mprotect(base, kGuardBytes, PROT_NONE);   // every guard page is a mapping of its own: only below the kernel's map-count limit
                f->stack = base + kGuardBytes;
            }
            f->task = i; f->done = false; f->express = false; f->broker = this; f->owner = w; f->error.clear();
#ifdef RTK_FIBER_ASM
            {   // stack image rtk_fiber_switch pops: r15 r14 r13 r12 rbx rbp, then `ret` into the trampoline (rsp % 16 == 8 there)
                void** sp = (void**)(((uintptr_t)f->stack + stack_bytes) & ~(uintptr_t)15);
                *--sp = nullptr;                       // the trampoline's (never used) return address: ends unwinder walks
                *--sp = (void*)&fiber_trampoline;
                for (int r6 = 0; r6 < 6; ++r6) *--sp = nullptr;
                f->sp = sp;
            }
#else
            getcontext(&f->uc);
            f->uc.uc_stack.ss_sp = f->stack;
            f->uc.uc_stack.ss_size = stack_bytes;
            f->uc.uc_link = &w->sched;
            const uintptr_t p = (uintptr_t)f;
            makecontext(&f->uc, (void (*)())fiber_entry, 2, (unsigned)(p & 0xffffffffu), (unsigned)(p >> 32));
#endif
            ++live; ++started;
            live_total.fetch_add(1, std::memory_order_acq_rel);
            enter(f);
        }
        if (!resume.empty() || started) continue;
        if (live == 0) {
            if (!tasks_left) break;
            std::lock_guard<std::mutex> g(mu_task);   // tasks_left may be stale after an error
            if (next_task >= n_tasks) break;
            continue;
        }
        const auto t_idle = std::chrono::steady_clock::now();
        {
            std::unique_lock<std::mutex> lk(w->mu);
            w->cv.wait(lk, [&] { return !w->ready.empty(); });
        }
        w->ns_idle += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_idle).count();
    }
    tl_broker = nullptr;
    tl_worker = nullptr;
}

void GpuBroker::run(size_t n, unsigned inflight, const std::function<void(size_t)>& task) {
    if (n == 0) return;
    const unsigned n_workers = (unsigned)std::min<size_t>(std::max(1u, thread_budget()), n);
    const size_t stack_bytes = fiber_stack_bytes();
    n_tasks = n; next_task = 0; task_fn = &task; task_error.clear();
    cap_per_worker = std::max<size_t>(1, (std::min<size_t>(std::max(1u, inflight), n) + n_workers - 1) / n_workers);
    cap_total = cap_per_worker * n_workers;
    run_budget = thread_budget();
    live_total.store(0);
    const auto t_begin = std::chrono::steady_clock::now();
    t_run_begin = t_begin;
    parked_total.store(0);
    uint64_t prof0[4][4], rs0[18];
    for (int k = 0; k < 4; ++k) for (int j = 0; j < 4; ++j) prof0[k][j] = g_prof[k][j];
    for (int j = 0; j < 18; ++j) rs0[j] = g_region_stats[j];
    const uint64_t kcells0 = g_region_kcells;
    for (unsigned t = 0; t < n_workers; ++t) {
        Worker* w = new Worker();
        w->slab_bytes = cap_per_worker * (stack_bytes + kGuardBytes);
        w->slab = (char*)mmap(nullptr, w->slab_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (w->slab == (char*)MAP_FAILED) { delete w; for (Worker* x : workers) { munmap(x->slab, x->slab_bytes); delete x; } workers.clear(); throw std::bad_alloc(); }
        w->rr = t;
        workers.push_back(w);
    }
    for (int k = 0; k < 4; ++k)
        for (Service* s : services[k]) { s->stop = false; s->th = std::thread([this, s] { service_main(s); }); }
    for (int k = 0; k < 3; ++k)
        for (Service* s : express[k]) { s->stop = false; s->th = std::thread([this, s] { service_main(s); }); }
    for (Worker* w : workers) w->th = std::thread([this, w] { worker_main(w); });
    uint64_t idle_ns = 0, resumes = 0;
    for (Worker* w : workers) {
        w->th.join();
        idle_ns += w->ns_idle; resumes += w->n_resumes;
        for (Fiber* f : w->pool) delete f;
        munmap(w->slab, w->slab_bytes);
        delete w;
    }
    workers.clear();
    for (int k = 0; k < 4; ++k)
        for (Service* s : services[k]) {
            { std::lock_guard<std::mutex> lk(s->mu); s->stop = true; }
            s->cv.notify_one();
            s->th.join();
            waves += s->batches; jobs += s->reqs;
        }
    for (int k = 0; k < 3; ++k)
        for (Service* s : express[k]) {
            { std::lock_guard<std::mutex> lk(s->mu); s->stop = true; }
            s->cv.notify_one();
            s->th.join();
            waves += s->batches; jobs += s->reqs;
            s->batches = s->reqs = s->ns_busy = 0;
        }
    task_fn = nullptr;
    uint64_t prof[4][4];
    for (int k = 0; k < 4; ++k) for (int j = 0; j < 4; ++j) prof[k][j] = g_prof[k][j] - prof0[k][j];
    for (int k = 0; k < 4; ++k) kernel_ns[k] += prof[k][3];
    region_calls += g_region_stats[0] - rs0[0]; region_bails += g_region_stats[1] - rs0[1];
    region_kcells += g_region_kcells - kcells0;
    for (int j = 0; j < 16; ++j) region_bail_reason[j] += g_region_stats[2 + j] - rs0[2 + j];
    if (getenv("RTK_BROKER_PROFILE")) {
        const double total_ms = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_begin).count() / 1e6;
        static const char* names[4] = {"dist", "path", "subgraph", "region"};
        fprintf(stderr, "[broker] tasks=%zu workers=%u inflight/worker=%zu total %.1f ms, workers idle %.1f%% |", n, n_workers, cap_per_worker, total_ms,
                100.0 * (idle_ns / 1e6) / (total_ms * n_workers));
        for (int k = 0; k < 4; ++k)
            for (Service* s : services[k])
                fprintf(stderr, " %s: %llu reqs in %llu batches, busy %.1f ms |", names[k], (unsigned long long)s->reqs, (unsigned long long)s->batches, s->ns_busy / 1e6);
        fprintf(stderr, "\n");
        fprintf(stderr, "[broker]   region engine: %llu calls, %llu declined (cycle %llu queue %llu vlist %llu arena %llu strcap %llu hirschberg %llu dfs %llu chain %llu logic %llu)\n",
                (unsigned long long)region_calls, (unsigned long long)region_bails, (unsigned long long)region_bail_reason[1], (unsigned long long)region_bail_reason[2],
                (unsigned long long)region_bail_reason[3], (unsigned long long)region_bail_reason[4], (unsigned long long)region_bail_reason[5],
                (unsigned long long)region_bail_reason[6], (unsigned long long)region_bail_reason[7], (unsigned long long)region_bail_reason[8],
                (unsigned long long)region_bail_reason[9]);
        for (int k = 0; k < 4; ++k) {
            fprintf(stderr, "[broker]   %s service host time: assemble %.1f ms, C-ABI call %.1f ms (GPU kernels %.1f ms), scatter %.1f ms\n", names[k],
                    prof[k][0] / 1e6, prof[k][1] / 1e6, prof[k][3] / 1e6, prof[k][2] / 1e6);
        }
        fprintf(stderr, "[broker]   request mix (cumulative) dist NW 1-job %llu multi %llu | SHW 1-job %llu multi %llu | HW 1-job %llu multi %llu || path NW %llu/%llu SHW %llu/%llu\n",
                (unsigned long long)g_mix[0][0][0], (unsigned long long)g_mix[0][0][1], (unsigned long long)g_mix[0][1][0], (unsigned long long)g_mix[0][1][1],
                (unsigned long long)g_mix[0][2][0], (unsigned long long)g_mix[0][2][1], (unsigned long long)g_mix[1][0][0], (unsigned long long)g_mix[1][0][1],
                (unsigned long long)g_mix[1][1][0], (unsigned long long)g_mix[1][1][1]);
#ifndef RTK_HOSTSIM
        fprintf(stderr, "[broker]   myers_run (all callers, cumulative): plan %.1f ms, H2D issue %.1f ms, launches %.1f ms, D2H + sync %.1f ms, ends phase %.1f ms\n",
                g_myers_prof[0] / 1e6, g_myers_prof[1] / 1e6, g_myers_prof[2] / 1e6, g_myers_prof[3] / 1e6, g_myers_prof[4] / 1e6);
#endif
    }
    for (int k = 0; k < 4; ++k) for (Service* s : services[k]) { s->batches = s->reqs = s->ns_busy = 0; }
    if (!task_error.empty()) throw std::runtime_error(task_error);
}

}  // namespace rtk
