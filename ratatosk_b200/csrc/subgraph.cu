// subgraph.cu — C ABI entry rtk_explore_subgraph_batch: K2/K3 enumeration (subgraph.cuh, COUNT then WRITE),
// K4 scoring of the spelled candidates (myers_run), selection on the scored list (subgraph_host.hpp).
#include <cstring>
#include <vector>

#include "rtk_host_common.hpp"
#include "subgraph.cuh"
#include "subgraph_host.hpp"

using namespace rtk;

extern "C" int rtk_explore_subgraph_batch(rtk_ctx* c, uint32_t n_calls, const rtk_subgraph_call_t* calls, const char* ref_pool,
                                          uint64_t ref_bytes, const uint32_t* pid_pool, uint64_t n_pids, double wrlf,
                                          rtk_subgraph_out* out, uint64_t* stats) {
    return guarded([&] {
        if (!c || !calls || !ref_pool || !out) throw std::invalid_argument("null argument");
        if (!c->has_graph) throw std::invalid_argument("no graph uploaded to this context");
        RTK_CUDA(cudaSetDevice(c->device));
        memset(out, 0, sizeof(*out));
        for (uint32_t i = 0; i < n_calls; ++i) {
            const rtk_subgraph_call_t& q = calls[i];
            if (q.level + 1 > RTK_DFS_MAX_NODES) throw std::invalid_argument("level too large");
            if (q.start_unitig >= c->hdr.n_unitigs || (q.end_unitig != RTK_NONE32 && q.end_unitig >= c->hdr.n_unitigs)) throw std::invalid_argument("bad unitig id");
            if (q.ref_off + q.ref_len > ref_bytes || q.pid_off + q.pid_len > n_pids) throw std::invalid_argument("call range outside its pool");
        }
        cudaStream_t st = c->stream;
        DevBuf* S = c->d_sub;  // [0] calls, [1] pids, [2] n_cand|n_chars, [3] cand_off|char_off, [4] cands, [5] refs + spelled paths
        S[0].reserve((size_t)n_calls * sizeof(rtk_subgraph_call_t) + 16);
        S[1].reserve(n_pids * 4 + 16);
        S[2].reserve((size_t)n_calls * 12 + 32);
        S[3].reserve((size_t)(n_calls + 1) * 16 + 16);
        // inputs through pinned staging (calls | pids | read windows): the uploads are then truly asynchronous
        const uint64_t b_calls = (uint64_t)n_calls * sizeof(rtk_subgraph_call_t), b_pids = n_pids * 4;
        PinBuf& H = c->h_pin[14];
        H.reserve(b_calls + b_pids + ref_bytes + 64);
        char* h_in = H.as<char>();
        memcpy(h_in, calls, b_calls);
        if (b_pids) memcpy(h_in + b_calls, pid_pool, b_pids);
        memcpy(h_in + b_calls + b_pids, ref_pool, ref_bytes);
        if (b_calls) RTK_CUDA(counted_memcpy_async(S[0].p, h_in, b_calls, cudaMemcpyHostToDevice, st));
        if (b_pids) RTK_CUDA(counted_memcpy_async(S[1].p, h_in + b_calls, b_pids, cudaMemcpyHostToDevice, st));
        rtk_dfs_params p;
        const rtk_graph_view& g = c->dview;
        p.unitig_off = g.unitig_off; p.pool = g.pool; p.shared = g.shared; p.adj = g.adj; p.gset_of = g.gset_of;
        p.gset_off = g.gset_off; p.gset_ids = g.gset_ids; p.loc_off = g.loc_off; p.loc_ids = g.loc_ids; p.k = c->hdr.k;
        p.calls = S[0].as<rtk_subgraph_call>(); p.pid_pool = S[1].as<uint32_t>(); p.n_calls = n_calls;
        uint64_t* d_nchars = S[2].as<uint64_t>();
        uint32_t* d_ncand = (uint32_t*)(d_nchars + n_calls);
        p.n_cand = d_ncand; p.n_chars = d_nchars; p.cand_off = nullptr; p.char_off = nullptr; p.cands = nullptr; p.chars = nullptr;
        p.overflow = d_ncand + n_calls;   // one flag after the counts, downloaded with them
        RTK_CUDA(cudaMemsetAsync(p.overflow, 0, 4, st));
        const uint32_t grid = (n_calls + RTK_DFS_WARPS - 1) / RTK_DFS_WARPS;
        RTK_CUDA(cudaEventRecord(c->ev0, st));
        if (n_calls) ++g_launches;
        if (n_calls) rtk_dfs_kernel<false><<<grid, RTK_DFS_WARPS * 32, 0, st>>>(p);
        RTK_CUDA(cudaGetLastError());
        std::vector<uint64_t> nchars(n_calls);
        std::vector<uint32_t> ncand(n_calls + 1, 0);
        PinnedD2H d2h(c, st);
        if (n_calls) {
            d2h.copy(2, nchars.data(), d_nchars, (size_t)n_calls * 8);
            d2h.copy(3, ncand.data(), d_ncand, (size_t)(n_calls + 1) * 4);
        }
        d2h.sync();
        if (ncand[n_calls]) throw std::runtime_error("exploreSubGraph: a burst exceeded the DFS capacity (RTK_DFS_MAX_NODES / RTK_DFS_STACK)");
        std::vector<uint64_t> cand_off(n_calls + 1, 0), char_off(n_calls + 1, 0);
        for (uint32_t i = 0; i < n_calls; ++i) { cand_off[i + 1] = cand_off[i] + ncand[i]; char_off[i + 1] = char_off[i] + nchars[i]; }
        const uint64_t n_cands = cand_off[n_calls], n_chars = char_off[n_calls];
        if (n_cands >= 0xFFFFFFFFull) throw std::runtime_error("too many candidate paths in one batch");
        S[4].reserve(n_cands * sizeof(rtk_cand) + 16);
        S[5].reserve(ref_bytes + n_chars + 16);
        uint64_t* d_off = S[3].as<uint64_t>();
        PinBuf& H2 = c->h_pin[15];
        H2.reserve((size_t)(n_calls + 1) * 16 + 64);
        memcpy(H2.p, cand_off.data(), (size_t)(n_calls + 1) * 8);
        memcpy(H2.as<char>() + (size_t)(n_calls + 1) * 8, char_off.data(), (size_t)(n_calls + 1) * 8);
        RTK_CUDA(counted_memcpy_async(d_off, H2.p, (size_t)(n_calls + 1) * 16, cudaMemcpyHostToDevice, st));
        if (ref_bytes) RTK_CUDA(counted_memcpy_async(S[5].p, h_in + b_calls + b_pids, ref_bytes, cudaMemcpyHostToDevice, st));
        p.cand_off = d_off; p.char_off = d_off + (n_calls + 1); p.cands = S[4].as<rtk_cand>(); p.chars = S[5].as<char>() + ref_bytes;
        if (n_calls) ++g_launches;
        if (n_calls) rtk_dfs_kernel<true><<<grid, RTK_DFS_WARPS * 32, 0, st>>>(p);
        RTK_CUDA(cudaGetLastError());
        RTK_CUDA(cudaEventRecord(c->ev1, st));
        std::vector<rtk_cand> cands(n_cands);
        if (n_cands) d2h.copy(4, cands.data(), S[4].p, n_cands * sizeof(rtk_cand));
        d2h.sync();
        float dfs_ms = 0.f;
        RTK_CUDA(cudaEventElapsedTime(&dfs_ms, c->ev0, c->ev1));
        // K4 on every candidate
        std::vector<CandAlign> plan(n_cands);
        std::vector<uint64_t> qb(n_cands + 1), tb(n_cands + 1);
        std::vector<uint32_t> ql(n_cands + 1), tl(n_cands + 1);
        std::vector<uint8_t> md(n_cands + 1);
        for (uint64_t i = 0; i < n_cands; ++i) {
            plan[i] = plan_candidate(cands[i], calls[cands[i].call], ref_bytes, wrlf);
            qb[i] = plan[i].q_beg; tb[i] = plan[i].t_beg; ql[i] = plan[i].q_len; tl[i] = plan[i].t_len; md[i] = plan[i].mode;
        }
        std::vector<int32_t> ed(n_cands + 1);
        float my_ms = 0.f;
        if (n_cands) {
            MyersJobs j{(uint32_t)n_cands, qb.data(), ql.data(), tb.data(), tl.data(), md.data(), nullptr};
            myers_run_lean(c, S[5].as<char>(), S[5].as<char>(), nullptr, 0, nullptr, 0, j, ed.data(), nullptr, nullptr, &my_ms);
        }
        fill_subgraph_out(n_calls, cands, cand_off, ed, plan, c->host_graph->view.unitig_off, c->hdr.k, out);
        if (stats) { stats[0] += n_cands; stats[1] += n_chars; stats[2] += (uint64_t)(dfs_ms * 1e6); stats[3] += (uint64_t)(my_ms * 1e6); }
    });
}
