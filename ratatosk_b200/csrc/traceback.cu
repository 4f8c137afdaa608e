// traceback.cu — C ABI entry rtk_edlib_path_batch: edit distance + alignment path for NW and SHW alignments
// (edlibAlign with EDLIB_TASK_PATH).  SHW: K4 finds the distance and its first end column, the path is the NW
// path against that target prefix (src/edlib.cpp:262-279).
#include <cstdlib>
#include <cstring>
#include <vector>

#include "rtk_host_common.hpp"
#include "traceback.cuh"
#include "traceback_host.hpp"

namespace rtk {

template <int G> static void launch_fill(rtk_ctx* c, rtk_fill_params p, const uint32_t* d_order, uint32_t n) {
    if (!n) return;
    p.order = d_order;
    p.n = n;
    const uint64_t threads = (uint64_t)n * G;
    rtk_myers_fill_kernel<G><<<(uint32_t)((threads + RTK_MYERS_THREADS - 1) / RTK_MYERS_THREADS), RTK_MYERS_THREADS, 0, c->stream>>>(p);
    RTK_CUDA(cudaGetLastError());
}

}  // namespace rtk

using namespace rtk;

extern "C" int rtk_edlib_path_batch(rtk_ctx* c, uint32_t n, const char* q_pool, const uint64_t* q_off, const char* t_pool,
                                    const uint64_t* t_off, const uint8_t* mode, int32_t* dist, int32_t* end_loc,
                                    uint8_t** ops, uint64_t** ops_off, uint8_t* flags, uint64_t* stats) {
    return guarded([&] {
        if (!c || !q_pool || !q_off || !t_pool || !t_off || !mode || !dist || !end_loc || !ops || !ops_off || !flags)
            throw std::invalid_argument("null argument");
        RTK_CUDA(cudaSetDevice(c->device));
        for (uint32_t a = 0; a < n; ++a) if (mode[a] > 1) throw std::invalid_argument("path mode must be 0 (NW) or 1 (SHW)");
        cudaStream_t st = c->stream;
        const uint64_t qb = q_off[n] - q_off[0], tb = t_off[n] - t_off[0];
        std::vector<uint64_t> qrel(n + 1), trel(n + 1);
        std::vector<uint32_t> qlen(n + 1, 0), tlen(n + 1, 0);
        for (uint32_t i = 0; i <= n; ++i) { qrel[i] = q_off[i] - q_off[0]; trel[i] = t_off[i] - t_off[0]; }
        for (uint32_t i = 0; i < n; ++i) { qlen[i] = (uint32_t)(q_off[i + 1] - q_off[i]); tlen[i] = (uint32_t)(t_off[i + 1] - t_off[i]); }
        c->d_aux[0].reserve(qb + 16);
        c->d_aux[1].reserve(tb + 16);
        RTK_CUDA(cudaMemcpyAsync(c->d_aux[0].p, q_pool + q_off[0], qb, cudaMemcpyHostToDevice, st));
        RTK_CUDA(cudaMemcpyAsync(c->d_aux[1].p, t_pool + t_off[0], tb, cudaMemcpyHostToDevice, st));
        // 1. SHW alignments: distance + first end column from K4
        std::vector<uint32_t> shw;
        for (uint32_t a = 0; a < n; ++a) if (mode[a] == 1) shw.push_back(a);
        std::vector<uint32_t> teff(tlen);
        for (uint32_t a = 0; a < n; ++a) end_loc[a] = (int32_t)tlen[a] - 1;
        if (!shw.empty()) {
            const uint32_t m = (uint32_t)shw.size();
            std::vector<uint64_t> qb2(m), tb2(m);
            std::vector<uint32_t> ql2(m), tl2(m);
            std::vector<uint8_t> md(m, 1);
            std::vector<int32_t> d2(m);
            for (uint32_t i = 0; i < m; ++i) { qb2[i] = qrel[shw[i]]; tb2[i] = trel[shw[i]]; ql2[i] = qlen[shw[i]]; tl2[i] = tlen[shw[i]]; }
            MyersJobs j{m, qb2.data(), ql2.data(), tb2.data(), tl2.data(), md.data(), nullptr};
            int32_t* e = nullptr; uint64_t* eo = nullptr;
            myers_run(c, c->d_aux[0].as<char>(), c->d_aux[1].as<char>(), j, d2.data(), true, &e, &eo, nullptr);
            for (uint32_t i = 0; i < m; ++i) {
                dist[shw[i]] = d2[i];
                end_loc[shw[i]] = (eo[i + 1] > eo[i]) ? e[eo[i]] : -1;
                teff[shw[i]] = (uint32_t)(end_loc[shw[i]] + 1);
            }
            free(e); free(eo);
        }
        // 2. classify
        for (uint32_t a = 0; a < n; ++a) {
            if (qlen[a] == 0 || tlen[a] == 0 || teff[a] == 0) flags[a] = 2;
            else if (tb_needs_hirschberg(qlen[a], teff[a]) || qlen[a] > 64 * 32) flags[a] = 1;
            else flags[a] = 0;
        }
        const TbPlan pl = plan_traceback(n, qlen.data(), teff.data(), flags);
        // 3. fill + walk (d_sub[0] offsets, [1] lens, [2] order, [3] matrix, [4] anchors, [5] ops, [6] ops_len|dist)
        DevBuf* S = c->d_sub;
        S[0].reserve((size_t)(n + 1) * 8 * 4);
        S[1].reserve((size_t)(n + 1) * 4 * 2);
        S[2].reserve((size_t)(pl.ids.size() + 1) * 4);
        S[3].reserve(pl.cells * 16 + 16);
        S[4].reserve(pl.cells * 4 + 16);
        S[5].reserve(pl.ops_off[n] + 16);
        S[6].reserve((size_t)(n + 1) * 8);
        uint64_t* d_off = S[0].as<uint64_t>();
        uint32_t* d_len = S[1].as<uint32_t>();
        RTK_CUDA(cudaMemcpyAsync(d_off, qrel.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
        RTK_CUDA(cudaMemcpyAsync(d_off + (n + 1), trel.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
        RTK_CUDA(cudaMemcpyAsync(d_off + 2 * (n + 1), pl.mat_off.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
        RTK_CUDA(cudaMemcpyAsync(d_off + 3 * (n + 1), pl.ops_off.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
        RTK_CUDA(cudaMemcpyAsync(d_len, qlen.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
        RTK_CUDA(cudaMemcpyAsync(d_len + (n + 1), teff.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
        if (!pl.ids.empty()) RTK_CUDA(cudaMemcpyAsync(S[2].p, pl.ids.data(), pl.ids.size() * 4, cudaMemcpyHostToDevice, st));
        uint32_t* d_opslen = S[6].as<uint32_t>();
        int32_t* d_dist = (int32_t*)(d_opslen + (n + 1));
        RTK_CUDA(cudaMemsetAsync(d_opslen, 0, (size_t)(n + 1) * 8, st));
        rtk_fill_params fp;
        fp.q_pool = c->d_aux[0].as<char>(); fp.q_beg = d_off; fp.q_len = d_len; fp.t_pool = c->d_aux[1].as<char>();
        fp.t_beg = d_off + (n + 1); fp.t_len = d_len + (n + 1); fp.order = nullptr; fp.n = 0; fp.mat_off = d_off + 2 * (n + 1);
        fp.mat = S[3].as<ulonglong2>(); fp.anchor = S[4].as<int32_t>(); fp.dist = d_dist;
        RTK_CUDA(cudaEventRecord(c->ev0, st));
        const uint32_t* d_ids = S[2].as<uint32_t>();
        uint32_t o = 0;
        launch_fill<1>(c, fp, d_ids + o, (uint32_t)pl.order[0].size()); o += (uint32_t)pl.order[0].size();
        launch_fill<2>(c, fp, d_ids + o, (uint32_t)pl.order[1].size()); o += (uint32_t)pl.order[1].size();
        launch_fill<4>(c, fp, d_ids + o, (uint32_t)pl.order[2].size()); o += (uint32_t)pl.order[2].size();
        launch_fill<8>(c, fp, d_ids + o, (uint32_t)pl.order[3].size()); o += (uint32_t)pl.order[3].size();
        launch_fill<16>(c, fp, d_ids + o, (uint32_t)pl.order[4].size()); o += (uint32_t)pl.order[4].size();
        launch_fill<32>(c, fp, d_ids + o, (uint32_t)pl.order[5].size());
        rtk_tb_params tp;
        tp.q_len = d_len; tp.t_len = d_len + (n + 1); tp.ids = d_ids; tp.n = (uint32_t)pl.ids.size(); tp.mat_off = d_off + 2 * (n + 1);
        tp.mat = S[3].as<ulonglong2>(); tp.anchor = S[4].as<int32_t>(); tp.dist = d_dist; tp.ops_off = d_off + 3 * (n + 1);
        tp.ops = S[5].as<uint8_t>(); tp.ops_len = d_opslen;
        if (tp.n) rtk_traceback_kernel<<<(tp.n + 127) / 128, 128, 0, st>>>(tp);
        RTK_CUDA(cudaGetLastError());
        RTK_CUDA(cudaEventRecord(c->ev1, st));
        std::vector<uint32_t> h_len(n + 1);
        std::vector<int32_t> h_dist(n + 1);
        std::vector<uint8_t> h_ops(pl.ops_off[n] + 1);
        RTK_CUDA(cudaMemcpyAsync(h_len.data(), d_opslen, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        RTK_CUDA(cudaMemcpyAsync(h_dist.data(), d_dist, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        if (pl.ops_off[n]) RTK_CUDA(cudaMemcpyAsync(h_ops.data(), S[5].p, pl.ops_off[n], cudaMemcpyDeviceToHost, st));
        RTK_CUDA(cudaStreamSynchronize(st));
        float kms = 0.f;
        RTK_CUDA(cudaEventElapsedTime(&kms, c->ev0, c->ev1));
        // 4. dense output
        uint64_t* off = (uint64_t*)malloc((size_t)(n + 1) * 8);
        if (!off) throw std::bad_alloc();
        off[0] = 0;
        for (uint32_t a = 0; a < n; ++a) {
            uint64_t len = 0;
            if (flags[a] == 0) len = h_len[a];
            else if (flags[a] == 2 && qlen[a] != 0 && tlen[a] != 0) len = qlen[a];  // SHW that ends before the target: all query bases unaligned
            off[a + 1] = off[a] + len;
        }
        uint8_t* out = (uint8_t*)malloc(off[n] + 1);
        if (!out) { free(off); throw std::bad_alloc(); }
        for (uint32_t a = 0; a < n; ++a) {
            if (flags[a] == 0) {
                const uint64_t cap = (uint64_t)qlen[a] + teff[a];
                memcpy(out + off[a], h_ops.data() + pl.ops_off[a] + (cap - h_len[a]), h_len[a]);
                if (mode[a] == 0) dist[a] = h_dist[a];
            } else if (flags[a] == 2) {
                memset(out + off[a], 1, off[a + 1] - off[a]);
                if (mode[a] == 0) dist[a] = (int32_t)std::max(qlen[a], tlen[a]);
                else if (qlen[a] == 0 || tlen[a] == 0) { dist[a] = (int32_t)qlen[a]; end_loc[a] = -1; }
                flags[a] = 0;
            } else if (mode[a] == 0) dist[a] = -1;
        }
        *ops = out;
        *ops_off = off;
        if (stats) stats[2] += (uint64_t)(kms * 1e6);
    });
}
