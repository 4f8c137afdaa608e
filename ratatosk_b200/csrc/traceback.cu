// traceback.cu — C ABI entry rtk_edlib_path_batch: edit distance + alignment path for NW and SHW alignments
// (edlibAlign with EDLIB_TASK_PATH).  SHW: K4 finds the distance and its first end column, the path is the NW
// path against that target prefix (src/edlib.cpp:262-279).  Large problems go through edlib's divide-and-
// conquer (traceback_host.hpp), whose per-level batches run on the device here.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "rtk_host_common.hpp"
#include "traceback.cuh"
#include "traceback_host.hpp"

namespace rtk {

// device pools: d_aux[0] queries, d_aux[1] targets (effective prefixes), d_sub[7] reversed queries | reversed targets
struct CudaTbBackend : TbBackend {
    rtk_ctx* c;
    const char* d_q; const char* d_t; const char* d_rq; const char* d_rt;
    const uint32_t* d_ids = nullptr;   // alignment ids of the current fill, device copy (class order G = 1 .. 32)
    float ms = 0.f;

    template <bool LC> void run_fill(const std::vector<TbItem>& items, const TbPlan& pl, rtk_fill_params& fp, uint64_t*& d_off, uint32_t*& d_len) {
        const uint32_t n = (uint32_t)items.size();
        cudaStream_t st = c->stream;
        DevBuf* S = c->d_sub;  // [0] u64 arrays, [1] u32 arrays, [2] order, [3] matrix, [4] anchors, [5] ops, [6] ops_len|dist + hbound
        std::vector<uint64_t> qb(n), tb(n);
        std::vector<uint32_t> ql(n), tl(n);
        // forward and reversed items may mix in one batch: rebase reversed ones onto one combined pool view
        // (forward pools and reversed pools live in different buffers, so offsets are made absolute device addresses
        // relative to d_q / d_t by using pointer differences)
        for (uint32_t i = 0; i < n; ++i) {
            const char* qp = items[i].rev ? d_rq : d_q;
            const char* tp = items[i].rev ? d_rt : d_t;
            qb[i] = (uint64_t)((qp + items[i].q_beg) - d_q);
            tb[i] = (uint64_t)((tp + items[i].t_beg) - d_t);
            ql[i] = items[i].q_len; tl[i] = items[i].t_len;
        }
        // one packed upload from pinned staging: u64 q_beg | t_beg | mat_off(n+1) | ops_off(n+1) | hb_off(n+1), u32 q_len | t_len | ids
        const uint64_t o_qb = 0, o_tb = o_qb + 8ull * n, o_mat = o_tb + 8ull * n, o_ops = o_mat + 8ull * (n + 1), o_hb = o_ops + 8ull * (n + 1),
                       o_ql = o_hb + 8ull * (n + 1), o_tl = o_ql + 4ull * (n + 1), o_ids = o_tl + 4ull * (n + 1), total = o_ids + 4ull * (n + 1);
        PinBuf& H = c->h_pin[11];
        H.reserve(total + 64);
        char* h = H.as<char>();
        memcpy(h + o_qb, qb.data(), 8ull * n);
        memcpy(h + o_tb, tb.data(), 8ull * n);
        memcpy(h + o_mat, pl.mat_off.data(), 8ull * (n + 1));
        memcpy(h + o_ops, pl.ops_off.data(), 8ull * (n + 1));
        memcpy(h + o_hb, pl.hb_off.data(), 8ull * (n + 1));
        memcpy(h + o_ql, ql.data(), 4ull * n);
        memcpy(h + o_tl, tl.data(), 4ull * n);
        memcpy(h + o_ids, pl.ids.data(), pl.ids.size() * 4);
        S[0].reserve(total + 64);
        S[3].reserve(pl.cells * 16 + 16);
        S[4].reserve(pl.cells * 4 + 16);
        S[5].reserve(pl.ops_off[n] + 16);
        S[6].reserve((size_t)(n + 1) * 8 + pl.hb_off[n] + 16);
        char* d = S[0].as<char>();
        RTK_CUDA(counted_memcpy_async(d, h, total, cudaMemcpyHostToDevice, st));
        d_off = (uint64_t*)d;                       // [q_beg | t_beg | mat_off | ops_off | hb_off]
        d_len = (uint32_t*)(d + o_ql);              // [q_len (n+1) | t_len (n+1)]
        d_ids = (const uint32_t*)(d + o_ids);
        uint32_t* d_opslen = S[6].as<uint32_t>();
        RTK_CUDA(cudaMemsetAsync(d_opslen, 0, (size_t)(n + 1) * 8, st));
        fp.q_pool = d_q; fp.q_beg = d_off; fp.q_len = d_len; fp.t_pool = d_t; fp.t_beg = d_off + n; fp.t_len = d_len + (n + 1);
        fp.order = d_ids; fp.n = n; fp.mat_off = (const uint64_t*)(d + o_mat); fp.mat = S[3].as<ulonglong2>(); fp.anchor = S[4].as<int32_t>();
        fp.dist = (int32_t*)(d_opslen + (n + 1)); fp.hbound = (int8_t*)(d_opslen + 2 * (n + 1)); fp.hb_off = (const uint64_t*)(d + o_hb);
        // pl.ids holds the classes in the order G = 1, 2, ..., 32; the fused kernel runs them 32 first
        uint32_t off_c[7] = {0}, n_blocks = 0;
        for (int k = 0; k < 6; ++k) off_c[k + 1] = off_c[k] + (uint32_t)pl.order[k].size();
        for (int jx = 0; jx < 6; ++jx) {
            const int k = 5 - jx;
            const uint32_t G = 1u << k;
            fp.cls_ord[jx] = off_c[k]; fp.cls_cnt[jx] = (uint32_t)pl.order[k].size(); fp.cls_blk[jx] = n_blocks;
            n_blocks += (uint32_t)(((uint64_t)pl.order[k].size() * G + RTK_MYERS_THREADS - 1) / RTK_MYERS_THREADS);
        }
        fp.cls_blk[6] = n_blocks;
        RTK_CUDA(cudaEventRecord(c->ev0, st));
        if (n_blocks) {
            ++g_launches;
            rtk_myers_fill_fused_kernel<LC><<<n_blocks, RTK_MYERS_THREADS, 0, st>>>(fp);
            RTK_CUDA(cudaGetLastError());
        }
    }

    void direct(const std::vector<TbItem>& items, std::vector<std::vector<uint8_t>>& ops, std::vector<int32_t>& dist) override {
        const uint32_t n = (uint32_t)items.size();
        ops.assign(n, {});
        dist.assign(n, -1);
        if (!n) return;
        const TbPlan pl = plan_items(items, false);
        rtk_fill_params fp;
        uint64_t* d_off; uint32_t* d_len;
        run_fill<false>(items, pl, fp, d_off, d_len);
        cudaStream_t st = c->stream;
        DevBuf* S = c->d_sub;
        rtk_tb_params tp;
        tp.q_len = d_len; tp.t_len = d_len + (n + 1); tp.ids = d_ids; tp.n = n; tp.mat_off = d_off + 2 * (size_t)n;
        tp.mat = S[3].as<ulonglong2>(); tp.anchor = S[4].as<int32_t>(); tp.dist = fp.dist; tp.ops_off = d_off + 2 * (size_t)n + (n + 1);
        tp.ops = S[5].as<uint8_t>(); tp.ops_len = S[6].as<uint32_t>();
        ++g_launches;
        rtk_traceback_kernel<<<(n + 127) / 128, 128, 0, st>>>(tp);
        RTK_CUDA(cudaGetLastError());
        RTK_CUDA(cudaEventRecord(c->ev1, st));
        std::vector<uint32_t> h_len(n + 1);
        std::vector<uint8_t> h_ops(pl.ops_off[n] + 1);
        PinnedD2H d2h(c, st);
        d2h.copy(5, h_len.data(), tp.ops_len, (size_t)n * 4);
        d2h.copy(6, dist.data(), fp.dist, (size_t)n * 4);
        if (pl.ops_off[n]) d2h.copy(7, h_ops.data(), S[5].p, pl.ops_off[n]);
        d2h.sync();
        float t = 0.f;
        RTK_CUDA(cudaEventElapsedTime(&t, c->ev0, c->ev1));
        ms += t;
        for (uint32_t a = 0; a < n; ++a) {
            const uint64_t cap = (uint64_t)items[a].q_len + items[a].t_len;
            const uint8_t* src = h_ops.data() + pl.ops_off[a] + (cap - h_len[a]);
            ops[a].assign(src, src + h_len[a]);
        }
    }

    void last_column(const std::vector<TbItem>& items, std::vector<std::vector<int32_t>>& rows) override {
        const uint32_t n = (uint32_t)items.size();
        rows.assign(n, {});
        if (!n) return;
        const TbPlan pl = plan_items(items, true);
        rtk_fill_params fp;
        uint64_t* d_off; uint32_t* d_len;
        run_fill<true>(items, pl, fp, d_off, d_len);
        cudaStream_t st = c->stream;
        RTK_CUDA(cudaEventRecord(c->ev1, st));
        std::vector<uint64_t> cells(pl.cells * 2 + 2);
        std::vector<int32_t> anchor(pl.cells + 1);
        PinnedD2H d2h(c, st);
        d2h.copy(8, cells.data(), c->d_sub[3].p, pl.cells * 16);
        d2h.copy(9, anchor.data(), c->d_sub[4].p, pl.cells * 4);
        d2h.sync();
        float t = 0.f;
        RTK_CUDA(cudaEventElapsedTime(&t, c->ev0, c->ev1));
        ms += t;
        parallel_for(n, [&](size_t ab, size_t ae) {
            std::vector<uint64_t> P, M;
            for (size_t a = ab; a < ae; ++a) {
                const uint32_t nb = (items[a].q_len + 63) / 64;
                P.resize(nb); M.resize(nb);
                for (uint32_t b = 0; b < nb; ++b) { P[b] = cells[2 * (pl.mat_off[a] + b)]; M[b] = cells[2 * (pl.mat_off[a] + b) + 1]; }
                tb_rows_from_column(items[a].q_len, P.data(), M.data(), anchor.data() + pl.mat_off[a], rows[a]);
            }
        });
    }
};

// NW paths of n problems as runs, restricted to the target columns the caller reads (TbNeed): phasing()'s whole-read alignments.
// Pools are used in place (problem a = q_pool[q_off[a], q_off[a+1]) vs t_pool[t_off[a], t_off[a+1])); forward and reversed copies
// go up once, every level of the divide-and-conquer is one fused launch.
void nw_path_runs_masked(rtk_ctx* c, uint32_t n, const char* q_pool, const uint64_t* q_off, const char* t_pool, const uint64_t* t_off,
                         const TbNeed& need, std::vector<std::vector<TbRun>>& runs, float* kernel_ms) {
    runs.assign(n, {});
    if (kernel_ms) *kernel_ms = 0.f;
    if (!n) return;
    DeviceBind bind(c);
    cudaStream_t st = c->stream;
    const uint64_t qb = q_off[n] - q_off[0], tb = t_off[n] - t_off[0];
    std::vector<uint64_t> qrel(n + 1), trel(n + 1);
    std::vector<uint32_t> qlen(n + 1, 0), tlen(n + 1, 0);
    for (uint32_t i = 0; i <= n; ++i) { qrel[i] = q_off[i] - q_off[0]; trel[i] = t_off[i] - t_off[0]; }
    for (uint32_t i = 0; i < n; ++i) {
        if (q_off[i + 1] - q_off[i] >= 0xFFFFFFFFull || t_off[i + 1] - t_off[i] >= 0xFFFFFFFFull) throw std::invalid_argument("alignment side longer than 4 Gbases");
        qlen[i] = (uint32_t)(q_off[i + 1] - q_off[i]); tlen[i] = (uint32_t)(t_off[i + 1] - t_off[i]);
    }
    c->d_aux[0].reserve(qb + 16);
    c->d_aux[1].reserve(tb + 16);
    c->d_sub[7].reserve(qb + tb + 32);
    PinBuf& H = c->h_pin[12];
    PinBuf& HR = c->h_pin[13];
    H.reserve(qb + tb + 64);
    HR.reserve(qb + tb + 64);
    char* hq = H.as<char>(); char* ht = hq + qb;
    char* rq = HR.as<char>(); char* rt = rq + qb;
    parallel_for(n, [&](size_t ab, size_t ae) {
        for (size_t a = ab; a < ae; ++a) {
            const char* qs = q_pool + q_off[a];
            const char* ts = t_pool + t_off[a];
            memcpy(hq + qrel[a], qs, qlen[a]);
            memcpy(ht + trel[a], ts, tlen[a]);
            if (need.any((uint32_t)a, 0, tlen[a])) {   // reversed copies feed the divide-and-conquer of solved problems only
                char* dq = rq + qrel[a]; char* dt = rt + trel[a];
                for (uint32_t i = 0; i < qlen[a]; ++i) dq[i] = qs[qlen[a] - 1 - i];
                for (uint32_t i = 0; i < tlen[a]; ++i) dt[i] = ts[tlen[a] - 1 - i];
            }
        }
    });
    if (qb) RTK_CUDA(counted_memcpy_async(c->d_aux[0].p, hq, qb, cudaMemcpyHostToDevice, st));
    if (tb) RTK_CUDA(counted_memcpy_async(c->d_aux[1].p, ht, tb, cudaMemcpyHostToDevice, st));
    if (qb) RTK_CUDA(counted_memcpy_async(c->d_sub[7].p, rq, qb, cudaMemcpyHostToDevice, st));
    if (tb) RTK_CUDA(counted_memcpy_async(c->d_sub[7].as<char>() + qb + 8, rt, tb, cudaMemcpyHostToDevice, st));
    CudaTbBackend be;
    be.c = c; be.d_q = c->d_aux[0].as<char>(); be.d_t = c->d_aux[1].as<char>();
    be.d_rq = c->d_sub[7].as<char>(); be.d_rt = c->d_sub[7].as<char>() + qb + 8;
    solve_nw_runs(be, n, qrel.data(), qlen.data(), trel.data(), tlen.data(), &need, runs);
    if (kernel_ms) *kernel_ms = be.ms;
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
        {   // pools through pinned staging: the copies are then truly asynchronous
            PinBuf& H = c->h_pin[12];
            H.reserve(qb + tb + 64);
            memcpy(H.as<char>(), q_pool + q_off[0], qb);
            memcpy(H.as<char>() + qb, t_pool + t_off[0], tb);
            if (qb) RTK_CUDA(counted_memcpy_async(c->d_aux[0].p, H.as<char>(), qb, cudaMemcpyHostToDevice, st));
            if (tb) RTK_CUDA(counted_memcpy_async(c->d_aux[1].p, H.as<char>() + qb, tb, cudaMemcpyHostToDevice, st));
        }
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
            std::vector<int32_t> fe(m, -1);
            myers_run_lean(c, c->d_aux[0].as<char>(), c->d_aux[1].as<char>(), nullptr, 0, nullptr, 0, j, d2.data(), fe.data(), nullptr, nullptr);
            for (uint32_t i = 0; i < m; ++i) {
                dist[shw[i]] = d2[i];
                end_loc[shw[i]] = fe[i];
                teff[shw[i]] = (uint32_t)(end_loc[shw[i]] + 1);
            }
        }
        // 2. reversed copies (each query, each effective target prefix, reversed in place of its own range): only edlib's
        //    divide-and-conquer reads them, i.e. only when some problem is above the 1 MiB switch
        bool need_rev = false;
        for (uint32_t a = 0; a < n && !need_rev; ++a) need_rev = qlen[a] != 0 && teff[a] != 0 && tb_needs_hirschberg(qlen[a], teff[a]);
        c->d_sub[7].reserve(qb + tb + 32);
        if (need_rev) {
            PinBuf& H = c->h_pin[13];
            H.reserve(qb + tb + 64);
            char* rq = H.as<char>();
            char* rt = rq + qb;
            memset(rq, 'N', qb + tb);
            for (uint32_t a = 0; a < n; ++a) {
                const char* qs = q_pool + q_off[a];
                for (uint32_t i = 0; i < qlen[a]; ++i) rq[qrel[a] + i] = qs[qlen[a] - 1 - i];
                const char* ts = t_pool + t_off[a];
                for (uint32_t i = 0; i < teff[a]; ++i) rt[trel[a] + i] = ts[teff[a] - 1 - i];
            }
            if (qb) RTK_CUDA(counted_memcpy_async(c->d_sub[7].p, rq, qb, cudaMemcpyHostToDevice, st));
            if (tb) RTK_CUDA(counted_memcpy_async(c->d_sub[7].as<char>() + qb + 8, rt, tb, cudaMemcpyHostToDevice, st));
        }
        // 3. solve the non-trivial problems
        std::vector<uint32_t> ids;
        for (uint32_t a = 0; a < n; ++a) { flags[a] = 0; if (qlen[a] != 0 && tlen[a] != 0 && teff[a] != 0) ids.push_back(a); }
        std::vector<uint64_t> sq(ids.size()), stt(ids.size());
        std::vector<uint32_t> sql(ids.size()), stl(ids.size());
        for (size_t i = 0; i < ids.size(); ++i) { sq[i] = qrel[ids[i]]; stt[i] = trel[ids[i]]; sql[i] = qlen[ids[i]]; stl[i] = teff[ids[i]]; }
        CudaTbBackend be;
        be.c = c; be.d_q = c->d_aux[0].as<char>(); be.d_t = c->d_aux[1].as<char>();
        be.d_rq = c->d_sub[7].as<char>(); be.d_rt = c->d_sub[7].as<char>() + qb + 8;
        std::vector<std::vector<uint8_t>> sops;
        std::vector<int32_t> sdist;
        solve_nw_paths(be, (uint32_t)ids.size(), sq.data(), sql.data(), stt.data(), stl.data(), sops, sdist);
        // 4. dense output
        std::vector<std::vector<uint8_t>> all(n);
        for (size_t i = 0; i < ids.size(); ++i) { all[ids[i]] = std::move(sops[i]); if (mode[ids[i]] == 0) dist[ids[i]] = sdist[i]; }
        for (uint32_t a = 0; a < n; ++a) {
            if (qlen[a] == 0 || tlen[a] == 0) {  // edlibAlign returns before any alignment is built (:160-176)
                if (mode[a] == 0) dist[a] = (int32_t)std::max(qlen[a], tlen[a]);
                else { dist[a] = (int32_t)qlen[a]; end_loc[a] = -1; }
            } else if (teff[a] == 0) all[a].assign(qlen[a], 1);  // SHW ending before the target starts: every query base unaligned
        }
        uint64_t* off = (uint64_t*)malloc((size_t)(n + 1) * 8);
        if (!off) throw std::bad_alloc();
        off[0] = 0;
        for (uint32_t a = 0; a < n; ++a) off[a + 1] = off[a] + all[a].size();
        uint8_t* out = (uint8_t*)malloc(off[n] + 1);
        if (!out) { free(off); throw std::bad_alloc(); }
        for (uint32_t a = 0; a < n; ++a) if (!all[a].empty()) memcpy(out + off[a], all[a].data(), all[a].size());
        *ops = out;
        *ops_off = off;
        if (stats) stats[2] += (uint64_t)(be.ms * 1e6);
    });
}
