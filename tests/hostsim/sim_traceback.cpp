// sim_traceback.cpp — TEST INFRASTRUCTURE: rtk_edlib_path_batch on the CPU simulator (kernel sources traceback.cuh,
// same host recursion traceback_host.hpp).
#include "cuda_sim.h"

#include <cstdlib>
#include <cstring>
#include <string>

#include "../../ratatosk_b200/csrc/rtk_host_common.hpp"
#include "../../ratatosk_b200/csrc/traceback.cuh"
#include "../../ratatosk_b200/csrc/traceback_host.hpp"

using namespace rtk;

extern "C" int rtk_edlib_batch(rtk_ctx*, uint32_t, const char*, const uint64_t*, const char*, const uint64_t*, const uint8_t*,
                               const int32_t*, int32_t*, int32_t**, uint64_t**, uint64_t*);

template <int G, bool LC> static void sim_fill(rtk_fill_params p, const std::vector<uint32_t>& order) {
    if (order.empty()) return;
    p.order = order.data();
    p.n = (uint32_t)order.size();
    const uint64_t threads = (uint64_t)p.n * G;
    sim_launch((unsigned)((threads + RTK_MYERS_THREADS - 1) / RTK_MYERS_THREADS), RTK_MYERS_THREADS, [&] { rtk_myers_fill_kernel<G, LC>(p); });
}

struct SimTbBackend : TbBackend {
    const char* q; const char* t; const char* rq; const char* rt;
    std::vector<ulonglong2> mat;
    std::vector<int32_t> anchor, dist_buf;
    std::vector<int8_t> hb;
    std::vector<uint64_t> qb, tb;
    std::vector<uint32_t> ql, tl;

    template <bool LC> void run_fill(const std::vector<TbItem>& items, const TbPlan& pl, rtk_fill_params& fp) {
        const uint32_t n = (uint32_t)items.size();
        qb.assign(n, 0); tb.assign(n, 0); ql.assign(n, 0); tl.assign(n, 0);
        for (uint32_t i = 0; i < n; ++i) {
            qb[i] = (uint64_t)(((items[i].rev ? rq : q) + items[i].q_beg) - q);
            tb[i] = (uint64_t)(((items[i].rev ? rt : t) + items[i].t_beg) - t);
            ql[i] = items[i].q_len; tl[i] = items[i].t_len;
        }
        mat.assign(pl.cells + 1, ulonglong2{0, 0});
        anchor.assign(pl.cells + 1, 0);
        dist_buf.assign(n + 1, -1);
        hb.assign(pl.hb_off[n] + 1, 0);
        fp.q_pool = q; fp.q_beg = qb.data(); fp.q_len = ql.data(); fp.t_pool = t; fp.t_beg = tb.data(); fp.t_len = tl.data();
        fp.order = nullptr; fp.n = 0; fp.mat_off = pl.mat_off.data(); fp.mat = mat.data(); fp.anchor = anchor.data(); fp.dist = dist_buf.data();
        fp.hbound = hb.data(); fp.hb_off = pl.hb_off.data();
        sim_fill<1, LC>(fp, pl.order[0]); sim_fill<2, LC>(fp, pl.order[1]); sim_fill<4, LC>(fp, pl.order[2]);
        sim_fill<8, LC>(fp, pl.order[3]); sim_fill<16, LC>(fp, pl.order[4]); sim_fill<32, LC>(fp, pl.order[5]);
    }

    void direct(const std::vector<TbItem>& items, std::vector<std::vector<uint8_t>>& ops, std::vector<int32_t>& dist) override {
        const uint32_t n = (uint32_t)items.size();
        ops.assign(n, {});
        dist.assign(n, -1);
        if (!n) return;
        const TbPlan pl = plan_items(items, false);
        rtk_fill_params fp;
        run_fill<false>(items, pl, fp);
        std::vector<uint8_t> h_ops(pl.ops_off[n] + 1);
        std::vector<uint32_t> h_len(n + 1, 0);
        rtk_tb_params tp;
        tp.q_len = ql.data(); tp.t_len = tl.data(); tp.ids = pl.ids.data(); tp.n = n; tp.mat_off = pl.mat_off.data();
        tp.mat = mat.data(); tp.anchor = anchor.data(); tp.dist = dist_buf.data(); tp.ops_off = pl.ops_off.data(); tp.ops = h_ops.data(); tp.ops_len = h_len.data();
        sim_launch((n + 127) / 128, 128, [&] { rtk_traceback_kernel(tp); });
        for (uint32_t a = 0; a < n; ++a) {
            const uint64_t cap = (uint64_t)items[a].q_len + items[a].t_len;
            const uint8_t* src = h_ops.data() + pl.ops_off[a] + (cap - h_len[a]);
            ops[a].assign(src, src + h_len[a]);
            dist[a] = dist_buf[a];
        }
    }

    void last_column(const std::vector<TbItem>& items, std::vector<std::vector<int32_t>>& rows) override {
        const uint32_t n = (uint32_t)items.size();
        rows.assign(n, {});
        if (!n) return;
        const TbPlan pl = plan_items(items, true);
        rtk_fill_params fp;
        run_fill<true>(items, pl, fp);
        for (uint32_t a = 0; a < n; ++a) {
            const uint32_t nb = (items[a].q_len + 63) / 64;
            std::vector<uint64_t> P(nb), M(nb);
            for (uint32_t b = 0; b < nb; ++b) { P[b] = mat[pl.mat_off[a] + b].x; M[b] = mat[pl.mat_off[a] + b].y; }
            tb_rows_from_column(items[a].q_len, P.data(), M.data(), anchor.data() + pl.mat_off[a], rows[a]);
        }
    }
};

namespace rtk {
void nw_path_runs_masked(rtk_ctx*, uint32_t n, const char* q_pool, const uint64_t* q_off, const char* t_pool, const uint64_t* t_off,
                         const TbNeed& need, std::vector<std::vector<TbRun>>& runs, float* kernel_ms) {
    runs.assign(n, {});
    if (kernel_ms) *kernel_ms = 0.f;
    if (!n) return;
    const uint64_t qb = q_off[n] - q_off[0], tb = t_off[n] - t_off[0];
    std::vector<uint64_t> qrel(n + 1), trel(n + 1);
    std::vector<uint32_t> qlen(n + 1, 0), tlen(n + 1, 0);
    for (uint32_t i = 0; i <= n; ++i) { qrel[i] = q_off[i] - q_off[0]; trel[i] = t_off[i] - t_off[0]; }
    for (uint32_t i = 0; i < n; ++i) { qlen[i] = (uint32_t)(q_off[i + 1] - q_off[i]); tlen[i] = (uint32_t)(t_off[i + 1] - t_off[i]); }
    std::string rq(qb + 1, 'N'), rt(tb + 1, 'N');
    for (uint32_t a = 0; a < n; ++a) {
        const char* qs = q_pool + q_off[a];
        for (uint32_t i = 0; i < qlen[a]; ++i) rq[qrel[a] + i] = qs[qlen[a] - 1 - i];
        const char* ts = t_pool + t_off[a];
        for (uint32_t i = 0; i < tlen[a]; ++i) rt[trel[a] + i] = ts[tlen[a] - 1 - i];
    }
    SimTbBackend be;
    be.q = q_pool + q_off[0]; be.t = t_pool + t_off[0]; be.rq = rq.data(); be.rt = rt.data();
    solve_nw_runs(be, n, qrel.data(), qlen.data(), trel.data(), tlen.data(), &need, runs);
}
}  // namespace rtk

extern "C" int rtk_edlib_path_batch(rtk_ctx* c, uint32_t n, const char* q_pool, const uint64_t* q_off, const char* t_pool,
                                    const uint64_t* t_off, const uint8_t* mode, int32_t* dist, int32_t* end_loc,
                                    uint8_t** ops, uint64_t** ops_off, uint8_t* flags, uint64_t*) {
    return guarded([&] {
        std::vector<uint32_t> qlen(n + 1, 0), tlen(n + 1, 0);
        std::vector<uint64_t> qrel(n + 1), trel(n + 1);
        for (uint32_t i = 0; i <= n; ++i) { qrel[i] = q_off[i] - q_off[0]; trel[i] = t_off[i] - t_off[0]; }
        for (uint32_t i = 0; i < n; ++i) { qlen[i] = (uint32_t)(q_off[i + 1] - q_off[i]); tlen[i] = (uint32_t)(t_off[i + 1] - t_off[i]); }
        std::vector<uint32_t> teff(tlen);
        for (uint32_t a = 0; a < n; ++a) end_loc[a] = (int32_t)tlen[a] - 1;
        {
            std::vector<int32_t> km(n, -1), d(n);
            int32_t* e = nullptr; uint64_t* eo = nullptr;
            std::vector<uint8_t> md(mode, mode + n);
            if (rtk_edlib_batch(c, n, q_pool, q_off, t_pool, t_off, md.data(), km.data(), d.data(), &e, &eo, nullptr) != 0) throw std::runtime_error("sim K4 failed");
            for (uint32_t a = 0; a < n; ++a) if (mode[a] == 1) {
                dist[a] = d[a];
                end_loc[a] = (eo[a + 1] > eo[a]) ? e[eo[a]] : -1;
                teff[a] = (uint32_t)(end_loc[a] + 1);
            }
            free(e); free(eo);
        }
        const uint64_t qb = q_off[n] - q_off[0], tb = t_off[n] - t_off[0];
        std::string rq(qb + 1, 'N'), rt(tb + 1, 'N');
        for (uint32_t a = 0; a < n; ++a) {
            const char* qs = q_pool + q_off[a];
            for (uint32_t i = 0; i < qlen[a]; ++i) rq[qrel[a] + i] = qs[qlen[a] - 1 - i];
            const char* ts = t_pool + t_off[a];
            for (uint32_t i = 0; i < teff[a]; ++i) rt[trel[a] + i] = ts[teff[a] - 1 - i];
        }
        std::vector<uint32_t> ids;
        for (uint32_t a = 0; a < n; ++a) { flags[a] = 0; if (qlen[a] != 0 && tlen[a] != 0 && teff[a] != 0) ids.push_back(a); }
        std::vector<uint64_t> sq(ids.size()), stt(ids.size());
        std::vector<uint32_t> sql(ids.size()), stl(ids.size());
        for (size_t i = 0; i < ids.size(); ++i) { sq[i] = qrel[ids[i]]; stt[i] = trel[ids[i]]; sql[i] = qlen[ids[i]]; stl[i] = teff[ids[i]]; }
        SimTbBackend be;
        be.q = q_pool + q_off[0]; be.t = t_pool + t_off[0]; be.rq = rq.data(); be.rt = rt.data();
        std::vector<std::vector<uint8_t>> sops;
        std::vector<int32_t> sdist;
        solve_nw_paths(be, (uint32_t)ids.size(), sq.data(), sql.data(), stt.data(), stl.data(), sops, sdist);
        std::vector<std::vector<uint8_t>> all(n);
        for (size_t i = 0; i < ids.size(); ++i) { all[ids[i]] = std::move(sops[i]); if (mode[ids[i]] == 0) dist[ids[i]] = sdist[i]; }
        for (uint32_t a = 0; a < n; ++a) {
            if (qlen[a] == 0 || tlen[a] == 0) {
                if (mode[a] == 0) dist[a] = (int32_t)std::max(qlen[a], tlen[a]);
                else { dist[a] = (int32_t)qlen[a]; end_loc[a] = -1; }
            } else if (teff[a] == 0) all[a].assign(qlen[a], 1);
        }
        uint64_t* off = (uint64_t*)malloc((size_t)(n + 1) * 8);
        off[0] = 0;
        for (uint32_t a = 0; a < n; ++a) off[a + 1] = off[a] + all[a].size();
        uint8_t* out = (uint8_t*)malloc(off[n] + 1);
        for (uint32_t a = 0; a < n; ++a) if (!all[a].empty()) memcpy(out + off[a], all[a].data(), all[a].size());
        *ops = out;
        *ops_off = off;
    });
}
