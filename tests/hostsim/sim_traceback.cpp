// sim_traceback.cpp — TEST INFRASTRUCTURE: rtk_edlib_path_batch on the CPU simulator (kernel sources traceback.cuh).
#include "cuda_sim.h"

#include <cstdlib>
#include <cstring>

#include "../../ratatosk_b200/csrc/rtk_host_common.hpp"
#include "../../ratatosk_b200/csrc/traceback.cuh"
#include "../../ratatosk_b200/csrc/traceback_host.hpp"

using namespace rtk;

extern "C" int rtk_edlib_batch(rtk_ctx*, uint32_t, const char*, const uint64_t*, const char*, const uint64_t*, const uint8_t*,
                               const int32_t*, int32_t*, int32_t**, uint64_t**, uint64_t*);

template <int G> static void sim_fill(rtk_fill_params p, const std::vector<uint32_t>& order) {
    if (order.empty()) return;
    p.order = order.data();
    p.n = (uint32_t)order.size();
    const uint64_t threads = (uint64_t)p.n * G;
    sim_launch((unsigned)((threads + RTK_MYERS_THREADS - 1) / RTK_MYERS_THREADS), RTK_MYERS_THREADS, [&] { rtk_myers_fill_kernel<G>(p); });
}

extern "C" int rtk_edlib_path_batch(rtk_ctx* c, uint32_t n, const char* q_pool, const uint64_t* q_off, const char* t_pool,
                                    const uint64_t* t_off, const uint8_t* mode, int32_t* dist, int32_t* end_loc,
                                    uint8_t** ops, uint64_t** ops_off, uint8_t* flags, uint64_t*) {
    return guarded([&] {
        std::vector<uint32_t> qlen(n + 1, 0), tlen(n + 1, 0);
        for (uint32_t i = 0; i < n; ++i) { qlen[i] = (uint32_t)(q_off[i + 1] - q_off[i]); tlen[i] = (uint32_t)(t_off[i + 1] - t_off[i]); }
        std::vector<uint32_t> teff(tlen);
        for (uint32_t a = 0; a < n; ++a) end_loc[a] = (int32_t)tlen[a] - 1;
        // SHW: distance + first end through the simulated K4
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
        for (uint32_t a = 0; a < n; ++a) {
            if (qlen[a] == 0 || tlen[a] == 0 || teff[a] == 0) flags[a] = 2;
            else if (tb_needs_hirschberg(qlen[a], teff[a]) || qlen[a] > 64 * 32) flags[a] = 1;
            else flags[a] = 0;
        }
        const TbPlan pl = plan_traceback(n, qlen.data(), teff.data(), flags);
        std::vector<ulonglong2> mat(pl.cells + 1);
        std::vector<int32_t> anchor(pl.cells + 1), h_dist(n + 1, -1);
        std::vector<uint8_t> h_ops(pl.ops_off[n] + 1);
        std::vector<uint32_t> h_len(n + 1, 0);
        rtk_fill_params fp;
        fp.q_pool = q_pool; fp.q_beg = q_off; fp.q_len = qlen.data(); fp.t_pool = t_pool; fp.t_beg = t_off; fp.t_len = teff.data();
        fp.order = nullptr; fp.n = 0; fp.mat_off = pl.mat_off.data(); fp.mat = mat.data(); fp.anchor = anchor.data(); fp.dist = h_dist.data();
        sim_fill<1>(fp, pl.order[0]); sim_fill<2>(fp, pl.order[1]); sim_fill<4>(fp, pl.order[2]);
        sim_fill<8>(fp, pl.order[3]); sim_fill<16>(fp, pl.order[4]); sim_fill<32>(fp, pl.order[5]);
        rtk_tb_params tp;
        tp.q_len = qlen.data(); tp.t_len = teff.data(); tp.ids = pl.ids.data(); tp.n = (uint32_t)pl.ids.size(); tp.mat_off = pl.mat_off.data();
        tp.mat = mat.data(); tp.anchor = anchor.data(); tp.dist = h_dist.data(); tp.ops_off = pl.ops_off.data(); tp.ops = h_ops.data(); tp.ops_len = h_len.data();
        if (tp.n) sim_launch((tp.n + 127) / 128, 128, [&] { rtk_traceback_kernel(tp); });
        uint64_t* off = (uint64_t*)malloc((size_t)(n + 1) * 8);
        off[0] = 0;
        for (uint32_t a = 0; a < n; ++a) {
            uint64_t len = 0;
            if (flags[a] == 0) len = h_len[a];
            else if (flags[a] == 2 && qlen[a] != 0 && tlen[a] != 0) len = qlen[a];
            off[a + 1] = off[a] + len;
        }
        uint8_t* out = (uint8_t*)malloc(off[n] + 1);
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
    });
}
