// sim_myers.cpp — TEST INFRASTRUCTURE: rtk_edlib_batch on the CPU simulator (same kernel source, same planner).
#include "cuda_sim.h"

#include <cstdlib>
#include <cstring>

#include "../../ratatosk_b200/csrc/myers.cuh"
#include "../../ratatosk_b200/csrc/myers_host.hpp"
#include "../../ratatosk_b200/csrc/rtk_host_common.hpp"

using namespace rtk;

template <int G> static void sim_class(rtk_myers_params p, const std::vector<uint32_t>& order) {
    if (order.empty()) return;
    p.order = order.data();
    p.n = (uint32_t)order.size();
    const uint64_t threads = (uint64_t)p.n * G;
    const unsigned grid = (unsigned)((threads + RTK_MYERS_THREADS - 1) / RTK_MYERS_THREADS);
    sim_launch(grid, RTK_MYERS_THREADS, [&] { rtk_myers_kernel<G>(p); });
}

extern "C" int rtk_edlib_batch(rtk_ctx*, uint32_t n, const char* q_pool, const uint64_t* q_off, const char* t_pool,
                               const uint64_t* t_off, const uint8_t* mode, const int32_t* kmax, int32_t* dist,
                               int32_t** end_loc, uint64_t** end_off, uint64_t*) {
    return guarded([&] {
        std::vector<uint32_t> qlen(n + 1, 0), tlen(n + 1, 0);
        for (uint32_t i = 0; i < n; ++i) { qlen[i] = (uint32_t)(q_off[i + 1] - q_off[i]); tlen[i] = (uint32_t)(t_off[i + 1] - t_off[i]); }
        const MyersPlan pl = plan_myers(n, qlen.data(), tlen.data());
        std::vector<int32_t> d(n, -1), ne(n, 0), ends(pl.ends_off[n] + 1, 0);
        std::vector<int8_t> hb(pl.hb_off[n] + 1, 0);
        rtk_myers_params p;
        p.q_pool = q_pool; p.q_beg = q_off; p.q_len = qlen.data(); p.t_pool = t_pool; p.t_beg = t_off; p.t_len = tlen.data(); p.mode = mode; p.kmax = kmax;
        p.order = nullptr; p.n = 0; p.dist = d.data(); p.n_ends = ne.data(); p.ends = ends.data(); p.ends_off = pl.ends_off.data();
        p.hbound = hb.data(); p.hb_off = pl.hb_off.data();
        sim_class<1>(p, pl.order[0]); sim_class<2>(p, pl.order[1]); sim_class<4>(p, pl.order[2]);
        sim_class<8>(p, pl.order[3]); sim_class<16>(p, pl.order[4]); sim_class<32>(p, pl.order[5]);
        for (uint32_t a : pl.trivial) {
            int32_t dd, e;
            myers_trivial(q_off[a + 1] - q_off[a], t_off[a + 1] - t_off[a], mode[a], dd, e);
            d[a] = dd; ne[a] = 1; ends[pl.ends_off[a]] = e;
        }
        uint64_t* off = (uint64_t*)malloc((size_t)(n + 1) * 8);
        off[0] = 0;
        for (uint32_t a = 0; a < n; ++a) off[a + 1] = off[a] + (uint64_t)ne[a];
        int32_t* out = (int32_t*)malloc((off[n] + 1) * 4);
        for (uint32_t a = 0; a < n; ++a) {
            for (int32_t i = 0; i < ne[a]; ++i) out[off[a] + i] = ends[pl.ends_off[a] + i];
            dist[a] = d[a];
        }
        *end_loc = out;
        *end_off = off;
    });
}
