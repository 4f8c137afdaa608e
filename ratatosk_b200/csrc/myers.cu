// myers.cu — K4 driver: plans a batch, runs rtk_myers_kernel (myers.cuh) once per lane-group class, compacts the
// end locations on the device.  C ABI entry rtk_edlib_batch (include/rtk.h); myers_run is reused by the
// exploreSubGraph driver for leaf scoring.
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "myers.cuh"
#include "myers_host.hpp"
#include "rtk_host_common.hpp"

namespace rtk {

// RTK_BROKER_PROFILE: where a K4 batch spends its host time (ns): plan, H2D issue, launches, first sync, ends phase
std::atomic<uint64_t> g_myers_prof[6];
struct MyersLap {
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(int slot) { const auto n = std::chrono::steady_clock::now(); g_myers_prof[slot] += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(n - t).count(); t = n; }
};

// ends (capacity layout) -> dense layout given the exclusive prefix sum of n_ends
__global__ void rtk_compact_ends_kernel(const int32_t* __restrict__ ends, const uint64_t* __restrict__ cap_off,
                                        const int32_t* __restrict__ n_ends, const uint64_t* __restrict__ out_off,
                                        int32_t* __restrict__ out, uint32_t n) {
    const uint32_t a = blockIdx.x;
    if (a >= n) return;
    const int32_t m = n_ends[a];
    const int32_t* src = ends + cap_off[a];
    int32_t* dst = out + out_off[a];
    for (int32_t i = threadIdx.x; i < m; i += blockDim.x) dst[i] = src[i];
}

// one class = one launch on its own side stream: the classes of a batch overlap on the device
template <int G> static void launch_class(rtk_ctx* c, int k, rtk_myers_params p, const uint32_t* d_order, uint32_t n) {
    if (!n) return;
    p.order = d_order;
    p.n = n;
    const uint64_t threads = (uint64_t)n * G;
    const uint32_t grid = (uint32_t)((threads + RTK_MYERS_THREADS - 1) / RTK_MYERS_THREADS);
    ++g_launches;
    rtk_myers_kernel<G><<<grid, RTK_MYERS_THREADS, 0, side_stream(c, k)>>>(p);
    RTK_CUDA(cudaGetLastError());
    fan_in(c, k);
}

void myers_run(rtk_ctx* c, const char* d_qpool, const char* d_tpool, const MyersJobs& j, int32_t* dist, bool want_ends,
               int32_t** ends_out, uint64_t** ends_off_out, float* kernel_ms) {
    const uint32_t n = j.n;
    if (kernel_ms) *kernel_ms = 0.f;
    if (want_ends) { *ends_out = nullptr; *ends_off_out = nullptr; }
    MyersLap lapt;
    const MyersPlan pl = plan_myers(n, j.q_len, j.t_len);
    lapt.lap(0);
    // device buffers: d_aux[2] u64 q_beg|t_beg|ends_off|hb_off then u32 q_len|t_len, [3] kmax+mode, [4] order,
    // [5] dist|n_ends, [6] ends (capacity layout), [7] hbound
    DevBuf* B = c->d_aux;
    B[2].reserve(5 * (size_t)(n + 1) * 8);
    B[3].reserve((size_t)n * 5 + 16);
    B[4].reserve((size_t)n * 4 + 16);
    B[5].reserve((size_t)n * 8 + 16);
    B[6].reserve(pl.ends_off[n] * 4 + 16);
    B[7].reserve(pl.hb_off[n] + 16);
    cudaStream_t st = c->stream;
    uint64_t* d_off = B[2].as<uint64_t>();
    RTK_CUDA(counted_memcpy_async(d_off, j.q_beg, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    RTK_CUDA(counted_memcpy_async(d_off + (n + 1), j.t_beg, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    RTK_CUDA(counted_memcpy_async(d_off + 2 * (n + 1), pl.ends_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
    RTK_CUDA(counted_memcpy_async(d_off + 3 * (n + 1), pl.hb_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
    uint32_t* d_len = (uint32_t*)(d_off + 4 * (n + 1));  // q_len[n+1] | t_len[n+1]
    RTK_CUDA(counted_memcpy_async(d_len, j.q_len, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    RTK_CUDA(counted_memcpy_async(d_len + (n + 1), j.t_len, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    int32_t* d_kmax = B[3].as<int32_t>();
    uint8_t* d_mode = (uint8_t*)(d_kmax + n);
    if (j.kmax) RTK_CUDA(counted_memcpy_async(d_kmax, j.kmax, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    else RTK_CUDA(cudaMemsetAsync(d_kmax, 0xff, (size_t)n * 4, st));
    RTK_CUDA(counted_memcpy_async(d_mode, j.mode, n, cudaMemcpyHostToDevice, st));
    std::vector<uint32_t> order_all;
    uint32_t cls_off[7] = {0};
    for (int k = 0; k < 6; ++k) { cls_off[k] = (uint32_t)order_all.size(); order_all.insert(order_all.end(), pl.order[k].begin(), pl.order[k].end()); }
    cls_off[6] = (uint32_t)order_all.size();
    if (!order_all.empty()) RTK_CUDA(counted_memcpy_async(B[4].p, order_all.data(), order_all.size() * 4, cudaMemcpyHostToDevice, st));
    int32_t* d_dist = B[5].as<int32_t>();
    int32_t* d_nends = d_dist + n;
    RTK_CUDA(cudaMemsetAsync(d_dist, 0xff, (size_t)n * 8, st));

    rtk_myers_params p;
    p.q_pool = d_qpool; p.q_beg = d_off; p.q_len = d_len; p.t_pool = d_tpool; p.t_beg = d_off + (n + 1); p.t_len = d_len + (n + 1);
    p.mode = d_mode; p.kmax = d_kmax; p.order = nullptr; p.n = 0;
    p.dist = d_dist; p.n_ends = d_nends; p.ends = B[6].as<int32_t>(); p.ends_off = d_off + 2 * (n + 1);
    p.hbound = B[7].as<int8_t>(); p.hb_off = d_off + 3 * (n + 1);
    lapt.lap(1);
    RTK_CUDA(cudaEventRecord(c->ev0, st));
    const uint32_t* d_order = B[4].as<uint32_t>();
    fan_out(c);
    launch_class<32>(c, 5, p, d_order + cls_off[5], cls_off[6] - cls_off[5]);
    launch_class<16>(c, 4, p, d_order + cls_off[4], cls_off[5] - cls_off[4]);
    launch_class<8>(c, 3, p, d_order + cls_off[3], cls_off[4] - cls_off[3]);
    launch_class<4>(c, 2, p, d_order + cls_off[2], cls_off[3] - cls_off[2]);
    launch_class<2>(c, 1, p, d_order + cls_off[1], cls_off[2] - cls_off[1]);
    launch_class<1>(c, 0, p, d_order + cls_off[0], cls_off[1] - cls_off[0]);
    RTK_CUDA(cudaEventRecord(c->ev1, st));
    lapt.lap(2);
    std::vector<int32_t> h_dn((size_t)n * 2 + 2);
    PinnedD2H d2h(c, st);
    d2h.copy(0, h_dn.data(), d_dist, (size_t)n * 8);
    d2h.sync();
    lapt.lap(3);
    if (kernel_ms) RTK_CUDA(cudaEventElapsedTime(kernel_ms, c->ev0, c->ev1));
    // alignments with an empty side are answered here (edlibAlign's special case)
    std::vector<int32_t> triv_end(pl.trivial.size());
    for (size_t i = 0; i < pl.trivial.size(); ++i) {
        const uint32_t a = pl.trivial[i];
        int32_t d, e;
        myers_trivial(j.q_len[a], j.t_len[a], j.mode[a], d, e);
        h_dn[a] = d; h_dn[n + a] = 1; triv_end[i] = e;
    }
    for (uint32_t a = 0; a < n; ++a) dist[a] = h_dn[a];
    if (!want_ends) return;
    uint64_t* off = (uint64_t*)malloc((size_t)(n + 1) * 8);
    if (!off) throw std::bad_alloc();
    off[0] = 0;
    for (uint32_t a = 0; a < n; ++a) off[a + 1] = off[a] + (uint64_t)std::max<int32_t>(0, h_dn[n + a]);
    int32_t* out = (int32_t*)malloc((off[n] + 1) * 4);
    if (!out) { free(off); throw std::bad_alloc(); }
    if (n) {
        c->d_hits.reserve((off[n] + 1) * 4 + (size_t)(n + 1) * 8);
        uint64_t* d_out_off = c->d_hits.as<uint64_t>();
        int32_t* d_out = (int32_t*)(d_out_off + (n + 1));
        // trivial alignments wrote nothing on the device: their n_ends is 0 for the gather
        std::vector<int32_t> ne(h_dn.begin() + n, h_dn.begin() + 2 * (size_t)n);
        for (uint32_t a : pl.trivial) ne[a] = 0;
        RTK_CUDA(counted_memcpy_async(d_nends, ne.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
        RTK_CUDA(counted_memcpy_async(d_out_off, off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
        ++g_launches;
        rtk_compact_ends_kernel<<<n, 64, 0, st>>>(p.ends, p.ends_off, d_nends, d_out_off, d_out, n);
        RTK_CUDA(cudaGetLastError());
        d2h.copy(1, out, d_out, off[n] * 4);
        d2h.sync();
    }
    for (size_t i = 0; i < pl.trivial.size(); ++i) out[off[pl.trivial[i]]] = triv_end[i];
    lapt.lap(4);
    *ends_out = out;
    *ends_off_out = off;
}

// Lean K4 batch for the library's own callers (broker services, leaf scoring, SHW pre-pass of K5): distance + first / last
// end column.  ONE packed H2D from pinned staging (descriptors + optional host pools), ONE fused launch, ONE D2H, one wait.
void myers_run_lean(rtk_ctx* c, const char* d_qpool, const char* d_tpool, const char* h_qpool, uint64_t q_bytes, const char* h_tpool,
                    uint64_t t_bytes, const MyersJobs& j, int32_t* dist, int32_t* first_end, int32_t* last_end, float* kernel_ms) {
    const uint32_t n = j.n;
    if (kernel_ms) *kernel_ms = 0.f;
    if (!n) return;
    MyersLap lapt;
    // classes in launch order: 32, 16, 8, 4, 2, 1 lanes
    std::vector<uint32_t> cls[6], trivial;
    std::vector<uint64_t> hb_off(n + 1, 0);
    for (uint32_t a = 0; a < n; ++a) {
        const uint64_t ql = j.q_len[a], tl = j.t_len[a];
        const uint64_t nb = (ql + 63) / 64;
        hb_off[a + 1] = hb_off[a] + (nb > 32 ? tl : 0);   // the spill row is only used by queries of more than one round
        if (ql == 0 || tl == 0) { trivial.push_back(a); continue; }
        int cc = 0;
        while (cc < 5 && (1u << cc) < nb) ++cc;
        cls[5 - cc].push_back(a);
    }
    rtk_myers_params p;
    uint32_t n_order = 0, n_blocks = 0;
    for (int k = 0; k < 6; ++k) {
        std::stable_sort(cls[k].begin(), cls[k].end(), [&](uint32_t x, uint32_t y) { return j.t_len[x] > j.t_len[y]; });
        const uint32_t G = 32u >> k;
        p.cls_ord[k] = n_order; p.cls_blk[k] = n_blocks;
        n_order += (uint32_t)cls[k].size();
        n_blocks += (uint32_t)(((uint64_t)cls[k].size() * G + RTK_MYERS_THREADS - 1) / RTK_MYERS_THREADS);
    }
    p.cls_ord[6] = n_order; p.cls_blk[6] = n_blocks;
    lapt.lap(0);
    // packed staging: u64 q_beg | t_beg | hb_off(n+1), u32 q_len | t_len | order, u8 mode (padded), then the pools
    auto al = [](uint64_t x) { return (x + 15) & ~15ull; };
    const uint64_t o_qbeg = 0, o_tbeg = o_qbeg + 8ull * n, o_hb = o_tbeg + 8ull * n, o_qlen = o_hb + 8ull * (n + 1), o_tlen = o_qlen + 4ull * n,
                   o_ord = o_tlen + 4ull * n, o_mode = o_ord + 4ull * n, o_qpool = al(o_mode + n), o_tpool = al(o_qpool + (h_qpool ? q_bytes : 0)),
                   total = al(o_tpool + (h_tpool ? t_bytes : 0));
    PinBuf& H = c->h_pin[10];
    H.reserve(total + 64);
    char* h = H.as<char>();
    uint64_t* hq = (uint64_t*)(h + o_qbeg); uint64_t* ht = (uint64_t*)(h + o_tbeg);
    for (uint32_t a = 0; a < n; ++a) { hq[a] = j.q_beg[a] + (h_qpool ? o_qpool : 0); ht[a] = j.t_beg[a] + (h_tpool ? o_tpool : 0); }
    memcpy(h + o_hb, hb_off.data(), 8ull * (n + 1));
    memcpy(h + o_qlen, j.q_len, 4ull * n);
    memcpy(h + o_tlen, j.t_len, 4ull * n);
    uint32_t* ho = (uint32_t*)(h + o_ord);
    for (int k = 0; k < 6; ++k) { if (!cls[k].empty()) memcpy(ho, cls[k].data(), cls[k].size() * 4); ho += cls[k].size(); }
    memcpy(h + o_mode, j.mode, n);
    if (h_qpool) memcpy(h + o_qpool, h_qpool, q_bytes);
    if (h_tpool) memcpy(h + o_tpool, h_tpool, t_bytes);
    DevBuf* B = c->d_aux;
    B[2].reserve(total + 64);
    B[5].reserve((size_t)n * 16 + 64);
    B[7].reserve(hb_off[n] + 64);
    cudaStream_t st = c->stream;
    char* d = B[2].as<char>();
    RTK_CUDA(counted_memcpy_async(d, h, total, cudaMemcpyHostToDevice, st));
    p.q_pool = h_qpool ? d : d_qpool; p.q_beg = (const uint64_t*)(d + o_qbeg); p.q_len = (const uint32_t*)(d + o_qlen);
    p.t_pool = h_tpool ? d : d_tpool; p.t_beg = (const uint64_t*)(d + o_tbeg); p.t_len = (const uint32_t*)(d + o_tlen);
    p.mode = (const uint8_t*)(d + o_mode); p.kmax = nullptr; p.order = (const uint32_t*)(d + o_ord); p.n = n_order;
    int32_t* d_res = B[5].as<int32_t>();
    p.dist = d_res; p.n_ends = d_res + n; p.first_end = d_res + 2 * (size_t)n; p.last_end = d_res + 3 * (size_t)n;
    p.ends = nullptr; p.ends_off = nullptr;
    p.hbound = B[7].as<int8_t>(); p.hb_off = (const uint64_t*)(d + o_hb);
    lapt.lap(1);
    if (kernel_ms) RTK_CUDA(cudaEventRecord(c->ev0, st));
    if (n_blocks) {
        ++g_launches;
        rtk_myers_fused_kernel<0><<<n_blocks, RTK_MYERS_THREADS, 0, st>>>(p);
        RTK_CUDA(cudaGetLastError());
    }
    if (kernel_ms) RTK_CUDA(cudaEventRecord(c->ev1, st));
    lapt.lap(2);
    std::vector<int32_t> res((size_t)n * 4);
    PinnedD2H d2h(c, st);
    d2h.copy(0, res.data(), d_res, (size_t)n * 16);
    d2h.sync();
    lapt.lap(3);
    if (kernel_ms) RTK_CUDA(cudaEventElapsedTime(kernel_ms, c->ev0, c->ev1));
    for (uint32_t a = 0; a < n; ++a) { dist[a] = res[a]; if (first_end) first_end[a] = res[2 * (size_t)n + a]; if (last_end) last_end[a] = res[3 * (size_t)n + a]; }
    for (const uint32_t a : trivial) {   // alignments with an empty side are answered here (edlibAlign's special case)
        int32_t dd, e;
        myers_trivial(j.q_len[a], j.t_len[a], j.mode[a], dd, e);
        dist[a] = dd;
        if (first_end) first_end[a] = e;
        if (last_end) last_end[a] = e;
    }
}

void dist_batch_lean(rtk_ctx* c, uint32_t n, const char* q_pool, uint64_t q_bytes, const uint64_t* q_beg, const uint32_t* q_len, const char* t_pool,
                     uint64_t t_bytes, const uint64_t* t_beg, const uint32_t* t_len, const uint8_t* mode, int32_t* dist, int32_t* first_end,
                     int32_t* last_end, uint64_t* stats) {
    RTK_CUDA(cudaSetDevice(c->device));
    MyersJobs j{n, q_beg, q_len, t_beg, t_len, mode, nullptr};
    float kms = 0.f;
    myers_run_lean(c, nullptr, nullptr, q_pool, q_bytes, t_pool, t_bytes, j, dist, first_end, last_end, stats ? &kms : nullptr);
    if (stats) stats[2] += (uint64_t)(kms * 1e6);
}

}  // namespace rtk

using namespace rtk;

extern "C" int rtk_edlib_batch(rtk_ctx* c, uint32_t n, const char* q_pool, const uint64_t* q_off, const char* t_pool,
                               const uint64_t* t_off, const uint8_t* mode, const int32_t* kmax, int32_t* dist,
                               int32_t** end_loc, uint64_t** end_off, uint64_t* stats) {
    return guarded([&] {
        if (!c || !q_pool || !q_off || !t_pool || !t_off || !mode || !kmax || !dist || !end_loc || !end_off)
            throw std::invalid_argument("null argument");
        RTK_CUDA(cudaSetDevice(c->device));
        for (uint32_t a = 0; a < n; ++a) if ((mode[a] & 3) > 2 || (mode[a] & ~7u)) throw std::invalid_argument("mode must be 0 (NW), 1 (SHW) or 2 (HW), optionally | 4 (plain equality)");
        const uint64_t qb = q_off[n] - q_off[0], tb = t_off[n] - t_off[0];
        std::vector<uint64_t> qrel(n + 1), trel(n + 1);
        std::vector<uint32_t> qlen(n + 1), tlen(n + 1);
        for (uint32_t i = 0; i <= n; ++i) { qrel[i] = q_off[i] - q_off[0]; trel[i] = t_off[i] - t_off[0]; }
        for (uint32_t i = 0; i < n; ++i) { qlen[i] = (uint32_t)(q_off[i + 1] - q_off[i]); tlen[i] = (uint32_t)(t_off[i + 1] - t_off[i]); }
        c->d_aux[0].reserve(qb + 16);
        c->d_aux[1].reserve(tb + 16);
        RTK_CUDA(counted_memcpy_async(c->d_aux[0].p, q_pool + q_off[0], qb, cudaMemcpyHostToDevice, c->stream));
        RTK_CUDA(counted_memcpy_async(c->d_aux[1].p, t_pool + t_off[0], tb, cudaMemcpyHostToDevice, c->stream));
        MyersJobs j{n, qrel.data(), qlen.data(), trel.data(), tlen.data(), mode, kmax};
        float kms = 0.f;
        myers_run(c, c->d_aux[0].as<char>(), c->d_aux[1].as<char>(), j, dist, true, end_loc, end_off, &kms);
        if (stats) stats[2] += (uint64_t)(kms * 1e6);
    });
}
