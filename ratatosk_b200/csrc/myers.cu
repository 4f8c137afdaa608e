// myers.cu — K4 driver: plans a batch, runs rtk_myers_kernel (myers.cuh) once per lane-group class, compacts the
// end locations on the device.  C ABI entry rtk_edlib_batch (include/rtk.h); myers_run is reused by the
// exploreSubGraph driver for leaf scoring.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "myers.cuh"
#include "myers_host.hpp"
#include "rtk_host_common.hpp"

namespace rtk {

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
    const MyersPlan pl = plan_myers(n, j.q_len, j.t_len);
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
    std::vector<int32_t> h_dn((size_t)n * 2 + 2);
    RTK_CUDA(counted_memcpy_async(h_dn.data(), d_dist, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    RTK_CUDA(cudaStreamSynchronize(st));
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
        RTK_CUDA(counted_memcpy_async(out, d_out, off[n] * 4, cudaMemcpyDeviceToHost, st));
        RTK_CUDA(cudaStreamSynchronize(st));
    }
    for (size_t i = 0; i < pl.trivial.size(); ++i) out[off[pl.trivial[i]]] = triv_end[i];
    *ends_out = out;
    *ends_off_out = off;
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
