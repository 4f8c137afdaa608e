// fixsnps.cu — launcher of the fixSNPs kernel (fixsnps.cuh): reads + the list of their non-ACGT positions go up in one packed
// upload, the kernel rewrites the resolved codes in place, the reads come back.  Batches without any code launch nothing.
#include <cstring>
#include <string>
#include <vector>

#include "fixsnps.cuh"
#include "rtk_host_common.hpp"

namespace rtk {

void fix_snps_host(rtk_ctx* c, uint32_t n_reads, char* seq_pool, const uint64_t* seq_off, const std::vector<uint32_t>& amb_pos,
                   const std::vector<uint64_t>& amb_off, uint64_t* n_fixed) {
    if (n_fixed) *n_fixed = 0;
    if (!n_reads || amb_pos.empty()) return;
    if (c->hdr.k > RTK_FS_MAXK) throw std::invalid_argument("fixSNPs: k > 64");
    DeviceBind bind(c);
    cudaStream_t st = c->stream;
    const uint64_t total = seq_off[n_reads] - seq_off[0];
    std::vector<uint64_t> off0(n_reads + 1);
    for (uint32_t i = 0; i <= n_reads; ++i) off0[i] = seq_off[i] - seq_off[0];
    // packed layout: [n_fixed u64][seq_off][amb_off][amb_pos u32][amb_done u8][seq]
    const uint64_t b_off = (uint64_t)(n_reads + 1) * 8, b_pos = (amb_pos.size() * 4 + 7) & ~7ull, b_done = (amb_pos.size() + 7) & ~7ull;
    const uint64_t o_soff = 8, o_aoff = o_soff + b_off, o_pos = o_aoff + b_off, o_done = o_pos + b_pos, o_seq = o_done + b_done;
    const uint64_t bytes = o_seq + total + 16;
    PinBuf& H = c->h_fs;
    H.reserve(bytes);
    char* h = H.as<char>();
    memset(h, 0, 8);
    memcpy(h + o_soff, off0.data(), b_off);
    memcpy(h + o_aoff, amb_off.data(), b_off);
    memcpy(h + o_pos, amb_pos.data(), amb_pos.size() * 4);
    memset(h + o_done, 0, b_done);
    memcpy(h + o_seq, seq_pool + seq_off[0], total);
    DevBuf& D = c->d_fs;
    D.reserve(bytes);
    char* d = D.as<char>();
    RTK_CUDA(counted_memcpy_async(d, h, bytes - 16, cudaMemcpyHostToDevice, st));
    rtk_fs_params p;
    p.table = c->dview.table; p.n_buckets = c->hdr.n_buckets; p.pool = c->dview.pool; p.k = c->hdr.k; p.n_reads = n_reads;
    p.seq = d + o_seq; p.seq_off = (const uint64_t*)(d + o_soff); p.amb_off = (const uint64_t*)(d + o_aoff);
    p.amb_pos = (const uint32_t*)(d + o_pos); p.amb_done = (uint8_t*)(d + o_done); p.n_fixed = (unsigned long long*)d;
    const uint32_t grid = (n_reads + RTK_FS_WARPS - 1) / RTK_FS_WARPS;
    ++g_launches;
    if (c->hdr.k <= 32) rtk_fixsnps_kernel<uint64_t><<<grid, RTK_FS_WARPS * 32, 0, st>>>(p);
    else rtk_fixsnps_kernel<rtk_u128><<<grid, RTK_FS_WARPS * 32, 0, st>>>(p);
    RTK_CUDA(cudaGetLastError());
    RTK_CUDA(counted_memcpy_async(h, d, 8, cudaMemcpyDeviceToHost, st));
    RTK_CUDA(counted_memcpy_async(h + o_seq, d + o_seq, total, cudaMemcpyDeviceToHost, st));
    stream_wait(st);
    memcpy(seq_pool + seq_off[0], h + o_seq, total);
    if (n_fixed) memcpy(n_fixed, h, 8);
}

}  // namespace rtk
