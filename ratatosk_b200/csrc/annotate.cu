// annotate.cu — launchers of the graph annotation kernels (annotate.cuh): one packed upload of the job lists, persistent
// grids (one arena per resident warp, the warps stride over the unitigs), results back through pinned memory.
#include <cstring>
#include <vector>

#include "annotate.cuh"
#include "rtk_host_common.hpp"

namespace rtk {

static rtk_an_graph an_graph(const rtk_ctx* c, uint32_t min_cov) {
    const rtk_graph_view& g = c->dview;
    rtk_an_graph G;
    G.unitig_off = g.unitig_off; G.pool = g.pool; G.shared = g.shared; G.adj = g.adj; G.gset_of = g.gset_of; G.gset_off = g.gset_off;
    G.gset_ids = g.gset_ids; G.loc_off = g.loc_off; G.loc_ids = g.loc_ids; G.k = c->hdr.k; G.min_cov = min_cov;
    return G;
}

// resident warps: enough CTAs to fill the SMs (16 CTAs of 4 warps each), fewer when the arenas are large or the jobs few
static uint32_t an_grid(const rtk_ctx* c, uint32_t n_jobs, uint64_t arena_bytes_per_warp) {
    uint64_t grid = (uint64_t)c->sm_count * 16;
    const uint64_t budget = 2ull << 30;   // arena memory per launch
    while (grid > 1 && grid * RTK_AN_WARPS * arena_bytes_per_warp > budget) grid /= 2;
    const uint64_t need = (n_jobs + RTK_AN_WARPS - 1) / RTK_AN_WARPS;
    return (uint32_t)std::max<uint64_t>(1, std::min(grid, need));
}

void cycles_run(rtk_ctx* c, uint32_t min_cov, const uint32_t* list, uint32_t first, uint32_t n, uint32_t arena_cap,
                std::vector<uint8_t>& status, std::vector<uint32_t>& records, float* kernel_ms) {
    status.assign(n, 0);
    records.clear();
    if (kernel_ms) *kernel_ms = 0.f;
    if (!n) return;
    DeviceBind bind(c);
    cudaStream_t st = c->stream;
    const uint32_t grid = an_grid(c, n, 12ull * arena_cap);
    DevBuf &d_list = c->d_aux[0], &d_arena = c->d_aux[1], &d_status = c->d_aux[2], &d_out = c->d_aux[3], &d_cnt = c->d_aux[4];
    if (list) {
        d_list.reserve((uint64_t)n * 4);
        RTK_CUDA(counted_memcpy_async(d_list.p, list, (uint64_t)n * 4, cudaMemcpyHostToDevice, st));
    }
    d_arena.reserve((uint64_t)grid * RTK_AN_WARPS * 12ull * arena_cap);
    d_status.reserve(n);
    d_cnt.reserve(8);
    uint64_t cap = 4ull * n + (1u << 16);
    float ms_sum = 0.f;
    for (;;) {
        d_out.reserve(cap * 4);
        RTK_CUDA(cudaMemsetAsync(d_cnt.p, 0, 8, st));
        rtk_cyc_params p;
        p.g = an_graph(c, min_cov);
        p.list = list ? d_list.as<uint32_t>() : nullptr; p.first = first; p.n = n;
        p.arena = d_arena.as<uint32_t>(); p.arena_cap = arena_cap; p.status = d_status.as<uint8_t>();
        p.out = d_out.as<uint32_t>(); p.out_used = d_cnt.as<unsigned long long>(); p.out_cap = cap;
        RTK_CUDA(cudaEventRecord(c->ev0, st));
        ++g_launches;
        rtk_cycles_kernel<<<grid, RTK_AN_WARPS * 32, 0, st>>>(p);
        RTK_CUDA(cudaGetLastError());
        RTK_CUDA(cudaEventRecord(c->ev1, st));
        unsigned long long used = 0;
        PinnedD2H back(c, st);
        back.copy(0, &used, d_cnt.p, 8);
        back.copy(1, status.data(), d_status.p, n);
        back.sync();
        float ms = 0.f;
        RTK_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        ms_sum += ms;
        if (used <= cap) {
            records.resize(used);
            if (used) {
                PinnedD2H rec(c, st);
                rec.copy(2, records.data(), d_out.p, used * 4);
                rec.sync();
            }
            break;
        }
        cap = used;   // the kernel kept counting: second run with the exact size
    }
    if (kernel_ms) *kernel_ms = ms_sum;
}

void snp_run(rtk_ctx* c, uint32_t min_cov, const std::vector<rtk_snp_job>& jobs, const std::vector<rtk_snp_cand>& cands,
             std::vector<uint8_t>& fin, uint32_t n_bslots, uint32_t arena_cap, std::vector<uint8_t>& status, uint64_t* n_walks,
             float* kernel_ms) {
    status.assign(jobs.size(), 0);
    if (kernel_ms) *kernel_ms = 0.f;
    if (jobs.empty()) return;
    DeviceBind bind(c);
    cudaStream_t st = c->stream;
    const uint32_t n = (uint32_t)jobs.size();
    const uint32_t grid = an_grid(c, n, 16ull * arena_cap);
    // packed upload: [walks u64][jobs][cands][fin][tried][verdict][status]
    auto pad = [](uint64_t x) { return (x + 15) & ~15ull; };
    const uint64_t b_jobs = pad((uint64_t)n * sizeof(rtk_snp_job)), b_cands = pad(cands.size() * sizeof(rtk_snp_cand)), b_fin = pad(fin.size()),
                   b_ver = pad((uint64_t)n_bslots + 1), b_status = pad(n);
    const uint64_t o_jobs = 16, o_cands = o_jobs + b_jobs, o_fin = o_cands + b_cands, o_tried = o_fin + b_fin, o_ver = o_tried + b_fin,
                   o_status = o_ver + b_ver, bytes = o_status + b_status;
    PinBuf& H = c->h_fs;
    H.reserve(bytes);
    char* h = H.as<char>();
    memset(h, 0, 16);
    memcpy(h + o_jobs, jobs.data(), (uint64_t)n * sizeof(rtk_snp_job));
    memcpy(h + o_cands, cands.data(), cands.size() * sizeof(rtk_snp_cand));
    memcpy(h + o_fin, fin.data(), fin.size());
    memcpy(h + o_tried, fin.data(), fin.size());
    memset(h + o_ver, 0, b_ver + b_status);
    DevBuf &D = c->d_aux[0], &d_arena = c->d_aux[1];
    D.reserve(bytes);
    d_arena.reserve((uint64_t)grid * RTK_AN_WARPS * 16ull * arena_cap);
    char* d = D.as<char>();
    RTK_CUDA(counted_memcpy_async(d, h, bytes, cudaMemcpyHostToDevice, st));
    rtk_snp_params p;
    p.g = an_graph(c, min_cov);
    p.jobs = (const rtk_snp_job*)(d + o_jobs); p.n_jobs = n; p.cands = (const rtk_snp_cand*)(d + o_cands);
    p.fin = (uint8_t*)(d + o_fin); p.tried = (uint8_t*)(d + o_tried); p.verdict = (uint8_t*)(d + o_ver);
    p.arena = d_arena.as<uint32_t>(); p.arena_cap = arena_cap; p.status = (uint8_t*)(d + o_status); p.n_walks = (unsigned long long*)d;
    RTK_CUDA(cudaEventRecord(c->ev0, st));
    ++g_launches;
    rtk_snp_kernel<<<grid, RTK_AN_WARPS * 32, 0, st>>>(p);
    RTK_CUDA(cudaGetLastError());
    RTK_CUDA(cudaEventRecord(c->ev1, st));
    unsigned long long walks = 0;
    PinnedD2H back(c, st);
    back.copy(0, &walks, d, 8);
    back.copy(1, fin.data(), d + o_fin, fin.size());
    back.copy(2, status.data(), d + o_status, n);
    back.sync();
    if (kernel_ms) RTK_CUDA(cudaEventElapsedTime(kernel_ms, c->ev0, c->ev1));
    if (n_walks) *n_walks += walks;
}

void edge_flags_run(rtk_ctx* c, uint32_t min_cov, const uint64_t* col_off, const uint32_t* col_ids, uint64_t* kmcov, uint64_t* shared,
                    float* kernel_ms) {
    if (kernel_ms) *kernel_ms = 0.f;
    const uint32_t n = (uint32_t)c->hdr.n_unitigs;
    if (!n) return;
    DeviceBind bind(c);
    cudaStream_t st = c->stream;
    const uint64_t n_ids = col_off[n];
    // packed: [col_off][kmcov][shared][col_ids]
    const uint64_t b_off = (uint64_t)(n + 1) * 8, b_w = (uint64_t)n * 8, b_ids = (n_ids * 4 + 15) & ~15ull;
    const uint64_t o_km = b_off, o_sh = o_km + b_w, o_ids = o_sh + b_w, bytes = o_ids + b_ids + 16;
    PinBuf& H = c->h_fs;
    H.reserve(bytes);
    char* h = H.as<char>();
    memcpy(h, col_off, b_off);
    memcpy(h + o_km, kmcov, b_w);
    memcpy(h + o_sh, shared, b_w);
    if (n_ids) memcpy(h + o_ids, col_ids, n_ids * 4);
    DevBuf& D = c->d_aux[0];
    D.reserve(bytes);
    char* d = D.as<char>();
    RTK_CUDA(counted_memcpy_async(d, h, bytes - 16, cudaMemcpyHostToDevice, st));
    rtk_edge_params p;
    p.adj = c->dview.adj; p.col_off = (const uint64_t*)d; p.col_ids = (const uint32_t*)(d + o_ids); p.kmcov = (uint64_t*)(d + o_km);
    p.shared = (uint64_t*)(d + o_sh); p.n = n; p.min_cov = min_cov;
    const uint32_t grid = an_grid(c, n, 0);
    RTK_CUDA(cudaEventRecord(c->ev0, st));
    ++g_launches;
    rtk_edge_flags_kernel<<<grid, RTK_AN_WARPS * 32, 0, st>>>(p);
    RTK_CUDA(cudaGetLastError());
    RTK_CUDA(cudaEventRecord(c->ev1, st));
    PinnedD2H back(c, st);
    back.copy(0, kmcov, d + o_km, b_w);
    back.copy(1, shared, d + o_sh, b_w);
    back.sync();
    if (kernel_ms) RTK_CUDA(cudaEventElapsedTime(kernel_ms, c->ev0, c->ev1));
}

}  // namespace rtk
