// sim_fixsnps.cpp — TEST INFRASTRUCTURE: the fixSNPs kernel source (fixsnps.cuh) on the CPU simulator.
#include "cuda_sim.h"

#include <cstring>
#include <string>
#include <vector>

#include "../../ratatosk_b200/csrc/fixsnps.cuh"
#include "../../ratatosk_b200/csrc/rtk_host_common.hpp"

namespace rtk {

void fix_snps_host(rtk_ctx* c, uint32_t n_reads, char* seq_pool, const uint64_t* seq_off, const std::vector<uint32_t>& amb_pos,
                   const std::vector<uint64_t>& amb_off, uint64_t* n_fixed) {
    if (n_fixed) *n_fixed = 0;
    if (!n_reads || amb_pos.empty()) return;
    const rtk_graph_view& g = c->host_graph->view;
    if (g.k > RTK_FS_MAXK) throw std::invalid_argument("fixSNPs: k > 64");
    std::vector<uint8_t> done(amb_pos.size(), 0);
    unsigned long long fixed = 0;
    rtk_fs_params p;
    p.table = g.table; p.n_buckets = g.n_buckets; p.pool = g.pool; p.k = g.k; p.n_reads = n_reads;
    p.seq = seq_pool; p.seq_off = seq_off; p.amb_off = amb_off.data(); p.amb_pos = amb_pos.data(); p.amb_done = done.data(); p.n_fixed = &fixed;
    const unsigned grid = (n_reads + RTK_FS_WARPS - 1) / RTK_FS_WARPS;
    if (g.k <= 32) sim_launch(grid, RTK_FS_WARPS * 32, [&] { rtk_fixsnps_kernel<uint64_t>(p); });
    else sim_launch(grid, RTK_FS_WARPS * 32, [&] { rtk_fixsnps_kernel<rtk_u128>(p); });
    if (n_fixed) *n_fixed = fixed;
}

}  // namespace rtk
