// graph_build.hpp — host-side graph containers and slab builder (see flat_graph.h).
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "flat_graph.h"

namespace rtk {

// Unflattened graph: what the index files (or a synthetic generator) provide per unitig.
struct HostGraph {
    int k = 31;
    double load_factor = 0.40;       // k-mer table fill (entries / capacity)
    double top_km_cov_ratio = 0.001; // Correct_Opt::top_km_cov_ratio (src/Common.hpp:124)
    std::vector<std::string> unitigs;                // forward spelling, upper-case ACGT
    std::vector<uint64_t> kmcov;                     // UnitigData::kmCov_cardBranches
    std::vector<uint64_t> shared;                    // UnitigData::shared_pids
    std::vector<std::vector<uint32_t>> global_ids;   // SharedPairID global set (sorted)
    std::vector<std::vector<uint32_t>> local_ids;    // SharedPairID local set (sorted)
    std::vector<std::vector<uint32_t>> amb_ids;      // (pos<<4)|iupac index (sorted)
    std::vector<std::vector<uint32_t>> hap_ids;
    std::vector<std::string> cycles;                 // raw compactedCycles blob ('\0'-separated)
};

struct rtk_slab {
    unsigned char* data = nullptr;  // 256-byte aligned, starts with rtk_slab_header
    uint64_t bytes = 0;
    bool mapped = false;            // data is a read-only file mapping (rtk_graph_open): released with munmap
};

HostGraph load_index(const std::string& fasta, const std::string& rtsk, int k);
rtk_slab build_slab(const HostGraph& hg);
HostGraph recolored_graph(const rtk_graph_view& g, const uint64_t* kmcov, const uint64_t* shared, const uint64_t* col_off, const uint32_t* col_ids);
void write_rtsk(const rtk_graph_view& g, const std::string& path, const uint64_t* amb_off, const uint32_t* amb_ids, const uint8_t* is_cycle,
                const uint64_t* cyc_off, const char* cyc_pool);
void patch_rtsk_annotations(const rtk_graph_view& g, const std::string& rtsk_in, const std::string& rtsk_out, const uint64_t* amb_off,
                            const uint32_t* amb_ids, const uint8_t* is_cycle, const uint64_t* cyc_off, const char* cyc_pool);

}  // namespace rtk
