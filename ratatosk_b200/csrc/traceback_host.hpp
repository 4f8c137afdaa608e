// traceback_host.hpp — host planning of a K5 batch, shared by the CUDA driver and tests/hostsim.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <vector>

namespace rtk {

// edlib switches from the stored-matrix traceback to Hirschberg's recursion when its AlignmentData would
// reach 1 MiB (src/edlib.cpp:1191-1193); the two can return different (equally optimal) paths, so sizes at
// or above the switch are reported as unsupported until the recursion is restated.
inline bool tb_needs_hirschberg(uint64_t qlen, uint64_t tlen) {
    const uint64_t nb = (qlen + 63) / 64;
    return (2ull * 8 + 4) * nb * tlen + 2ull * 4 * tlen >= 1024ull * 1024ull;
}

struct TbPlan {
    std::vector<uint32_t> order[6];   // per lane-group class (G = 1 << c), ids into the batch
    std::vector<uint32_t> ids;        // all alignments that get a matrix, in class order
    std::vector<uint64_t> mat_off;    // [n] cells
    std::vector<uint64_t> ops_off;    // [n+1] capacity q_len + t_len
    uint64_t cells = 0;
};

// flags[a]: 0 = run on the device, 1 = unsupported size, 2 = trivial (answered on the host)
inline TbPlan plan_traceback(uint32_t n, const uint32_t* q_len, const uint32_t* t_len, const uint8_t* flags) {
    TbPlan pl;
    pl.mat_off.assign(n, 0);
    pl.ops_off.assign(n + 1, 0);
    for (uint32_t a = 0; a < n; ++a) {
        pl.ops_off[a + 1] = pl.ops_off[a] + (flags[a] == 0 ? (uint64_t)q_len[a] + t_len[a] : 0);
        if (flags[a] != 0) continue;
        const uint64_t nb = ((uint64_t)q_len[a] + 63) / 64;
        int c = 0;
        while ((1u << c) < nb) ++c;
        pl.order[c].push_back(a);
        pl.mat_off[a] = pl.cells;
        pl.cells += nb * t_len[a];
    }
    for (int c = 0; c < 6; ++c) {
        std::stable_sort(pl.order[c].begin(), pl.order[c].end(), [&](uint32_t x, uint32_t y) { return t_len[x] > t_len[y]; });
        pl.ids.insert(pl.ids.end(), pl.order[c].begin(), pl.order[c].end());
    }
    return pl;
}

}  // namespace rtk
