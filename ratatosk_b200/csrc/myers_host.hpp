// myers_host.hpp — host-side planning of a K4 batch, shared by the CUDA driver and tests/hostsim.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <vector>

namespace rtk {

struct MyersPlan {
    std::vector<uint32_t> order[6];   // alignment ids per lane-group class G = 1,2,4,8,16,32
    std::vector<uint64_t> ends_off;   // n+1, capacity tlen + 1 per alignment (position -1 may be reported too)
    std::vector<uint64_t> hb_off;     // n+1
    std::vector<uint32_t> trivial;    // alignments with an empty query or target (answered on the host)
};

// class c uses G = 1<<c lanes per alignment: the smallest power of two covering the query's 64-row blocks
inline MyersPlan plan_myers(uint32_t n, const uint32_t* q_len, const uint32_t* t_len) {
    MyersPlan pl;
    pl.ends_off.assign(n + 1, 0);
    pl.hb_off.assign(n + 1, 0);
    for (uint32_t a = 0; a < n; ++a) {
        const uint64_t ql = q_len[a], tl = t_len[a];
        pl.ends_off[a + 1] = pl.ends_off[a] + tl + 1;
        pl.hb_off[a + 1] = pl.hb_off[a] + tl;
        if (ql == 0 || tl == 0) { pl.trivial.push_back(a); continue; }
        const uint64_t nb = (ql + 63) / 64;
        int c = 0;
        while (c < 5 && (1u << c) < nb) ++c;
        pl.order[c].push_back(a);
    }
    for (int c = 0; c < 6; ++c)  // longest targets first: the groups packed into one warp finish together
        std::stable_sort(pl.order[c].begin(), pl.order[c].end(), [&](uint32_t x, uint32_t y) {
            return t_len[x] > t_len[y];
        });
    return pl;
}

// edlibAlign's special case (src/edlib.cpp:160-176): no k check on this path
inline void myers_trivial(uint64_t ql, uint64_t tl, int mode, int32_t& dist, int32_t& end) {
    if ((mode & 3) == 0) { dist = (int32_t)std::max(ql, tl); end = (int32_t)tl - 1; }
    else { dist = (int32_t)ql; end = -1; }
}

}  // namespace rtk
