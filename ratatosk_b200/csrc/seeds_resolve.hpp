// seeds_resolve.hpp — replay of searchSequence's order-dependent bookkeeping on the sparse
// hit list produced by the K1 kernels (Bifrost/src/Search.tcc:565-683).
//
// The kernels answer "is variant k-mer (variant, pos_s) in the graph, and where"; what the
// reference returns additionally depends on the ORDER in which it meets those k-mers:
//   * findUnitig extends a hit along the unitig (CompactedDBG.tcc:4479-4548) and the iterator
//     then jumps over the extended k-mers (Search.tcc:674), so they skip the tests of :666;
//   * every reported (position, mapped k-mer) is remembered in `us_pos_km` and reported once,
//     but UnitigMap::getMappedKmer(j) (UnitigMap.tcc:204-235) yields the EMPTY k-mer unless
//     j < um.len, so most keys collapse to (position, empty) and only the first such hit per
//     read position survives.
// Both effects are functions of the sorted sparse hit list only, so they are replayed here
// in O(#hits) per read.
#pragma once
#include <stdint.h>

#include <memory>
#include <utility>
#include <vector>

#include "../../include/rtk.h"
#include "flat_graph.h"

namespace rtk {

struct RawHit {
    uint64_t a;  // read(24) | variant(10) | pos_s(30)
    uint64_t b;  // P(40) | strand<<40
};

// vector whose resize / sized constructor leaves trivially-constructible elements uninitialised: the hit lists of a batch are
// hundreds of MB that are about to be overwritten by a device-to-host copy or a counting sort
template <typename T> struct DefaultInitAlloc : std::allocator<T> {
    template <typename U> struct rebind { typedef DefaultInitAlloc<U> other; };
    template <typename U> void construct(U* p) { ::new ((void*)p) U; }
    template <typename U, typename... A> void construct(U* p, A&&... a) { ::new ((void*)p) U(std::forward<A>(a)...); }
};
typedef std::vector<RawHit, DefaultInitAlloc<RawHit>> RawHitVec;

// raw: hits of ONE read, sorted ascending by `a`. Appends to out in reference order.
void resolve_exact(const rtk_graph_view& g, const RawHit* raw, size_t n, std::vector<rtk_hit>& out);
void resolve_inexact(const rtk_graph_view& g, const char* s, uint32_t slen, bool or_exclusive, const RawHit* raw,
                     size_t n, std::vector<rtk_hit>& out);

// dense exact sweep (rtk_k1_params::dense): dense[seq_off[r] - seq_off[0] + l] = P | strand << 40 or ~0.  Same output as
// resolve_exact on the sorted hit list of every read; reads in parallel.
void resolve_exact_dense(const rtk_graph_view& g, uint32_t n_reads, const uint64_t* seq_off, const uint64_t* dense,
                         std::vector<std::vector<rtk_hit>>& per_read);

}  // namespace rtk
