// graph_build.cpp — host side: read a Ratatosk index (unitig FASTA[.gz] + .rtsk) or a
// plain list of unitigs and flatten it into the slab described in flat_graph.h.
//
// File formats consumed (written by the reference, re-parsed here from their layout):
//   *.fasta[.gz]  one record per unitig (CompactedDBG::write, Bifrost/src/IO.tcc)
//   *.rtsk        per unitig: head Kmer (MAX_KMER_SIZE=64 => 2 x u64, first base at bit 62 of
//                 word 0; Bifrost/src/Kmer.cpp:313-335), then UnitigData::write
//                 (src/UnitigData.hpp:493-517): kmCov_cardBranches u64, shared_pids u64,
//                 SharedPairID {global PairID, local PairID} (src/SharedPairID.cpp:445-463),
//                 ambiguity PairID, hap PairID, cycles_len u64, cycles bytes.
//   PairID        u64 word, low 3 bits = kind (src/PairID.cpp:1137-1174): 0 TinyBitmap follows,
//                 1 61-bit vector in bits 3..63, 2 single id in bits 3..63,
//                 3 CRoaring portable blob of (word>>3) bytes follows.
#include "graph_build.hpp"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "lookup.cuh"

namespace rtk {

// ------------------------------------------------------------------ small byte reader
struct ByteReader {
    const unsigned char* p;
    const unsigned char* e;
    bool eof() const { return p >= e; }
    void need(size_t n) const {
        if ((size_t)(e - p) < n) throw std::runtime_error("rtsk: truncated file");
    }
    uint64_t u64() { need(8); uint64_t v; memcpy(&v, p, 8); p += 8; return v; }
    uint32_t u32() { need(4); uint32_t v; memcpy(&v, p, 4); p += 4; return v; }
    uint16_t u16() { need(2); uint16_t v; memcpy(&v, p, 2); p += 2; return v; }
    const unsigned char* bytes(size_t n) { need(n); const unsigned char* r = p; p += n; return r; }
};

// TinyBitmap payload (Bifrost/src/TinyBitmap.cpp:825-878, iterator :1309-1395)
static void parse_tiny_bitmap(ByteReader& r, std::vector<uint32_t>& out) {
    const uint16_t header = r.u16();
    const uint16_t sz = header >> 3;
    if (sz == 0) return;
    std::vector<uint16_t> w(sz, 0);
    w[0] = header;
    for (uint16_t i = 1; i < sz; ++i) w[i] = r.u16();
    if (sz < 3) return;
    const uint16_t mode = header & 0x6;
    const uint16_t card = w[1];
    const uint32_t offset = ((uint32_t)w[2]) << 16;
    if (mode == 0x0) {  // bitmap
        for (uint32_t i = 3; i < sz; ++i)
            for (uint32_t j = 0; j < 16; ++j)
                if ((w[i] >> j) & 1) out.push_back(offset | (((i - 3) << 4) + j));
    } else if (mode == 0x2) {  // sorted list
        for (uint32_t i = 3; i < (uint32_t)card + 3 && i < sz; ++i) out.push_back(offset | w[i]);
    } else {  // run-length list: pairs [start, end] inclusive, `card` = words used
        for (uint32_t i = 3; i + 1 < (uint32_t)card + 3 && i + 1 < sz; i += 2)
            for (uint32_t v = w[i]; v <= w[i + 1]; ++v) out.push_back(offset | v);
    }
}

// CRoaring portable serialisation (Bifrost/src/roaring.c:10554-10700)
static void parse_roaring(const unsigned char* buf, size_t n, std::vector<uint32_t>& out) {
    ByteReader r{buf, buf + n};
    const uint32_t cookie = r.u32();
    uint32_t size;
    bool hasrun = false;
    if ((cookie & 0xFFFF) == 12347) { hasrun = true; size = (cookie >> 16) + 1; }
    else if (cookie == 12346) size = r.u32();
    else throw std::runtime_error("rtsk: bad roaring cookie");
    const unsigned char* runbm = nullptr;
    if (hasrun) runbm = r.bytes((size + 7) / 8);
    std::vector<uint16_t> keys(size), cards(size);
    for (uint32_t i = 0; i < size; ++i) { keys[i] = r.u16(); cards[i] = r.u16(); }
    if (!hasrun || size >= 4) r.bytes((size_t)size * 4);
    for (uint32_t i = 0; i < size; ++i) {
        const uint32_t hi = ((uint32_t)keys[i]) << 16;
        const uint32_t card = (uint32_t)cards[i] + 1;
        const bool isrun = hasrun && ((runbm[i / 8] >> (i % 8)) & 1);
        if (isrun) {
            const uint16_t nruns = r.u16();
            for (uint16_t j = 0; j < nruns; ++j) {
                const uint32_t s = r.u16(), l = r.u16();
                for (uint32_t v = s; v <= s + l; ++v) out.push_back(hi | v);
            }
        } else if (card > 4096) {
            for (uint32_t wi = 0; wi < 1024; ++wi) {
                uint64_t w = r.u64();
                while (w) { const int b = __builtin_ctzll(w); out.push_back(hi | (wi * 64 + b)); w &= w - 1; }
            }
        } else {
            for (uint32_t j = 0; j < card; ++j) out.push_back(hi | r.u16());
        }
    }
}

static void parse_pairid(ByteReader& r, std::vector<uint32_t>& out) {
    out.clear();
    const uint64_t w = r.u64();
    const uint64_t flag = w & 7;
    if (flag == 0) parse_tiny_bitmap(r, out);
    else if (flag == 1) { for (uint32_t i = 0; i < 61; ++i) if ((w >> (i + 3)) & 1) out.push_back(i); }
    else if (flag == 2) out.push_back((uint32_t)(w >> 3));
    else if (flag == 3) { const size_t n = (size_t)(w >> 3); parse_roaring(r.bytes(n), n, out); }
    else throw std::runtime_error("rtsk: unknown PairID kind");
    if (!std::is_sorted(out.begin(), out.end())) std::sort(out.begin(), out.end());
}

// ------------------------------------------------------------------ file helpers
static std::vector<unsigned char> slurp(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    std::vector<unsigned char> buf;
    unsigned char tmp[1 << 16];
    size_t n;
    while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    fclose(f);
    return buf;
}

static void read_fasta(const std::string& path, std::vector<std::string>& seqs) {
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    gzbuffer(f, 1 << 20);
    std::string line, cur;
    bool in_rec = false;
    std::vector<char> buf(1 << 20);
    std::string pending;
    auto flush = [&]() { if (in_rec) seqs.push_back(std::move(cur)); cur.clear(); };
    int n;
    while ((n = gzread(f, buf.data(), (unsigned)buf.size())) > 0) {
        pending.append(buf.data(), (size_t)n);
        size_t start = 0, nl;
        while ((nl = pending.find('\n', start)) != std::string::npos) {
            size_t end = nl;
            if (end > start && pending[end - 1] == '\r') --end;
            if (end > start && pending[start] == '>') { flush(); in_rec = true; }
            else if (in_rec) cur.append(pending, start, end - start);
            start = nl + 1;
        }
        pending.erase(0, start);
    }
    if (!pending.empty()) {
        if (pending[0] == '>') { flush(); in_rec = true; }
        else if (in_rec) cur.append(pending);
    }
    flush();
    gzclose(f);
    for (auto& s : seqs) for (auto& c : s) c = (char)(c & 0xDF);
}

// ------------------------------------------------------------------ slab assembly
static inline uint64_t align256(uint64_t x) { return (x + 255) & ~(uint64_t)255; }

template <typename KT>
static void build_table_and_adj(const HostGraph& hg, rtk_slab_header& h, unsigned char* slab) {
    const int k = (int)h.k;
    uint64_t* table = (uint64_t*)(slab + h.off_table);
    uint32_t* t32 = (uint32_t*)table;
    const uint64_t* pool = (const uint64_t*)(slab + h.off_pool);
    const uint64_t* uoff = (const uint64_t*)(slab + h.off_unitig_off);
    const KT mask = KmerOps<KT>::mask(k);
    // insert every k-mer of every unitig
    for (uint64_t u = 0; u < h.n_unitigs; ++u) {
        const std::string& s = hg.unitigs[u];
        KT fw = 0, rc = 0;
        for (size_t i = 0; i < s.size(); ++i) {
            const KT c = (KT)rtk_base_code(s[i]);
            fw = ((fw << 2) | c) & mask;
            rc = (rc >> 2) | ((KT)(3 - c) << (2 * (k - 1)));
            if (i + 1 >= (size_t)k) {
                const uint64_t P = uoff[u] + (i + 1 - k);
                const KT canon = fw < rc ? fw : rc;
                const uint64_t hh = rtk_hash_kmer<KT>(canon);
                uint64_t b = rtk_bucket_of(hh, h.n_buckets);
                const uint32_t hi = (rtk_tag_of(hh) << 8) | (uint32_t)(P >> 32);  // aux nibble (bits 4..7) stays 0
                for (;;) {
                    bool done = false;
                    for (int e = 0; e < RTK_BUCKET_ENTRIES; ++e) {
                        if ((t32[8 * b + e] >> 8) == 0) {  // tag 0 = empty (aux bits of entry 0 are only set on full buckets)
                            t32[8 * b + e] = hi;
                            t32[8 * b + 4 + e] = (uint32_t)P;
                            done = true;
                            break;
                        }
                    }
                    if (done) break;
                    t32[8 * b] |= 1u << (4 + rtk_class_of(hh));  // a key of this class was bumped past this (full) bucket
                    b = (b + 1 == h.n_buckets) ? 0 : b + 1;
                }
            }
        }
    }
    // explicit adjacency: 4 fw successors then 4 bw predecessors per unitig, A,C,G,T order
    // (NeighborIterator.tcc:25-47: find(tail.forwardBase(c), extremities_only=true))
    uint32_t* adj = (uint32_t*)(slab + h.off_adj);
    const uint32_t* blk = (const uint32_t*)(slab + h.off_blk2unitig);
    for (uint64_t u = 0; u < h.n_unitigs; ++u) {
        const uint64_t len = uoff[u + 1] - uoff[u];
        const KT head = rtk_pool_kmer<KT>(pool, uoff[u], k);
        const KT tail = rtk_pool_kmer<KT>(pool, uoff[u] + len - k, k);
        for (int side = 0; side < 2; ++side) {
            for (int c = 0; c < 4; ++c) {
                KT fw;
                if (side == 0) fw = ((tail << 2) | (KT)c) & mask;              // tail.forwardBase(c)
                else fw = (head >> 2) | ((KT)c << (2 * (k - 1)));              // head.backwardBase(c)
                const KT rc = KmerOps<KT>::rc(fw, k);
                rtk_kmer_hit hit;
                uint32_t val = RTK_NONE32;
                if (rtk_lookup<KT>(table, h.n_buckets, pool, k, fw, rc, hit)) {
                    const uint32_t v = rtk_unitig_of(blk, uoff, hit.P);
                    const uint64_t off = hit.P - uoff[v];
                    const uint64_t vlen = uoff[v + 1] - uoff[v];
                    // successor must be entered through an extremity: its head if same strand,
                    // its tail if opposite strand (mirror for predecessors)
                    const bool at_head = (off == 0), at_tail = (off == vlen - k);
                    const bool ok = (side == 0) ? (hit.strand ? at_head : at_tail) : (hit.strand ? at_tail : at_head);
                    if (ok) val = v | (hit.strand ? 0x80000000u : 0u);
                }
                adj[8 * u + 4 * side + c] = val;
            }
        }
    }
}

rtk_slab build_slab(const HostGraph& hg) {
    const int k = hg.k;
    if (k < 3 || k > 64) throw std::runtime_error("k must be in [3,64]");
    const uint64_t n = hg.unitigs.size();
    if (n >= 0x7FFFFFFFull) throw std::runtime_error("too many unitigs");
    rtk_slab_header h;
    memset(&h, 0, sizeof(h));
    h.magic = RTK_SLAB_MAGIC; h.version = RTK_SLAB_VERSION; h.k = (uint32_t)k;
    h.n_unitigs = n;
    uint64_t bases = 0, kmers = 0;
    for (const auto& s : hg.unitigs) {
        if (s.size() < (size_t)k) throw std::runtime_error("unitig shorter than k");
        bases += s.size(); kmers += s.size() - k + 1;
    }
    h.pool_bases = bases; h.n_kmers = kmers;
    h.pool_words = (bases + 31) / 32 + 4;  // +4: k-mer extraction may touch two words past the end
    if (bases >= RTK_POS_MASK) throw std::runtime_error("pool exceeds 36-bit positions");
    if (kmers / 2 >= 0xFFFFFFFFull) throw std::runtime_error("k-mer table exceeds 2^32 buckets");
    h.n_buckets = std::max<uint64_t>(16, (uint64_t)std::ceil((double)kmers / (RTK_BUCKET_ENTRIES * hg.load_factor)));
    // de-duplicate global colour sets by content (src/Graph.cpp:756-769)
    std::vector<uint32_t> gset_of(n, RTK_NONE32);
    std::vector<const std::vector<uint32_t>*> gsets;
    {
        std::unordered_map<uint64_t, std::vector<uint32_t>> by_hash;  // content hash -> gset ids
        for (uint64_t u = 0; u < n && u < hg.global_ids.size(); ++u) {
            const std::vector<uint32_t>& g = hg.global_ids[u];
            if (g.empty()) continue;
            uint64_t hh = 0x9E3779B97F4A7C15ULL ^ g.size();
            for (uint32_t id : g) hh = rtk_mix64(hh ^ id);
            std::vector<uint32_t>& cand = by_hash[hh];
            uint32_t found = RTK_NONE32;
            for (uint32_t gi : cand) if (*gsets[gi] == g) { found = gi; break; }
            if (found == RTK_NONE32) { found = (uint32_t)gsets.size(); gsets.push_back(&g); cand.push_back(found); }
            gset_of[u] = found;
        }
    }
    h.n_gsets = gsets.size();
    auto csr_total = [&](const std::vector<std::vector<uint32_t>>& v) {
        uint64_t t = 0; for (const auto& x : v) t += x.size(); return t;
    };
    uint64_t gset_total = 0; for (auto* g : gsets) gset_total += g->size();
    const uint64_t loc_total = csr_total(hg.local_ids), amb_total = csr_total(hg.amb_ids), hap_total = csr_total(hg.hap_ids);
    uint64_t cyc_total = 0; for (const auto& c : hg.cycles) cyc_total += c.size();

    uint64_t off = align256(sizeof(rtk_slab_header));
    auto place = [&](uint64_t bytes) { const uint64_t o = off; off = align256(off + bytes); return o; };
    h.off_unitig_off = place((n + 1) * 8);
    h.off_pool = place(h.pool_words * 8);
    h.off_table = place(h.n_buckets * 32);
    h.off_blk2unitig = place(((bases >> 7) + 2) * 4);
    h.off_kmcov = place(n * 8);
    h.off_shared = place(n * 8);
    h.off_adj = place(n * 32);
    h.off_gset_of = place(n * 4);
    h.off_gset_off = place((h.n_gsets + 1) * 8);
    h.off_gset_ids = place(gset_total * 4 + 4);
    h.off_loc_off = place((n + 1) * 8);
    h.off_loc_ids = place(loc_total * 4 + 4);
    h.off_amb_off = place((n + 1) * 8);
    h.off_amb_ids = place(amb_total * 4 + 4);
    h.off_hap_off = place((n + 1) * 8);
    h.off_hap_ids = place(hap_total * 4 + 4);
    h.off_cyc_off = place((n + 1) * 8);
    h.off_cyc_pool = place(cyc_total + 4);
    h.total_bytes = off;

    rtk_slab slab;
    slab.bytes = h.total_bytes;
    slab.data = (unsigned char*)aligned_alloc(256, h.total_bytes);
    if (!slab.data) throw std::runtime_error("out of host memory for slab");
    memset(slab.data, 0, h.total_bytes);
    unsigned char* S = slab.data;

    uint64_t* uoff = (uint64_t*)(S + h.off_unitig_off);
    uint64_t* pool = (uint64_t*)(S + h.off_pool);
    uint32_t* blk = (uint32_t*)(S + h.off_blk2unitig);
    {
        uint64_t P = 0;
        for (uint64_t u = 0; u < n; ++u) {
            uoff[u] = P;
            for (char c : hg.unitigs[u]) {
                const uint32_t code = rtk_base_code(c);
                if (code > 3) throw std::runtime_error("non-ACGT base in unitig");
                pool[P >> 5] |= ((uint64_t)code) << (62 - 2 * (P & 31));
                ++P;
            }
        }
        uoff[n] = P;
        uint64_t u = 0;
        for (uint64_t b = 0; b <= (bases >> 7) + 1; ++b) {
            const uint64_t pos = b << 7;
            while (u + 1 < n && uoff[u + 1] <= pos) ++u;
            blk[b] = (uint32_t)u;
        }
    }
    auto fill_csr = [&](const std::vector<std::vector<uint32_t>>& v, uint64_t o_off, uint64_t o_ids) {
        uint64_t* o = (uint64_t*)(S + o_off);
        uint32_t* ids = (uint32_t*)(S + o_ids);
        uint64_t t = 0;
        for (uint64_t u = 0; u < n; ++u) {
            o[u] = t;
            if (u < v.size()) { for (uint32_t id : v[u]) ids[t++] = id; }
        }
        o[n] = t;
    };
    fill_csr(hg.local_ids, h.off_loc_off, h.off_loc_ids);
    fill_csr(hg.amb_ids, h.off_amb_off, h.off_amb_ids);
    fill_csr(hg.hap_ids, h.off_hap_off, h.off_hap_ids);
    {
        uint64_t* o = (uint64_t*)(S + h.off_gset_off);
        uint32_t* ids = (uint32_t*)(S + h.off_gset_ids);
        uint64_t t = 0;
        for (uint64_t g = 0; g < h.n_gsets; ++g) { o[g] = t; for (uint32_t id : *gsets[g]) ids[t++] = id; }
        o[h.n_gsets] = t;
        memcpy(S + h.off_gset_of, gset_of.data(), n * 4);
    }
    {
        uint64_t* o = (uint64_t*)(S + h.off_cyc_off);
        char* cp = (char*)(S + h.off_cyc_pool);
        uint64_t t = 0;
        for (uint64_t u = 0; u < n; ++u) {
            o[u] = t;
            if (u < hg.cycles.size()) { memcpy(cp + t, hg.cycles[u].data(), hg.cycles[u].size()); t += hg.cycles[u].size(); }
        }
        o[n] = t;
    }
    uint64_t* kmcov = (uint64_t*)(S + h.off_kmcov);
    uint64_t* shared = (uint64_t*)(S + h.off_shared);
    for (uint64_t u = 0; u < n; ++u) {
        kmcov[u] = u < hg.kmcov.size() ? hg.kmcov[u] : 0;
        shared[u] = u < hg.shared.size() ? hg.shared[u] : 0;
    }
    // getMaxKmerCoverage (src/Graph.cpp:825-841) with top_ratio = 0.001 (Correct_Opt default)
    {
        std::vector<double> v(n);
        for (uint64_t u = 0; u < n; ++u) {
            const uint64_t w = kmcov[u];
            const double cov = (double)((w & 0x7fffffffULL) + ((w >> 31) & 0x7fffffffULL));
            v[u] = std::round(cov / (double)(hg.unitigs[u].size() - k + 1));
        }
        std::sort(v.begin(), v.end(), [](double a, double b) { return a > b; });
        h.max_km_cov_graph = n ? (uint64_t)v[(size_t)(n * hg.top_km_cov_ratio)] : 0;
    }
    if (k <= 32) build_table_and_adj<uint64_t>(hg, h, S);
    else build_table_and_adj<rtk_u128>(hg, h, S);
    memcpy(S, &h, sizeof(h));
    return slab;
}

// ------------------------------------------------------------------ index loader
HostGraph load_index(const std::string& fasta, const std::string& rtsk, int k) {
    HostGraph hg;
    hg.k = k;
    read_fasta(fasta, hg.unitigs);
    const uint64_t n = hg.unitigs.size();
    hg.kmcov.assign(n, 0); hg.shared.assign(n, 0);
    hg.global_ids.assign(n, {}); hg.local_ids.assign(n, {}); hg.amb_ids.assign(n, {}); hg.hap_ids.assign(n, {});
    hg.cycles.assign(n, {});
    if (rtsk.empty()) return hg;
    // head k-mer -> unitig.  The .rtsk stores each unitig's head in the orientation the
    // reference holds it in memory; the same unitig may be spelled reverse-complemented in the
    // FASTA we read, so both extremities are indexed.
    std::unordered_map<std::string, std::pair<uint32_t, bool>> head2u;
    head2u.reserve(2 * n);
    auto revcomp = [](const std::string& s) {
        std::string r(s.rbegin(), s.rend());
        for (auto& c : r) c = (c == 'A') ? 'T' : (c == 'C') ? 'G' : (c == 'G') ? 'C' : (c == 'T') ? 'A' : c;
        return r;
    };
    for (uint64_t u = 0; u < n; ++u) {
        const std::string& s = hg.unitigs[u];
        head2u[s.substr(0, k)] = {(uint32_t)u, true};
        const std::string rh = revcomp(s.substr(s.size() - k, k));
        if (!head2u.count(rh)) head2u[rh] = {(uint32_t)u, false};
    }
    const std::vector<unsigned char> buf = slurp(rtsk);
    ByteReader r{buf.data(), buf.data() + buf.size()};
    std::vector<uint32_t> tmp;
    std::vector<bool> seen(n, false);
    while (!r.eof()) {
        const uint64_t l0 = r.u64(), l1 = r.u64();
        std::string head(k, 'A');
        for (int i = 0; i < k; ++i) {
            const uint64_t w = (i < 32) ? l0 : l1;
            head[i] = "ACGT"[(w >> (62 - 2 * (i & 31))) & 3];
        }
        const auto it = head2u.find(head);
        if (it == head2u.end()) throw std::runtime_error("rtsk: head k-mer " + head + " not found in graph");
        const uint32_t u = it->second.first;
        if (!it->second.second) {
            // stored orientation differs from the FASTA spelling: flip our copy so that strand-
            // specific payload (edge masks, ambiguity positions) keeps its meaning
            hg.unitigs[u] = revcomp(hg.unitigs[u]);
        }
        seen[u] = true;
        hg.kmcov[u] = r.u64();
        hg.shared[u] = r.u64();
        parse_pairid(r, hg.global_ids[u]);
        parse_pairid(r, hg.local_ids[u]);
        parse_pairid(r, hg.amb_ids[u]);
        parse_pairid(r, hg.hap_ids[u]);
        const uint64_t cl = r.u64();
        if (cl) { const unsigned char* c = r.bytes(cl); hg.cycles[u].assign((const char*)c, cl); }
    }
    return hg;
}


// ------------------------------------------------------------------ .rtsk writer for the annotation fields
// PairID::write (src/PairID.cpp:1135-1172) of a sorted id list, in container kinds PairID::read accepts: the 61-bit vector
// (empty set included), the single id, or a CRoaring portable blob of array / bitset containers without runs
// (Bifrost/src/roaring.c:10554-10700).  The reference's behaviour depends on the set only, not on the container kind.
static void write_pairid(std::string& out, const uint32_t* ids, uint64_t n) {
    auto put = [&](const void* p, size_t b) { out.append((const char*)p, b); };
    if (n == 0 || ids[n - 1] < 61) {
        uint64_t w = 1;
        for (uint64_t i = 0; i < n; ++i) w |= 1ULL << (ids[i] + 3);
        put(&w, 8);
        return;
    }
    if (n == 1) { const uint64_t w = ((uint64_t)ids[0] << 3) | 2ULL; put(&w, 8); return; }
    std::vector<std::pair<uint16_t, std::pair<uint64_t, uint64_t>>> cont;   // key, [begin, end)
    for (uint64_t i = 0; i < n;) {
        uint64_t j = i;
        while (j < n && (ids[j] >> 16) == (ids[i] >> 16)) ++j;
        cont.push_back({(uint16_t)(ids[i] >> 16), {i, j}});
        i = j;
    }
    std::string blob;
    auto bput = [&](const void* p, size_t b) { blob.append((const char*)p, b); };
    const uint32_t cookie = 12346, nc = (uint32_t)cont.size();
    bput(&cookie, 4); bput(&nc, 4);
    for (const auto& c : cont) { const uint16_t card1 = (uint16_t)(c.second.second - c.second.first - 1); bput(&c.first, 2); bput(&card1, 2); }
    uint32_t off = 8 + 8 * nc;
    for (const auto& c : cont) { bput(&off, 4); const uint64_t card = c.second.second - c.second.first; off += (card > 4096) ? 8192u : (uint32_t)(2 * card); }
    for (const auto& c : cont) {
        const uint64_t card = c.second.second - c.second.first;
        if (card > 4096) {
            std::vector<uint64_t> bits(1024, 0);
            for (uint64_t i = c.second.first; i < c.second.second; ++i) bits[(ids[i] & 0xffff) >> 6] |= 1ULL << (ids[i] & 63);
            bput(bits.data(), 8192);
        } else {
            for (uint64_t i = c.second.first; i < c.second.second; ++i) { const uint16_t v = (uint16_t)(ids[i] & 0xffff); bput(&v, 2); }
        }
    }
    const uint64_t w = ((uint64_t)blob.size() << 3) | 3ULL;
    put(&w, 8);
    out += blob;
}

template <typename KT>
static uint32_t unitig_of_head(const rtk_graph_view& g, const uint64_t l0, const uint64_t l1) {
    const int k = (int)g.k;
    KT fw;
    if (sizeof(KT) == 8) fw = (KT)(l0 >> (64 - 2 * k));
    else fw = (KT)(((((rtk_u128)l0) << 64) | (rtk_u128)l1) >> (128 - 2 * k));
    rtk_kmer_hit h;
    if (!rtk_lookup<KT>(g.table, g.n_buckets, g.pool, k, fw, KmerOps<KT>::rc(fw, k), h)) throw std::runtime_error("rtsk: head k-mer not found in graph");
    const uint32_t u = rtk_unitig_of(g.blk2unitig, g.unitig_off, h.P);
    if (!h.strand || h.P != g.unitig_off[u]) throw std::runtime_error("rtsk: record does not start a unitig of this graph (index files of another graph?)");
    return u;
}

// Copy of rtsk_in with, per unitig, the fields detectSNPs / detectShortCycles own replaced: bit 8 of shared_pids (isShortCycle),
// the ambiguity PairID and the compactedCycles blob (UnitigData::write, src/UnitigData.hpp:493-517).  Everything else is copied
// byte for byte.  g must be the graph loaded from rtsk_in (unitig ids and orientation).
void patch_rtsk_annotations(const rtk_graph_view& g, const std::string& rtsk_in, const std::string& rtsk_out, const uint64_t* amb_off,
                            const uint32_t* amb_ids, const uint8_t* is_cycle, const uint64_t* cyc_off, const char* cyc_pool) {
    const std::vector<unsigned char> buf = slurp(rtsk_in);
    ByteReader r{buf.data(), buf.data() + buf.size()};
    std::string out;
    out.reserve(buf.size());
    std::vector<uint32_t> tmp;
    std::vector<bool> seen(g.n_unitigs, false);
    while (!r.eof()) {
        const unsigned char* rec = r.p;
        const uint64_t l0 = r.u64(), l1 = r.u64();
        const uint32_t u = (g.k <= 32) ? unitig_of_head<uint64_t>(g, l0, l1) : unitig_of_head<rtk_u128>(g, l0, l1);
        if (seen[u]) throw std::runtime_error("rtsk: unitig listed twice");
        seen[u] = true;
        const uint64_t kmcov = r.u64();
        uint64_t shared = r.u64();
        shared = (shared & ~0x100ULL) | (is_cycle[u] ? 0x100ULL : 0ULL);
        out.append((const char*)rec, 16);
        out.append((const char*)&kmcov, 8);
        out.append((const char*)&shared, 8);
        const unsigned char* c0 = r.p;
        parse_pairid(r, tmp);   // global set
        parse_pairid(r, tmp);   // local set
        out.append((const char*)c0, (size_t)(r.p - c0));
        parse_pairid(r, tmp);   // stored ambiguity ids: dropped
        write_pairid(out, amb_ids + amb_off[u], amb_off[u + 1] - amb_off[u]);
        const unsigned char* h0 = r.p;
        parse_pairid(r, tmp);   // hap ids
        out.append((const char*)h0, (size_t)(r.p - h0));
        const uint64_t old_cl = r.u64();
        if (old_cl) r.bytes(old_cl);
        const uint64_t cl = cyc_off[u + 1] - cyc_off[u];
        out.append((const char*)&cl, 8);
        if (cl) out.append(cyc_pool + cyc_off[u], cl);
    }
    const std::string tmp_path = rtsk_out + ".tmp";
    FILE* f = fopen(tmp_path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write " + tmp_path);
    const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    if (fclose(f) != 0 || !ok) { remove(tmp_path.c_str()); throw std::runtime_error("short write to " + tmp_path); }
    if (rename(tmp_path.c_str(), rtsk_out.c_str()) != 0) { remove(tmp_path.c_str()); throw std::runtime_error("cannot rename to " + rtsk_out); }
}


// The graph of a slab with other per-unitig words and colours (one local set per unitig, no annotations): what addCoverage
// leaves before detectSNPs / detectShortCycles run.
HostGraph recolored_graph(const rtk_graph_view& g, const uint64_t* kmcov, const uint64_t* shared, const uint64_t* col_off, const uint32_t* col_ids) {
    HostGraph hg;
    hg.k = (int)g.k;
    const uint64_t n = g.n_unitigs;
    hg.unitigs.resize(n);
    hg.kmcov.assign(kmcov, kmcov + n);
    hg.shared.assign(shared, shared + n);
    hg.global_ids.assign(n, {}); hg.local_ids.assign(n, {}); hg.amb_ids.assign(n, {}); hg.hap_ids.assign(n, {});
    hg.cycles.assign(n, {});
    for (uint64_t u = 0; u < n; ++u) {
        const uint64_t ub = g.unitig_off[u], len = g.unitig_off[u + 1] - ub;
        std::string& s = hg.unitigs[u];
        s.resize(len);
        for (uint64_t i = 0; i < len; ++i) s[i] = "ACGT"[rtk_pool_base(g.pool, ub + i)];
        hg.local_ids[u].assign(col_ids + col_off[u], col_ids + col_off[u + 1]);
    }
    return hg;
}

// writeGraphData (src/Graph.cpp:786-801) for a whole graph: per unitig its head k-mer (Kmer::write: MAX_KMER_SIZE = 64 -> two
// words, first base in the top bits of word 0) and UnitigData::write (src/UnitigData.hpp:493-517) - the words and colours of the
// slab, the given annotations, no haplotype ids.
void write_rtsk(const rtk_graph_view& g, const std::string& path, const uint64_t* amb_off, const uint32_t* amb_ids, const uint8_t* is_cycle,
                const uint64_t* cyc_off, const char* cyc_pool) {
    std::string out;
    const uint32_t k = g.k;
    for (uint64_t u = 0; u < g.n_unitigs; ++u) {
        uint64_t l[2] = {0, 0};
        const uint64_t ub = g.unitig_off[u];
        for (uint32_t i = 0; i < k; ++i) l[i >> 5] |= (uint64_t)rtk_pool_base(g.pool, ub + i) << (62 - 2 * (i & 31));
        out.append((const char*)l, 16);
        const uint64_t kmcov = g.kmcov[u] & ~(1ULL << 62);   // the visit mark is scratch
        const uint64_t shared = (g.shared[u] & 0xffULL) | (is_cycle[u] ? 0x100ULL : 0ULL);
        out.append((const char*)&kmcov, 8);
        out.append((const char*)&shared, 8);
        const uint32_t gs = g.gset_of[u];
        if (gs == RTK_NONE32) write_pairid(out, nullptr, 0);
        else write_pairid(out, g.gset_ids + g.gset_off[gs], g.gset_off[gs + 1] - g.gset_off[gs]);
        write_pairid(out, g.loc_ids + g.loc_off[u], g.loc_off[u + 1] - g.loc_off[u]);
        write_pairid(out, amb_ids + amb_off[u], amb_off[u + 1] - amb_off[u]);
        write_pairid(out, nullptr, 0);
        const uint64_t cl = cyc_off[u + 1] - cyc_off[u];
        out.append((const char*)&cl, 8);
        if (cl) out.append(cyc_pool + cyc_off[u], cl);
    }
    const std::string tmp_path = path + ".tmp";
    FILE* f = fopen(tmp_path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write " + tmp_path);
    const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    if (fclose(f) != 0 || !ok) { remove(tmp_path.c_str()); throw std::runtime_error("short write to " + tmp_path); }
    if (rename(tmp_path.c_str(), path.c_str()) != 0) { remove(tmp_path.c_str()); throw std::runtime_error("cannot rename to " + path); }
}

}  // namespace rtk
