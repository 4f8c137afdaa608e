"""ctypes binding of the C ABI in include/rtk.h (librtk_b200.so).

Host-side mirror of the reference's correction entry points, with the reference's names:

    Graph.load(fasta, rtsk, k)          CompactedDBG::read + readGraphData (src/Ratatosk.cpp:1087-1089)
    Context.search_sequence(reads, ...) CompactedDBG::searchSequence       (Bifrost/src/Search.tcc:526)
    Context.get_seeds(reads, opt, pass) getSeeds                           (src/Graph.cpp:3)
    Context.edlib_batch(...)            edlibAlign                         (src/edlib.cpp:141)

There is no CPU implementation behind these calls: the library needs a CUDA device and
`Context()` raises if none is present.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "librtk_b200.so")

SEARCH_EXACT, SEARCH_INS, SEARCH_DEL, SEARCH_SUBST, SEARCH_OR_EXCL = 1, 2, 4, 8, 16


class RtkError(RuntimeError):
    pass


class RtkOpt(C.Structure):
    _fields_ = [("k", C.c_uint32), ("insert_sz", C.c_uint32), ("min_cov_vertices", C.c_uint32),
                ("max_km_cov", C.c_uint32), ("max_len_weak_region1", C.c_uint32),
                ("max_len_weak_region2", C.c_uint32), ("nb_correction_rounds", C.c_uint32),
                ("out_qual", C.c_int32), ("max_qual", C.c_int32), ("trim_qual", C.c_int32),
                ("weak_region_len_factor", C.c_double), ("large_k_factor", C.c_double), ("min_score", C.c_double),
                ("min_confidence_snp_corr", C.c_double), ("force_unres_snp_corr", C.c_uint32),
                ("reserved", C.c_uint32)]


class RtkHit(C.Structure):
    _fields_ = [("pos", C.c_uint32), ("unitig", C.c_uint32), ("dist", C.c_uint32), ("strand", C.c_uint32)]


class RtkGraphInfo(C.Structure):
    _fields_ = [("k", C.c_uint32), ("n_unitigs", C.c_uint64), ("n_kmers", C.c_uint64), ("pool_bases", C.c_uint64),
                ("n_buckets", C.c_uint64), ("n_gsets", C.c_uint64), ("slab_bytes", C.c_uint64),
                ("max_km_cov_graph", C.c_uint64)]


class RtkSeeds(C.Structure):
    _fields_ = [("solid", C.POINTER(RtkHit)), ("solid_off", C.POINTER(C.c_uint64)),
                ("weak", C.POINTER(RtkHit)), ("weak_off", C.POINTER(C.c_uint64))]


class RtkSubgraphCall(C.Structure):
    _fields_ = [("start_unitig", C.c_uint32), ("start_strand", C.c_uint32), ("end_unitig", C.c_uint32),
                ("end_strand", C.c_uint32), ("end_dist", C.c_uint32), ("level", C.c_uint32),
                ("max_len_path", C.c_uint32), ("ref_len", C.c_uint32), ("ref_off", C.c_uint64),
                ("pid_off", C.c_uint64), ("pid_len", C.c_uint32), ("min_cov", C.c_uint32),
                ("max_len_subpath", C.c_uint32), ("reserved", C.c_uint32)]


class RtkPathNode(C.Structure):
    _fields_ = [("unitig", C.c_uint32), ("strand", C.c_uint32), ("dist", C.c_uint32), ("len", C.c_uint32)]


class RtkSubgraphOut(C.Structure):
    _fields_ = [("scores", C.POINTER(C.c_double)), ("path_off", C.POINTER(C.c_uint64)),
                ("n_terminal", C.POINTER(C.c_uint32)), ("node_off", C.POINTER(C.c_uint64)),
                ("nodes", C.POINTER(RtkPathNode)), ("path_len", C.POINTER(C.c_uint32)),
                ("path_ed", C.POINTER(C.c_int32))]


class RtkRegionCall(C.Structure):
    _fields_ = [("win_off", C.c_uint64), ("weak_off", C.c_uint64), ("pid_off", C.c_uint64), ("win_len", C.c_uint32),
                ("n_weak", C.c_uint32), ("pid_len", C.c_uint32), ("start_pos", C.c_uint32), ("start_unitig", C.c_uint32),
                ("start_dist", C.c_uint32), ("start_strand", C.c_uint32), ("has_end", C.c_uint32), ("end_pos", C.c_uint32),
                ("end_unitig", C.c_uint32), ("end_dist", C.c_uint32), ("end_strand", C.c_uint32), ("s_len", C.c_uint32),
                ("reserved", C.c_uint32)]


class RtkRegionResult(C.Structure):
    _fields_ = [("status", C.c_uint32), ("bail", C.c_uint32), ("n_nodes", C.c_uint32), ("len", C.c_uint32),
                ("node_off", C.c_uint64), ("str_off", C.c_uint64), ("n_hops", C.c_uint32), ("n_pops", C.c_uint32),
                ("n_cands", C.c_uint32), ("n_aligns", C.c_uint32), ("seg_off", C.c_uint64), ("n_segs", C.c_uint32),
                ("reserved", C.c_uint32)]


class RtkRegionSeg(C.Structure):
    _fields_ = [("status", C.c_uint32), ("start_weak", C.c_uint32), ("n_nodes", C.c_uint32), ("len", C.c_uint32),
                ("node_off", C.c_uint64), ("str_off", C.c_uint64), ("shw_dist", C.c_int32), ("shw_first_end", C.c_int32)]


class RtkRegionOut(C.Structure):
    _fields_ = [("results", C.POINTER(RtkRegionResult)), ("nodes", C.POINTER(RtkPathNode)), ("chars", C.c_void_p),
                ("segs", C.POINTER(RtkRegionSeg)), ("n_nodes", C.c_uint64), ("n_chars", C.c_uint64), ("n_segs", C.c_uint64)]


HIT_DTYPE = np.dtype([("pos", "<u4"), ("unitig", "<u4"), ("dist", "<u4"), ("strand", "<u4")])

_libs = {}


def load_library(path=None):
    """dlopen the C-ABI library (cached per path). Raises if the built .so is missing."""
    path = os.path.abspath(path or os.environ.get("RTK_LIB", DEFAULT_LIB))
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise RtkError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no fallback implementation)" % path)
    L = C.CDLL(path)
    L.rtk_version.restype = C.c_int
    L.rtk_last_error.restype = C.c_char_p
    L.rtk_free.argtypes = [C.c_void_p]
    L.rtk_opt_default.argtypes = [C.POINTER(RtkOpt), C.c_int]
    L.rtk_graph_load.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    L.rtk_graph_from_unitigs.argtypes = [C.c_int, C.c_uint64, C.POINTER(C.c_char_p), C.POINTER(C.c_void_p)]
    L.rtk_graph_free.argtypes = [C.c_void_p]
    L.rtk_graph_get_info.argtypes = [C.c_void_p, C.POINTER(RtkGraphInfo)]
    L.rtk_graph_slab.restype = C.c_void_p
    L.rtk_graph_slab.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.rtk_graph_save.argtypes = [C.c_void_p, C.c_char_p]
    L.rtk_graph_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    L.rtk_graph_load_cached.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
    L.rtk_graph_unitig_seq.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.rtk_graph_unitig_words.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                         C.POINTER(C.c_uint32 * 8)]
    L.rtk_graph_unitig_colors.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.POINTER(C.c_uint32)),
                                          C.POINTER(C.c_uint64), C.POINTER(C.POINTER(C.c_uint32)),
                                          C.POINTER(C.c_uint64)]
    L.rtk_graph_unitig_annotations.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_uint64),
                                               C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.rtk_detect_snps.argtypes = [C.c_void_p, C.POINTER(RtkOpt), C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.POINTER(C.c_uint32)),
                                  C.POINTER(C.c_uint64)]
    L.rtk_detect_short_cycles.argtypes = [C.c_void_p, C.POINTER(RtkOpt), C.POINTER(C.POINTER(C.c_uint8)),
                                          C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    u64pp, u32pp = C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.POINTER(C.c_uint32))
    L.rtk_color_long_reads.argtypes = [C.c_void_p, C.POINTER(RtkOpt), C.c_uint32, C.c_char_p, C.POINTER(C.c_uint64), C.c_char_p,
                                       C.POINTER(C.c_uint64), C.c_char_p, C.POINTER(C.c_uint64), C.c_uint32, C.c_double,
                                       u64pp, u64pp, u64pp, u32pp, u32pp, C.POINTER(C.c_uint64)]
    u64p__, u32p__ = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
    L.rtk_graph_recolor.argtypes = [C.c_void_p, u64p__, u64p__, u64p__, u32p__, C.POINTER(C.c_void_p)]
    L.rtk_rtsk_write.argtypes = [C.c_void_p, C.c_char_p, u64p__, u32p__, C.POINTER(C.c_uint8), u64p__, C.c_char_p]
    L.rtk_rtsk_write_annotations.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, u64p__, u32p__, C.POINTER(C.c_uint8), u64p__, C.c_char_p]
    L.rtk_detect_snps_range.argtypes = [C.c_void_p, C.POINTER(RtkOpt), C.c_uint64, C.c_uint64, C.POINTER(C.POINTER(C.c_uint64)),
                                        C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_uint64)]
    L.rtk_detect_short_cycles_range.argtypes = [C.c_void_p, C.POINTER(RtkOpt), C.c_uint64, C.c_uint64, C.POINTER(C.POINTER(C.c_uint8)),
                                                C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.rtk_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.rtk_ctx_destroy.argtypes = [C.c_void_p]
    L.rtk_graph_upload.argtypes = [C.c_void_p, C.c_void_p]
    L.rtk_ctx_sync.argtypes = [C.c_void_p]
    L.rtk_search_sequence.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.POINTER(C.c_uint64), C.c_uint32,
                                      C.POINTER(C.POINTER(RtkHit)), C.POINTER(C.POINTER(C.c_uint64)),
                                      C.POINTER(C.c_uint64)]
    L.rtk_get_seeds.argtypes = [C.c_void_p, C.POINTER(RtkOpt), C.c_int, C.c_uint32, C.c_char_p,
                                C.POINTER(C.c_uint64), C.POINTER(RtkSeeds), C.POINTER(C.c_uint64)]
    L.rtk_seeds_free.argtypes = [C.POINTER(RtkSeeds)]
    L.rtk_explore_subgraph_batch.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(RtkSubgraphCall), C.c_char_p, C.c_uint64,
                                             C.POINTER(C.c_uint32), C.c_uint64, C.c_double,
                                             C.POINTER(RtkSubgraphOut), C.POINTER(C.c_uint64)]
    L.rtk_subgraph_out_free.argtypes = [C.POINTER(RtkSubgraphOut)]
    L.rtk_correct_batch.argtypes = [C.c_void_p, C.POINTER(RtkOpt), C.c_int, C.c_uint32, C.c_char_p, C.POINTER(C.c_uint64),
                                    C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                    C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.c_uint64)]
    if hasattr(L, "rtk_correct_batch_resident"):   # product library only (the CPU simulator has no device memory)
        L.rtk_correct_batch_resident.argtypes = [C.c_void_p, C.POINTER(RtkOpt), C.c_int, C.c_uint32, C.c_char_p,
                                                 C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p, C.c_char_p,
                                                 C.POINTER(C.c_uint64), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                                 C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.c_uint64)]
    L.rtk_fix_snps_batch.argtypes = [C.c_void_p, C.POINTER(RtkOpt), C.c_uint32, C.c_char_p, C.POINTER(C.c_uint64),
                                     C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    u64p_ = C.POINTER(C.c_uint64)
    L.rtk_correct_two_pass_batch.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(RtkOpt), C.POINTER(RtkOpt), C.c_uint32, C.c_char_p, u64p_, C.c_char_p, u64p_,
                                             C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(u64p_), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                             C.POINTER(u64p_), u64p_, u64p_, u64p_]
    L.rtk_ctx_resident_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64]
    L.rtk_phasing_batch.argtypes = [C.c_void_p, C.POINTER(RtkOpt), C.c_uint32, C.c_char_p, C.POINTER(C.c_uint64), C.c_char_p,
                                    C.POINTER(C.c_uint64), C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(C.c_void_p),
                                    C.POINTER(C.c_void_p), C.POINTER(C.POINTER(C.c_uint64))]
    L.rtk_explore_paths.argtypes = [C.c_void_p, C.POINTER(RtkOpt), C.POINTER(RtkHit), C.POINTER(RtkHit), C.c_char_p, C.c_uint32,
                                    C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.POINTER(RtkPathNode)), C.POINTER(C.c_uint32),
                                    C.POINTER(C.c_char_p), C.POINTER(C.c_uint32)]
    L.rtk_region_paths_batch.argtypes = [C.c_void_p, C.POINTER(RtkOpt), C.c_int, C.c_uint32, C.POINTER(RtkRegionCall), C.c_char_p,
                                         C.c_uint64, C.POINTER(RtkHit), C.c_uint64, C.POINTER(C.c_uint32), C.c_uint64,
                                         C.POINTER(RtkRegionOut), C.POINTER(C.c_uint64)]
    L.rtk_region_out_free.argtypes = [C.POINTER(RtkRegionOut)]
    L.rtk_edlib_path_batch.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.POINTER(C.c_uint64), C.c_char_p,
                                       C.POINTER(C.c_uint64), C.POINTER(C.c_uint8), C.POINTER(C.c_int32),
                                       C.POINTER(C.c_int32), C.POINTER(C.POINTER(C.c_uint8)),
                                       C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)]
    for name, args in (("rtk_graph_adopt_device", [C.c_void_p, C.c_void_p, C.c_uint64]),
                       ("rtk_k1_sweep_device", [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                                C.POINTER(C.c_uint64), C.c_uint32, C.POINTER(C.c_uint64),
                                                C.POINTER(C.c_uint64), C.POINTER(C.c_float)]),
                       ("rtk_edlib_batch", [C.c_void_p, C.c_uint32, C.c_char_p, C.POINTER(C.c_uint64), C.c_char_p,
                                            C.POINTER(C.c_uint64), C.POINTER(C.c_uint8), C.POINTER(C.c_int32),
                                            C.POINTER(C.c_int32), C.POINTER(C.POINTER(C.c_int32)),
                                            C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.c_uint64)])):
        if hasattr(L, name):
            getattr(L, name).argtypes = args
    _libs[path] = L
    return L


def _check(L, rc):
    if rc != 0:
        raise RtkError("rtk error %d: %s" % (rc, (L.rtk_last_error() or b"").decode()))


def default_opt(pass_no=1, lib=None):
    L = load_library(lib)
    o = RtkOpt()
    L.rtk_opt_default(C.byref(o), pass_no)
    return o


def pack_reads(reads):
    """list of str/bytes -> (pool bytes, uint64 offsets) in the layout the C ABI takes."""
    bs = [r.encode() if isinstance(r, str) else bytes(r) for r in reads]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    return b"".join(bs), off


class Graph:
    """Flat graph slab in host memory."""

    def __init__(self, handle, lib=None):
        self.L = load_library(lib)
        self.lib = lib
        self.h = handle

    @classmethod
    def load(cls, fasta, rtsk, k, lib=None):
        L = load_library(lib)
        h = C.c_void_p()
        _check(L, L.rtk_graph_load(fasta.encode(), (rtsk or "").encode() if rtsk else None, k, C.byref(h)))
        return cls(h, lib)

    @classmethod
    def from_unitigs(cls, unitigs, k, lib=None):
        L = load_library(lib)
        arr = (C.c_char_p * len(unitigs))(*[u.encode() if isinstance(u, str) else u for u in unitigs])
        h = C.c_void_p()
        _check(L, L.rtk_graph_from_unitigs(k, len(unitigs), arr, C.byref(h)))
        return cls(h, lib)

    @classmethod
    def open(cls, path, lib=None):
        L = load_library(lib)
        h = C.c_void_p()
        _check(L, L.rtk_graph_open(path.encode(), C.byref(h)))
        return cls(h, lib)

    @classmethod
    def load_cached(cls, fasta, rtsk, k, cache=None, lib=None):
        """rtk_graph_load through the flat cache file (mmap-ed in place when valid) -> (Graph, came_from_cache)"""
        L = load_library(lib)
        h, fc = C.c_void_p(), C.c_int()
        _check(L, L.rtk_graph_load_cached(fasta.encode(), (rtsk or "").encode() if rtsk else None, k, cache.encode() if cache else None,
                                          C.byref(h), C.byref(fc)))
        return cls(h, lib), bool(fc.value)

    def save(self, path):
        _check(self.L, self.L.rtk_graph_save(self.h, path.encode()))

    def info(self):
        i = RtkGraphInfo()
        _check(self.L, self.L.rtk_graph_get_info(self.h, C.byref(i)))
        return {f[0]: getattr(i, f[0]) for f in RtkGraphInfo._fields_}

    def slab(self):
        """numpy uint8 view of the slab (no copy)."""
        n = C.c_uint64()
        p = self.L.rtk_graph_slab(self.h, C.byref(n))
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n.value,))

    def unitig_seq(self, u):
        n = C.c_uint64()
        _check(self.L, self.L.rtk_graph_unitig_seq(self.h, u, None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        _check(self.L, self.L.rtk_graph_unitig_seq(self.h, u, buf, n.value, C.byref(n)))
        return buf.raw[:n.value].decode()

    def unitig_words(self, u):
        a, b = C.c_uint64(), C.c_uint64()
        adj = (C.c_uint32 * 8)()
        _check(self.L, self.L.rtk_graph_unitig_words(self.h, u, C.byref(a), C.byref(b), C.byref(adj)))
        return a.value, b.value, list(adj)

    def unitig_colors(self, u):
        pg, pl = C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)()
        ng, nl = C.c_uint64(), C.c_uint64()
        _check(self.L, self.L.rtk_graph_unitig_colors(self.h, u, C.byref(pg), C.byref(ng), C.byref(pl), C.byref(nl)))
        return [pg[i] for i in range(ng.value)], [pl[i] for i in range(nl.value)]

    def recolor(self, kmcov, shared, col_off, col_ids):
        """new Graph: same unitigs, the given words and colours (rtk_color_long_reads output), no annotations"""
        u64p, u32p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
        a = [np.ascontiguousarray(kmcov, dtype=np.uint64), np.ascontiguousarray(shared, dtype=np.uint64),
             np.ascontiguousarray(col_off, dtype=np.uint64), np.ascontiguousarray(np.append(np.asarray(col_ids, dtype=np.uint32), np.uint32(0)))]
        h = C.c_void_p()
        _check(self.L, self.L.rtk_graph_recolor(self.h, a[0].ctypes.data_as(u64p), a[1].ctypes.data_as(u64p), a[2].ctypes.data_as(u64p),
                                                a[3].ctypes.data_as(u32p), C.byref(h)))
        return Graph(h, self.lib)

    def write_rtsk(self, path, amb_off, amb_ids, is_cycle, cyc_off, cyc_pool):
        """writeGraphData: the words and colours of this graph + the given annotations -> .rtsk"""
        u64p, u32p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
        a = [np.ascontiguousarray(amb_off, dtype=np.uint64), np.ascontiguousarray(np.append(np.asarray(amb_ids, dtype=np.uint32), np.uint32(0))),
             np.ascontiguousarray(np.append(np.asarray(is_cycle, dtype=np.uint8), np.uint8(0))), np.ascontiguousarray(cyc_off, dtype=np.uint64)]
        _check(self.L, self.L.rtk_rtsk_write(self.h, path.encode(), a[0].ctypes.data_as(u64p), a[1].ctypes.data_as(u32p),
                                             a[2].ctypes.data_as(C.POINTER(C.c_uint8)), a[3].ctypes.data_as(u64p), bytes(cyc_pool) + b"\0"))

    def unitig_annotations(self, u):
        """what the index stores for unitig u: (ambiguity ids [(pos << 4) | base set], compacted-cycles blob bytes)"""
        pa, na, pc, nc = C.POINTER(C.c_uint32)(), C.c_uint64(), C.c_void_p(), C.c_uint64()
        _check(self.L, self.L.rtk_graph_unitig_annotations(self.h, u, C.byref(pa), C.byref(na), C.byref(pc), C.byref(nc)))
        return [pa[i] for i in range(na.value)], (C.string_at(pc, nc.value) if nc.value else b"")

    def close(self):
        if self.h:
            self.L.rtk_graph_free(self.h)
            self.h = None


class Context:
    """One GPU: stream, resident graph, scratch."""

    def __init__(self, device=0, lib=None):
        self.L = load_library(lib)
        self.h = C.c_void_p()
        _check(self.L, self.L.rtk_ctx_create(device, C.byref(self.h)))
        self.graph = None

    def upload(self, graph):
        _check(self.L, self.L.rtk_graph_upload(self.h, graph.h))
        self.graph = graph

    def adopt_device_slab(self, dev_ptr, nbytes):
        _check(self.L, self.L.rtk_graph_adopt_device(self.h, C.c_void_p(dev_ptr), nbytes))

    def sync(self):
        _check(self.L, self.L.rtk_ctx_sync(self.h))

    def _split(self, ptr, off, n):
        offs = np.ctypeslib.as_array(off, shape=(n + 1,)).copy()
        total = int(offs[-1])
        if total:
            raw = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint32)), shape=(total, 4)).copy()
        else:
            raw = np.zeros((0, 4), dtype=np.uint32)
        return [raw[int(offs[i]):int(offs[i + 1])] for i in range(n)]

    def search_sequence(self, reads, exact=True, insertion=False, deletion=False, substitution=False,
                        or_exclusive_match=False, stats=None):
        """per read: uint32 array [n,4] of (pos, unitig, dist, strand) in the reference's order"""
        pool, off = pack_reads(reads)
        flags = (SEARCH_EXACT if exact else 0) | (SEARCH_INS if insertion else 0) | (SEARCH_DEL if deletion else 0) \
            | (SEARCH_SUBST if substitution else 0) | (SEARCH_OR_EXCL if or_exclusive_match else 0)
        ph, po = C.POINTER(RtkHit)(), C.POINTER(C.c_uint64)()
        st = (C.c_uint64 * 8)()
        _check(self.L, self.L.rtk_search_sequence(self.h, len(reads), pool, off.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                  flags, C.byref(ph), C.byref(po), st))
        out = self._split(ph, po, len(reads))
        self.L.rtk_free(C.cast(ph, C.c_void_p))
        self.L.rtk_free(C.cast(po, C.c_void_p))
        if stats is not None:
            stats.extend(list(st))
        return out

    def get_seeds(self, reads, opt=None, pass_no=1, stats=None):
        """-> (solid, weak): per read uint32 arrays [n,4] of (pos, unitig, dist, strand)"""
        opt = opt or default_opt(pass_no)
        pool, off = pack_reads(reads)
        s = RtkSeeds()
        st = (C.c_uint64 * 8)()
        _check(self.L, self.L.rtk_get_seeds(self.h, C.byref(opt), pass_no, len(reads), pool,
                                            off.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(s), st))
        solid = self._split(s.solid, s.solid_off, len(reads))
        weak = self._split(s.weak, s.weak_off, len(reads))
        self.L.rtk_seeds_free(C.byref(s))
        if stats is not None:
            stats.extend(list(st))
        return solid, weak

    def edlib_batch(self, queries, targets, modes, kmax=None, stats=None):
        """Myers edit distance for pairs; modes: 0 NW, 1 SHW, 2 HW -> (dist int32[n], list of end-location arrays)"""
        n = len(queries)
        qp, qo = pack_reads(queries)
        tp, to = pack_reads(targets)
        m = np.asarray(modes, dtype=np.uint8)
        km = np.full(n, -1, dtype=np.int32) if kmax is None else np.asarray(kmax, dtype=np.int32)
        dist = np.zeros(n, dtype=np.int32)
        pe, po = C.POINTER(C.c_int32)(), C.POINTER(C.c_uint64)()
        st = (C.c_uint64 * 8)()
        _check(self.L, self.L.rtk_edlib_batch(self.h, n, qp, qo.ctypes.data_as(C.POINTER(C.c_uint64)), tp,
                                              to.ctypes.data_as(C.POINTER(C.c_uint64)),
                                              m.ctypes.data_as(C.POINTER(C.c_uint8)),
                                              km.ctypes.data_as(C.POINTER(C.c_int32)),
                                              dist.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(pe), C.byref(po), st))
        offs = np.ctypeslib.as_array(po, shape=(n + 1,)).copy()
        total = int(offs[-1])
        ends = np.ctypeslib.as_array(pe, shape=(max(total, 1),)).copy()[:total]
        self.L.rtk_free(C.cast(pe, C.c_void_p))
        self.L.rtk_free(C.cast(po, C.c_void_p))
        if stats is not None:
            stats.extend(list(st))
        return dist, [ends[int(offs[i]):int(offs[i + 1])] for i in range(n)]

    def correct(self, reads, quals=None, opt=None, pass_no=1, stats=None):
        """getSeeds + correctSequence for a batch (the per-read body of the reference's search()):
        -> list of (corrected sequence, quality string)"""
        opt = opt or default_opt(pass_no)
        pool, off = pack_reads(reads)
        if quals is not None:
            qpool, qoff = pack_reads(quals)
            qp, qo = qpool, qoff.ctypes.data_as(C.POINTER(C.c_uint64))
        else:
            qp, qo = None, None
        os_, oq_ = C.c_void_p(), C.c_void_p()
        oo = C.POINTER(C.c_uint64)()
        st = (C.c_uint64 * 24)()
        _check(self.L, self.L.rtk_correct_batch(self.h, C.byref(opt), pass_no, len(reads), pool, off.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                qp, qo, C.byref(os_), C.byref(oq_), C.byref(oo), st))
        n = len(reads)
        offs = [oo[i] for i in range(n + 1)]
        sbuf = C.string_at(os_, offs[-1])
        qbuf = C.string_at(oq_, offs[-1])
        out = [(sbuf[offs[i]:offs[i + 1]].decode("latin1"), qbuf[offs[i]:offs[i + 1]].decode("latin1")) for i in range(n)]
        self.L.rtk_free(os_)
        self.L.rtk_free(oq_)
        self.L.rtk_free(C.cast(oo, C.c_void_p))
        if stats is not None:
            stats.extend(list(st))
        return out

    def correct_two_pass(self, ctx2, reads, quals, opt1=None, opt2=None, want_pass1=False):
        """pass 1 on this context (k1 graph) + phasing + pass 2 on ctx2 (k2 graph) as one pipeline (rtk_correct_two_pass_batch)
        -> list of (sequence, quality) [, the same for the pass-1 output]"""
        opt1, opt2 = opt1 or default_opt(1), opt2 or default_opt(2)
        pool, off = pack_reads(reads)
        qpool, qoff = pack_reads(quals)
        u64p = C.POINTER(C.c_uint64)
        os_, oq_, oo = C.c_void_p(), C.c_void_p(), u64p()
        ps_, pq_, po = C.c_void_p(), C.c_void_p(), u64p()
        _check(self.L, self.L.rtk_correct_two_pass_batch(self.h, ctx2.h, C.byref(opt1), C.byref(opt2), len(reads), pool, off.ctypes.data_as(u64p), qpool,
                                                         qoff.ctypes.data_as(u64p), C.byref(os_), C.byref(oq_), C.byref(oo),
                                                         C.byref(ps_) if want_pass1 else None, C.byref(pq_) if want_pass1 else None,
                                                         C.byref(po) if want_pass1 else None, None, None, None))

        def take(sp, qp, op):
            n = len(reads)
            offs = [op[i] for i in range(n + 1)]
            sbuf, qbuf = C.string_at(sp, offs[-1]), C.string_at(qp, offs[-1])
            out = [(sbuf[offs[i]:offs[i + 1]].decode("latin1"), qbuf[offs[i]:offs[i + 1]].decode("latin1")) for i in range(n)]
            self.L.rtk_free(sp); self.L.rtk_free(qp); self.L.rtk_free(C.cast(op, C.c_void_p))
            return out
        fin = take(os_, oq_, oo)
        return (fin, take(ps_, pq_, po)) if want_pass1 else fin

    def fix_snps(self, reads, opt=None):
        """fixSNPs (src/Alignment.cpp:846) for a batch of pass-1 reads on the k = 63 graph -> (list of reads, codes replaced)"""
        opt = opt or default_opt(2)
        pool, off = pack_reads(reads)
        out, nf = C.c_void_p(), C.c_uint64()
        _check(self.L, self.L.rtk_fix_snps_batch(self.h, C.byref(opt), len(reads), pool, off.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                 C.byref(out), C.byref(nf)))
        buf = C.string_at(out, int(off[-1]))
        self.L.rtk_free(out)
        return [buf[int(off[i]):int(off[i + 1])].decode("latin1") for i in range(len(reads))], nf.value

    def color_long_reads(self, reads, quals=None, names=None, opt=None, min_len=3000, min_conf=0.0, stats=None):
        """addCoverage for long reads (`Ratatosk index -2`) -> (kmcov u64[n], shared u64[n], col_off u64[n+1], col_ids u32[], read ids)"""
        u64p = C.POINTER(C.c_uint64)
        sp, so = pack_reads(reads)
        qp, qo = pack_reads(quals) if quals is not None else (None, None)
        np_, no = pack_reads(names) if names is not None else (None, None)
        pk, ps, po, pi, pr = u64p(), u64p(), u64p(), C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)()
        st = (C.c_uint64 * 10)()
        rc = self.L.rtk_color_long_reads(self.h, C.byref(opt) if opt else None, len(reads), sp, so.ctypes.data_as(u64p),
                                         qp, qo.ctypes.data_as(u64p) if qo is not None else None,
                                         np_, no.ctypes.data_as(u64p) if no is not None else None, min_len, min_conf,
                                         C.byref(pk), C.byref(ps), C.byref(po), C.byref(pi), C.byref(pr), st)
        if stats is not None:
            stats[:] = list(st)
        _check(self.L, rc)
        n = self.graph.info()["n_unitigs"]
        kmcov = np.ctypeslib.as_array(pk, shape=(n + 1,))[:n].copy()
        shared = np.ctypeslib.as_array(ps, shape=(n + 1,))[:n].copy()
        off = np.ctypeslib.as_array(po, shape=(n + 1,)).copy()
        ids = np.ctypeslib.as_array(pi, shape=(int(off[-1]) + 1,))[:int(off[-1])].copy()
        rid = np.ctypeslib.as_array(pr, shape=(len(reads) + 1,))[:len(reads)].copy()
        for p in (pk, ps, po, pi, pr):
            self.L.rtk_free(C.cast(p, C.c_void_p))
        return kmcov, shared, off, ids, rid

    def detect_snps(self, opt=None, stats=None, first=0, count=None):
        """detectSNPs (src/Graph.cpp:484) on the resident graph -> per unitig the sorted ambiguity ids (pos << 4 | base set);
        first / count: only that range of unitigs (the share of one rank)"""
        po, pi = C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint32)()
        st = (C.c_uint64 * 10)()
        _check(self.L, self.L.rtk_detect_snps_range(self.h, C.byref(opt) if opt else None, first, (1 << 64) - 1 if count is None else count,
                                                    C.byref(po), C.byref(pi), st))
        n = self.graph.info()["n_unitigs"]
        off = np.ctypeslib.as_array(po, shape=(n + 1,)).copy()
        ids = np.ctypeslib.as_array(pi, shape=(int(off[-1]) + 1,))[:int(off[-1])].copy()
        self.L.rtk_free(C.cast(po, C.c_void_p)); self.L.rtk_free(C.cast(pi, C.c_void_p))
        if stats is not None:
            stats[:] = list(st)
        return off, ids

    def detect_short_cycles(self, opt=None, stats=None, first=0, count=None):
        """detectShortCycles (src/Graph.cpp:4660) on the resident graph -> (is_cycle u8[n], blob offsets u64[n+1], blob bytes)"""
        pf, po, pp = C.POINTER(C.c_uint8)(), C.POINTER(C.c_uint64)(), C.c_void_p()
        st = (C.c_uint64 * 10)()
        _check(self.L, self.L.rtk_detect_short_cycles_range(self.h, C.byref(opt) if opt else None, first, (1 << 64) - 1 if count is None else count,
                                                            C.byref(pf), C.byref(po), C.byref(pp), st))
        n = self.graph.info()["n_unitigs"]
        flags = np.ctypeslib.as_array(pf, shape=(n + 1,))[:n].copy()
        off = np.ctypeslib.as_array(po, shape=(n + 1,)).copy()
        pool = C.string_at(pp, int(off[-1]))
        self.L.rtk_free(C.cast(pf, C.c_void_p)); self.L.rtk_free(C.cast(po, C.c_void_p)); self.L.rtk_free(pp)
        if stats is not None:
            stats[:] = list(st)
        return flags, off, pool

    def phasing(self, raw_reads, corr_reads, corr_quals, opt=None):
        """phasing() of the second pass (src/Graph.cpp:869) for a batch -> list of (sequence, quality string)"""
        opt = opt or default_opt(2)
        rp, ro = pack_reads(raw_reads)
        cp, co = pack_reads(corr_reads)
        qp, qo = pack_reads(corr_quals)
        os_, oq_ = C.c_void_p(), C.c_void_p()
        oo = C.POINTER(C.c_uint64)()
        u64p = C.POINTER(C.c_uint64)
        _check(self.L, self.L.rtk_phasing_batch(self.h, C.byref(opt), len(raw_reads), rp, ro.ctypes.data_as(u64p), cp,
                                                co.ctypes.data_as(u64p), qp, qo.ctypes.data_as(u64p), C.byref(os_),
                                                C.byref(oq_), C.byref(oo)))
        n = len(raw_reads)
        offs = [oo[i] for i in range(n + 1)]
        sbuf = C.string_at(os_, offs[-1])
        qbuf = C.string_at(oq_, offs[-1])
        out = [(sbuf[offs[i]:offs[i + 1]].decode("latin1"), qbuf[offs[i]:offs[i + 1]].decode("latin1")) for i in range(n)]
        self.L.rtk_free(os_)
        self.L.rtk_free(oq_)
        self.L.rtk_free(C.cast(oo, C.c_void_p))
        return out

    def explore_paths(self, start, end, ref, pids, opt=None):
        """explorePathsBFS2 (end given) / explorePathsBFS (end None); start/end = (unitig, strand, dist) anchors.
        -> None or dict(nodes=[(unitig, strand, dist, len)], qual=str, length=int)"""
        opt = opt or default_opt(1)
        hs = RtkHit(0, start[0], start[2], start[1])
        he = RtkHit(0, end[0], end[2], end[1]) if end is not None else None
        arr = (C.c_uint32 * max(len(pids), 1))(*pids)
        pn = C.POINTER(RtkPathNode)()
        nn, pl = C.c_uint32(), C.c_uint32()
        q = C.c_char_p()
        r = ref.encode()
        _check(self.L, self.L.rtk_explore_paths(self.h, C.byref(opt), C.byref(hs), C.byref(he) if he is not None else None, r,
                                                len(r), arr, len(pids), C.byref(pn), C.byref(nn), C.byref(q), C.byref(pl)))
        if nn.value == 0:
            return None
        res = {"nodes": [(pn[i].unitig, pn[i].strand, pn[i].dist, pn[i].len) for i in range(nn.value)],
               "qual": q.value.decode("latin1"), "length": pl.value}
        self.L.rtk_free(C.cast(pn, C.c_void_p))
        self.L.rtk_free(C.cast(q, C.c_void_p))
        return res

    def region_paths(self, calls, opt=None, pass_no=1, stats=None):
        """Device-resident region engine (extractSemiWeakPaths and everything below it, one warp per call).
        calls: dicts with window (str = s[start_pos, pos2 + k)), start=(pos, unitig, strand, dist),
        end=(pos, unitig, strand, dist) or None (open end: s_len required), weak=[(pos, unitig, strand, dist)], pids (sorted).
        -> per call dict(status, bail, nodes=[(unitig, strand, dist, len)], seq, qual, hops, pops, cands, aligns)"""
        opt = opt or default_opt(pass_no)
        n = len(calls)
        arr = (RtkRegionCall * max(n, 1))()
        wins, weak, pids = [], [], []
        wo = 0
        for i, c in enumerate(calls):
            w = c["window"].encode()
            a = arr[i]
            a.win_off, a.win_len = wo, len(w)
            a.weak_off, a.n_weak = len(weak), len(c.get("weak", []))
            a.pid_off, a.pid_len = len(pids), len(c["pids"])
            a.start_pos, a.start_unitig, a.start_strand, a.start_dist = c["start"]
            a.reserved = 1 if c.get("follow_dead_ends") else 0
            if c.get("end") is None:
                a.has_end, a.s_len = 0, c["s_len"]
            else:
                a.has_end = 1
                a.end_pos, a.end_unitig, a.end_strand, a.end_dist = c["end"]
                a.s_len = c.get("s_len", a.end_pos + opt.k)
            wins.append(w); wo += len(w)
            weak.extend(c.get("weak", []))
            pids.extend(c["pids"])
        win_pool = b"".join(wins)
        weak_arr = (RtkHit * max(len(weak), 1))(*[RtkHit(w[0], w[1], w[3], w[2]) for w in weak])
        pid_arr = (C.c_uint32 * max(len(pids), 1))(*pids)
        out = RtkRegionOut()
        st = (C.c_uint64 * 8)()
        _check(self.L, self.L.rtk_region_paths_batch(self.h, C.byref(opt), pass_no, n, arr, win_pool, len(win_pool), weak_arr, len(weak),
                                                     pid_arr, len(pids), C.byref(out), st))
        chars = C.string_at(out.chars, out.n_chars) if out.n_chars else b""
        res = []
        for i in range(n):
            r = out.results[i]
            d = {"status": r.status, "bail": r.bail, "nodes": [], "seq": "", "qual": "", "hops": r.n_hops, "pops": r.n_pops,
                 "cands": r.n_cands, "aligns": r.n_aligns}
            if r.status != 2:
                d["nodes"] = [(out.nodes[j].unitig, out.nodes[j].strand, out.nodes[j].dist, out.nodes[j].len)
                              for j in range(r.node_off, r.node_off + r.n_nodes)]
                pad = (r.len + 7) & ~7
                d["seq"] = chars[r.str_off:r.str_off + r.len].decode("latin1")
                d["qual"] = chars[r.str_off + pad:r.str_off + pad + r.len].decode("latin1")
                d["segments"] = [{"status": out.segs[j].status, "start_weak": out.segs[j].start_weak, "len": out.segs[j].len,
                                  "shw_dist": out.segs[j].shw_dist, "shw_first_end": out.segs[j].shw_first_end}
                                 for j in range(r.seg_off, r.seg_off + r.n_segs)]
            res.append(d)
        self.L.rtk_region_out_free(C.byref(out))
        if stats is not None:
            stats.extend(list(st))
        return res

    def edlib_path_batch(self, queries, targets, modes, stats=None):
        """edlibAlign with TASK_PATH; modes: 0 NW, 1 SHW -> (dist, end, [ops bytes], flags)"""
        n = len(queries)
        qp, qo = pack_reads(queries)
        tp, to = pack_reads(targets)
        m = np.asarray(modes, dtype=np.uint8)
        dist = np.zeros(n, dtype=np.int32)
        end = np.zeros(n, dtype=np.int32)
        flags = np.zeros(n, dtype=np.uint8)
        po, poff = C.POINTER(C.c_uint8)(), C.POINTER(C.c_uint64)()
        st = (C.c_uint64 * 8)()
        _check(self.L, self.L.rtk_edlib_path_batch(self.h, n, qp, qo.ctypes.data_as(C.POINTER(C.c_uint64)), tp,
                                                   to.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                   m.ctypes.data_as(C.POINTER(C.c_uint8)),
                                                   dist.ctypes.data_as(C.POINTER(C.c_int32)),
                                                   end.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(po), C.byref(poff),
                                                   flags.ctypes.data_as(C.POINTER(C.c_uint8)), st))
        offs = np.ctypeslib.as_array(poff, shape=(n + 1,)).copy()
        total = int(offs[-1])
        raw = np.ctypeslib.as_array(po, shape=(max(total, 1),)).copy()[:total]
        self.L.rtk_free(C.cast(po, C.c_void_p))
        self.L.rtk_free(C.cast(poff, C.c_void_p))
        if stats is not None:
            stats.extend(list(st))
        return dist, end, [raw[int(offs[i]):int(offs[i + 1])].tolist() for i in range(n)], flags

    def explore_subgraph(self, calls, weak_region_len_factor=0.25, stats=None):
        """exploreSubGraph for a batch.  calls: dicts with start=(unitig, strand), end=(unitig, strand, dist) or None,
        ref (str), level, max_len_path, pids (sorted ints), min_cov (default 2).
        -> per call dict(scores=(t1, t2, nt1, nt2), terminal=[path], nonterminal=[path]); path = dict(nodes=[(unitig,
        strand, dist, len)], length, ed)"""
        n = len(calls)
        arr = (RtkSubgraphCall * max(n, 1))()
        refs, pids = [], []
        ro = po = 0
        for i, c in enumerate(calls):
            r = c["ref"].encode()
            a = arr[i]
            a.start_unitig, a.start_strand = c["start"]
            if c.get("end") is None:
                a.end_unitig, a.end_strand, a.end_dist = 0xFFFFFFFF, 0, 0
            else:
                a.end_unitig, a.end_strand, a.end_dist = c["end"]
            a.level, a.max_len_path = c["level"], c["max_len_path"]
            a.ref_off, a.ref_len = ro, len(r)
            a.pid_off, a.pid_len = po, len(c["pids"])
            a.min_cov = c.get("min_cov", 2)
            refs.append(r); ro += len(r)
            pids.extend(c["pids"]); po += len(c["pids"])
        ref_pool = b"".join(refs)
        pid_arr = (C.c_uint32 * max(len(pids), 1))(*pids)
        out = RtkSubgraphOut()
        st = (C.c_uint64 * 8)()
        _check(self.L, self.L.rtk_explore_subgraph_batch(self.h, n, arr, ref_pool, len(ref_pool), pid_arr, len(pids),
                                                         weak_region_len_factor, C.byref(out), st))
        res = []
        for i in range(n):
            paths = []
            for pi in range(out.path_off[i], out.path_off[i + 1]):
                nodes = [(out.nodes[j].unitig, out.nodes[j].strand, out.nodes[j].dist, out.nodes[j].len)
                         for j in range(out.node_off[pi], out.node_off[pi + 1])]
                paths.append({"nodes": nodes, "length": out.path_len[pi], "ed": out.path_ed[pi]})
            nt = out.n_terminal[i]
            res.append({"scores": tuple(out.scores[4 * i + j] for j in range(4)), "terminal": paths[:nt], "nonterminal": paths[nt:]})
        self.L.rtk_subgraph_out_free(C.byref(out))
        if stats is not None:
            stats.extend(list(st))
        return res

    def close(self):
        if self.h:
            self.L.rtk_ctx_destroy(self.h)
            self.h = None
