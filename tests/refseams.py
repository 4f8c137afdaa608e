"""ctypes wrapper over oracle/_ref/libref_seams.so (the UNMODIFIED reference objects).

TEST INFRASTRUCTURE: used by tests/golden/make_golden.py (in the build container,
where /root/reference exists) and by optional differential tests when the library
is present.  Never imported by the product package.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "..", "oracle", "_ref", "libref_seams.so")


class RefHit(C.Structure):
    _fields_ = [("pos", C.c_uint64), ("unitig", C.c_uint64), ("dist", C.c_uint32),
                ("len", C.c_uint32), ("size", C.c_uint32), ("strand", C.c_uint32)]


def available():
    return os.path.exists(LIB)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        L.ref_graph_load.restype = C.c_void_p
        L.ref_graph_load.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
        L.ref_graph_free.argtypes = [C.c_void_p]
        L.ref_graph_num_unitigs.restype = C.c_uint64
        L.ref_graph_num_unitigs.argtypes = [C.c_void_p]
        L.ref_graph_max_km_cov.restype = C.c_uint64
        L.ref_graph_max_km_cov.argtypes = [C.c_void_p]
        L.ref_set_opt.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.ref_graph_dump.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_search_sequence.restype = C.c_int64
        L.ref_search_sequence.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(C.POINTER(RefHit))]
        L.ref_get_seeds.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int,
                                    C.POINTER(C.POINTER(RefHit)), C.POINTER(C.c_int64),
                                    C.POINTER(C.POINTER(RefHit)), C.POINTER(C.c_int64)]
        L.ref_correct_read.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int,
                                       C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.ref_phasing.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.ref_fix_snps.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]
        L.ref_edlib.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_int)),
                                C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.POINTER(C.c_ubyte)), C.POINTER(C.c_int)]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_annotate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p]
        L.ref_explore_subgraph.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.c_int, C.c_uint32, C.c_char_p,
                                           C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32,
                                           C.POINTER(C.c_double), C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_uint64)]
        L.ref_explore_paths.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_uint32, C.c_uint64, C.c_int, C.c_uint32, C.c_char_p,
                                        C.POINTER(C.c_uint32), C.c_uint32, C.c_int, C.POINTER(C.POINTER(C.c_uint32)),
                                        C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def _hits(ptr, n):
    out = [(ptr[i].pos, ptr[i].unitig, ptr[i].dist, ptr[i].len, ptr[i].size, ptr[i].strand) for i in range(n)]
    lib().ref_free(C.cast(ptr, C.c_void_p))
    return out


class RefGraph:
    def __init__(self, fasta, rtsk, k, threads=1):
        self.h = lib().ref_graph_load(fasta.encode(), (rtsk or "").encode(), k, 0, threads)
        if not self.h:
            raise RuntimeError("reference failed to load graph %s" % fasta)
        self.k = k

    def close(self):
        if self.h:
            lib().ref_graph_free(self.h)
            self.h = None

    def num_unitigs(self):
        return lib().ref_graph_num_unitigs(self.h)

    def max_km_cov(self):
        return lib().ref_graph_max_km_cov(self.h)

    def set_opt(self, name, v):
        lib().ref_set_opt(self.h, name.encode(), float(v))

    def dump(self, path):
        if lib().ref_graph_dump(self.h, path.encode()) != 0:
            raise RuntimeError("dump failed")

    def annotate(self, min_cov, threads=1):
        """re-run detectSNPs + detectShortCycles with min_cov_vertices = min_cov -> {unitig sequence: (ambiguity ids, flag, blob)}"""
        import tempfile
        with tempfile.NamedTemporaryFile(suffix=".tsv") as t:
            if lib().ref_annotate(self.h, int(min_cov), int(threads), t.name.encode()) != 0:
                raise RuntimeError("annotate failed")
            out = {}
            for line in open(t.name):
                seq, amb, flag, cyc = line.rstrip("\n").split("\t")
                out[seq] = ([int(x) for x in amb.split(",") if x], int(flag), cyc.replace(";", "\0").encode())
            return out

    def search_sequence(self, s, exact, ins, dele, subst, or_excl):
        p = C.POINTER(RefHit)()
        n = lib().ref_search_sequence(self.h, s.encode(), int(exact), int(ins), int(dele), int(subst), int(or_excl),
                                      C.byref(p))
        return _hits(p, n)

    def get_seeds(self, s, q="", pass2=False):
        ps, pw = C.POINTER(RefHit)(), C.POINTER(RefHit)()
        ns, nw = C.c_int64(), C.c_int64()
        lib().ref_get_seeds(self.h, s.encode(), q.encode(), int(pass2), C.byref(ps), C.byref(ns), C.byref(pw), C.byref(nw))
        return _hits(ps, ns.value), _hits(pw, nw.value)

    def explore_subgraph(self, start_key, start_strand, end_key, end_strand, end_dist, ref, level, max_len_path, pids):
        """-> (score_t1, score_nt1, terminal paths, non-terminal paths); path = ([(key, strand, dist, len)...], qual)"""
        import numpy as np
        arr = (C.c_uint32 * max(1, len(pids)))(*pids)
        sc = (C.c_double * 2)()
        po = C.POINTER(C.c_uint32)()
        nw = C.c_uint64()
        ek = 0xFFFFFFFFFFFFFFFF if end_key is None else end_key
        rc = lib().ref_explore_subgraph(self.h, start_key, int(start_strand), ek, int(end_strand), int(end_dist), ref.encode(),
                                        level, max_len_path, arr, len(pids), sc, C.byref(po), C.byref(nw))
        if rc != 0:
            raise RuntimeError("ref_explore_subgraph failed")
        w = [po[i] for i in range(nw.value)]
        lib().ref_free(C.cast(po, C.c_void_p))
        pos = 0
        groups = []
        for _ in range(2):
            n = w[pos]; pos += 1
            paths = []
            for _p in range(n):
                m = w[pos]; pos += 1
                ums = []
                for _u in range(m):
                    key = w[pos] | (w[pos + 1] << 32)
                    ums.append((key, w[pos + 2], w[pos + 3], w[pos + 4])); pos += 5
                ql = w[pos]; pos += 1
                nq = (ql + 3) // 4
                qb = b"".join(int(x).to_bytes(4, "little") for x in w[pos:pos + nq])[:ql]; pos += nq
                paths.append((ums, qb.decode("latin1")))
            groups.append(paths)
        return sc[0], sc[1], groups[0], groups[1]

    def explore_paths(self, start, end, ref, pids, pass2=False):
        """explorePathsBFS2 (end given) / explorePathsBFS (end None); start/end = (key, strand, dist) -> [(ums, qual)]"""
        arr = (C.c_uint32 * max(1, len(pids)))(*pids)
        po = C.POINTER(C.c_uint32)()
        nw = C.c_uint64()
        ek, es, ed = (0xFFFFFFFFFFFFFFFF, 0, 0) if end is None else end
        rc = lib().ref_explore_paths(self.h, start[0], int(start[1]), int(start[2]), ek, int(es), int(ed), ref.encode(), arr,
                                     len(pids), int(pass2), C.byref(po), C.byref(nw))
        if rc != 0:
            raise RuntimeError("ref_explore_paths failed")
        w = [po[i] for i in range(nw.value)]
        lib().ref_free(C.cast(po, C.c_void_p))
        pos = 0
        n = w[pos]; pos += 1
        paths = []
        for _p in range(n):
            m = w[pos]; pos += 1
            ums = []
            for _u in range(m):
                key = w[pos] | (w[pos + 1] << 32)
                ums.append((key, w[pos + 2], w[pos + 3], w[pos + 4])); pos += 5
            ql = w[pos]; pos += 1
            nq = (ql + 3) // 4
            qb = b"".join(int(x).to_bytes(4, "little") for x in w[pos:pos + nq])[:ql]; pos += nq
            paths.append((ums, qb.decode("latin1")))
        return paths

    def correct_read(self, s, q, pass2=False):
        so, qo = C.c_void_p(), C.c_void_p()
        lib().ref_correct_read(self.h, s.encode(), q.encode(), int(pass2), C.byref(so), C.byref(qo))
        rs = C.string_at(so).decode()
        rq = C.string_at(qo).decode()
        lib().ref_free(so)
        lib().ref_free(qo)
        return rs, rq


def _phasing(self, raw, corr, qual):
    so, qo = C.c_void_p(), C.c_void_p()
    lib().ref_phasing(self.h, raw.encode(), corr.encode(), qual.encode(), C.byref(so), C.byref(qo))
    rs, rq = C.string_at(so).decode(), C.string_at(qo).decode()
    lib().ref_free(so)
    lib().ref_free(qo)
    return rs, rq


RefGraph.phasing = _phasing


def _fix_snps(self, corr):
    so = C.c_void_p()
    lib().ref_fix_snps(self.h, corr.encode(), C.byref(so))
    rs = C.string_at(so).decode()
    lib().ref_free(so)
    return rs


RefGraph.fix_snps = _fix_snps


def edlib(q, t, mode, task=0, k=-1, iupac=True):
    """mode: 0 NW, 1 SHW, 2 HW; task: 0 distance, 1 loc, 2 path -> (dist, ends, starts, alignment bytes)"""
    if isinstance(q, str):
        q = q.encode()
    if isinstance(t, str):
        t = t.encode()
    d, n, al = C.c_int(), C.c_int(), C.c_int()
    pe, ps = C.POINTER(C.c_int)(), C.POINTER(C.c_int)()
    pa = C.POINTER(C.c_ubyte)()
    lib().ref_edlib(q, len(q), t, len(t), mode, task, k, int(iupac), C.byref(d), C.byref(n), C.byref(pe), C.byref(ps),
                    C.byref(pa), C.byref(al))
    ends = [pe[i] for i in range(n.value)] if pe else []
    starts = [ps[i] for i in range(n.value)] if ps else []
    aln = bytes(pa[i] for i in range(al.value)) if pa else b""
    for p in (pe, ps, pa):
        if p:
            lib().ref_free(C.cast(p, C.c_void_p))
    return d.value, ends, starts, aln
