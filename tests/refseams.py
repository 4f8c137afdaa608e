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
        L.ref_edlib.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_int)),
                                C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.POINTER(C.c_ubyte)), C.POINTER(C.c_int)]
        L.ref_free.argtypes = [C.c_void_p]
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

    def correct_read(self, s, q, pass2=False):
        so, qo = C.c_void_p(), C.c_void_p()
        lib().ref_correct_read(self.h, s.encode(), q.encode(), int(pass2), C.byref(so), C.byref(qo))
        rs = C.string_at(so).decode()
        rq = C.string_at(qo).decode()
        lib().ref_free(so)
        lib().ref_free(qo)
        return rs, rq


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
