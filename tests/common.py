"""Shared helpers for the test-suite, smoke() and bench.py's checker legs (TEST INFRASTRUCTURE)."""
import ctypes as C
import gzip
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, ".."))
GOLDEN = os.path.join(HERE, "golden")
SIM_LIB = os.path.join(HERE, "hostsim", "_build", "librtk_hostsim.so")
ORACLE_LIB = os.path.join(ROOT, "oracle", "librtk_oracle.so")
PRODUCT_LIB = os.path.join(ROOT, "ratatosk_b200", "librtk_b200.so")


def ensure_built():
    """Build the oracle / simulator / product libraries if a fresh checkout has not yet."""
    if not os.path.exists(ORACLE_LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    if not os.path.exists(SIM_LIB):
        subprocess.check_call(["make", "-C", os.path.join(HERE, "hostsim")])
    if not os.path.exists(PRODUCT_LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "ratatosk_b200", "csrc"), "-j8"])


def read_fastq(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        L = f.read().split("\n")
    return [(L[i][1:], L[i + 1], L[i + 3]) for i in range(0, len(L) - 3, 4)]


def load_golden_reads(recipe):
    return read_fastq(os.path.join(GOLDEN, recipe, "reads.fastq.gz"))


def golden_paths(recipe):
    d = os.path.join(GOLDEN, recipe)
    return os.path.join(d, "index.k31.fasta.gz"), os.path.join(d, "index.k31.rtsk")


def revcomp(s):
    return s[::-1].translate(str.maketrans("ACGTacgt", "TGCAtgca"))


# ------------------------------------------------------------------ oracle binding
_orc = None


def oracle():
    global _orc
    if _orc is None:
        ensure_built()
        L = C.CDLL(ORACLE_LIB)
        L.orc_graph_create.restype = C.c_void_p
        L.orc_graph_create.argtypes = [C.c_int, C.c_uint64, C.POINTER(C.c_char_p)]
        L.orc_graph_free.argtypes = [C.c_void_p]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_search_sequence.restype = C.c_int64
        L.orc_search_sequence.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_uint64)]
        L.orc_edit_distance.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.c_int)]
        _orc = L
    return _orc


class OracleGraph:
    def __init__(self, unitigs, k):
        L = oracle()
        arr = (C.c_char_p * len(unitigs))(*[u.encode() for u in unitigs])
        self.h = L.orc_graph_create(k, len(unitigs), arr)
        self.k = k

    def search(self, s, exact=True, ins=False, dele=False, subst=False, or_excl=False):
        L = oracle()
        p = C.POINTER(C.c_uint32)()
        nl = C.c_uint64()
        n = L.orc_search_sequence(self.h, s.encode(), int(exact), int(ins), int(dele), int(subst), int(or_excl),
                                  C.byref(p), C.byref(nl))
        a = np.ctypeslib.as_array(p, shape=(max(n, 1), 4)).copy()[:n] if n else np.zeros((0, 4), np.uint32)
        if p:
            L.orc_free(C.cast(p, C.c_void_p))
        self.last_lookups = nl.value
        return a.astype(np.uint32)

    def close(self):
        if self.h:
            oracle().orc_graph_free(self.h)
            self.h = None


_ograph_cache = []   # [(graph, OracleGraph)]: the strong reference keeps the Graph alive, so its identity cannot be recycled


def oracle_graph_for(graph):
    """OracleGraph over the unitigs of a loaded product Graph (unitig ids coincide).  Cached per Graph OBJECT: the
    cache holds the graph itself (an `id()` key could be reused by a later Graph once the first one is collected,
    which is what made round 1's F2 parity test compare against the F1 oracle)."""
    for g, og in _ograph_cache:
        if g is graph:
            return og
    n = graph.info()["n_unitigs"]
    og = OracleGraph([graph.unitig_seq(u) for u in range(n)], graph.info()["k"])
    _ograph_cache.append((graph, og))
    return og


def oracle_search(graph, s, exact=True):
    og = oracle_graph_for(graph)
    if exact:
        return og.search(s, True, False, False, False, False)
    return og.search(s, False, True, True, True, True)


def oracle_edit_distance(q, t, mode, kmax=-1, iupac=True):
    L = oracle()
    q = q.encode() if isinstance(q, str) else q
    t = t.encode() if isinstance(t, str) else t
    d, n = C.c_int(), C.c_int()
    pe = C.POINTER(C.c_int)()
    L.orc_edit_distance(q, len(q), t, len(t), mode, kmax, int(iupac), C.byref(d), C.byref(pe), C.byref(n))
    ends = [pe[i] for i in range(n.value)]
    if pe:
        L.orc_free(C.cast(pe, C.c_void_p))
    return d.value, ends
