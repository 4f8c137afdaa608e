"""Graph annotation: detectSNPs (src/Graph.cpp:484-720) and detectShortCycles (src/Graph.cpp:4660-4854).

Golden = the index files the UNMODIFIED reference wrote (tests/golden/F1, F2; bench_data/F3 when present): their .rtsk holds, per
unitig, the ambiguity ids detectSNPs added and the short-cycle flag / compacted-cycles blob detectShortCycles stored
(src/UnitigData.hpp:493-553), computed by the reference from the same colours and edge flags the slab carries.  The kernels
(ratatosk_b200/csrc/annotate.cuh) must reproduce them bit for bit on every unitig.  CPU: the kernel sources on the simulator;
GPU: the product library."""
import os

import numpy as np
import pytest

import ratatosk_b200 as rb
from common import GOLDEN, ROOT

F3 = os.path.join(ROOT, "bench_data", "F3")
CASES = [("F1", 31), ("F1", 63), ("F2", 31), ("F2", 63)]


def _dir(recipe):
    return F3 if recipe == "F3" else os.path.join(GOLDEN, recipe)


def _load(recipe, k, lib):
    d = _dir(recipe)
    g = rb.Graph.load(os.path.join(d, "index.k%d.fasta.gz" % k), os.path.join(d, "index.k%d.rtsk" % k), k, lib=lib)
    ctx = rb.Context(0, lib=lib)
    ctx.upload(g)
    return g, ctx


def _stored(g):
    n = g.info()["n_unitigs"]
    amb, cyc, flag = [], [], []
    for u in range(n):
        a, c = g.unitig_annotations(u)
        amb.append(a)
        cyc.append(c)
        flag.append((g.unitig_words(u)[1] >> 8) & 1)
    return amb, cyc, flag


def _check_cycles(recipe, k, lib, need_cycles):
    g, ctx = _load(recipe, k, lib)
    _, want_blob, want_flag = _stored(g)
    stats = [0] * 10
    flags, off, pool = ctx.detect_short_cycles(stats=stats)
    n = g.info()["n_unitigs"]
    assert len(flags) == n
    bad = [u for u in range(n) if int(flags[u]) != want_flag[u] or pool[int(off[u]):int(off[u + 1])] != want_blob[u]]
    assert not bad, (recipe, k, bad[:10], [(want_blob[u], pool[int(off[u]):int(off[u + 1])]) for u in bad[:3]])
    if need_cycles:
        assert sum(want_flag) >= need_cycles                                 # the case holds real cycles ...
        assert any(b.count(b"\0") > 1 for b in want_blob)                    # ... some unitigs with several
        assert any(b == b"\0" for b in want_blob)                            # ... and self loops (empty middle path)
    ctx.close()
    g.close()


def _check_snps(recipe, k, lib, need_marks):
    g, ctx = _load(recipe, k, lib)
    want, _, _ = _stored(g)
    stats = [0] * 10
    off, ids = ctx.detect_snps(stats=stats)
    n = g.info()["n_unitigs"]
    got = [list(map(int, ids[int(off[u]):int(off[u + 1])])) for u in range(n)]
    bad = [u for u in range(n) if got[u] != want[u]]
    assert not bad, (recipe, k, len(bad), [(u, want[u], got[u]) for u in bad[:5]])
    n_marks = sum(len(w) for w in want)
    assert n_marks >= need_marks, n_marks
    assert stats[4] >= n_marks and stats[6] > 0                               # candidates were replayed, traversals ran
    ctx.close()
    g.close()


@pytest.mark.parametrize("recipe,k", CASES)
def test_short_cycles_kernel_source_matches_reference_index(recipe, k, sim_lib):
    _check_cycles(recipe, k, sim_lib, 100 if (recipe, k) == ("F2", 31) else 0)


@pytest.mark.parametrize("recipe,k", CASES)
def test_detect_snps_kernel_source_matches_reference_index(recipe, k, sim_lib):
    _check_snps(recipe, k, sim_lib, 90 if recipe == "F2" else 0)


def _check_min_cov_vectors(lib):
    """min_cov_vertices 1, 2, 3, 5 against the reference re-run through the seam probe (tests/golden/make_golden_annotate.py)"""
    import gzip
    import json
    vec = json.load(gzip.open(os.path.join(GOLDEN, "annotate_vectors.json.gz"), "rt"))
    n_marks = 0
    for fx in ("F1", "F2"):
        for k in (31, 63):
            g, ctx = _load(fx, k, lib)
            n = g.info()["n_unitigs"]
            for mc, want in sorted(vec[fx][str(k)].items()):
                opt = rb.default_opt(1 if k == 31 else 2, lib=lib)
                opt.min_cov_vertices = int(mc)
                off, ids = ctx.detect_snps(opt=opt)
                flags, coff, pool = ctx.detect_short_cycles(opt=opt)
                got = {}
                for u in range(n):
                    a = list(map(int, ids[int(off[u]):int(off[u + 1])]))
                    b = pool[int(coff[u]):int(coff[u + 1])]
                    if a or b or flags[u]:
                        got[str(u)] = [a, int(flags[u]), b.decode("latin1")]
                assert got == want, (fx, k, mc, [u for u in set(got) | set(want) if got.get(u) != want.get(u)][:10])
                n_marks += sum(len(v[0]) for v in want.values())
            ctx.close()
            g.close()
    assert n_marks > 10000


def test_annotation_min_cov_variants_match_reference(sim_lib):
    _check_min_cov_vectors(sim_lib)


@pytest.mark.gpu
def test_annotation_min_cov_variants_match_reference_cuda():
    _check_min_cov_vectors(None)


def _check_small_arena(lib, monkeypatch):
    """first-attempt arenas of 3 entries: most unitigs overflow and are re-run with the large arena; same bytes out"""
    monkeypatch.setenv("RTK_AN_ARENA", "3")
    g, ctx = _load("F2", 31, lib)
    want_amb, want_blob, _ = _stored(g)
    n = g.info()["n_unitigs"]
    st1, st2 = [0] * 10, [0] * 10
    off, ids = ctx.detect_snps(stats=st1)
    _, coff, pool = ctx.detect_short_cycles(stats=st2)
    assert st1[8] > 100 and st2[8] > 100                                      # the re-run path was taken
    assert all(list(map(int, ids[int(off[u]):int(off[u + 1])])) == want_amb[u] for u in range(n))
    assert all(pool[int(coff[u]):int(coff[u + 1])] == want_blob[u] for u in range(n))
    ctx.close()
    g.close()


def test_annotation_rerun_with_large_arena(sim_lib, monkeypatch):
    _check_small_arena(sim_lib, monkeypatch)


def test_annotation_needs_graph_and_min_cov(sim_lib):
    """no graph on the context: error, not a crash; a min_cov_vertices nobody reaches: no marks, no cycles"""
    ctx = rb.Context(0, lib=sim_lib)
    ctx.graph = rb.Graph.from_unitigs(["ACGTTGCATGGACCAGTTAGACCATGACCAGTAGGACCATAG"], 31, lib=sim_lib)
    with pytest.raises(rb.RtkError):
        ctx.detect_snps()
    with pytest.raises(rb.RtkError):
        ctx.detect_short_cycles()
    ctx.upload(ctx.graph)                                                       # uncoloured graph: nothing to annotate
    off, ids = ctx.detect_snps()
    flags, coff, pool = ctx.detect_short_cycles()
    assert len(ids) == 0 and int(flags.sum()) == 0 and pool == b""
    ctx.close()
    g, ctx = _load("F2", 31, sim_lib)
    opt = rb.default_opt(1, lib=sim_lib)
    opt.min_cov_vertices = 1 << 30
    off, ids = ctx.detect_snps(opt=opt)
    flags, coff, pool = ctx.detect_short_cycles(opt=opt)
    assert len(ids) == 0 and int(flags.sum()) == 0 and pool == b""
    ctx.close()
    g.close()


@pytest.mark.gpu
def test_annotation_rerun_with_large_arena_cuda(monkeypatch):
    _check_small_arena(None, monkeypatch)


@pytest.mark.gpu
@pytest.mark.parametrize("recipe,k", CASES + [("F3", 31), ("F3", 63)])
def test_short_cycles_cuda_matches_reference_index(recipe, k):
    if recipe == "F3" and not os.path.exists(os.path.join(F3, "index.k31.rtsk")):
        pytest.skip("bench_data/F3 not generated")
    _check_cycles(recipe, k, None, 100 if (recipe, k) == ("F2", 31) else 0)


@pytest.mark.gpu
@pytest.mark.parametrize("recipe,k", CASES + [("F3", 31), ("F3", 63)])
def test_detect_snps_cuda_matches_reference_index(recipe, k):
    if recipe == "F3" and not os.path.exists(os.path.join(F3, "index.k31.rtsk")):
        pytest.skip("bench_data/F3 not generated")
    _check_snps(recipe, k, None, 90 if recipe == "F2" else 0)
