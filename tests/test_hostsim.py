"""CPU tests of the product's own sources: the index loader / slab builder, the host-side anchor
logic, and the K1 kernel SOURCES executed on the CPU simulator (tests/hostsim), all compared with
golden vectors recorded from the unmodified reference."""
import gzip
import os

import numpy as np
import pytest

import ratatosk_b200 as rb
from common import GOLDEN, golden_paths, load_golden_reads


@pytest.fixture(scope="module", params=["F1", "F2"])
def loaded(request, sim_lib):
    fa, rt = golden_paths(request.param)
    g = rb.Graph.load(fa, rt, 31, lib=sim_lib)
    ctx = rb.Context(0, lib=sim_lib)
    ctx.upload(g)
    yield request.param, g, ctx
    ctx.close()
    g.close()


def test_index_loader_matches_reference_dump(loaded):
    """FASTA + .rtsk parser (PairID kinds, TinyBitmap, Roaring) vs the reference's in-memory graph."""
    recipe, g, _ = loaded
    info = g.info()
    import json
    meta = json.load(open(os.path.join(GOLDEN, recipe, "meta.json")))
    assert info["n_unitigs"] == meta["n_unitigs"]
    assert max(info["max_km_cov_graph"], 128) == meta["max_km_cov"]
    n = 0
    with gzip.open(os.path.join(GOLDEN, recipe, "ref_unitigs.tsv.gz"), "rt") as f:
        for line in f:
            c = line.rstrip("\n").split("\t")
            u = int(c[0])
            seq = g.unitig_seq(u)
            assert seq == c[1] or seq == c[1][::-1].translate(str.maketrans("ACGT", "TGCA"))
            w0, w1, _ = g.unitig_words(u)
            assert (w0, w1) == (int(c[2]), int(c[3]))
            gi, li = g.unitig_colors(u)
            assert gi == [int(x) for x in c[4].split(",") if x]
            assert li == [int(x) for x in c[5].split(",") if x]
            n += 1
    assert n == info["n_unitigs"]


def test_adjacency_is_consistent(loaded):
    """successor in A,C,G,T order: the (k-1)-overlap must hold and the reverse edge must exist"""
    _, g, _ = loaded
    k = 31
    rc = lambda s: s[::-1].translate(str.maketrans("ACGT", "TGCA"))
    n = g.info()["n_unitigs"]
    seqs = [g.unitig_seq(u) for u in range(n)]
    for u in range(0, n, max(1, n // 400)):
        _, _, adj = g.unitig_words(u)
        for c in range(4):
            v = adj[c]
            if v == 0xFFFFFFFF:
                continue
            vs = seqs[v & 0x7FFFFFFF] if (v >> 31) else rc(seqs[v & 0x7FFFFFFF])
            assert vs[:k - 1] == seqs[u][-(k - 1):] and vs[k - 1] == "ACGT"[c]
        for c in range(4):
            v = adj[4 + c]
            if v == 0xFFFFFFFF:
                continue
            vs = seqs[v & 0x7FFFFFFF] if (v >> 31) else rc(seqs[v & 0x7FFFFFFF])
            assert vs[-(k - 1):] == seqs[u][:k - 1] and vs[-k] == "ACGT"[c]


def test_k1_kernel_source_matches_reference(loaded):
    recipe, _, ctx = loaded
    gold = np.load(os.path.join(GOLDEN, recipe, "golden_hits.npz"))
    reads = [s for _, s, _ in load_golden_reads(recipe)][:6]
    ex = ctx.search_sequence(reads)
    ix = ctx.search_sequence(reads, exact=False, insertion=True, deletion=True, substitution=True, or_exclusive_match=True)
    for i in range(len(reads)):
        assert np.array_equal(ex[i], gold["exact_%d" % i]), (recipe, i)
        assert np.array_equal(ix[i], gold["inexact_%d" % i]), (recipe, i)


def test_get_seeds_matches_reference(loaded):
    recipe, _, ctx = loaded
    gold = np.load(os.path.join(GOLDEN, recipe, "golden_hits.npz"))
    reads = [s for _, s, _ in load_golden_reads(recipe)][:6]
    solid, weak = ctx.get_seeds(reads)
    for i in range(len(reads)):
        assert np.array_equal(solid[i], gold["solid_%d" % i]), (recipe, i)
        assert np.array_equal(weak[i], gold["weak_%d" % i]), (recipe, i)


def test_slab_roundtrip(loaded, tmp_path, sim_lib):
    _, g, _ = loaded
    p = str(tmp_path / "g.rtkflat")
    g.save(p)
    g2 = rb.Graph.open(p, lib=sim_lib)
    assert g2.info() == g.info()
    assert np.array_equal(g2.slab(), g.slab())
    g2.close()


def test_flat_cache_is_written_once_then_mapped_in_place(tmp_path, sim_lib):
    """rtk_graph_load_cached: first call parses the index and writes the cache, the second maps the file (same slab bytes, same
    anchors); a stale cache (older than the index) and a cache of another k are rebuilt, a corrupt one is rejected by open"""
    import shutil
    import time
    d = tmp_path / "idx"
    d.mkdir()
    fa, rt = golden_paths("F1")
    fa2, rt2 = str(d / "g.fasta.gz"), str(d / "g.rtsk")
    shutil.copy(fa, fa2); shutil.copy(rt, rt2)
    g1, c1 = rb.Graph.load_cached(fa2, rt2, 31, lib=sim_lib)
    cache = rt2 + ".k31.rtkflat"
    assert not c1 and os.path.exists(cache)
    g2, c2 = rb.Graph.load_cached(fa2, rt2, 31, lib=sim_lib)
    assert c2 and np.array_equal(g1.slab(), g2.slab())
    ctx = rb.Context(0, lib=sim_lib)
    ctx.upload(g2)
    gold = np.load(os.path.join(GOLDEN, "F1", "golden_hits.npz"))
    reads = [s for _, s, _ in load_golden_reads("F1")][:3]
    solid, weak = ctx.get_seeds(reads)
    for i in range(len(reads)):
        assert np.array_equal(solid[i], gold["solid_%d" % i]) and np.array_equal(weak[i], gold["weak_%d" % i])
    ctx.close()
    g1.close(); g2.close()
    # stale: index newer than the cache
    os.utime(cache, (time.time() - 1000, time.time() - 1000))
    g3, c3 = rb.Graph.load_cached(fa2, rt2, 31, lib=sim_lib)
    assert not c3
    g3.close()
    # explicit cache path; a truncated file is not a slab
    other = str(tmp_path / "other.rtkflat")
    g4, c4 = rb.Graph.load_cached(fa2, rt2, 31, cache=other, lib=sim_lib)
    assert not c4 and os.path.exists(other)
    g4.close()
    open(other, "r+b").truncate(1000)
    with pytest.raises(rb.RtkError):
        rb.Graph.open(other, lib=sim_lib)


def test_masked_reads_with_isolated_n_keep_their_deletion_hits(loaded):
    """The K1 driver skips tiles that cannot hold a valid window (the masked copies of getSeeds are mostly 'N').  A deletion
    window may drop an 'N' that sits between two short valid stretches: such tiles must stay.  Checked against the CPU oracle."""
    from common import oracle_search
    recipe, g, ctx = loaded
    n = g.info()["n_unitigs"]
    u = max(range(min(n, 50)), key=lambda x: len(g.unitig_seq(x)))
    s = g.unitig_seq(u)
    assert len(s) >= 60
    pad = "N" * 700
    reads = [pad + s[:20] + "N" + s[20:48] + pad,            # valid stretches of 20 and 28 < k - 1, joined by deleting the N
             pad + s[:20] + "N" + s[20:48],                   # the same at the end of the read
             s[:20] + "N" + s[20:48] + pad + s[5:40] + pad,   # ... at the start, plus an ordinary stretch
             pad + s[:29] + pad]                              # k - 2 valid bases: nothing can hit
    got = ctx.search_sequence(reads, exact=False, insertion=True, deletion=True, substitution=True, or_exclusive_match=True)
    for i, r in enumerate(reads):
        want = oracle_search(g, r, exact=False)
        assert np.array_equal(got[i], want), (recipe, i)
    assert len(got[0]) > 0 and len(got[3]) == 0
