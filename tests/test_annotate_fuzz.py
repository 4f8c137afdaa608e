"""Differential test of the annotation kernels against the REAL reference on colourings the fixtures do not contain.

The F2 graphs (k = 31: 3,024 unitigs with repeats, tandem repeats and SNP bubbles; k = 63) are re-coloured with synthetic reads -
random walks through the graph, at several depths - so that walks and cycle searches meet far denser and far sparser read support
than a real 30x data set gives: tangles where most edges are flagged (the first-attempt arenas overflow and unitigs are re-run),
unitigs without any support, colour sets of every container kind.  The colouring is written as an index (rtk_graph_recolor +
rtk_rtsk_write), the UNMODIFIED reference loads it (readGraphData) and runs its own detectSNPs + detectShortCycles through the seam
probe ref_annotate; the kernel sources (CPU simulator) must give the same ambiguity ids, flags and cycle blobs for every unitig.
CPU only: needs oracle/_ref/libref_seams.so (build container)."""
import os

import numpy as np
import pytest

import ratatosk_b200 as rb
import refseams
from common import GOLDEN

pytestmark = pytest.mark.skipif(not refseams.available(), reason="reference seam library not built (oracle/_ref)")


def _walk_colouring(g, n, n_reads, walk_len, seed):
    rng = np.random.RandomState(seed)
    adj = [g.unitig_words(u)[2] for u in range(n)]
    sets = [set() for _ in range(n)]
    for r in range(n_reads):
        u, s = int(rng.randint(n)), int(rng.randint(2))
        for _ in range(int(rng.randint(1, walk_len + 1))):
            sets[u].add(r)
            nxt = [adj[u][b] if s else adj[u][4 + (3 - b)] for b in range(4)]
            nxt = [x for x in nxt if x != 0xFFFFFFFF]
            if not nxt:
                break
            x = nxt[int(rng.randint(len(nxt)))]
            u, s = x & 0x7fffffff, (x >> 31) if s else 1 - (x >> 31)
    return [sorted(x) for x in sets], adj


def _edge_flags(sets, adj, n, min_cov):
    """postProcessUnitigs (src/Graph.cpp:1986-2023) in plain Python: the input of the functions under test"""
    ss = [set(x) for x in sets]
    flags = np.zeros(n, dtype=np.uint64)
    for u in range(n):
        f = 0
        for b in range(4):
            x = adj[u][b]
            if x != 0xFFFFFFFF and len(ss[u] & ss[x & 0x7fffffff]) >= min_cov:
                f |= (1 << b) << 4
            x = adj[u][4 + (3 - b)]
            if x != 0xFFFFFFFF and len(ss[u] & ss[x & 0x7fffffff]) >= min_cov:
                f |= 1 << b
        flags[u] = f
    return flags


@pytest.mark.parametrize("k,n_reads,walk_len,seed,min_cov,arena", [
    (31, 400, 12, 1, 2, None),      # sparse: most unitigs carry 0-3 reads
    (31, 4000, 25, 2, 2, None),     # dense
    (31, 4000, 25, 3, 3, "5"),      # dense, min_cov 3, tiny first-attempt arenas: the re-run path on real tangles
    (31, 20000, 40, 4, 2, None),    # very dense: nearly every edge flagged, long walks
    (63, 3000, 30, 5, 2, None),
    (63, 3000, 30, 6, 1, "7"),      # min_cov 1
])
def test_annotation_kernels_match_reference_on_synthetic_colourings(k, n_reads, walk_len, seed, min_cov, arena, sim_lib, tmp_path, monkeypatch):
    if arena:
        monkeypatch.setenv("RTK_AN_ARENA", arena)
    fa = os.path.join(GOLDEN, "F2", "index.k%d.fasta.gz" % k)
    g = rb.Graph.load(fa, "", k, lib=sim_lib)
    n = g.info()["n_unitigs"]
    sets, adj = _walk_colouring(g, n, n_reads, walk_len, seed)
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s in sets])
    ids = np.array([x for s in sets for x in s], dtype=np.uint32)
    g2 = g.recolor(np.zeros(n, dtype=np.uint64), _edge_flags(sets, adj, n, min_cov), off, ids)
    path = str(tmp_path / "synthetic.rtsk")
    zero = np.zeros(n + 1, dtype=np.uint64)
    g2.write_rtsk(path, zero, np.zeros(0, dtype=np.uint32), np.zeros(n, dtype=np.uint8), zero, b"")
    ref = refseams.RefGraph(fa, path, k)
    want = ref.annotate(min_cov, threads=4)
    ref.close()
    ctx = rb.Context(0, lib=sim_lib)
    ctx.upload(g2)
    opt = rb.default_opt(1 if k == 31 else 2, lib=sim_lib)
    opt.min_cov_vertices = min_cov
    st1, st2 = [0] * 10, [0] * 10
    a_off, a_ids = ctx.detect_snps(opt=opt, stats=st1)
    flags, c_off, pool = ctx.detect_short_cycles(opt=opt, stats=st2)
    got = {}
    for u in range(n):
        a = list(map(int, a_ids[int(a_off[u]):int(a_off[u + 1])]))
        b = pool[int(c_off[u]):int(c_off[u + 1])]
        if a or b or flags[u]:
            got[g2.unitig_seq(u)] = (a, int(flags[u]), b)
    bad = [s for s in set(got) | set(want) if got.get(s) != want.get(s)]
    assert not bad, (len(bad), [(got.get(s), want.get(s)) for s in bad[:3]])
    assert len(want) > 5                                   # the colouring produced annotations at all
    if arena:
        assert st1[8] + st2[8] > 0                         # ... and the re-run path was taken
    ctx.close(); g.close(); g2.close()
