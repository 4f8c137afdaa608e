"""Colouring a graph with long reads (the long_read_correct branch of addCoverage, src/Graph.cpp:1561-3366 =
`Ratatosk index -2`): rtk_color_long_reads against the k = 63 indexes the UNMODIFIED reference built from the same inputs
(tests/golden/F1, F2: index.k63.fasta.gz + the pass-1 corrected reads corrected_pass1.fastq.gz -> index.k63.rtsk).

The reference deals read ids in an order that depends on thread timing (its indexes differ run to run, and its single-thread
branch colours nothing), so colours are compared up to a relabelling of the ids: a bijection between our ids and the stored ids
must exist under which EVERY unitig's colour set is identical; coverage words (unphased k-mer coverage, isBranching) and the
edge flags are compared exactly.  CPU: kernel sources on the simulator; GPU: the product library."""
import os

import pytest

import ratatosk_b200 as rb
from common import GOLDEN, ROOT, read_fastq

F3 = os.path.join(ROOT, "bench_data", "F3")   # E. coli-scale k = 63 graph (31 k unitigs), coloured by the reference with 200 corrected reads


def _dir(fx):
    return (F3, "corrected200_pass1.fastq.gz") if fx == "F3" else (os.path.join(GOLDEN, fx), "corrected_pass1.fastq.gz")


def _color(fx, lib, opt=None, **kw):
    d, reads = _dir(fx)
    fa = os.path.join(d, "index.k63.fasta.gz")
    g = rb.Graph.load(fa, "", 63, lib=lib)                      # the graph only: no colours, no flags
    ctx = rb.Context(0, lib=lib)
    ctx.upload(g)
    recs = read_fastq(os.path.join(d, reads))
    stats = [0] * 10
    res = ctx.color_long_reads([r[1] for r in recs], [r[2] for r in recs], [r[0] for r in recs], opt=opt, stats=stats, **kw)
    return g, ctx, recs, res, stats


def _bijection(ours, stored):
    """ours / stored: per unitig a set of ids.  Returns the relabelling our id -> stored id, or raises."""
    n_ours = len(set().union(*ours)) if ours else 0
    n_stored = len(set().union(*stored)) if stored else 0
    assert n_ours == n_stored, (n_ours, n_stored)
    m = {}
    changed = True
    while changed:
        changed = False
        for a, b in zip(ours, stored):
            assert len(a) == len(b)
            un = [x for x in a if x not in m]
            rem = b - set(m[x] for x in a if x in m)
            if len(un) == 1 and len(rem) == 1:
                m[un[0]] = next(iter(rem))
                changed = True
    return m


def _check_color(fx, lib):
    g, ctx, recs, (kmcov, shared, off, ids, rid), stats = _color(fx, lib)
    d = _dir(fx)[0]
    want = rb.Graph.load(os.path.join(d, "index.k63.fasta.gz"), os.path.join(d, "index.k63.rtsk"), 63, lib=lib)
    n = g.info()["n_unitigs"]
    assert n == want.info()["n_unitigs"] and all(g.unitig_seq(u) == want.unitig_seq(u) for u in range(0, n, 97))
    ours = [set(map(int, ids[int(off[u]):int(off[u + 1])])) for u in range(n)]
    stored, bad_words = [], []
    for u in range(n):
        a, b = want.unitig_colors(u)
        stored.append(set(a) | set(b))
        kc, sh, _ = want.unitig_words(u)
        if int(kmcov[u]) != kc or (int(shared[u]) & 0xff) != (sh & 0xff):
            bad_words.append((u, hex(int(kmcov[u])), hex(kc), hex(int(shared[u])), hex(sh)))
    assert not bad_words, bad_words[:5]
    m = _bijection(ours, stored)
    assert len(set(m.values())) == len(m)
    resolved = [u for u in range(n) if all(x in m for x in ours[u])]
    assert all(set(m[x] for x in ours[u]) == stored[u] for u in resolved)
    if fx != "F1":
        assert len(resolved) == n and stats[5] == len(m) >= 20             # every unitig, every id
    if fx == "F2":
        assert sum(1 for r in recs if len(r[1]) < 3000) > 0                 # the case holds reads below min_len_2nd_pass ...
        assert all((len(recs[i][1]) >= 3000) == (int(rid[i]) != 0xFFFFFFFF) for i in range(len(recs)))
    assert stats[8] < 10
    ctx.close(); g.close(); want.close()


@pytest.mark.parametrize("fx", ["F1", "F2", "F3"])
def test_color_long_reads_kernel_source_matches_reference_index(fx, sim_lib):
    _check_color(fx, sim_lib)


def test_color_long_reads_options(sim_lib):
    """min_len keeps short reads out; without qualities nothing is masked (more k-mers mapped); duplicate names share an id; a
    min_cov_vertices nobody reaches clears every edge flag but leaves coverage and branching alone"""
    g, ctx, recs, (km0, sh0, off0, ids0, rid0), st0 = _color("F2", sim_lib)
    seqs, quals, names = [r[1] for r in recs], [r[2] for r in recs], [r[0] for r in recs]
    km1, _, off1, ids1, rid1 = ctx.color_long_reads(seqs, quals, names, min_len=10 ** 9)
    assert len(ids1) == 0 and int(km1.max()) >> 63 in (0, 1) and all(int(x) == 0xFFFFFFFF for x in rid1)
    km2, _, off2, ids2, _ = ctx.color_long_reads(seqs, None, names)
    assert int((km2 & ((1 << 62) - 1)).sum()) > int((km0 & ((1 << 62) - 1)).sum())
    _, _, off3, ids3, rid3 = ctx.color_long_reads(seqs + seqs[:3], quals + quals[:3], names + names[:3])
    assert list(rid3[-3:]) == list(rid3[:3]) and list(ids3) == list(ids0) and list(off3) == list(off0)
    opt = rb.default_opt(2, lib=sim_lib)
    opt.min_cov_vertices = 1 << 30
    km4, sh4, _, ids4, _ = ctx.color_long_reads(seqs, quals, names, opt=opt)
    assert int((sh4 & 0xff).max()) == 0 and list(km4) == list(km0) and list(ids4) == list(ids0)
    ctx.close(); g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("fx", ["F1", "F2", "F3"])
def test_color_long_reads_cuda_matches_reference_index(fx):
    _check_color(fx, None)


def test_color_long_reads_refuses_where_the_reference_subsamples(sim_lib):
    """the pass-1 reads of F2 given 30 times under different names: the estimated haplotype coverage passes 10, where the reference
    starts drawing reads at random (src/Graph.cpp:2312) - the call refuses, unless the caller opts into keeping every read"""
    g, ctx, recs, (km0, sh0, off0, ids0, _), st0 = _color("F2", sim_lib)
    seqs = [r[1] for r in recs] * 30
    quals = [r[2] for r in recs] * 30
    names = ["%s_%d" % (r[0], c) for c in range(30) for r in recs]
    with pytest.raises(rb.RtkError, match="haplotype coverage"):
        ctx.color_long_reads(seqs, quals, names)
    opt = rb.default_opt(2, lib=sim_lib)
    opt.reserved = 1
    st = [0] * 10
    km, sh, off, ids, rid = ctx.color_long_reads(seqs, quals, names, opt=opt, stats=st)
    assert st[8] >= 10 and st[5] == 30 * st0[5]
    cov = lambda w: (w >> 31) & 0x7fffffff
    assert all(cov(int(a)) == 30 * cov(int(b)) for a, b in zip(km, km0))
    assert all(int(off[u + 1] - off[u]) == 30 * int(off0[u + 1] - off0[u]) for u in range(len(km0)))
    ctx.close(); g.close()


def test_haplotype_coverage_estimate_switches_where_the_reference_does(sim_lib, tmp_path):
    """estimateHaplotypeCoverage (src/Graph.cpp:4185-4233) restated on the adjacency table: with 23 copies of the F2 reads the
    reference's `index -2 -v` does not subsample, with 24 it does - the estimate must cross 10 between the two"""
    import subprocess
    ref_cli = os.path.join(ROOT, "oracle", "_ref", "Ratatosk")
    if not os.path.exists(ref_cli):
        pytest.skip("reference CLI not built")
    g, ctx, recs, _, _ = _color("F2", sim_lib)
    fa = os.path.join(GOLDEN, "F2", "index.k63.fasta.gz")
    opt = rb.default_opt(2, lib=sim_lib)
    opt.reserved = 1
    for copies, want in ((23, False), (24, True)):
        st = [0] * 10
        ctx.color_long_reads([r[1] for r in recs] * copies, [r[2] for r in recs] * copies,
                             ["%s_%d" % (r[0], c) for c in range(copies) for r in recs], opt=opt, stats=st)
        assert (st[8] >= 10) == want, (copies, st[8])
        reads = str(tmp_path / ("rep%d.fastq" % copies))
        with open(reads, "w") as f:
            for c in range(copies):
                for n, s, q in recs:
                    f.write("@%s_%d\n%s\n+\n%s\n" % (n, c, s, q))
        out = subprocess.run([ref_cli, "index", "-2", "-v", "-c", "4", "-g", fa, "-l", reads, "-o", str(tmp_path / "rep")],
                             capture_output=True, text=True).stdout
        assert ("Subsampling reads" in out) == want, copies
    ctx.close(); g.close()
