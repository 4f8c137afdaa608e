"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path through the C ABI against
(a) golden vectors recorded from the unmodified reference and (b) the CPU oracle on seeded inputs,
plus size-independent properties at larger sizes and the edge cases of the domain."""
import os
import random

import numpy as np
import pytest

import ratatosk_b200 as rb
from common import GOLDEN, golden_paths, load_golden_reads, oracle_search, revcomp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["F1", "F2"])
def loaded(request):
    fa, rt = golden_paths(request.param)
    g = rb.Graph.load(fa, rt, 31)
    ctx = rb.Context(0)
    ctx.upload(g)
    yield request.param, g, ctx
    ctx.close()
    g.close()


def _inexact(ctx, reads, **kw):
    return ctx.search_sequence(reads, exact=False, insertion=True, deletion=True, substitution=True,
                               or_exclusive_match=True, **kw)


def test_search_sequence_matches_reference_golden(loaded):
    recipe, _, ctx = loaded
    gold = np.load(os.path.join(GOLDEN, recipe, "golden_hits.npz"))
    reads = [s for _, s, _ in load_golden_reads(recipe)]
    ex = ctx.search_sequence(reads)
    ix = _inexact(ctx, reads)
    for i in range(len(reads)):
        assert np.array_equal(ex[i], gold["exact_%d" % i]), (recipe, i)
        assert np.array_equal(ix[i], gold["inexact_%d" % i]), (recipe, i)


def test_get_seeds_matches_reference_golden(loaded):
    recipe, _, ctx = loaded
    gold = np.load(os.path.join(GOLDEN, recipe, "golden_hits.npz"))
    reads = [s for _, s, _ in load_golden_reads(recipe)]
    solid, weak = ctx.get_seeds(reads)
    for i in range(len(reads)):
        assert np.array_equal(solid[i], gold["solid_%d" % i]), (recipe, i)
        assert np.array_equal(weak[i], gold["weak_%d" % i]), (recipe, i)


def test_batching_is_transparent(loaded):
    """one read per call == all reads in one call (no cross-read state)"""
    recipe, _, ctx = loaded
    reads = [s for _, s, _ in load_golden_reads(recipe)][:5]
    together = _inexact(ctx, reads)
    for i, r in enumerate(reads):
        assert np.array_equal(_inexact(ctx, [r])[0], together[i])


def _mutate(rng, s, rate):
    out = []
    for c in s:
        r = rng.random()
        if r < rate * 0.4:
            continue
        if r < rate * 0.7:
            c = rng.choice("ACGT")
        out.append(c)
        if rng.random() < rate * 0.3:
            out.append(rng.choice("ACGT"))
    return "".join(out)


def test_cuda_matches_oracle_on_seeded_reads(loaded):
    """fresh seeded reads (not in the golden set), incl. N's, lower-error and ragged lengths"""
    recipe, g, ctx = loaded
    rng = random.Random(20261017 + len(recipe))
    n = g.info()["n_unitigs"]
    reads = []
    for t in range(10):
        u = g.unitig_seq(rng.randrange(n))
        while len(u) < 200:
            u += g.unitig_seq(rng.randrange(n))
        s = _mutate(rng, u[:rng.randint(64, 2500)], [0.0, 0.02, 0.1, 0.15][t % 4])
        if t % 3 == 0 and len(s) > 80:
            for _ in range(3):
                p = rng.randrange(len(s))
                s = s[:p] + "N" + s[p + 1:]
        if t % 2:
            s = revcomp(s)
        reads.append(s)
    reads += ["", "ACGT", "A" * 31, "ACGTN" * 20, g.unitig_seq(0)[:31], g.unitig_seq(0)[:32], "N" * 300]
    ex = ctx.search_sequence(reads)
    ix = _inexact(ctx, reads)
    for i, s in enumerate(reads):
        assert np.array_equal(ex[i], oracle_search(g, s, exact=True)), (recipe, i, "exact")
        assert np.array_equal(ix[i], oracle_search(g, s, exact=False)), (recipe, i, "inexact")


def test_every_graph_kmer_finds_itself(loaded):
    """property at full graph size: each unitig, searched exactly, maps onto itself end to end"""
    _, g, ctx = loaded
    n = g.info()["n_unitigs"]
    ids = list(range(0, n, max(1, n // 500)))
    seqs = [g.unitig_seq(u) for u in ids]
    hits = ctx.search_sequence(seqs)
    for u, s, h in zip(ids, seqs, hits):
        assert len(h) == len(s) - 30
        assert (h[:, 1] == u).all() and (h[:, 3] == 1).all()
        assert (h[:, 0] == np.arange(len(h))).all() and (h[:, 2] == np.arange(len(h))).all()
    hits_rc = ctx.search_sequence([revcomp(s) for s in seqs])
    for u, s, h in zip(ids, seqs, hits_rc):
        assert len(h) == len(s) - 30 and (h[:, 1] == u).all() and (h[:, 3] == 0).all()
        assert sorted((h[:, 0] + h[:, 2]).tolist()) == [len(s) - 31] * len(h)


def test_single_edit_is_recovered(loaded):
    """property: a unitig window with ONE substitution / insertion / deletion in the middle is found by the
    inexact sweep on the windows spanning the edit (what the sweep exists for)"""
    _, g, ctx = loaded
    rng = random.Random(7)
    n = g.info()["n_unitigs"]
    cands = [u for u in range(n) if len(g.unitig_seq(u)) >= 120][:40]
    reads, kinds = [], []
    for u in cands:
        s = g.unitig_seq(u)[:100]
        p = 50
        kind = rng.randrange(3)
        if kind == 0:
            s2 = s[:p] + rng.choice([c for c in "ACGT" if c != s[p]]) + s[p + 1:]
        elif kind == 1:
            s2 = s[:p] + s[p + 1:]
        else:
            s2 = s[:p] + rng.choice("ACGT") + s[p:]
        reads.append(s2)
        kinds.append(kind)
    ix = _inexact(ctx, reads)
    for u, h in zip(cands, ix):
        on_u = h[h[:, 1] == u]
        assert len(on_u) >= 20, (u, len(on_u))


def test_masked_reads_with_isolated_n_keep_their_deletion_hits(loaded):
    """tile skipping of the K1 driver (mostly-'N' masked copies) must keep windows whose deletion drops an isolated 'N'"""
    recipe, g, ctx = loaded
    n = g.info()["n_unitigs"]
    u = max(range(min(n, 50)), key=lambda x: len(g.unitig_seq(x)))
    s = g.unitig_seq(u)
    pad = "N" * 700
    reads = [pad + s[:20] + "N" + s[20:48] + pad, pad + s[:20] + "N" + s[20:48],
             s[:20] + "N" + s[20:48] + pad + s[5:40] + pad, pad + s[:29] + pad]
    got = ctx.search_sequence(reads, exact=False, insertion=True, deletion=True, substitution=True, or_exclusive_match=True)
    for i, r in enumerate(reads):
        assert np.array_equal(got[i], oracle_search(g, r, exact=False)), (recipe, i)
    assert len(got[0]) > 0 and len(got[3]) == 0
