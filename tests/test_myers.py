"""K4 (edit distance) parity: oracle vs reference golden, kernel source on the CPU simulator vs golden."""
import gzip
import json
import os

import numpy as np
import pytest

import ratatosk_b200 as rb
from common import GOLDEN, oracle_edit_distance


def _vectors():
    with gzip.open(os.path.join(GOLDEN, "edlib_vectors.json.gz"), "rt") as f:
        return json.load(f)


def test_oracle_edit_distance_matches_reference():
    for i, c in enumerate(_vectors()):
        d, ends = oracle_edit_distance(c["q"], c["t"], c["mode"], c["k"])
        assert d == c["dist"], (i, c["mode"], c["k"], len(c["q"]), len(c["t"]), d, c["dist"])
        assert ends == c["ends"], (i, c["mode"])


def test_myers_kernel_source_matches_reference(sim_lib):
    ctx = rb.Context(0, lib=sim_lib)
    vec = _vectors()
    vec = vec[:25] + vec[25::4]  # all edge cases + a quarter of the random ones (simulator speed)
    dist, ends = ctx.edlib_batch([c["q"] for c in vec], [c["t"] for c in vec], [c["mode"] for c in vec], [c["k"] for c in vec])
    for i, c in enumerate(vec):
        assert int(dist[i]) == c["dist"], (i, c["mode"], c["k"], len(c["q"]), len(c["t"]))
        assert ends[i].tolist() == c["ends"], (i, c["mode"])
    ctx.close()


@pytest.mark.gpu
def test_myers_cuda_matches_reference_golden():
    ctx = rb.Context(0)
    vec = _vectors()
    dist, ends = ctx.edlib_batch([c["q"] for c in vec], [c["t"] for c in vec], [c["mode"] for c in vec], [c["k"] for c in vec])
    for i, c in enumerate(vec):
        assert int(dist[i]) == c["dist"], (i, c["mode"], c["k"], len(c["q"]), len(c["t"]))
        assert ends[i].tolist() == c["ends"], (i, c["mode"])
    ctx.close()


@pytest.mark.gpu
def test_myers_cuda_matches_oracle_on_seeded_pairs():
    """larger seeded set incl. pass-2 sized queries (> 2048 rows: multi-round sweep)"""
    import random
    rng = random.Random(99)
    qs, ts, ms = [], [], []
    for i in range(300):
        tl = rng.choice([rng.randint(1, 200), rng.randint(200, 1300), rng.randint(2100, 5200) if i % 10 == 0 else 500])
        t = "".join(rng.choice("ACGT") for _ in range(tl))
        q = "".join(c for c in t if rng.random() > 0.05)
        q = "".join(rng.choice("ACGTN") if rng.random() < 0.08 else c for c in q) or "A"
        qs.append(q); ts.append(t); ms.append(i % 3)
    ctx = rb.Context(0)
    dist, ends = ctx.edlib_batch(qs, ts, ms)
    for i in range(len(qs)):
        d, e = oracle_edit_distance(qs[i], ts[i], ms[i])
        assert int(dist[i]) == d and ends[i].tolist() == e, (i, ms[i], len(qs[i]), len(ts[i]))
    # properties: symmetry of NW, identity
    d2, _ = ctx.edlib_batch(ts, qs, [0] * len(qs))
    d1, _ = ctx.edlib_batch(qs, ts, [0] * len(qs))
    assert np.array_equal(d1, d2)
    d0, _ = ctx.edlib_batch(qs, qs, [0] * len(qs))
    assert (d0 == 0).all()
    ctx.close()


def _iupac_pairs(n, seed):
    """pairs dense in ambiguity codes on BOTH sides (the second pass aligns pass-1 reads that carry unresolved SNP codes)"""
    import random
    rng = random.Random(seed)
    codes = "MRSVWYHKDBN"
    qs, ts, ms = [], [], []
    for i in range(n):
        tl = rng.choice([rng.randint(1, 70), rng.randint(60, 200), rng.randint(200, 400)])
        t = "".join(rng.choice(codes) if rng.random() < 0.12 else rng.choice("ACGT") for _ in range(tl))
        q = "".join(c for c in t if rng.random() > 0.06)
        q = "".join(rng.choice(codes + "ACGT") if rng.random() < 0.10 else c for c in q) or "R"
        if i % 17 == 0:
            t = t[:len(t) // 2] + "." + t[len(t) // 2:]   # a character outside the alphabet: equal to itself only
        qs.append(q); ts.append(t); ms.append(i % 3)
    return qs, ts, ms


def test_myers_kernel_source_iupac_codes_on_both_sides(sim_lib):
    """the bit-parallel treatment of ambiguity codes in the TARGET (rtk_iupac_eq_word) against the oracle's per-pair equality"""
    qs, ts, ms = _iupac_pairs(60, 4242)
    ctx = rb.Context(0, lib=sim_lib)
    dist, ends = ctx.edlib_batch(qs, ts, ms)
    for i in range(len(qs)):
        d, e = oracle_edit_distance(qs[i], ts[i], ms[i])
        assert int(dist[i]) == d and ends[i].tolist() == e, (i, ms[i], qs[i], ts[i])
    ctx.close()


@pytest.mark.gpu
def test_myers_cuda_iupac_codes_on_both_sides():
    qs, ts, ms = _iupac_pairs(600, 777)
    ctx = rb.Context(0)
    dist, ends = ctx.edlib_batch(qs, ts, ms)
    for i in range(len(qs)):
        d, e = oracle_edit_distance(qs[i], ts[i], ms[i])
        assert int(dist[i]) == d and ends[i].tolist() == e, (i, ms[i], len(qs[i]), len(ts[i]))
    ctx.close()
