"""Pin the CPU oracle (oracle/rtk_oracle.cpp) against golden vectors recorded from the UNMODIFIED
reference (tests/golden/make_golden.py)."""
import gzip
import os

import numpy as np
import pytest

from common import GOLDEN, OracleGraph, load_golden_reads


def _unitigs(recipe):
    seqs = []
    with gzip.open(os.path.join(GOLDEN, recipe, "index.k31.fasta.gz"), "rt") as f:
        for line in f:
            if line.startswith(">"):
                seqs.append("")
            else:
                seqs[-1] += line.strip()
    return seqs


@pytest.mark.parametrize("recipe,nreads", [("F1", 18), ("F2", 8)])
def test_oracle_search_sequence_matches_reference(recipe, nreads):
    og = OracleGraph(_unitigs(recipe), 31)
    gold = np.load(os.path.join(GOLDEN, recipe, "golden_hits.npz"))
    reads = load_golden_reads(recipe)[:nreads]
    for i, (_, s, _) in enumerate(reads):
        assert np.array_equal(og.search(s, True, False, False, False, False), gold["exact_%d" % i]), (recipe, i, "exact")
        assert np.array_equal(og.search(s, False, True, True, True, True), gold["inexact_%d" % i]), (recipe, i, "inexact")
    og.close()


def test_oracle_search_edge_cases():
    og = OracleGraph(_unitigs("F1"), 31)
    assert len(og.search("ACGT")) == 0                      # shorter than k
    assert len(og.search("N" * 100, False, True, True, True, True)) == 0
    u = _unitigs("F1")[0]
    h = og.search(u[:31])                                     # exactly one k-mer
    assert h.tolist() == [[0, 0, 0, 1]]
    og.close()
