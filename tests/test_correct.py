"""End-to-end parity of the per-read correction (getSeeds + correctSequence, pass 1) with the reference's own corrected
FASTQ records (tests/golden/*/corrected_pass1.fastq.gz, recorded from the unmodified reference and asserted equal to the
reference CLI's output when they were made): sequence AND quality strings byte for byte."""
import os

import pytest

import ratatosk_b200 as rb
from common import GOLDEN, golden_paths, load_golden_reads, read_fastq


def _run(ctx, recipe, idx):
    reads = load_golden_reads(recipe)
    gold = read_fastq(os.path.join(GOLDEN, recipe, "corrected_pass1.fastq.gz"))
    out = ctx.correct([reads[i][1] for i in idx], [reads[i][2] for i in idx])
    bad = [i for i, (cs, cq) in zip(idx, out) if (cs, cq) != (gold[i][1], gold[i][2])]
    assert not bad, (recipe, bad)
    return sum(len(reads[i][1]) for i in idx)


def test_correction_kernel_sources_match_reference_fastq(sim_lib):
    for recipe, idx in (("F1", [0, 17]), ("F2", [1])):
        fa, rt = golden_paths(recipe)
        g = rb.Graph.load(fa, rt, 31, lib=sim_lib)
        ctx = rb.Context(0, lib=sim_lib)
        ctx.upload(g)
        _run(ctx, recipe, idx)
        ctx.close()
        g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("recipe", ["F1", "F2"])
def test_correction_cuda_matches_reference_fastq(recipe):
    fa, rt = golden_paths(recipe)
    g = rb.Graph.load(fa, rt, 31)
    ctx = rb.Context(0)
    ctx.upload(g)
    n = len(load_golden_reads(recipe))
    _run(ctx, recipe, list(range(n)))
    # edge cases of the driver: empty batch entries, reads shorter than k, reads without any anchor
    out = ctx.correct(["", "ACGT", "N" * 100, "ACGTTGCA" * 20], ["", "IIII", "#" * 100, "5" * 160])
    assert out[0] == ("", "")
    assert out[1] == ("ACGT", "!!!!")
    assert out[2] == ("N" * 100, "!" * 100)
    assert out[3][0] == "ACGTTGCA" * 20 and set(out[3][1]) == {"!"}
    ctx.close()
    g.close()
