"""fixSNPs (src/Alignment.cpp:846-964; `Ratatosk correct -2 -f`): the device kernel (ratatosk_b200/csrc/fixsnps.cuh) against
golden vectors recorded from the unmodified reference through the seam probe ref_fix_snps and the CLI
(tests/golden/make_golden_fixsnps.py).  CPU: the kernel source on the simulator; GPU: the product library."""
import gzip
import json
import os

import pytest

import ratatosk_b200 as rb
from common import GOLDEN, ROOT, read_fastq

F3 = os.path.join(ROOT, "bench_data", "F3")
CASES = {"F2": (os.path.join(GOLDEN, "F2"), "corrected_pass1.fastq.gz", "reads.fastq.gz"),
         "F3": (F3, "corrected200_pass1.fastq.gz", "reads200.fastq.gz")}


def _apply(reads, changes):
    out = []
    for i, s in enumerate(reads):
        b = list(s)
        for pos, base in changes.get(str(i), []):
            b[pos] = base
        out.append("".join(b))
    return out


def _check_fix(recipe, lib):
    d, p1, _ = CASES[recipe]
    g = rb.Graph.load(os.path.join(d, "index.k63.fasta.gz"), os.path.join(d, "index.k63.rtsk"), 63, lib=lib)
    ctx = rb.Context(0, lib=lib)
    ctx.upload(g)
    reads = [r[1] for r in read_fastq(os.path.join(d, p1))]
    gold = json.load(gzip.open(os.path.join(d, "fixsnps.json.gz"), "rt"))
    assert gold["n_fixed"] > 20 and gold["n_ambiguous"] > gold["n_fixed"]   # the case exercises both outcomes
    out, n_fixed = ctx.fix_snps(reads)
    want = _apply(reads, gold["changes"])
    bad = [i for i in range(len(reads)) if out[i] != want[i]]
    assert not bad, (recipe, bad[:10])
    assert n_fixed == gold["n_fixed"]
    # batching is transparent and a second application of the step is what the reference gives on its own output
    one_by_one = [ctx.fix_snps([r])[0][0] for r in reads[:6]]
    assert one_by_one == want[:6]
    ctx.close()
    g.close()


@pytest.mark.parametrize("recipe", ["F2", "F3"])
def test_fixsnps_kernel_source_matches_reference(recipe, sim_lib):
    _check_fix(recipe, sim_lib)


def test_fixsnps_edge_cases(sim_lib):
    """reads shorter than k, without codes, empty; codes whose window holds >= 64 combinations (left alone); foreign characters"""
    d = os.path.join(GOLDEN, "F2")
    g = rb.Graph.load(os.path.join(d, "index.k63.fasta.gz"), os.path.join(d, "index.k63.rtsk"), 63, lib=sim_lib)
    ctx = rb.Context(0, lib=sim_lib)
    ctx.upload(g)
    u = g.unitig_seq(0)
    assert len(u) >= 63
    km = u[:63]
    one = km[:30] + "N" + km[31:]                    # one code, the graph k-mer resolves it
    many = "NNN" + km[3:]                            # 4*4*4 = 64 combinations: not tried
    foreign = km[:10] + "." + km[11:20] + "N" + km[21:]   # '.' zeroes the product: nothing is valid
    reads = ["", "ACGT", km, one, many, foreign, one[:40]]
    out, n = ctx.fix_snps(reads)
    assert out[:3] == reads[:3]
    assert out[3] == km and n == 1
    assert out[4] == many and out[5] == foreign and out[6] == one[:40]
    ctx.close()
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("recipe", ["F2", "F3"])
def test_fixsnps_cuda_matches_reference(recipe):
    _check_fix(recipe, None)


@pytest.mark.gpu
@pytest.mark.parametrize("recipe", ["F2", "F3"])
def test_two_pass_force_snp_cuda_matches_reference_cli(recipe):
    """rtk_opt.force_unres_snp_corr: fixSNPs + phasing + getSeeds + correctSequence == `Ratatosk correct -2 -O --force-correct-snp -c 8`"""
    d, p1f, rawf = CASES[recipe]
    g = rb.Graph.load(os.path.join(d, "index.k63.fasta.gz"), os.path.join(d, "index.k63.rtsk"), 63)
    ctx = rb.Context(0)
    ctx.upload(g)
    raw = read_fastq(os.path.join(d, rawf))
    p1 = read_fastq(os.path.join(d, p1f))
    gold = read_fastq(os.path.join(d, "corrected_pass2_forcesnp.fastq.gz"))
    plain = read_fastq(os.path.join(d, "corrected_pass2.fastq.gz" if recipe == "F2" else "corrected200_pass2.fastq.gz"))
    opt = rb.default_opt(2)
    opt.force_unres_snp_corr = 1
    ph = ctx.phasing([r[1].upper() for r in raw], [r[1] for r in p1], [r[2] for r in p1], opt=opt)
    fin = ctx.correct([o[0] for o in ph], [o[1] for o in ph], opt=opt, pass_no=2)
    bad = [i for i in range(len(p1)) if fin[i] != (gold[i][1], gold[i][2])]
    assert not bad, (recipe, bad[:10])
    assert any(gold[i][1] != plain[i][1] for i in range(len(gold)))   # -f does change the outcome on this fixture
    ctx.close()
    g.close()
