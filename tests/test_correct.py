"""End-to-end parity of the per-read correction (getSeeds + correctSequence) with the reference's own corrected FASTQ
records: sequence AND quality strings byte for byte.
  pass 1 (k = 31, short-read colours): tests/golden/*/corrected_pass1.fastq.gz, recorded from the unmodified reference and
         asserted equal to the reference CLI's output when they were made; bench_data/F3 at E. coli scale
  pass 2 (k = 63, long-read colours, exploreSubGraphLong bursts, qualities carried over): corrected_pass2_nophasing.fastq.gz =
         `Ratatosk correct -2 -c 1` on the pass-1 output (tests/golden/make_golden_pass2.sh), i.e. getSeeds + correctSequence
         without the multi-thread branch's phasing()"""
import os

import pytest

import ratatosk_b200 as rb
from common import GOLDEN, ROOT, golden_paths, load_golden_reads, read_fastq


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


@pytest.mark.gpu
def test_correction_cuda_matches_reference_cli_at_ecoli_scale():
    """BASELINE configs[1] scale: the 4.64 Mbp index built by the unmodified reference (bench_data/F3, 52,968 unitigs,
    6,365 shared colour sets) and the first 200 long reads of the F3 recipe; expected output = the reference CLI's
    `correct -1 -c 8` FASTQ.  The batch is submitted twice in one call (3,200-region broker run, every service batched)
    and in two different splits: batching must be transparent."""
    d = os.path.join(ROOT, "bench_data", "F3")
    g = rb.Graph.load(os.path.join(d, "index.k31.fasta.gz"), os.path.join(d, "index.k31.rtsk"), 31)
    ctx = rb.Context(0)
    ctx.upload(g)
    reads = read_fastq(os.path.join(d, "reads200.fastq.gz"))
    gold = read_fastq(os.path.join(d, "corrected200_pass1.fastq.gz"))
    assert len(reads) == len(gold) == 200
    seqs, quals = [r[1] for r in reads], [r[2] for r in reads]
    out = ctx.correct(seqs + seqs, quals + quals)
    bad = [i for i in range(400) if out[i] != (gold[i % 200][1], gold[i % 200][2])]
    assert not bad, bad[:10]
    part = ctx.correct(seqs[150:], quals[150:]) + ctx.correct(seqs[:3], quals[:3])
    assert part == [(gr[1], gr[2]) for gr in gold[150:] + gold[:3]]
    ctx.close()
    g.close()


def _run_pass2(ctx, recipe, idx):
    d = os.path.join(GOLDEN, recipe)
    p1 = read_fastq(os.path.join(d, "corrected_pass1.fastq.gz"))
    gold = read_fastq(os.path.join(d, "corrected_pass2_nophasing.fastq.gz"))
    out = ctx.correct([p1[i][1] for i in idx], [p1[i][2] for i in idx], pass_no=2)
    bad = [i for i, o in zip(idx, out) if o != (gold[i][1], gold[i][2])]
    assert not bad, (recipe, bad)
    return sum(1 for i in idx if (p1[i][1], p1[i][2]) != (gold[i][1], gold[i][2]))


def test_pass2_correction_kernel_sources_match_reference_fastq(sim_lib):
    d = os.path.join(GOLDEN, "F2")
    g = rb.Graph.load(os.path.join(d, "index.k63.fasta.gz"), os.path.join(d, "index.k63.rtsk"), 63, lib=sim_lib)
    ctx = rb.Context(0, lib=sim_lib)
    ctx.upload(g)
    assert _run_pass2(ctx, "F2", [15, 9]) == 2   # two (short) reads the second pass changes
    ctx.close()
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("recipe", ["F1", "F2"])
def test_pass2_correction_cuda_matches_reference_fastq(recipe):
    d = os.path.join(GOLDEN, recipe)
    g = rb.Graph.load(os.path.join(d, "index.k63.fasta.gz"), os.path.join(d, "index.k63.rtsk"), 63)
    ctx = rb.Context(0)
    ctx.upload(g)
    n = len(read_fastq(os.path.join(d, "corrected_pass1.fastq.gz")))
    changed = _run_pass2(ctx, recipe, list(range(n)))
    assert changed == (0 if recipe == "F1" else 18)   # what the reference's second pass changes on these fixtures
    ctx.close()
    g.close()


# ---- phasing (second pass, multi-thread branch of the reference): golden = phasing() recorded through the seam probe
# (oracle/ref_seams.cpp: ref_phasing) on the unmodified reference; phasing + correct(pass 2) = the CLI's `correct -2 -O -c 8`
def _phasing_inputs(recipe):
    d = os.path.join(GOLDEN, recipe)
    raw = read_fastq(os.path.join(d, "reads.fastq.gz"))
    p1 = read_fastq(os.path.join(d, "corrected_pass1.fastq.gz"))
    ph = read_fastq(os.path.join(d, "phasing.fastq.gz"))
    return raw, p1, ph


def test_phasing_kernel_sources_match_reference(sim_lib):
    d = os.path.join(GOLDEN, "F1")
    g = rb.Graph.load(os.path.join(d, "index.k63.fasta.gz"), os.path.join(d, "index.k63.rtsk"), 63, lib=sim_lib)
    ctx = rb.Context(0, lib=sim_lib)
    ctx.upload(g)
    raw, p1, ph = _phasing_inputs("F1")
    idx = [7, 0]   # read 7 is the one phasing() changes on this fixture
    out = ctx.phasing([raw[i][1].upper() for i in idx], [p1[i][1] for i in idx], [p1[i][2] for i in idx])
    assert out == [(ph[i][1], ph[i][2]) for i in idx]
    assert out[0] != (p1[7][1], p1[7][2])
    ctx.close()
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("recipe", ["F1", "F2"])
def test_two_pass_cuda_matches_reference_cli(recipe):
    """phasing + getSeeds + correctSequence of the second pass == `Ratatosk correct -2 -O -c 8` (corrected_pass2.fastq.gz),
    with the intermediate phasing() output checked on the way"""
    d = os.path.join(GOLDEN, recipe)
    g = rb.Graph.load(os.path.join(d, "index.k63.fasta.gz"), os.path.join(d, "index.k63.rtsk"), 63)
    ctx = rb.Context(0)
    ctx.upload(g)
    raw, p1, ph = _phasing_inputs(recipe)
    gold = read_fastq(os.path.join(d, "corrected_pass2.fastq.gz"))
    out = ctx.phasing([r[1].upper() for r in raw], [r[1] for r in p1], [r[2] for r in p1])
    bad = [i for i in range(len(p1)) if out[i] != (ph[i][1], ph[i][2])]
    assert not bad, ("phasing", recipe, bad)
    fin = ctx.correct([o[0] for o in out], [o[1] for o in out], pass_no=2)
    bad = [i for i in range(len(p1)) if fin[i] != (gold[i][1], gold[i][2])]
    assert not bad, ("two-pass", recipe, bad)
    ctx.close()
    g.close()


@pytest.mark.gpu
def test_two_pass_cuda_matches_reference_cli_at_ecoli_scale():
    """Second pass on the E. coli-scale k = 63 graph (bench_data/F3, 4.3 M k-mers): phasing + getSeeds + correctSequence of the
    200 pass-1 corrected reads == `Ratatosk correct -2 -O -c 8` (whole-read NW paths of ~10 kb x 10 kb go through the
    divide-and-conquer traceback; exploreSubGraphLong bursts on a real-size graph)."""
    d = os.path.join(ROOT, "bench_data", "F3")
    g = rb.Graph.load(os.path.join(d, "index.k63.fasta.gz"), os.path.join(d, "index.k63.rtsk"), 63)
    ctx = rb.Context(0)
    ctx.upload(g)
    raw = read_fastq(os.path.join(d, "reads200.fastq.gz"))
    p1 = read_fastq(os.path.join(d, "corrected200_pass1.fastq.gz"))
    gold = read_fastq(os.path.join(d, "corrected200_pass2.fastq.gz"))
    ph = ctx.phasing([r[1].upper() for r in raw], [r[1] for r in p1], [r[2] for r in p1])
    gph = read_fastq(os.path.join(d, "phasing200.fastq.gz"))   # phasing() of the reference through the seam probe
    bad = [i for i in range(len(p1)) if ph[i] != (gph[i][1], gph[i][2])]
    assert not bad, ("phasing", bad[:10])
    fin = ctx.correct([o[0] for o in ph], [o[1] for o in ph], pass_no=2)
    bad = [i for i in range(len(p1)) if fin[i] != (gold[i][1], gold[i][2])]
    assert not bad, bad[:10]
    assert sum(1 for i in range(len(p1)) if fin[i] != (p1[i][1], p1[i][2])) > 10   # the second pass does change reads here
    ctx.close()
    g.close()


@pytest.mark.gpu
def test_correction_cuda_matches_reference_library_on_fresh_reads():
    """Differential test against the UNMODIFIED reference objects (oracle/_ref/libref_seams.so travels to the GPU box): a fresh
    seeded batch of ONT-like reads drawn from the F3 genome with bench.py's generator (~1.5 Mbases, not a committed fixture),
    corrected by both; skipped when the reference library was not built.  The reference breaks ties between colour sets of
    equal cardinality by pointer hash (tests/ref_worker.py), so its output is collected from four processes and a read may
    match any of them; at least 99 % of the reads must be identical in all."""
    import pickle
    import subprocess
    import sys
    import tempfile
    sys.path.insert(0, ROOT)
    import refseams as R
    if not R.available():
        pytest.skip("oracle/_ref/libref_seams.so not built")
    import bench
    haps = bench.load_haplotypes()
    seq, qual, off = bench.make_reads(haps, 1_500_000, seed=20261017)
    reads = [(seq[int(off[i]):int(off[i + 1])].tobytes().decode(), qual[int(off[i]):int(off[i + 1])].tobytes().decode())
             for i in range(len(off) - 1)]
    d = os.path.join(ROOT, "bench_data", "F3")
    fa, rt = os.path.join(d, "index.k31.fasta.gz"), os.path.join(d, "index.k31.rtsk")
    tmp = tempfile.mkdtemp(prefix="rtk_ref_")
    pickle.dump(reads, open(os.path.join(tmp, "reads.pkl"), "wb"))
    variants = []
    for i in range(4):
        dst = os.path.join(tmp, "out%d.pkl" % i)
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tests", "ref_worker.py"), fa, rt, "31", os.path.join(tmp, "reads.pkl"), dst])
        variants.append(pickle.load(open(dst, "rb")))
    g = rb.Graph.load(fa, rt, 31)
    ctx = rb.Context(0)
    ctx.upload(g)
    got = ctx.correct([r[0] for r in reads], [r[1] for r in reads])
    unstable = [i for i in range(len(reads)) if any(v[i] != variants[0][i] for v in variants)]
    bad = [i for i in range(len(reads)) if all(got[i] != v[i] for v in variants)]
    # A read whose tie the reference breaks by pointer order can come out in a variant none of the four runs happened to produce;
    # the strict pins are the committed golden files above.  Here: at most 1 % of the reads may be of that kind.
    if bad:
        print("differs from every reference run:", bad[:10], "| reads on which the reference runs disagree:", unstable)
    assert len(bad) <= max(1, len(reads) // 100), (bad[:10], unstable)
    assert len(unstable) <= 0.03 * len(reads), unstable
    ctx.close()
    g.close()


F4 = os.path.join(ROOT, "bench_data", "F4")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(F4, "index.k31.rtsk")), reason="bench_data/F4 not generated (scripts/make_f4.sh)")
def test_two_pass_cuda_matches_reference_cli_at_chr20_scale():
    """BASELINE configs[2] scale: the 64 Mbp diploid index built by the unmodified reference (bench_data/F4: k = 31 graph
    coloured by 30x PE150, k = 63 graph coloured by 3x pass-1-corrected long reads; slabs of 1.3 GB / 0.9 GB, far beyond the
    L2) and the first 200 long reads of the recipe.  Expected = the reference CLI's own files: `correct -1` and
    `correct -2 -O` (phasing + second pass)."""
    raw = read_fastq(os.path.join(F4, "reads200.fastq.gz"))
    gold1 = read_fastq(os.path.join(F4, "corrected200_pass1.fastq.gz"))
    gold2 = read_fastq(os.path.join(F4, "corrected200_pass2.fastq.gz"))
    g = rb.Graph.load(os.path.join(F4, "index.k31.fasta.gz"), os.path.join(F4, "index.k31.rtsk"), 31)
    ctx = rb.Context(0)
    ctx.upload(g)
    p1 = ctx.correct([r[1] for r in raw], [r[2] for r in raw])
    bad = [i for i in range(len(raw)) if p1[i] != (gold1[i][1], gold1[i][2])]
    assert len(bad) <= 2, ("pass 1", bad[:10])   # colour-set ties the reference breaks by pointer order (see the fresh-reads test)
    ctx.close()
    g.close()
    g = rb.Graph.load(os.path.join(F4, "index.k63.fasta.gz"), os.path.join(F4, "index.k63.rtsk"), 63)
    ctx = rb.Context(0)
    ctx.upload(g)
    ph = ctx.phasing([r[1].upper() for r in raw], [r[1] for r in gold1], [r[2] for r in gold1])
    fin = ctx.correct([o[0] for o in ph], [o[1] for o in ph], pass_no=2)
    bad = [i for i in range(len(raw)) if fin[i] != (gold2[i][1], gold2[i][2])]
    assert len(bad) <= 2, ("pass 2", bad[:10])
    ctx.close()
    g.close()


# ---- both passes as one pipeline (rtk_correct_two_pass_batch): same bytes as pass 1, phasing, pass 2 called one after the other
def _two_pass_fused(recipe, lib, idx=None, gangs=None):
    d = os.path.join(GOLDEN, recipe)
    g1 = rb.Graph.load(os.path.join(d, "index.k31.fasta.gz"), os.path.join(d, "index.k31.rtsk"), 31, lib=lib)
    g2 = rb.Graph.load(os.path.join(d, "index.k63.fasta.gz"), os.path.join(d, "index.k63.rtsk"), 63, lib=lib)
    c1, c2 = rb.Context(0, lib=lib), rb.Context(0, lib=lib)
    c1.upload(g1); c2.upload(g2)
    raw = read_fastq(os.path.join(d, "reads.fastq.gz"))
    gp1 = read_fastq(os.path.join(d, "corrected_pass1.fastq.gz"))
    gp2 = read_fastq(os.path.join(d, "corrected_pass2.fastq.gz"))
    idx = list(range(len(raw))) if idx is None else idx
    if gangs:
        os.environ["RTK_GANGS2"] = str(gangs)
    try:
        fin, p1 = c1.correct_two_pass(c2, [raw[i][1] for i in idx], [raw[i][2] for i in idx], want_pass1=True)
    finally:
        os.environ.pop("RTK_GANGS2", None)
    bad1 = [i for j, i in enumerate(idx) if p1[j] != (gp1[i][1], gp1[i][2])]
    bad2 = [i for j, i in enumerate(idx) if fin[j] != (gp2[i][1], gp2[i][2])]
    c1.close(); c2.close(); g1.close(); g2.close()
    return bad1, bad2


def test_two_pass_pipeline_kernel_sources_match_reference_cli(sim_lib):
    """reads 7 (changed by phasing) and 0-2 of F1 through the fused pipeline on the simulator == the reference CLI's two files"""
    assert _two_pass_fused("F1", sim_lib, idx=[7, 0, 1, 2]) == ([], [])


@pytest.mark.gpu
@pytest.mark.parametrize("recipe,gangs", [("F1", None), ("F2", None), ("F2", 3)])
def test_two_pass_pipeline_cuda_matches_reference_cli(recipe, gangs):
    """whole fixture through rtk_correct_two_pass_batch (several gangs forced on the small fixture): <out>.2.fastq and <out>.fastq
    of `Ratatosk correct -1` + `correct -2 -O`"""
    os.environ["RTK_GANGS_MIN_BASES"] = "1"
    try:
        assert _two_pass_fused(recipe, None, gangs=gangs) == ([], [])
    finally:
        os.environ.pop("RTK_GANGS_MIN_BASES", None)
