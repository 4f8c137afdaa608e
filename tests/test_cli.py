"""The host driver (ratatosk_b200/host/rtk_cli.cpp): FASTQ -> library -> FASTQ, compared as FILES with what the
unmodified reference CLI writes (`Ratatosk correct -1` -> <out>.2.fastq, `correct -2 -O` -> <out>.fastq; src/Ratatosk.cpp:
510-616 writeCorrectedOutput incl. -t trimming, :727-906 ticket loop, :919-999 block re-ordering).

CPU: the same driver source linked against the kernel-source simulator library (tests/hostsim), a few reads.
GPU: the product binary ratatosk_b200/rtk_correct on whole fixtures, against the committed golden files (recorded from the
reference CLI) and, where the reference binary travelled to the box (oracle/_ref), against a fresh reference run with -t."""
import gzip
import os
import subprocess

import pytest

from common import GOLDEN, HERE, ROOT, ensure_built

SIM_CLI = os.path.join(HERE, "hostsim", "_build", "rtk_correct_sim")
GPU_CLI = os.path.join(ROOT, "ratatosk_b200", "rtk_correct")
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "Ratatosk")
F3 = os.path.join(ROOT, "bench_data", "F3")


def _head_fastq(src_gz, dst, n_reads):
    with gzip.open(src_gz, "rb") as f, open(dst, "wb") as o:
        for _ in range(4 * n_reads):
            line = f.readline()
            if not line:
                break
            o.write(line)


def _golden_bytes(path_gz, n_reads=None):
    with gzip.open(path_gz, "rb") as f:
        data = f.read()
    if n_reads is None:
        return data
    return b"\n".join(data.split(b"\n")[:4 * n_reads]) + b"\n"


def _run(cli, args):
    r = subprocess.run([cli, "correct"] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert r.returncode == 0, r.stdout.decode(errors="replace")[-2000:]
    return r.stdout.decode(errors="replace")


def _two_pass(cli, d, tmp, reads, extra1=(), extra2=(), tag="ours"):
    o = os.path.join(tmp, tag)
    _run(cli, ["-1", "-g", os.path.join(d, "index.k31.fasta.gz"), "-d", os.path.join(d, "index.k31.rtsk"), "-l", reads, "-o", o] + list(extra1))
    _run(cli, ["-2", "-O", "-c", "8", "-g", os.path.join(d, "index.k63.fasta.gz"), "-d", os.path.join(d, "index.k63.rtsk"), "-l", o + ".2.fastq", "-L", reads,
               "-o", o] + list(extra2))
    return o + ".2.fastq", o + ".fastq"


@pytest.fixture(scope="module")
def sim_cli():
    ensure_built()
    if not os.path.exists(SIM_CLI):
        subprocess.check_call(["make", "-C", os.path.join(HERE, "hostsim")])
    return SIM_CLI


def test_cli_two_pass_files_match_reference_cli_on_simulator(sim_cli, tmp_path):
    """12 F1 reads, tickets of ~20 kb dealt to two contexts: <out>.2.fastq and <out>.fastq equal the reference CLI's files"""
    d = os.path.join(GOLDEN, "F1")
    reads = str(tmp_path / "r.fastq")
    _head_fastq(os.path.join(d, "reads.fastq.gz"), reads, 12)
    p1, p2 = _two_pass(sim_cli, d, str(tmp_path), reads, extra1=["--ticket-bases", "20000", "--gpus", "2"], extra2=["--ticket-bases", "20000", "--gpus", "2"])
    assert open(p1, "rb").read() == _golden_bytes(os.path.join(d, "corrected_pass1.fastq.gz"), 12)
    assert open(p2, "rb").read() == _golden_bytes(os.path.join(d, "corrected_pass2.fastq.gz"), 12)


def test_cli_reads_gzip_and_multiline_input_and_writes_gzip(sim_cli, tmp_path):
    """gz input, a FASTQ with wrapped sequence / quality lines, header comments dropped (kseq name), -G output"""
    d = os.path.join(GOLDEN, "F1")
    plain = str(tmp_path / "r.fastq")
    _head_fastq(os.path.join(d, "reads.fastq.gz"), plain, 4)
    recs = open(plain).read().split("\n")
    wrapped = str(tmp_path / "wrapped_in.fastq.gz")
    with gzip.open(wrapped, "wt") as f:
        for i in range(0, len(recs) - 3, 4):
            s, q = recs[i + 1], recs[i + 3]
            f.write(recs[i] + " some comment\n")
            f.write("\n".join(s[j:j + 70] for j in range(0, len(s), 70)) + "\n+" + recs[i][1:] + "\n")
            f.write("\n".join(q[j:j + 70] for j in range(0, len(q), 70)) + "\n")
    o = str(tmp_path / "w")
    _run(sim_cli, ["-1", "-g", os.path.join(d, "index.k31.fasta.gz"), "-d", os.path.join(d, "index.k31.rtsk"), "-l", wrapped, "-o", o])
    assert open(o + ".2.fastq", "rb").read() == _golden_bytes(os.path.join(d, "corrected_pass1.fastq.gz"), 4)
    _run(sim_cli, ["-2", "-G", "-c", "8", "-g", os.path.join(d, "index.k63.fasta.gz"), "-d", os.path.join(d, "index.k63.rtsk"), "-l", o + ".2.fastq", "-L", wrapped,
                   "-o", o, "--ticket-bases", "10000"])
    assert gzip.open(o + ".fastq.gz", "rb").read() == _golden_bytes(os.path.join(d, "corrected_pass2.fastq.gz"), 4)


def _trim_reference(name, seq, qual, k, trim):
    """writeCorrectedOutput with trim > 0, restated (src/Ratatosk.cpp:518-555)"""
    out, start, run, sub = [], -1, -1, 1
    cmin = chr(trim + 33)
    for p, c in enumerate(qual):
        if c >= cmin:
            if start == -1:
                start, run = p, 0
            run += 1
        else:
            if run >= k:
                out.append("@%s/%d\n%s\n+\n%s\n" % (name, sub, seq[start:start + run], qual[start:start + run]))
                sub += 1
            start, run = -1, -1
    if run >= k:
        out.append("@%s/%d\n%s\n+\n%s\n" % (name, sub, seq[start:start + run], qual[start:start + run]))
    return "".join(out)


def test_cli_trim_split_matches_write_corrected_output(sim_cli, tmp_path):
    """-t 20 on pass 2: sub-reads name/1, name/2 ... of >= k bases at or above Q20, derived from the golden untrimmed output"""
    d = os.path.join(GOLDEN, "F1")
    reads = str(tmp_path / "r.fastq")
    _head_fastq(os.path.join(d, "reads.fastq.gz"), reads, 6)
    _, p2 = _two_pass(sim_cli, d, str(tmp_path), reads, extra2=["-t", "20"])
    gold = _golden_bytes(os.path.join(d, "corrected_pass2.fastq.gz"), 6).decode().split("\n")
    want = "".join(_trim_reference(gold[i][1:], gold[i + 1], gold[i + 3], 63, 20) for i in range(0, len(gold) - 3, 4))
    got = open(p2).read()
    assert got == want
    assert "/2\n" in got   # the case splits at least one read


def test_cli_rejects_bad_command_lines(sim_cli, tmp_path):
    d = os.path.join(GOLDEN, "F1")
    base = ["-g", os.path.join(d, "index.k31.fasta.gz"), "-d", os.path.join(d, "index.k31.rtsk"), "-l", "x.fastq", "-o", str(tmp_path / "o")]
    for args in (["-1", "-2"] + base, base, ["-1", "-t", "99"] + base, ["-2"] + base, ["-1", "-g", "a", "-l", "x", "-o", "y"]):
        r = subprocess.run([sim_cli, "correct"] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        assert r.returncode != 0
    # pass 2 with raw reads in another order: the reference aborts (src/Ratatosk.cpp:788-799)
    reads = str(tmp_path / "r.fastq")
    _head_fastq(os.path.join(d, "reads.fastq.gz"), reads, 2)
    L = open(reads).read().split("\n")
    swapped = str(tmp_path / "s.fastq")
    open(swapped, "w").write("\n".join(L[4:8] + L[0:4]) + "\n")
    r = subprocess.run([sim_cli, "correct", "-2", "-g", os.path.join(d, "index.k63.fasta.gz"), "-d", os.path.join(d, "index.k63.rtsk"), "-l", reads, "-L", swapped,
                        "-o", str(tmp_path / "o")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert r.returncode != 0 and b"not in the same order" in r.stdout


# ----------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("recipe", ["F1", "F2"])
def test_cli_cuda_files_identical_to_reference_cli(recipe, tmp_path):
    """whole fixture through ratatosk_b200/rtk_correct: both output FILES equal the reference CLI's (golden, recorded from
    `Ratatosk correct -1 -c 8` and `correct -2 -O -c 8`)"""
    d = os.path.join(GOLDEN, recipe)
    reads = str(tmp_path / "reads.fastq")
    with gzip.open(os.path.join(d, "reads.fastq.gz"), "rb") as f:
        open(reads, "wb").write(f.read())
    p1, p2 = _two_pass(GPU_CLI, d, str(tmp_path), reads, extra1=["--ticket-bases", "200000"], extra2=["--ticket-bases", "200000"])
    assert open(p1, "rb").read() == _golden_bytes(os.path.join(d, "corrected_pass1.fastq.gz"))
    assert open(p2, "rb").read() == _golden_bytes(os.path.join(d, "corrected_pass2.fastq.gz"))


@pytest.mark.gpu
def test_cli_cuda_ecoli_scale_files_identical_to_reference_cli(tmp_path):
    """E. coli-scale indexes (bench_data/F3), 200 reads: cmp against the reference CLI's files"""
    reads = str(tmp_path / "reads.fastq")
    with gzip.open(os.path.join(F3, "reads200.fastq.gz"), "rb") as f:
        open(reads, "wb").write(f.read())
    p1, p2 = _two_pass(GPU_CLI, F3, str(tmp_path), reads)
    assert open(p1, "rb").read() == _golden_bytes(os.path.join(F3, "corrected200_pass1.fastq.gz"))
    assert open(p2, "rb").read() == _golden_bytes(os.path.join(F3, "corrected200_pass2.fastq.gz"))


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF_CLI), reason="oracle/_ref/Ratatosk did not travel to this box")
def test_cli_cuda_trim_and_gzip_against_fresh_reference_run(tmp_path):
    """-t 15 and -G on F2 against the reference binary run here with the same flags"""
    d = os.path.join(GOLDEN, "F2")
    reads = str(tmp_path / "reads.fastq")
    with gzip.open(os.path.join(d, "reads.fastq.gz"), "rb") as f:
        open(reads, "wb").write(f.read())
    p1, p2 = _two_pass(GPU_CLI, d, str(tmp_path), reads, extra2=["-t", "15", "-G"])
    ref = str(tmp_path / "ref")
    cores = str(max(2, min(8, os.cpu_count() or 2)))
    subprocess.check_call([REF_CLI, "correct", "-2", "-O", "-G", "-t", "15", "-c", cores, "-g", os.path.join(d, "index.k63.fasta.gz"), "-d",
                           os.path.join(d, "index.k63.rtsk"), "-l", p1, "-L", reads, "-o", ref], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    assert gzip.open(p2 + ".gz", "rb").read() == gzip.open(ref + ".fastq.gz", "rb").read()


@pytest.mark.gpu
def test_cli_cuda_force_snp_correction_matches_reference_cli(tmp_path):
    """--force-correct-snp (fixSNPs before phasing, src/Ratatosk.cpp:828) on F2 == the reference CLI with the same flag"""
    d = os.path.join(GOLDEN, "F2")
    reads, p1 = str(tmp_path / "reads.fastq"), str(tmp_path / "p1.fastq")
    open(reads, "wb").write(gzip.open(os.path.join(d, "reads.fastq.gz"), "rb").read())
    open(p1, "wb").write(gzip.open(os.path.join(d, "corrected_pass1.fastq.gz"), "rb").read())
    o = str(tmp_path / "o")
    _run(GPU_CLI, ["-2", "--force-correct-snp", "-c", "8", "-g", os.path.join(d, "index.k63.fasta.gz"), "-d", os.path.join(d, "index.k63.rtsk"), "-l", p1, "-L", reads, "-o", o])
    assert open(o + ".fastq", "rb").read() == _golden_bytes(os.path.join(d, "corrected_pass2_forcesnp.fastq.gz"))


def _two_pass_mode(cli, d, tmp, reads, extra=()):
    o = os.path.join(tmp, "tp")
    _run(cli, ["-g", os.path.join(d, "index.k31.fasta.gz"), "-d", os.path.join(d, "index.k31.rtsk"), "--in-graph2", os.path.join(d, "index.k63.fasta.gz"),
               "--in-unitig-data2", os.path.join(d, "index.k63.rtsk"), "-l", reads, "-o", o] + list(extra))
    return o + ".fastq"


def test_cli_two_pass_mode_equals_the_two_reference_commands_on_simulator(sim_cli, tmp_path):
    """both indexes given, neither -1 nor -2: one pipelined library call per ticket (rtk_correct_two_pass_batch); the file equals
    what `Ratatosk correct -1` followed by `correct -2 -O` writes"""
    d = os.path.join(GOLDEN, "F1")
    reads = str(tmp_path / "r.fastq")
    _head_fastq(os.path.join(d, "reads.fastq.gz"), reads, 9)
    out = _two_pass_mode(sim_cli, d, str(tmp_path), reads, extra=["--ticket-bases", "25000"])
    assert open(out, "rb").read() == _golden_bytes(os.path.join(d, "corrected_pass2.fastq.gz"), 9)


@pytest.mark.gpu
@pytest.mark.parametrize("recipe", ["F1", "F2"])
def test_cli_cuda_two_pass_mode_identical_to_reference_cli(recipe, tmp_path):
    d = os.path.join(GOLDEN, recipe)
    reads = str(tmp_path / "reads.fastq")
    open(reads, "wb").write(gzip.open(os.path.join(d, "reads.fastq.gz"), "rb").read())
    out = _two_pass_mode(GPU_CLI, d, str(tmp_path), reads)
    assert open(out, "rb").read() == _golden_bytes(os.path.join(d, "corrected_pass2.fastq.gz"))


# ---------------------------------------------------------------------------------------------- rtk_correct annotate
def _dump_lines(fasta, rtsk, k, tmp, tag):
    from refseams import RefGraph
    g = RefGraph(fasta, rtsk, k)
    p = os.path.join(tmp, tag + ".dump")
    g.dump(p)
    g.close()
    return sorted(open(p).read().split("\n"))


def _check_annotate_cli(cli, tmp, lib):
    """`rtk_correct annotate` (detectSNPs + detectShortCycles + the .rtsk fields they own) on the F2 k = 31 index: an index stripped
    of its annotations gets them back; the REFERENCE reads the rewritten file as the same graph (per-unitig dump through the seam
    library: flags, colours, ambiguity characters, cycles) and corrects reads with it to the same bytes."""
    import ratatosk_b200 as rb
    import refseams
    d = os.path.join(GOLDEN, "F2")
    fa, rt = os.path.join(d, "index.k31.fasta.gz"), os.path.join(d, "index.k31.rtsk")
    stripped, redone = os.path.join(tmp, "stripped.rtsk"), os.path.join(tmp, "redone.rtsk")
    r = subprocess.run([cli, "annotate", "-g", fa, "-d", rt, "-o", stripped, "-k", "31", "--no-snp", "--min-cov", "1000000000"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert r.returncode == 0, r.stdout.decode(errors="replace")[-2000:]
    r = subprocess.run([cli, "annotate", "-g", fa, "-d", stripped, "-o", redone, "-k", "31", "-v"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert r.returncode == 0, r.stdout.decode(errors="replace")[-2000:]
    assert b"2997 SNP marks" in r.stdout and b"250 unitigs in short cycles" in r.stdout, r.stdout
    g0, g1, g2 = (rb.Graph.load(fa, x, 31, lib=lib) for x in (rt, stripped, redone))
    n = g0.info()["n_unitigs"]
    assert all(g1.unitig_annotations(u) == ([], b"") and not (g1.unitig_words(u)[1] >> 8) & 1 for u in range(n))
    assert all(g2.unitig_annotations(u) == g0.unitig_annotations(u) and g2.unitig_words(u) == g0.unitig_words(u)
               and g2.unitig_colors(u) == g0.unitig_colors(u) for u in range(n))
    for g in (g0, g1, g2):
        g.close()
    if refseams.available():
        want = _dump_lines(fa, rt, 31, tmp, "orig")
        assert _dump_lines(fa, redone, 31, tmp, "redone") == want
        assert _dump_lines(fa, stripped, 31, tmp, "stripped") != want
    if os.path.exists(REF_CLI):
        reads = os.path.join(tmp, "r.fastq")
        _head_fastq(os.path.join(d, "reads.fastq.gz"), reads, 1000)   # all 26 reads
        o = os.path.join(tmp, "ref_on_redone")
        subprocess.check_call([REF_CLI, "correct", "-1", "-c", "4", "-g", fa, "-d", redone, "-l", reads, "-o", o], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        assert open(o + ".2.fastq", "rb").read() == _golden_bytes(os.path.join(d, "corrected_pass1.fastq.gz"))


def test_cli_annotate_rewrites_index_the_reference_accepts(sim_cli, sim_lib, tmp_path):
    _check_annotate_cli(sim_cli, str(tmp_path), sim_lib)


@pytest.mark.gpu
def test_cli_annotate_cuda_rewrites_index_the_reference_accepts(tmp_path):
    _check_annotate_cli(GPU_CLI, str(tmp_path), None)


# ---------------------------------------------------------------------------------------------- rtk_correct index2
def _dump_records(fasta, rtsk, k, tmp, tag):
    """per unitig sequence -> (kmcov, shared, colour set, ambiguity chars, cycles, branching, short-cycle) as the REFERENCE reads the index"""
    from refseams import RefGraph
    g = RefGraph(fasta, rtsk, k)
    p = os.path.join(tmp, tag + ".dump")
    g.dump(p)
    g.close()
    out = {}
    for line in open(p):
        f = line.rstrip("\n").split("\t")
        ids = set(int(x) for x in (f[4] + f[5]).split(",") if x)
        out[f[1]] = (int(f[2]), int(f[3]), ids, f[6], f[7], f[8], f[9], f[10])
    return out


def _check_index2_cli(cli, tmp):
    """`rtk_correct index2` = `Ratatosk index -2` (colour the k2 graph with the pass-1 reads, detectSNPs, detectShortCycles, write the
    .rtsk).  The REFERENCE reads both files (readGraphData); per unitig everything must agree with the index the reference built from
    the same inputs - coverage word, flags word (edge flags + short-cycle bit), ambiguity characters, cycles - and the colour sets under
    one relabelling of the read ids (the reference's own ids depend on thread timing)."""
    import refseams
    if not refseams.available():
        pytest.skip("reference seam library not built")
    d = os.path.join(GOLDEN, "F2")
    fa = os.path.join(d, "index.k63.fasta.gz")
    reads = os.path.join(tmp, "p1.fastq")
    _head_fastq(os.path.join(d, "corrected_pass1.fastq.gz"), reads, 1000)
    r = subprocess.run([cli, "index2", "-g", fa, "-l", reads, "-o", os.path.join(tmp, "ours"), "-v"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert r.returncode == 0, r.stdout.decode(errors="replace")[-2000:]
    ours = _dump_records(fa, os.path.join(tmp, "ours.index.k63.rtsk"), 63, tmp, "ours")
    want = _dump_records(fa, os.path.join(d, "index.k63.rtsk"), 63, tmp, "want")
    assert set(ours) == set(want) and len(want) == 1926
    assert all(ours[s][:2] == want[s][:2] and ours[s][3:] == want[s][3:] for s in want), [s[:20] for s in want if ours[s][:2] != want[s][:2] or ours[s][3:] != want[s][3:]][:5]
    assert sum(1 for s in want if want[s][3]) > 50 and sum(1 for s in want if want[s][5]) > 50      # SNP marks and cycles are there
    m, changed = {}, True
    while changed:
        changed = False
        for s in want:
            a, b = ours[s][2], want[s][2]
            assert len(a) == len(b)
            un = [x for x in a if x not in m]
            rem = b - set(m[x] for x in a if x in m)
            if len(un) == 1 and len(rem) == 1:
                m[un[0]] = next(iter(rem))
                changed = True
    assert len(m) == 23 and len(set(m.values())) == 23
    assert all(set(m[x] for x in ours[s][2]) == want[s][2] for s in want)


def test_cli_index2_writes_the_index_the_reference_builds(sim_cli, tmp_path):
    _check_index2_cli(sim_cli, str(tmp_path))


@pytest.mark.gpu
def test_cli_index2_cuda_writes_the_index_the_reference_builds(tmp_path):
    _check_index2_cli(GPU_CLI, str(tmp_path))


def _mutated_reads(path, seed):
    """pass-1 reads of F2 made awkward: random '!' qualities (masked bases), lower case, IUPAC codes, lengths around min_len_2nd_pass,
    a duplicated name, a read without any graph k-mer"""
    import numpy as np
    from common import read_fastq
    rng = np.random.RandomState(seed)
    recs = read_fastq(os.path.join(GOLDEN, "F2", "corrected_pass1.fastq.gz"))
    out = []
    for i, (name, s, q) in enumerate(recs):
        s, q = list(s), list(q)
        for j in rng.randint(0, len(s), len(s) // 50):
            q[j] = "!"
        for j in rng.randint(0, len(s), 3):
            s[j] = "RYKMSWN"[int(rng.randint(7))]
        if i % 3 == 0:
            s = [c.lower() for c in s]
        if i % 5 == 1:
            cut = (2999, 3000, 3001, 200)[(i // 5) % 4]
            s, q = s[:cut], q[:cut]
        out.append((name, "".join(s), "".join(q)))
    out.append((out[2][0], out[7][1], out[7][2]))                       # the name of read 2 again, with other bases
    out.append(("junk", "ACGT" * 1000, "I" * 4000))
    with open(path, "w") as f:
        for name, s, q in out:
            f.write("@%s\n%s\n+\n%s\n" % (name, s, q))


@pytest.mark.parametrize("seed,extra", [(1, []), (2, ["-M", "0.5", "-C", "1000"])])
def test_cli_index2_matches_reference_cli_on_awkward_reads(seed, extra, sim_cli, tmp_path):
    """`rtk_correct index2` against a fresh `Ratatosk index -2` run on the same awkward reads (and with -M / -C): per unitig the
    reference reads the same words, marks and cycles from both files, and the colour sets agree under one relabelling"""
    import refseams
    if not (refseams.available() and os.path.exists(REF_CLI)):
        pytest.skip("reference not built")
    tmp = str(tmp_path)
    fa = os.path.join(GOLDEN, "F2", "index.k63.fasta.gz")
    reads = os.path.join(tmp, "reads.fastq")
    _mutated_reads(reads, seed)
    subprocess.check_call([REF_CLI, "index", "-2", "-c", "4", "-g", fa, "-l", reads, "-o", os.path.join(tmp, "ref")] + extra,
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    r = subprocess.run([sim_cli, "index2", "-g", fa, "-l", reads, "-o", os.path.join(tmp, "ours")] + extra, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert r.returncode == 0, r.stdout.decode(errors="replace")[-2000:]
    ours = _dump_records(fa, os.path.join(tmp, "ours.index.k63.rtsk"), 63, tmp, "ours")
    want = _dump_records(fa, os.path.join(tmp, "ref.index.k63.rtsk"), 63, tmp, "want")
    bad = [s[:24] for s in want if ours[s][:2] != want[s][:2] or ours[s][3:] != want[s][3:] or len(ours[s][2]) != len(want[s][2])]
    assert not bad, (len(bad), bad[:5])
    m, changed = {}, True
    while changed:
        changed = False
        for s in want:
            a, b = ours[s][2], want[s][2]
            un = [x for x in a if x not in m]
            rem = b - set(m[x] for x in a if x in m)
            if len(un) == 1 and len(rem) == 1:
                m[un[0]] = next(iter(rem))
                changed = True
    n_ids = len(set().union(*[want[s][2] for s in want]))
    assert len(m) == n_ids and len(set(m.values())) == n_ids and n_ids >= 15
    assert all(set(m[x] for x in ours[s][2]) == want[s][2] for s in want)


def test_cli_single_thread_branch_has_no_phasing(sim_cli, tmp_path):
    """`-c 1` (the reference's default): the single-thread branch of search() (src/Ratatosk.cpp:660-703) corrects pass-2 reads
    without phasing(); files equal the reference CLI's `correct -2 -c 1` golden, and with --force-correct-snp a fresh reference run"""
    d = os.path.join(GOLDEN, "F1")
    tmp = str(tmp_path)
    p1, raw = os.path.join(tmp, "p1.fastq"), os.path.join(tmp, "raw.fastq")
    _head_fastq(os.path.join(d, "corrected_pass1.fastq.gz"), p1, 12)
    _head_fastq(os.path.join(d, "reads.fastq.gz"), raw, 12)
    idx = ["-g", os.path.join(d, "index.k63.fasta.gz"), "-d", os.path.join(d, "index.k63.rtsk")]
    _run(sim_cli, ["-2", "-c", "1"] + idx + ["-l", p1, "-L", raw, "-o", os.path.join(tmp, "st")])
    got = open(os.path.join(tmp, "st.fastq"), "rb").read()
    assert got == _golden_bytes(os.path.join(d, "corrected_pass2_nophasing.fastq.gz"), 12)
    assert got != _golden_bytes(os.path.join(d, "corrected_pass2.fastq.gz"), 12)          # phasing changes these reads
    if os.path.exists(REF_CLI):
        d2 = os.path.join(GOLDEN, "F2")
        _head_fastq(os.path.join(d2, "corrected_pass1.fastq.gz"), p1, 8)
        _head_fastq(os.path.join(d2, "reads.fastq.gz"), raw, 8)
        idx2 = ["-g", os.path.join(d2, "index.k63.fasta.gz"), "-d", os.path.join(d2, "index.k63.rtsk")]
        args = ["-2", "-c", "1", "--force-correct-snp"] + idx2 + ["-l", p1, "-L", raw]
        _run(sim_cli, args + ["-o", os.path.join(tmp, "fs")])
        subprocess.check_call([REF_CLI, "correct"] + args + ["-o", os.path.join(tmp, "ref_fs")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        assert open(os.path.join(tmp, "fs.fastq"), "rb").read() == open(os.path.join(tmp, "ref_fs.fastq"), "rb").read()


def _walk_reads(path, fa, seed, n_reads=60):
    """long reads spelled along random walks of the k = 63 graph (both orientations, through short unitigs and cycles), with a
    sprinkle of substitutions and '!' qualities"""
    import numpy as np
    import ratatosk_b200 as rb
    from common import revcomp
    sim = os.path.join(HERE, "hostsim", "_build", "librtk_hostsim.so")
    g = rb.Graph.load(fa, "", 63, lib=sim)
    n = g.info()["n_unitigs"]
    rng = np.random.RandomState(seed)
    with open(path, "w") as f:
        for r in range(n_reads):
            u, s = int(rng.randint(n)), int(rng.randint(2))
            seq = g.unitig_seq(u) if s else revcomp(g.unitig_seq(u))
            target = int(rng.randint(2500, 20000))
            while len(seq) < target:
                adj = g.unitig_words(u)[2]
                nxt = [adj[b] if s else adj[4 + (3 - b)] for b in range(4)]
                nxt = [x for x in nxt if x != 0xFFFFFFFF]
                if not nxt:
                    break
                x = nxt[int(rng.randint(len(nxt)))]
                u, s = x & 0x7fffffff, (x >> 31) if s else 1 - (x >> 31)
                t = g.unitig_seq(u) if s else revcomp(g.unitig_seq(u))
                assert seq[-62:] == t[:62]
                seq += t[62:]
            seq, q = list(seq), ["I"] * len(seq)
            for j in rng.randint(0, len(seq), max(1, len(seq) // 1500)):
                seq[j] = "ACGT"[("ACGT".index(seq[j]) + 1 + int(rng.randint(3))) % 4]
            for j in rng.randint(0, len(seq), max(1, len(seq) // 300)):
                q[j] = "!"
            f.write("@w%d\n%s\n+\n%s\n" % (r, "".join(seq), "".join(q)))
    g.close()


@pytest.mark.parametrize("seed", [11, 12])
def test_cli_index2_matches_reference_cli_on_graph_walk_reads(seed, sim_cli, tmp_path):
    """`rtk_correct index2` against a fresh `Ratatosk index -2` on reads spelled along random walks of the graph"""
    import refseams
    if not (refseams.available() and os.path.exists(REF_CLI)):
        pytest.skip("reference not built")
    tmp = str(tmp_path)
    fa = os.path.join(GOLDEN, "F2", "index.k63.fasta.gz")
    reads = os.path.join(tmp, "reads.fastq")
    _walk_reads(reads, fa, seed)
    subprocess.check_call([REF_CLI, "index", "-2", "-c", "4", "-g", fa, "-l", reads, "-o", os.path.join(tmp, "ref")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    r = subprocess.run([sim_cli, "index2", "-g", fa, "-l", reads, "-o", os.path.join(tmp, "ours")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert r.returncode == 0, r.stdout.decode(errors="replace")[-2000:]
    ours = _dump_records(fa, os.path.join(tmp, "ours.index.k63.rtsk"), 63, tmp, "ours")
    want = _dump_records(fa, os.path.join(tmp, "ref.index.k63.rtsk"), 63, tmp, "want")
    bad = [s[:24] for s in want if ours[s][:2] != want[s][:2] or ours[s][3:] != want[s][3:] or len(ours[s][2]) != len(want[s][2])]
    assert not bad, (len(bad), bad[:5], [(ours[s][:2], want[s][:2], ours[s][3:], want[s][3:]) for s in want if s[:24] in bad[:2]])
    m, changed = {}, True
    while changed:
        changed = False
        for s in want:
            a, b = ours[s][2], want[s][2]
            un = [x for x in a if x not in m]
            rem = b - set(m[x] for x in a if x in m)
            if len(un) == 1 and len(rem) == 1:
                m[un[0]] = next(iter(rem))
                changed = True
    n_ids = len(set().union(*[want[s][2] for s in want]))
    assert len(m) == n_ids and n_ids >= 30
    assert all(set(m[x] for x in ours[s][2]) == want[s][2] for s in want)


def _ont_walk_reads(path, fa, k, seed, n_reads):
    """ONT-like reads (3 % substitutions, 3 % insertions, 4 % deletions, Q5-29) of sequences spelled along random walks of the graph"""
    import numpy as np
    import ratatosk_b200 as rb
    from common import revcomp
    sim = os.path.join(HERE, "hostsim", "_build", "librtk_hostsim.so")
    g = rb.Graph.load(fa, "", k, lib=sim)
    n = g.info()["n_unitigs"]
    rng = np.random.RandomState(seed)
    with open(path, "w") as f:
        for r in range(n_reads):
            u, s = int(rng.randint(n)), int(rng.randint(2))
            seq = g.unitig_seq(u) if s else revcomp(g.unitig_seq(u))
            target = int(rng.randint(1500, 9000))
            while len(seq) < target:
                adj = g.unitig_words(u)[2]
                nxt = [x for x in (adj[b] if s else adj[4 + (3 - b)] for b in range(4)) if x != 0xFFFFFFFF]
                if not nxt:
                    break
                x = nxt[int(rng.randint(len(nxt)))]
                u, s = x & 0x7fffffff, (x >> 31) if s else 1 - (x >> 31)
                seq += (g.unitig_seq(u) if s else revcomp(g.unitig_seq(u)))[k - 1:]
            out = []
            for c in seq[:target]:
                p = rng.rand()
                if p < 0.03:
                    out.append("ACGT"[int(rng.randint(4))])
                elif p < 0.06:
                    out.append(c)
                    out.append("ACGT"[int(rng.randint(4))])
                elif p >= 0.10:
                    out.append(c)
            q = "".join(chr(33 + int(x)) for x in rng.randint(5, 30, len(out)))
            f.write("@g%d\n%s\n+\n%s\n" % (r, "".join(out), q))
    g.close()


@pytest.mark.parametrize("seed", [21, 22])
def test_cli_fresh_graph_walk_reads_match_fresh_reference_runs(seed, sim_cli, tmp_path):
    """both passes on reads no fixture holds: noisy reads along random walks of the F2 graph (repeats, bubbles, cycles), the host
    driver on the kernel-source simulator against the reference CLI run here on the same file; byte-identical outputs.  The
    reference itself can flip a colour-set tie from process to process, so a difference is re-checked against two more
    reference runs before it counts"""
    if not os.path.exists(REF_CLI):
        pytest.skip("reference CLI not built")
    d, tmp = os.path.join(GOLDEN, "F2"), str(tmp_path)
    reads = os.path.join(tmp, "reads.fastq")
    _ont_walk_reads(reads, os.path.join(d, "index.k31.fasta.gz"), 31, seed, 40)
    i1 = ["-g", os.path.join(d, "index.k31.fasta.gz"), "-d", os.path.join(d, "index.k31.rtsk")]
    i2 = ["-g", os.path.join(d, "index.k63.fasta.gz"), "-d", os.path.join(d, "index.k63.rtsk")]

    def ref(args, out_file):
        subprocess.check_call([REF_CLI, "correct"] + args, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return open(out_file, "rb").read()

    def same_as_some_reference_run(ours, args, out_file):
        return any(ours == ref(args, out_file) for _ in range(3))

    o, r = os.path.join(tmp, "ours"), os.path.join(tmp, "ref")
    _run(sim_cli, ["-1", "--no-cache"] + i1 + ["-l", reads, "-o", o])
    assert same_as_some_reference_run(open(o + ".2.fastq", "rb").read(), ["-1", "-c", "4"] + i1 + ["-l", reads, "-o", r], r + ".2.fastq")
    _run(sim_cli, ["-2", "-O", "-c", "8", "--no-cache"] + i2 + ["-l", o + ".2.fastq", "-L", reads, "-o", o])
    assert same_as_some_reference_run(open(o + ".fastq", "rb").read(), ["-2", "-O", "-c", "4"] + i2 + ["-l", o + ".2.fastq", "-L", reads, "-o", r], r + ".fastq")


def test_cli_odd_records_match_fresh_reference_runs(sim_cli, tmp_path):
    """records the fixtures do not hold - shorter than k, exactly k - 1 / k / 2k, all N, a homopolymer, header comments, FASTA input
    - next to ordinary noisy reads: pass 1 (FASTQ and FASTA), pass 2 and the two-pass mode against the reference CLI run here"""
    if not os.path.exists(REF_CLI):
        pytest.skip("reference CLI not built")
    d, tmp = os.path.join(GOLDEN, "F2"), str(tmp_path)
    walk = os.path.join(tmp, "walk.fastq")
    _ont_walk_reads(walk, os.path.join(d, "index.k31.fasta.gz"), 31, 31, 8)
    w = open(walk).read().split("\n")
    recs = [("short", "ACGTACGTAC", "I" * 10), ("k30", "A" * 30, "I" * 30), ("k31", "ACGT" * 7 + "ACG", "I" * 31), ("allN", "N" * 500, "#" * 500)]
    recs += [(w[i][1:], w[i + 1], w[i + 3]) for i in range(0, len(w) - 3, 4)]
    recs += [("k62", "ACGT" * 15 + "AC", "I" * 62), ("poly", "A" * 3000, "5" * 3000)]
    fq, fa = os.path.join(tmp, "odd.fastq"), os.path.join(tmp, "odd.fasta")
    with open(fq, "w") as f:
        for n, s, q in recs:
            f.write("@%s extra comment\n%s\n+\n%s\n" % (n, s, q))
    with open(fa, "w") as f:
        for n, s, q in recs:
            f.write(">%s\n%s\n" % (n, s))
    i1 = ["-g", os.path.join(d, "index.k31.fasta.gz"), "-d", os.path.join(d, "index.k31.rtsk")]
    i2 = ["-g", os.path.join(d, "index.k63.fasta.gz"), "-d", os.path.join(d, "index.k63.rtsk")]
    o, r = os.path.join(tmp, "ours"), os.path.join(tmp, "ref")

    def ref(args):
        subprocess.check_call([REF_CLI, "correct"] + args, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)

    for inp in (fa, fq):
        _run(sim_cli, ["-1", "--no-cache"] + i1 + ["-l", inp, "-o", o])
        ref(["-1", "-c", "4"] + i1 + ["-l", inp, "-o", r])
        assert open(o + ".2.fastq", "rb").read() == open(r + ".2.fastq", "rb").read(), inp
    _run(sim_cli, ["-2", "-O", "-c", "8", "--no-cache"] + i2 + ["-l", r + ".2.fastq", "-L", fq, "-o", o])
    ref(["-2", "-O", "-c", "4"] + i2 + ["-l", r + ".2.fastq", "-L", fq, "-o", r])
    want = open(r + ".fastq", "rb").read()
    assert open(o + ".fastq", "rb").read() == want
    _run(sim_cli, ["--no-cache"] + i1 + ["--in-graph2", i2[1], "--in-unitig-data2", i2[3], "-l", fq, "-o", os.path.join(tmp, "tp")])
    assert open(os.path.join(tmp, "tp.fastq"), "rb").read() == want
