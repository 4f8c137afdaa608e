"""k values other than the defaults (31 / 63): the reference builds an index of the F1 recipe with -k 25 -K 47 and with -k 21 -K 33
(33: the smallest k on the 128-bit k-mer path), then both correction passes, the annotation steps and the colouring are compared
with it - the fixtures only hold k = 31 / 63.  CPU only (kernel sources on the simulator); needs the reference binary (oracle/_ref)."""
import os
import subprocess
import sys

import pytest

import ratatosk_b200 as rb
from common import HERE, ROOT, ensure_built

REF_CLI = os.path.join(ROOT, "oracle", "_ref", "Ratatosk")
SIM_CLI = os.path.join(HERE, "hostsim", "_build", "rtk_correct_sim")
pytestmark = pytest.mark.skipif(not os.path.exists(REF_CLI), reason="reference CLI not built (oracle/_ref)")


def _ref(args, cwd):
    subprocess.check_call([REF_CLI] + args, cwd=cwd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def _ours(args, cwd):
    r = subprocess.run([SIM_CLI] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert r.returncode == 0, r.stdout.decode(errors="replace")[-2000:]


@pytest.fixture(scope="module")
def recipe(tmp_path_factory):
    ensure_built()
    d = str(tmp_path_factory.mktemp("otherk"))
    subprocess.check_call([sys.executable, os.path.join(HERE, "golden", "make_fixtures.py"), "--recipe", "F1", "--out", os.path.join(d, "F1")],
                          stdout=subprocess.DEVNULL)
    with open(os.path.join(d, "F1.lr.fastq")) as f, open(os.path.join(d, "lr.fastq"), "w") as o:
        o.writelines(f.readlines()[:4 * 30])
    return d


@pytest.mark.parametrize("k1,k2", [(25, 47), (21, 33)])
def test_other_k_values_match_reference(k1, k2, recipe, sim_lib):
    d = recipe
    ks = ["-k", str(k1), "-K", str(k2)]
    _ref(["index", "-1", "-c", "4"] + ks + ["-s", "F1.sr.fastq", "-l", "F1.lr.fastq", "-o", "i%d" % k1], d)
    g1, d1 = "i%d.index.k%d.fasta.gz" % (k1, k1), "i%d.index.k%d.rtsk" % (k1, k1)
    g2 = "i%d.index.k%d.fasta.gz" % (k1, k2)
    # pass 1
    _ours(["correct", "-1", "--no-cache"] + ks + ["-g", g1, "-d", d1, "-l", "lr.fastq", "-o", "o%d" % k1], d)
    _ref(["correct", "-1", "-c", "4"] + ks + ["-g", g1, "-d", d1, "-l", "lr.fastq", "-o", "r%d" % k1], d)
    p1 = open(os.path.join(d, "r%d.2.fastq" % k1), "rb").read()
    assert open(os.path.join(d, "o%d.2.fastq" % k1), "rb").read() == p1 and len(p1) > 100000
    # second index by the reference, pass 2
    _ref(["index", "-2", "-c", "4"] + ks + ["-g", g2, "-l", "r%d.2.fastq" % k1, "-o", "j%d" % k1], d)
    d2 = "j%d.index.k%d.rtsk" % (k1, k2)
    _ours(["correct", "-2", "-O", "-c", "8", "--no-cache"] + ks + ["-g", g2, "-d", d2, "-l", "r%d.2.fastq" % k1, "-L", "lr.fastq", "-o", "o%d" % k1], d)
    _ref(["correct", "-2", "-O", "-c", "4"] + ks + ["-g", g2, "-d", d2, "-l", "r%d.2.fastq" % k1, "-L", "lr.fastq", "-o", "r%d" % k1], d)
    assert open(os.path.join(d, "o%d.fastq" % k1), "rb").read() == open(os.path.join(d, "r%d.fastq" % k1), "rb").read()
    # annotation of both indexes against what the reference stored
    for k, fa, rt in ((k1, g1, d1), (k2, g2, d2)):
        g = rb.Graph.load(os.path.join(d, fa), os.path.join(d, rt), k, lib=sim_lib)
        ctx = rb.Context(0, lib=sim_lib)
        ctx.upload(g)
        n = g.info()["n_unitigs"]
        off, ids = ctx.detect_snps()
        flags, coff, pool = ctx.detect_short_cycles()
        for u in range(n):
            amb, blob = g.unitig_annotations(u)
            assert list(map(int, ids[int(off[u]):int(off[u + 1])])) == amb and pool[int(coff[u]):int(coff[u + 1])] == blob, (k, u)
            assert int(flags[u]) == (g.unitig_words(u)[1] >> 8) & 1, (k, u)
        ctx.close()
        g.close()
    # colouring of the k2 graph: coverage words and edge flags as in the reference's second index
    from common import read_fastq
    recs = read_fastq(os.path.join(d, "r%d.2.fastq" % k1))
    gs = rb.Graph.load(os.path.join(d, g2), "", k2, lib=sim_lib)
    want = rb.Graph.load(os.path.join(d, g2), os.path.join(d, d2), k2, lib=sim_lib)
    ctx = rb.Context(0, lib=sim_lib)
    ctx.upload(gs)
    kmcov, shared, off, ids, _ = ctx.color_long_reads([r[1] for r in recs], [r[2] for r in recs], [r[0] for r in recs])
    for u in range(gs.info()["n_unitigs"]):
        kc, sh, _ = want.unitig_words(u)
        a, b = want.unitig_colors(u)
        assert int(kmcov[u]) == kc and (int(shared[u]) & 0xff) == (sh & 0xff) and int(off[u + 1] - off[u]) == len(set(a) | set(b)), (k2, u)
    ctx.close(); gs.close(); want.close()
