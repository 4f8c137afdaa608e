#!/usr/bin/env python3
"""Freeze reference outputs into committed golden fixtures (run in the build container only).

Needs /root/reference (compiled by `make -C oracle ref` into oracle/_ref/).  For each recipe:
  1. generate the seeded synthetic data (make_fixtures.py),
  2. build the pass-1 index ONCE with the reference binary (`Ratatosk index -1`) - the index
     is not reproducible (std::random_device, src/Graph.cpp:2093), so the index files
     themselves are committed,
  3. take a subset of the long reads and record, from the UNMODIFIED reference objects
     (oracle/_ref/libref_seams.so):
        searchSequence exact / inexact hit lists     (Bifrost/src/Search.tcc:526)
        getSeeds solid / weak anchors                (src/Graph.cpp:3)
        pass-1 corrected read + quality              (src/Ratatosk.cpp:808-867)
     plus the per-unitig dump used to check the index loader,
  4. record md5 of the reference CLI's full pass-1 output (`Ratatosk correct -1 -c 2`).
Unitigs are identified by their rank in the committed index FASTA, which is how the product
numbers them.
"""
import argparse, gzip, hashlib, json, os, shutil, subprocess, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refseams as R  # noqa: E402

RATATOSK = os.path.join(ROOT, "oracle", "_ref", "Ratatosk")


def read_fastq(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        L = f.read().split("\n")
    return [(L[i][1:], L[i + 1], L[i + 3]) for i in range(0, len(L) - 3, 4)]


def read_fasta_gz(path):
    seqs = []
    with gzip.open(path, "rt") as f:
        for line in f:
            if line.startswith(">"):
                seqs.append("")
            else:
                seqs[-1] += line.strip()
    return seqs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--recipe", required=True, choices=["F1", "F2"])
    ap.add_argument("--n-reads", type=int, default=24)
    ap.add_argument("--work", default="/tmp/rtk_golden")
    a = ap.parse_args()
    work = os.path.join(a.work, a.recipe)
    os.makedirs(work, exist_ok=True)
    out = os.path.join(HERE, a.recipe)
    os.makedirs(out, exist_ok=True)
    pre = os.path.join(work, a.recipe)
    subprocess.check_call([sys.executable, os.path.join(HERE, "make_fixtures.py"), "--recipe", a.recipe, "--out", pre])
    subprocess.check_call([RATATOSK, "index", "-1", "-c", "4", "-s", pre + ".sr.fastq", "-l", pre + ".lr.fastq", "-o", pre + "i"],
                          stdout=subprocess.DEVNULL)
    fasta, rtsk = pre + "i.index.k31.fasta.gz", pre + "i.index.k31.rtsk"
    # full-run md5 of the reference CLI from this index (pass 1; -c 2 because -c 1 is broken)
    subprocess.check_call([RATATOSK, "correct", "-1", "-c", "2", "-g", fasta, "-d", rtsk, "-l", pre + ".lr.fastq", "-o", pre + "o"],
                          stdout=subprocess.DEVNULL)
    full_md5 = hashlib.md5(open(pre + "o.2.fastq", "rb").read()).hexdigest()
    full = read_fastq(pre + "o.2.fastq")

    shutil.copy(fasta, os.path.join(out, "index.k31.fasta.gz"))
    shutil.copy(rtsk, os.path.join(out, "index.k31.rtsk"))
    reads = read_fastq(pre + ".lr.fastq")
    # spread the subset over the file; always keep the shortest and the longest read
    idx = sorted(set(list(range(0, len(reads), max(1, len(reads) // a.n_reads)))[:a.n_reads]
                     + [min(range(len(reads)), key=lambda i: len(reads[i][1])),
                        max(range(len(reads)), key=lambda i: len(reads[i][1]))]))
    sub = [reads[i] for i in idx]
    with gzip.open(os.path.join(out, "reads.fastq.gz"), "wt") as f:
        for n, s, q in sub:
            f.write("@%s\n%s\n+\n%s\n" % (n, s, q))

    g = R.RefGraph(fasta, rtsk, 31)
    dump = os.path.join(work, "dump.txt")
    g.dump(dump)
    seqs = read_fasta_gz(fasta)
    rc = lambda s: s[::-1].translate(str.maketrans("ACGT", "TGCA"))
    seq2id = {}
    for i, s in enumerate(seqs):
        seq2id[s] = i
        seq2id.setdefault(rc(s), i)
    key2id = {}
    with open(dump) as f, gzip.open(os.path.join(out, "ref_unitigs.tsv.gz"), "wt") as fo:
        for line in f:
            c = line.rstrip("\n").split("\t")
            key2id[int(c[0])] = seq2id[c[1]]
            fo.write("\t".join([str(seq2id[c[1]])] + c[1:]) + "\n")

    def conv(h):
        return np.array([(p, key2id[u], d, s) for (p, u, d, l, sz, s) in h], dtype=np.uint32).reshape(-1, 4)

    arrays = {}
    corrected = []
    for i, (n, s, q) in enumerate(sub):
        arrays["exact_%d" % i] = conv(g.search_sequence(s, 1, 0, 0, 0, 0))
        arrays["inexact_%d" % i] = conv(g.search_sequence(s, 0, 1, 1, 1, 1))
        so, we = g.get_seeds(s, q, False)
        arrays["solid_%d" % i] = conv(so)
        arrays["weak_%d" % i] = conv(we)
        cs, cq = g.correct_read(s, q, False)
        assert (n, cs, cq) == full[idx[i]], "seam output differs from CLI output for read %s" % n
        corrected.append((n, cs, cq))
    np.savez_compressed(os.path.join(out, "golden_hits.npz"), **arrays)
    with gzip.open(os.path.join(out, "corrected_pass1.fastq.gz"), "wt") as f:
        for n, s, q in corrected:
            f.write("@%s\n%s\n+\n%s\n" % (n, s, q))
    meta = {"recipe": a.recipe, "k": 31, "n_unitigs": int(g.num_unitigs()), "max_km_cov": int(g.max_km_cov()),
            "n_reads_total": len(reads), "subset_indices": idx, "reference_cli_pass1_md5_full": full_md5,
            "reference": "DecodeGenetics/Ratatosk @156b750, Bifrost @d2ff315, g++ -O3 -mno-avx2 (oracle/Makefile)"}
    json.dump(meta, open(os.path.join(out, "meta.json"), "w"), indent=1)
    print(json.dumps(meta))


if __name__ == "__main__":
    main()
