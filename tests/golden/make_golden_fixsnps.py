#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — golden vectors for fixSNPs (src/Alignment.cpp:846; `Ratatosk correct -2 -f`), recorded from the
UNMODIFIED reference (oracle/_ref: seam probe ref_fix_snps + the CLI).  Run in the build container only.

  <fixture>/fixsnps.json.gz                  per pass-1 read that changes: [[position, base], ...]  (k = 63 graph)
  <fixture>/corrected_pass2_forcesnp.fastq.gz  `Ratatosk correct -2 -O --force-correct-snp -c 8` (the short form -f is missing from the reference's getopt string, src/Ratatosk.cpp:149) (fixSNPs -> phasing -> getSeeds -> correctSequence)
"""
import gzip
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import read_fastq  # noqa: E402
from refseams import RefGraph  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "Ratatosk")
FIX = {"F1": (os.path.join(HERE, "F1"), "corrected_pass1.fastq.gz", "reads.fastq.gz"),
       "F2": (os.path.join(HERE, "F2"), "corrected_pass1.fastq.gz", "reads.fastq.gz"),
       "F3": (os.path.join(ROOT, "bench_data", "F3"), "corrected200_pass1.fastq.gz", "reads200.fastq.gz")}

for name, (d, p1, raw) in FIX.items():
    g = RefGraph(os.path.join(d, "index.k63.fasta.gz"), os.path.join(d, "index.k63.rtsk"), 63)
    reads = read_fastq(os.path.join(d, p1))
    out, n_amb, n_fix = {}, 0, 0
    for i, (_, s, _) in enumerate(reads):
        n_amb += sum(c not in "ACGT" for c in s)
        f = g.fix_snps(s)
        assert len(f) == len(s)
        ch = [[j, f[j]] for j in range(len(s)) if f[j] != s[j]]
        if ch:
            out[str(i)] = ch
            n_fix += len(ch)
    with gzip.open(os.path.join(d, "fixsnps.json.gz"), "wt") as fo:
        json.dump({"n_reads": len(reads), "n_ambiguous": n_amb, "n_fixed": n_fix, "changes": out}, fo)
    print(name, "reads", len(reads), "ambiguity codes", n_amb, "fixed", n_fix)
    with tempfile.TemporaryDirectory() as t:
        a, b = os.path.join(t, "p1.fastq"), os.path.join(t, "raw.fastq")
        open(a, "wb").write(gzip.open(os.path.join(d, p1), "rb").read())
        open(b, "wb").write(gzip.open(os.path.join(d, raw), "rb").read())
        subprocess.check_call([REF, "correct", "-2", "-O", "--force-correct-snp", "-c", "8", "-g", os.path.join(d, "index.k63.fasta.gz"), "-d",
                               os.path.join(d, "index.k63.rtsk"), "-l", a, "-L", b, "-o", os.path.join(t, "o")],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        with gzip.GzipFile(os.path.join(d, "corrected_pass2_forcesnp.fastq.gz"), "wb", 9, mtime=0) as fo:
            fo.write(open(os.path.join(t, "o.fastq"), "rb").read())
