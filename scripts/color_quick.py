#!/usr/bin/env python
"""DEVELOPMENT AID: rtk_color_long_reads of the product library on bench_data/F3 (k = 63 graph + the 200 pass-1 corrected reads the
reference coloured it with), timed, words checked against the reference's index.  -> one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ratatosk_b200 as rb
from common import read_fastq

lib = sys.argv[1] if len(sys.argv) > 1 else None
d = os.path.join(ROOT, "bench_data", "F3")
g = rb.Graph.load(os.path.join(d, "index.k63.fasta.gz"), "", 63, lib=lib)
want = rb.Graph.load(os.path.join(d, "index.k63.fasta.gz"), os.path.join(d, "index.k63.rtsk"), 63, lib=lib)
ctx = rb.Context(0, lib=lib)
ctx.upload(g)
recs = read_fastq(os.path.join(d, "corrected200_pass1.fastq.gz"))
seqs, quals, names = [r[1] for r in recs], [r[2] for r in recs], [r[0] for r in recs]
for rep in range(2):
    st = [0] * 10
    t = time.time()
    kmcov, shared, off, ids, rid = ctx.color_long_reads(seqs, quals, names, stats=st)
    dt = time.time() - t
n = g.info()["n_unitigs"]
ok = all(int(kmcov[u]) == want.unitig_words(u)[0] and (int(shared[u]) & 0xff) == (want.unitig_words(u)[1] & 0xff) for u in range(n))
print(json.dumps({"n_unitigs": n, "reads": len(recs), "bases": sum(map(len, seqs)), "color_s": round(dt, 4), "k1_probes": st[0], "k1_kernel_ms": st[2] / 1e6,
                  "k1_stage_ms": st[3] / 1e6, "unitig_read_pairs": st[4], "ids": st[5], "edge_flag_kernel_ms": st[7] / 1e6, "hap_cov": st[8],
                  "words_identical": ok}))
