import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ratatosk_b200 as rb
import refseams as R
from common import read_fastq
d = os.path.join(ROOT, "bench_data", "F3")
g = rb.Graph.load(os.path.join(d, "index.k63.fasta.gz"), os.path.join(d, "index.k63.rtsk"), 63)
ctx = rb.Context(0); ctx.upload(g)
raw = read_fastq(os.path.join(d, "reads200.fastq.gz")); p1 = read_fastq(os.path.join(d, "corrected200_pass1.fastq.gz"))
idx = [14, 30, 73, 76, 0, 1]
qs = [raw[i][1].upper() for i in idx]; ts = [p1[i][1] for i in idx]
dist, ends, ops, _flags = ctx.edlib_path_batch(qs, ts, [0] * len(idx))
for j, i in enumerate(idx):
    rd, re, rs, raln = R.edlib(qs[j], ts[j], 0, task=2)
    o = bytes(ops[j])
    same = (o == raln)
    fd = next((x for x in range(min(len(o), len(raln))) if o[x] != raln[x]), None)
    print("read", i, "qlen", len(qs[j]), "tlen", len(ts[j]), "dist ours/ref", dist[j], rd, "ops len", len(o), len(raln), "identical", same, "first diff at op", fd)
    if fd is not None:
        print("   ours", list(o[max(0, fd - 5):fd + 12]), " ref", list(raln[max(0, fd - 5):fd + 12]))
