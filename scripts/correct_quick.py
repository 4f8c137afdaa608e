"""Quick end-to-end correction timing + parity on a golden recipe (F1 / F2) or on bench_data/F3 (development aid)."""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ratatosk_b200 as rb
from common import GOLDEN, golden_paths, load_golden_reads, read_fastq

recipe = sys.argv[1] if len(sys.argv) > 1 else "F2"
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 8
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
if recipe == "F3":
    d = os.path.join(ROOT, "bench_data", "F3")
    fa, rt = os.path.join(d, "index.k31.fasta.gz"), os.path.join(d, "index.k31.rtsk")
    reads = read_fastq(os.path.join(d, "reads200.fastq.gz"))
    gold = read_fastq(os.path.join(d, "corrected200_pass1.fastq.gz"))
else:
    fa, rt = golden_paths(recipe)
    reads = load_golden_reads(recipe)
    gold = read_fastq(os.path.join(GOLDEN, recipe, "corrected_pass1.fastq.gz"))
t = time.time()
g = rb.Graph.load(fa, rt, 31)
ctx = rb.Context(0)
ctx.upload(g)
print("graph load + upload %.2fs" % (time.time() - t), g.info())
seqs = [r[1] for r in reads] * rep
quals = [r[2] for r in reads] * rep
nb = sum(len(s) for s in seqs)
for it in range(iters):
    st = []
    t = time.time()
    out = ctx.correct(seqs, quals, stats=st)
    dt = time.time() - t
    bad = [i for i in range(len(out)) if (out[i][0], out[i][1]) != (gold[i % len(reads)][1], gold[i % len(reads)][2])]
    print("correct: reads=%d bases=%d time=%.2fs -> %.3f Mbases/s  waves=%d jobs=%d  identical_to_reference=%s %s" %
          (len(seqs), nb, dt, nb / dt / 1e6, st[5], st[6], not bad, bad[:10]))
